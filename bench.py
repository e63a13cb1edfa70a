#!/usr/bin/env python
"""Benchmark of the signature-k-mer calling path (BASELINE.json metric: proteins/s and k-mer probes/s per
B200, % of the HBM gather roofline), with the reference's CPU path timed beside it.

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3] [--impl reference] [--skip c3,fq,matrix]

A step = one pass of the hot path (encode -> probe -> ordered scoring scan -> find_best_call) over one batch
of synthetic proteins against a synthetic signature image in the reference's file format.  Default workload
is BASELINE.json configs[1] ("c2": 1M proteins, mean 300 aa, vs a 100M-k-mer image = 508,000,037 buckets).

Keys of the JSON line: see DESIGN.md "Measurement".  `value` is kernel-resident (inputs already in HBM); `e2e` goes
through the C ABI with pinned HOST buffers (H2D + kernels + D2H inside the timed region) -- ckm_call_batch_packed, the
entry point a parser feeds (seven residues per 32-bit word), with `e2e_ascii` (ckm_call_batch, the reference's char* strings) beside it.
Sub-records of the same line: `sparse_signatures` (a signature set that is a sparse subset of the windows: the neighbour copy
pinned on, plain probing, and what the automatic fall-back picks), `c3_stream` (BASELINE configs[2]: 100M proteins streamed against an 80M-k-mer image, strong-scaled
over the ranks), `fq` (configs[3]: reads -> 6 frames -> calling -> family voting -> best frame) and `matrix` (configs[4]: 50k
proteins, row blocks over the ranks, NCCL tile gather).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from close_kmers_b200 import synth  # noqa: E402

WORKLOADS = {
    # name: (n_signature_kmers, n_proteins per GPU per step, prototype length sd, description)
    "c1": (1_000_000, 10_000, 0.0, "10k proteins (300 aa) vs 1M-k-mer image (4,000,037 buckets, 96 MB; fits L2)"),
    "c2": (100_000_000, 1_000_000, 60.0, "1M proteins (mean 300 aa) vs 100M-k-mer image (508,000,037 buckets, 12.19 GB file)"),
    "c3": (80_000_000, 4_000_000, 60.0, "4M-protein batches (of a 100M-protein stream) vs 80M-k-mer image (248,000,009 buckets, 5.95 GB file)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def stop(self, t_begin=None, t_end=None):
        """Statistics over the samples that arrived inside [t_begin, t_end] (the timed regions); the sampler itself is
        started earlier so that nvidia-smi is already reporting when a short timed region begins."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r[1:] for r in self.rows if t_begin is None or t_begin <= r[0] <= t_end + 0.05]
        window = "timed regions"
        if not rows:  # a timed region shorter than one nvidia-smi period: the samples of the same run around it
            rows = [r[1:] for r in self.rows]
            window = "whole run (no sample fell inside the timed regions)"
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def build_world(n_sigs, n_proteins, sd, rank, seed=12345, what="bench"):
    """Seeded synthetic world: prototypes -> (signature k-mers), plus this rank's proteins."""
    t0 = time.time()
    n_protos = max(64, -(-n_sigs // 293) + 8)
    protos = synth.make_prototypes(seed, n_protos, 300, sd)
    batch = synth.make_proteins_parallel(seed + 1 + rank, protos, n_proteins)
    log(f"[{what} r{rank}] prototypes+proteins: {time.time() - t0:.1f}s  ({batch.n} proteins, {batch.residues.nbytes / 1e6:.0f} MB)")
    return protos, batch


def signatures_of(protos, n_sigs):
    sig = synth.make_signatures(protos, n_sigs, dedupe=n_sigs <= 2_000_000)
    return sig, synth.bucket_count(len(sig.keys))


def image_dir(tag, need_bytes):
    for base in ("/dev/shm", "/tmp"):
        try:
            if shutil.disk_usage(base).free > need_bytes * 1.1 + (1 << 28):
                d = os.path.join(base, f"ckm_bench_{tag}")
                os.makedirs(d, exist_ok=True)
                return d
        except OSError:
            pass
    return None


def write_image_dir(kdir, protos, n_sigs, rank, with_reference_builder):
    """<kdir>/kmer.table.mem_map + function.index + otu.index in the reference's format.  The reference arm builds the file
    with the reference's own builder (KmerGuts(dir, nbuckets) + insert_kmer + save_kmer_hash_table, oracle/_ref); our arm
    with the library's host builder (byte-identical, tests/test_oracle_vs_ref.py)."""
    t0 = time.time()
    sig, nb = signatures_of(protos, n_sigs)
    if with_reference_builder:
        import cpu_checkers as cc
        with stdout_to_stderr():
            cc.Ref().build_image(kdir, nb, sig)
    else:
        from close_kmers_b200 import api
        api.build_image(nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt).tofile(os.path.join(kdir, "kmer.table.mem_map"))
    synth.write_index_files(kdir, sig.n_functions, 0)
    log(f"[bench r{rank}] image: {len(sig.keys)} k-mers, {nb} buckets, {24 * nb / 1e9:.2f} GB in {kdir} in {time.time() - t0:.1f}s"
        f" ({'reference builder' if with_reference_builder else 'ckm_image_build'})")
    return sig, nb


def host_info(local):
    """Where this rank's GPU hangs and what the process may run on: the end-to-end figures are host-memory / PCIe figures."""
    info = {"cpus_allowed": cpu_threads()}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        info["gpu_pci"] = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        info["gpu_numa_node"] = int(open(path).read().strip())
    except Exception as e:  # not every box exposes it
        info["gpu_numa_node"] = None
        info["note"] = f"{type(e).__name__}"
    try:
        info["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except OSError:
        pass
    return info


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class stdout_to_stderr:
    """The reference's own code prints to stdout (kmer_image.cc:44, kmer.cc:46); bench.py's stdout carries exactly one
    JSON line, so the CPU arm runs with file descriptor 1 pointed at stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def sub_batch(batch, n):
    n = min(n, batch.n)
    return synth.Batch(batch.residues[: int(batch.offsets[n])], batch.offsets[: n + 1])


class CpuEngine:
    """The reference's CPU path for this workload: oracle/_ref (the reference's own object code, one KmerGuts
    per thread sharing one mmapped image, like threadpool.cc:18-45) when present, else the plain-C port."""

    def __init__(self, kdir, threads):
        import cpu_checkers as cc
        cc.ensure_built()
        self.threads = threads
        self.kind = "reference" if os.path.exists(cc.REF_SO) else "port"
        if self.kind == "reference":
            self.eng = cc.Ref().open(kdir, threads)
            self.eng.set_params()
            self.run = lambda b: self.eng.bench_calls(b, True)
        else:
            self.eng = cc.Oracle()
            self.eng.open(kdir)
            self.run = lambda b: self.eng.bench_calls(b, threads, True)
        self.rate = None

    def sample(self, batch, target_s):
        """One bounded sample sized for ~target_s of wall time on all threads; returns the cpu_baseline dict."""
        if self.rate is None:  # calibrate once on a small slice
            self.run(sub_batch(batch, 50 * self.threads))
            cal_n = min(batch.n, 500 * self.threads)
            self.rate = cal_n / max(self.run(sub_batch(batch, cal_n)), 1e-6)
        n = int(max(min(batch.n, 500 * self.threads), min(batch.n, self.rate * target_s)))
        b = sub_batch(batch, n)
        t = self.run(b)
        probes = synth.n_probes_expected(b) if n <= 200_000 else None
        return {"value": n / t, "unit": "proteins/s", "cores": self.threads, "kind": self.kind,
                "sample": f"first {n} proteins of the same batch, {t:.2f}s wall on {self.threads} threads "
                          f"(one KmerGuts per thread, shared image, process_aa_seq + find_best_call)",
                "probes_per_s": (probes / t) if probes else None, "seconds": t, "n": n}

    def close(self):
        self.eng.close()


_JSON_FD = None


def claim_stdout():
    """bench.py's stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner, the
    reference's own std::cout lines), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is
    written to a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


class Pinned:
    """A batch in page-locked host memory: ASCII residues + offsets, and the packed form (seven residues per word) + word offsets."""

    def __init__(self, api, batch, packed=True):
        self.api, self.L = api, api.lib()
        self.n, self.total = batch.n, int(batch.offsets[-1])
        self.bufs = []
        self.res = self._copy(batch.residues, self.total + 64)
        self.off = self._copy(batch.offsets.astype(np.uint64), (self.n + 1) * 8)
        self.pk = self.woff = None
        if packed:
            pk, woff = api.pack_residues(batch.residues, batch.offsets)
            self.words = int(woff[-1])
            self.pk = self._copy(pk, (self.words + 2) * 4)
            self.woff = self._copy(woff, (self.n + 1) * 8)

    def _copy(self, arr, nbytes):
        p = C.c_void_p()
        self.api._check(self.L.ckm_host_alloc(C.byref(p), max(nbytes, 64)))
        a = np.ascontiguousarray(arr)
        C.memmove(p.value, a.ctypes.data, a.nbytes)
        self.bufs.append(p)
        return p.value

    def free(self):
        for p in self.bufs:
            self.L.ckm_host_free(p)
        self.bufs = []


def timed_host_calls(fn, reps, barrier):
    for _ in range(2):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return time.perf_counter() - t0, out


def allreduce(vals, op, world, torch, dist):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return [float(x) for x in t]


# ----------------------------------------------------------------------------------------------------------------------
# C3: a 100M-protein stream against the ~6 GB image, strong-scaled over the ranks, end to end through the C ABI
# ----------------------------------------------------------------------------------------------------------------------
def run_c3_stream(api, torch, dist, rank, world, local, barrier, peak, args, tag):
    n_sigs, _, sd, _ = WORKLOADS["c3"]
    n_sigs = args.c3_sigs or n_sigs
    total_proteins = args.c3_proteins
    pool_n = min(args.c3_pool, max(1, total_proteins // world))
    protos, pool = build_world(n_sigs, pool_n, sd, rank, seed=54321, what="c3")
    kdir = image_dir(tag + "_c3", 24 * synth.bucket_count(n_sigs))
    if rank == 0:
        write_image_dir(kdir, protos, n_sigs, rank, False)
    barrier()
    guts = api.KmerGuts(kmer_dir=kdir, device=local)
    guts.set_default_parameters()
    pin = Pinned(api, pool)
    mine = total_proteins // world + (1 if rank < total_proteins % world else 0)  # this rank's share of the stream
    passes, rem = divmod(mine, pool_n)

    def one_pass(n=pool_n):
        return guts.call_batch_packed_raw(pin.pk, pin.woff, n, api.WANT_BEST)

    for _ in range(2):
        out = one_pass()
    probes_pool = int(out.n_probes)
    barrier()
    t0 = time.perf_counter()
    for _ in range(passes):
        one_pass()
    probes_rem = 0
    if rem:
        probes_rem = int(one_pass(rem).n_probes)
    barrier()
    wall = time.perf_counter() - t0
    wall_max, = allreduce([wall], "MAX", world, torch, dist)
    probes_all, prot_all, bytes_all = allreduce([passes * probes_pool + probes_rem, mine,
                                                 passes * (pin.words * 4 + (pool_n + 1) * 8)], "SUM", world, torch, dist)
    rec = None
    if rank == 0:
        alg = 32.0 * probes_all + prot_all * 300.0
        rec = {"workload": f"c3: {total_proteins} proteins (a {pool_n}-protein pool per rank, cycled) vs {n_sigs}-k-mer image "
                           f"({guts.num_sigs} buckets), strong-scaled: every rank streams 1/{world} of the proteins",
               "entry": "ckm_call_batch_packed, pinned host buffers, WANT_BEST (H2D + unpack + K1 + D2H of 28-byte best calls in the timed region)",
               "value": prot_all / wall_max, "unit": "proteins/s", "probes_per_s": probes_all / wall_max, "wall_s": wall_max,
               "proteins": prot_all, "scaling": "strong", "timing": "wall clock, barrier on both sides, max over ranks",
               "h2d_GBps_whole_job": bytes_all / wall_max / 1e9,
               "roofline": {"bound": "hbm", "achieved": alg / wall_max / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                            "frac": alg / wall_max / 1e9 / world / peak,
                            "note": "end-to-end wall time (PCIe included), not kernel time: a lower bound of the kernel's fraction"}}
        if not args.no_cpu_baseline:
            with stdout_to_stderr():
                eng = CpuEngine(kdir, cpu_threads())
                cb = eng.sample(pool, target_s=5.0)
                eng.close()
            rec["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    barrier()
    guts.close()
    pin.free()
    if rank == 0 and kdir:
        shutil.rmtree(kdir, ignore_errors=True)
    return rec


# ----------------------------------------------------------------------------------------------------------------------
# C4: fastq reads -> six frames -> fragments -> calling + family voting -> best frame (FqProcessRequest::on_parsed_seq)
# ----------------------------------------------------------------------------------------------------------------------
def run_fq(api, torch, dist, rank, world, local, barrier, peak, args):
    n_sigs, pool_n, reads_total = args.fq_sigs, args.fq_pool, args.fq_reads
    t0 = time.time()
    protos = synth.make_prototypes(4242, -(-n_sigs // 293) + 8, 300, 60.0)
    sig = synth.make_signatures(protos, n_sigs, dedupe=True)  # family tables are keyed by k-mer: keys must be distinct
    from close_kmers_b200 import api as _api
    img = _api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    fam = synth.make_families(7, sig)
    chunk = 250_000
    parts = [synth.make_reads(100 + 7919 * rank + k, protos, min(chunk, pool_n - k * chunk)) for k in range(-(-pool_n // chunk))]
    reads = synth.Batch(np.concatenate([p.residues for p in parts]), np.arange(pool_n + 1, dtype=np.uint64) * np.uint64(150))
    log(f"[fq r{rank}] world ({len(sig.keys)} k-mers, {fam.n_fams} families, {pool_n} reads) in {time.time() - t0:.1f}s")
    g = api.KmerGuts(image=img, device=local, function_names=synth.function_names(sig.n_functions))
    g.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    pin = Pinned(api, reads, packed=False)
    o = api.FqOutC()
    L = api.lib()

    def one_pass():
        api._check(L.ckm_fq_batch(g._h, pin.res, pin.off, pool_n, C.byref(o)))

    passes = max(1, reads_total // world // pool_n)
    for _ in range(2):
        one_pass()
    n_frag, n_probes = int(o.n_fragments), int(o.n_probes)
    barrier()
    t0 = time.perf_counter()
    for _ in range(passes):
        one_pass()
    barrier()
    wall = time.perf_counter() - t0
    wall_max, = allreduce([wall], "MAX", world, torch, dist)
    reads_all, probes_all = allreduce([passes * pool_n, passes * n_probes], "SUM", world, torch, dist)
    rec = None
    if rank == 0:
        res = g.fq_batch(reads.residues[: 150 * 20_000], reads.offsets[: 20_001])
        fan_out = float(np.mean(np.diff(fam.fam_off.astype(np.int64))))
        alg = reads_all * 150.0 + 32.0 * probes_all  # SURVEY 8d (the 4 B x fan-out per hit of the family lists is not counted)
        rec = {"workload": f"c4: {int(reads_all)} reads of 150 bp ({passes} passes over a {pool_n}-read pool per rank) vs a {len(sig.keys)}-k-mer "
                           f"image ({g.num_sigs} buckets, larger than L2) + family tables ({fam.n_fams} families, mean list {fan_out:.2f}); "
                           f"BASELINE names 50M reads: the path is linear in reads (every read is independent), x{50e6 / max(reads_all, 1):.0f}",
               "entry": "ckm_fq_batch, pinned host buffers (H2D of the bases, 6-frame translation, calling, family voting, best frame, D2H of the matches)",
               "value": reads_all / wall_max, "unit": "reads/s", "probes_per_s": probes_all / wall_max, "wall_s": wall_max,
               "fragments_per_read": n_frag / pool_n, "probes_per_read": n_probes / pool_n, "scaling": "weak",
               "reads_with_output_in_sample": int((res["best_frame"] != 0).sum()),
               "roofline": {"bound": "hbm", "achieved": alg / wall_max / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                            "frac": alg / wall_max / 1e9 / world / peak,
                            "note": "150 B + 32 B x probes per read over end-to-end wall time"}}
        if not args.no_cpu_baseline:
            import cpu_checkers as cc
            cc.ensure_built()
            m = 4000
            sub = synth.Batch(reads.residues[: 150 * m], reads.offsets[: m + 1])
            if os.path.exists(cc.REF_SO):
                d = image_dir(f"fq_{os.getpid()}", img.nbytes)
                img.tofile(os.path.join(d, "kmer.table.mem_map"))
                synth.write_index_files(d, sig.n_functions, 0)
                with stdout_to_stderr():
                    ref = cc.Ref().open(d, 1)
                    ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
                    ref.fq_batch(synth.Batch(sub.residues[: 150 * 200], sub.offsets[:201]))
                    t0 = time.perf_counter()
                    want = ref.fq_batch(sub)
                    dt = time.perf_counter() - t0
                    ref.close()
                shutil.rmtree(d, ignore_errors=True)
                got = g.fq_batch(sub.residues, sub.offsets)
                same = all(int(got["best_frame"][i]) == want[i][0] and float(got["best_score"][i]) == want[i][1] for i in range(m))
                kind = "reference"
                how = "the reference's DNASequence / TranslationTable / FamilyMapper / KmerGuts object code under the restated on_parsed_seq loop"
            else:
                orc = cc.Oracle().open_image(img)
                orc.family_load(fam)
                t0 = time.perf_counter()
                want = orc.fq_batch(sub)
                dt = time.perf_counter() - t0
                got = g.fq_batch(sub.residues, sub.offsets)
                same = bool(np.array_equal(got["best_frame"], want["best_frame"]) and np.array_equal(got["best_score"], want["best_score"]))
                kind, how = "port", "the plain-C oracle"
            rec["cpu_baseline"] = {"value": m / dt, "unit": "reads/s", "cores": 1, "kind": kind,
                                   "sample": f"first {m} reads of the pool, {dt:.2f}s on one thread ({how}; the handler serves one request per worker thread)"}
            rec["parity_on_sample"] = bool(same)
    barrier()
    g.close()
    pin.free()
    return rec


# ----------------------------------------------------------------------------------------------------------------------
# C5: /add + /matrix for 50k proteins, row blocks over the ranks, tiles gathered with NCCL
# ----------------------------------------------------------------------------------------------------------------------
def run_matrix(api, torch, dist, rank, world, local, barrier, peak, args):
    from close_kmers_b200 import parallel
    n = args.matrix_proteins
    t0 = time.time()
    protos = synth.make_prototypes(777, max(n // 10, 8), 300, 60.0)
    sig = synth.make_signatures(protos, min(1_000_000, int(protos.offsets[-1]) - 8 * protos.n), dedupe=True)
    img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    batch = synth.make_proteins_parallel(778, protos, n, mix=(0.9, 0.1, 0.0, 0.0))
    eids = np.arange(batch.n, dtype=np.uint32)
    log(f"[matrix r{rank}] world in {time.time() - t0:.1f}s")
    g = api.KmerGuts(image=img, device=local)
    job = parallel.MatrixJob(g, eids, batch, rank, world, device=torch.device("cuda", local) if world > 1 else None)
    times = []
    for rep in range(1 + max(2, min(args.steps, 5))):  # first pass grows every buffer
        barrier()
        t0 = time.perf_counter()
        merged, stats = job.run()
        torch.cuda.synchronize()
        barrier()
        times.append(time.perf_counter() - t0)
    wall, = allreduce([min(times[1:])], "MAX", world, torch, dist)
    rec = None
    if rank == 0:
        rec = {"workload": f"c5: /add of {n} proteins ({n // 10} prototypes x 10 mutated copies) + one /matrix request over all of them, "
                           f"{len(sig.keys)}-k-mer image",
               "entry": "ckm_postings_add_block + ckm_postings_import (NCCL all-gather of the postings) + ckm_matrix_rows per row block + "
                        "all-gather of the COO tiles + merge",
               "value": n / wall, "unit": "proteins/s", "wall_s": wall, "pairs": int(len(merged)), "postings": int(stats["postings"]),
               "postings_walked": int(stats["walked"]), "scaling": "strong",
               "row_blocks": stats["row_blocks"], "phase_ms": stats["phase_ms"],
               "roofline": {"bound": "hbm", "achieved": 4.0 * stats["walked"] / wall / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                            "frac": 4.0 * stats["walked"] / wall / 1e9 / world / peak,
                            "note": "4 B per posting walked (SURVEY 8d) over the whole request's wall time; the path is bound by "
                                    "shared-memory atomics and launch latency at this size, not by HBM"}}
        if not args.no_cpu_baseline:
            import cpu_checkers as cc
            cc.ensure_built()
            m = min(n, 4000)
            sub = sub_batch(batch, m)
            ids = [f"fig|{i}.peg" for i in range(m)]
            if os.path.exists(cc.REF_SO):
                d = image_dir(f"mx_{os.getpid()}", img.nbytes)
                img.tofile(os.path.join(d, "kmer.table.mem_map"))
                synth.write_index_files(d, sig.n_functions, 0)
                with stdout_to_stderr():
                    ref = cc.Ref().open(d, 1)
                    ref.mapping_new()
                    t0 = time.perf_counter()
                    ref.add_text(ids, sub, silent=1)
                    want = ref.matrix_text(ids, sub)
                    dt = time.perf_counter() - t0
                    ref.close()
                shutil.rmtree(d, ignore_errors=True)
                g2 = api.KmerGuts(image=img, device=local)
                mp = api.KmerPegMapping()
                g2.add_text(mp, ids, sub.residues, sub.offsets, silent=1)
                got = g2.matrix_text(mp, ids, sub.residues, sub.offsets)
                g2.close()
                rec["cpu_baseline"] = {"value": m / dt, "unit": "proteins/s", "cores": 1, "kind": "reference",
                                       "sample": f"/add + /matrix of the first {m} proteins, {dt:.2f}s on one thread (the reference's KmerGuts "
                                                 f"object code under the restated AddRequest / MatrixRequest loops; cost grows with n^2 / prototypes)"}
                rec["equals_reference_text_on_sample"] = bool(got == want)
            # the whole request on one GPU, in a context of its own (/add of every protein, then every row)
            g1 = api.KmerGuts(image=img, device=local)
            g1.postings_add(eids, batch.residues, batch.offsets)
            whole = api.merge_pairs(g1.matrix_rows(eids, batch.residues, batch.offsets))
            g1.close()
            rec["equals_single_gpu"] = bool(whole.tobytes() == merged.tobytes())
    barrier()
    job.close()
    g.close()
    return rec


# ----------------------------------------------------------------------------------------------------------------------
# Signature sets that are a sparse subset of the windows, some of another function (what build_signature_kmers leaves): the
# worst case for the neighbour copy, and the test of the automatic fall-back to plain hash probing
# ----------------------------------------------------------------------------------------------------------------------
def run_sparse(api, torch, rank, local, peak, args):
    if rank != 0:
        return None
    n_sigs, n_prot, keep, foreign = args.sparse_sigs, args.sparse_proteins, 0.5, 0.05
    t0 = time.time()
    protos = synth.make_prototypes(2468, int(-(-n_sigs // 293) / keep) + 8, 300, 60.0)
    batch = synth.make_proteins_parallel(2469, protos, n_prot)
    sig = synth.make_signatures_sparse(protos, n_sigs, keep=keep, foreign=foreign, dedupe=False)
    img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    log(f"[sparse r{rank}] world ({len(sig.keys)} k-mers = {keep:.0%} of the prototypes' windows, {foreign:.0%} of another function, "
        f"{n_prot} proteins) in {time.time() - t0:.1f}s")
    total = int(batch.offsets[-1])
    max_len = int(np.diff(batch.offsets.astype(np.int64)).max())
    d_res = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    d_res[:total] = torch.from_numpy(batch.residues).cuda()
    d_off = torch.from_numpy(batch.offsets.astype(np.int64)).cuda()
    g = api.KmerGuts(image=img, device=local)
    rec = {"workload": f"{n_prot} proteins vs a {len(sig.keys)}-k-mer image ({g.num_sigs} buckets, larger than L2) whose signatures are a random "
                       f"{keep:.0%} of every prototype's windows, {foreign:.0%} of them of another function, avg_from_end jittered "
                       f"(synth.make_signatures_sparse); device-resident batch, WANT_BEST",
           "neighbour_copy": {"entries": g.chain_info["entries"], "chains": g.chain_info["chains"]}}
    best = {}
    for name, tuning in (("copy_pinned", api.TUNE_NO_FALLBACK), ("plain_pinned", api.TUNE_NO_FALLBACK | api.TUNE_PLAIN_PROBE), ("automatic", 0)):
        g.set_tuning(tuning)
        for _ in range(3):
            g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
            g.read_totals()  # the counters reach the host: the automatic fall-back judges the batch
        g.profile_enable(True)
        g.profile_read()
        for _ in range(max(3, min(args.steps, 10))):
            g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
        p, s_ms, nb = g.profile_read()
        g.profile_enable(False)
        probes, hits, calls = g.read_totals()
        ms = (p + s_ms) / nb
        rec[name] = {"ms_per_step": ms, "proteins_per_s": batch.n / (ms * 1e-3), "probes_per_s": probes / (ms * 1e-3),
                     "roofline_frac": (32.0 * probes + total) / (ms * 1e-3) / 1e9 / peak}
        if name == "copy_pinned":
            rec["per_step"] = {"proteins": batch.n, "probes": probes, "hits": hits, "calls": calls}
            rec["share_of_probes_answered_from_copy"] = g.chain_info["hits_from_copy"] / max(probes, 1)
        if name == "automatic":
            rec["automatic"]["state"] = g.copy_state
        o = g.device_results()
        from close_kmers_b200 import parallel
        best[name] = parallel._alias(o.d_best, batch.n * 28, "|u1", torch, torch.device("cuda", local)).cpu().numpy().tobytes()
    rec["identical_best_calls_on_all_three_paths"] = len(set(best.values())) == 1
    rec["automatic_never_slower_than_plain"] = rec["automatic"]["ms_per_step"] <= 1.03 * rec["plain_pinned"]["ms_per_step"]
    g.close()
    del d_res, d_off
    torch.cuda.empty_cache()
    return rec


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--sigs", type=int, default=0, help="override signature k-mer count")
    ap.add_argument("--proteins", type=int, default=0, help="override proteins per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip", default="", help="comma-separated sub-records to leave out: c3,fq,matrix,sparse,plain")
    ap.add_argument("--c3-proteins", type=int, default=100_000_000)
    ap.add_argument("--c3-pool", type=int, default=2_000_000, help="proteins in the host pool each rank cycles through")
    ap.add_argument("--c3-sigs", type=int, default=0)
    ap.add_argument("--fq-reads", type=int, default=10_000_000)
    ap.add_argument("--fq-pool", type=int, default=2_000_000)
    ap.add_argument("--fq-sigs", type=int, default=20_000_000)
    ap.add_argument("--matrix-proteins", type=int, default=50_000)
    ap.add_argument("--sparse-sigs", type=int, default=20_000_000)
    ap.add_argument("--sparse-proteins", type=int, default=300_000)
    args = ap.parse_args()
    skip = set(x for x in args.skip.split(",") if x)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n_sigs, n_prot, sd, desc = WORKLOADS[args.workload]
    n_sigs = args.sigs or n_sigs
    n_prot = args.proteins or n_prot
    K, W = args.steps, max(args.warmup, 0)
    tag = os.environ.get("MASTER_PORT", str(os.getpid())) + f"_{args.workload}_{n_sigs}"
    config = {"workload": f"{args.workload}: {desc}", "signature_kmers": n_sigs, "proteins_per_gpu_per_step": n_prot,
              "flags": "WANT_BEST (process_aa_seq + find_best_call)", "params": "defaults (min_hits 5, max_gap 200)",
              "l2": "inputs larger than L2 (table, residues and the neighbour copy each exceed 126 MB)" if args.workload != "c1"
                    else "table fits L2 (96 MB); parity config, not an HBM test", "seed": 12345}

    # ------------------------------------------------------------------ reference arm (CPU only; libckm.so is never loaded)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import cpu_checkers as cc
        cc.ensure_built()
        protos, batch = build_world(n_sigs, n_prot, sd, 0)
        kdir = image_dir(tag + "_ref", 24 * synth.bucket_count(n_sigs))
        if kdir is None:
            emit({"impl": "reference", "unavailable": "no room for the image file under /dev/shm or /tmp"})
            return 0
        T = cpu_threads()
        try:
            write_image_dir(kdir, protos, n_sigs, 0, os.path.exists(cc.REF_SO))
            with stdout_to_stderr():
                eng = CpuEngine(kdir, T)
                vals = []
                for s in range(W + K):
                    r = eng.sample(batch, target_s=max(1.0, 40.0 / (W + K)))
                    if s >= W:
                        vals.append(r)
                eng.close()
        finally:
            shutil.rmtree(kdir, ignore_errors=True)
        v = float(np.mean([r["value"] for r in vals]))
        line = {"impl": "reference", "metric": "proteins/sec", "value": v, "unit": "proteins/s", "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": n_prot / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 keys / f32 scores", "data": "synthetic", "config": config,
                "ms_per_step_note": f"time of the stated {n_prot}-protein step at the measured rate; every timed step ran a bounded sample of "
                                    f"{vals[-1]['n']} proteins ({float(np.mean([r['seconds'] for r in vals])):.2f}s)",
                "cpu_baseline": {k: vals[-1][k] for k in ("value", "unit", "cores", "kind", "sample")} | {"value": v},
                "e2e": {"value": v, "unit": "proteins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from close_kmers_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    protos, batch = build_world(n_sigs, n_prot, sd, rank)
    # rank 0 builds the image once and publishes it as a reference-format kmer dir; every rank opens that
    # directory with ckm_open (the reference's own load path: mmap + validate, kmer_image.cc:41-108)
    kdir = image_dir(tag, 24 * synth.bucket_count(n_sigs))
    if kdir is None:
        raise SystemExit("bench.py: no room for the image file under /dev/shm or /tmp")
    if rank == 0:
        write_image_dir(kdir, protos, n_sigs, rank, False)
    barrier()
    t0 = time.time()
    guts = api.KmerGuts(kmer_dir=kdir, device=local)
    log(f"[bench r{rank}] table resident in HBM: {guts.table_buckets} buckets x {guts.slot_bytes} B (image: {guts.num_sigs} buckets) in {time.time() - t0:.1f}s")
    guts.set_default_parameters()
    flags = api.WANT_BEST
    n = batch.n
    total = int(batch.offsets[-1])
    max_len = int(np.diff(batch.offsets.astype(np.int64)).max())

    # device-resident copy of the batch (+ slack) for the kernel-resident measurement
    d_res = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    d_res[:total] = torch.from_numpy(batch.residues).cuda()
    d_off = torch.from_numpy(batch.offsets.astype(np.int64)).cuda()
    pin = Pinned(api, batch)  # pinned host copies for the end-to-end measurements

    stream = torch.cuda.ExternalStream(guts.stream, device=torch.device("cuda", local))

    def step_resident():
        guts.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), n, total, max_len, flags)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(W, 3)):
        step_resident()
    guts.synchronize()
    # parity spot check on the live workload (never in the timed region): the first 2000 proteins against the reference's own
    # object code when oracle/_ref is here, else the C port -- through the ASCII and the packed entry point
    if rank == 0:
        import cpu_checkers as cc
        cc.ensure_built()
        sb = sub_batch(batch, 2000)
        with stdout_to_stderr():
            if os.path.exists(cc.REF_SO):
                chk = cc.Ref().open(kdir, 1)
                chk.set_params()
                checker = "oracle/_ref (reference object code)"
            else:
                chk = cc.Oracle()
                chk.open(kdir)
                checker = "oracle port"
            ref_out = chk.call_batch(sb, cc.WANT_BEST)
            chk.close()
        want, names = ref_out["best"], ref_out.get("best_function")

        def same(got):  # the reference reports an ambiguous call as a string, not as a pair of indices (kguts.cc:1176-1196)
            a, b = got.copy(), want.copy()
            if names is not None:
                for f in ("ambig_a", "ambig_b"):
                    a[f] = 0
                    b[f] = 0
            return a.tobytes() == b.tobytes() and (names is None or [guts.best_function(r) for r in got] == names)

        got = guts.process_aa_seq_batch(sb.residues, sb.offsets, flags)["best"]
        assert same(got), "bench: CUDA best calls differ from " + checker
        pk, woff = api.pack_residues(sb.residues, sb.offsets)
        got = guts.process_packed_batch(pk, woff, flags)["best"]
        assert same(got), "bench: CUDA best calls (packed entry) differ from " + checker
        log(f"[bench] parity spot check ok vs {checker} ({sb.n} proteins, {int((got['function_index'] >= 0).sum())} confident calls)")
        config["parity_spot_check"] = f"2000 proteins, best-call records bit-identical to {checker}, ASCII and packed entry"

    # ---- kernel-resident: K steps, device-timed on the ctx stream, max over ranks
    guts.profile_enable(True)
    guts.profile_read()
    launches0 = guts.launch_count
    barrier()
    t_clk0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step_resident()
    e1.record(stream)
    guts.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = guts.launch_count - launches0
    probe_ms, scan_ms, nb_prof = guts.profile_read()
    guts.profile_enable(False)
    fused = guts.last_batch_was_fused
    n_probes, n_hits, n_calls = guts.read_totals()
    chain = guts.chain_info
    ab = {}
    if "plain" not in skip:
        # A/B, outside the timed region: the same steps (i) through K1 + scan_kernel with 16-byte hit records in HBM (the round-1
        # path) and (ii) with plain hash probing instead of the neighbour copy -- identical results
        for name, bits in (("unfused_K1_plus_scan_kernel", api.TUNE_UNFUSED), ("plain_hash_probing", api.TUNE_PLAIN_PROBE | api.TUNE_NO_FALLBACK)):
            if name == "plain_hash_probing" and not chain["entries"]:
                continue
            guts.set_tuning(bits)
            step_resident()
            guts.synchronize()
            guts.profile_enable(True)
            guts.profile_read()
            for _ in range(min(K, 5)):
                step_resident()
            guts.synchronize()
            pm, sm, nbp = guts.profile_read()
            guts.profile_enable(False)
            ab[name + "_ms"] = {"K1": pm / max(nbp, 1), "K2": sm / max(nbp, 1)}
        guts.set_tuning(0)

    # ---- end to end through the C ABI with HOST buffers (H2D + kernels + D2H in the timed region)
    t_ascii, out = timed_host_calls(lambda: guts.call_batch_raw(pin.res, pin.off, n, flags), K, barrier)
    assert out.n_probes == n_probes, (out.n_probes, n_probes)
    t_packed, out = timed_host_calls(lambda: guts.call_batch_packed_raw(pin.pk, pin.woff, n, flags), K, barrier)
    assert out.n_probes == n_probes, (out.n_probes, n_probes)
    clocks = sampler.stop(t_clk0, time.time())  # all timed regions (device-resident steps, then end-to-end steps)

    ms_total, ms_ascii, ms_packed = allreduce([ms_total, t_ascii * 1e3, t_packed * 1e3], "MAX", world, torch, dist)
    probes_all, prot_all = allreduce([n_probes, n], "SUM", world, torch, dist)

    line = None
    if rank == 0:
        alg_bytes = 32.0 * n_probes + float(total)  # one 32 B sector per probe + 1 B per residue
        probe_ms_avg = probe_ms / max(nb_prof, 1)
        achieved = alg_bytes / (probe_ms_avg * 1e-3) / 1e9
        step_ms = ms_total / K
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tj.get(args.workload + ("_fused" if fused else "_chain" if chain["entries"] else ""))
            if isinstance(ent, dict):
                traffic, traffic_src = ent.get("bytes"), {k: v for k, v in ent.items() if k != "bytes"}
            elif ent is not None:
                traffic, traffic_src = ent, {"note": "round-1 capture"}
        except (OSError, ValueError):
            pass
        # independent random 16 B reads over the resident table: best of a few occupancies
        cal_rate = max(guts.calibrate_gather(16, u, 64, b)[0] for u, b in ((1, 8), (4, 4), (4, 8)))
        value = prot_all * K / (ms_total * 1e-3)
        k1 = ("hint_kernel+" if chain["entries"] else "") + ("probe_pc_kernel+best_fixup_kernel" if fused else
                                                            "probe_hint_kernel" if chain["entries"] else "probe_kernel")
        h2d_packed = pin.words * 4 + (n + 1) * 8
        line = {
            "metric": "proteins/sec", "value": value, "unit": "proteins/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 keys / f32 scores", "data": "synthetic", "config": config,
            "probes_per_s": probes_all * K / (ms_total * 1e-3),
            "per_step": {"proteins": n, "residues": total, "probes": n_probes, "hits": n_hits, "calls": n_calls},
            "kernels_ms": {k1: probe_ms_avg, "scan_kernel": scan_ms / max(nb_prof, 1)} | ab,
            "roofline": {"bound": "hbm", "kernel": k1, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_step": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                         "frac_step_note": "the same algorithmic bytes over the whole device-timed step (what `value` implies)",
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "gather_calibration": {"accesses_per_s": cal_rate, "as_32B_sectors_GBps": cal_rate * 32 / 1e9,
                                                "probe_kernel_probes_per_s": n_probes / (probe_ms_avg * 1e-3),
                                                "note": "ceiling of independent random reads over this table on this GPU "
                                                        "(DESIGN.md section 6); probes/s / accesses_per_s = fraction of it"}},
            "e2e": {"value": prot_all * K / (ms_packed * 1e-3), "unit": "proteins/s", "ms_per_step": ms_packed / K,
                    "h2d_bytes_per_step": h2d_packed, "d2h_bytes_per_step": n * 28 + 64,
                    "entry": "ckm_call_batch_packed: residues packed seven to a 32-bit word (base 22) in pinned host memory (what a parser emits in the "
                             "pass it makes over every byte), unpacked on the device behind each chunk's copy",
                    "host_read_GBps_per_rank": h2d_packed / (ms_packed / K * 1e-3) / 1e9},
            "e2e_ascii": {"value": prot_all * K / (ms_ascii * 1e-3), "unit": "proteins/s", "ms_per_step": ms_ascii / K,
                          "h2d_bytes_per_step": total + (n + 1) * 8, "d2h_bytes_per_step": n * 28 + 64,
                          "entry": "ckm_call_batch: the reference's char* strings concatenated, pinned host memory",
                          "host_read_GBps_per_rank": (total + (n + 1) * 8) / (ms_ascii / K * 1e-3) / 1e9},
            "gpu_launches": int(launches),
            "host": host_info(local) | {"pinned_buffers": "cudaMallocHost (ckm_host_alloc); one pool per rank"},
            "clocks": clocks,
            "table": {"buckets": guts.num_sigs, "buckets_in_hbm": guts.table_buckets, "slot_bytes": guts.slot_bytes, "l2_fetch_granularity": guts.l2_fetch_granularity,
                      "occupancy_bitmap": guts.has_occupancy_bitmap, "scan_inside_K1": bool(fused),
                      "neighbour_copy": {"entries": chain["entries"], "chains": chain["chains"], "build_ms": chain["build_ms"],
                                         "hits_answered_from_copy": chain["hits_from_copy"],
                                         "share_of_probes_answered_from_copy": chain["hits_from_copy"] / max(n_probes, 1),
                                         "automatic_fallback": guts.copy_state}},
        }
        if not args.no_cpu_baseline and world == 1:
            with stdout_to_stderr():
                eng = CpuEngine(kdir, cpu_threads())
                cb = eng.sample(batch, target_s=8.0)
                eng.close()
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "probes_per_s")}
    barrier()
    guts.close()
    pin.free()
    del d_res, d_off
    torch.cuda.empty_cache()
    if rank == 0:
        shutil.rmtree(kdir, ignore_errors=True)

    # ---- the other BASELINE configs, as sub-records of the same line
    subs = {}
    for name, fn in (("c3_stream", lambda: run_c3_stream(api, torch, dist, rank, world, local, barrier, peak, args, tag)),
                     ("fq", lambda: run_fq(api, torch, dist, rank, world, local, barrier, peak, args)),
                     ("matrix", lambda: run_matrix(api, torch, dist, rank, world, local, barrier, peak, args)),
                     ("sparse_signatures", lambda: run_sparse(api, torch, rank, local, peak, args))):
        if name.split("_")[0] in skip or args.workload != "c2":
            continue
        t0 = time.time()
        try:
            subs[name] = fn()
        except Exception as e:  # a sub-record must not take the headline line down with it
            if world > 1:
                raise
            subs[name] = {"error": f"{type(e).__name__}: {e}"}
        log(f"[bench r{rank}] {name}: {time.time() - t0:.1f}s")
    if rank == 0:
        line.update({k: v for k, v in subs.items() if v is not None})
        emit(line)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
