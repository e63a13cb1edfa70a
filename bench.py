#!/usr/bin/env python
"""Benchmark of the signature-k-mer calling path (BASELINE.json metric: proteins/s and k-mer probes/s per
B200, % of the HBM gather roofline), with the reference's CPU path timed beside it.

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3] [--impl reference]

A step = one pass of the hot path (encode -> probe -> ordered scoring scan -> find_best_call) over one batch
of synthetic proteins against a synthetic signature image in the reference's file format.  Default workload
is BASELINE.json configs[1] ("c2": 1M proteins, mean 300 aa, vs a 100M-k-mer image = 508,000,037 buckets).

Keys of the JSON line: see DESIGN.md "Measurement".  `value` is kernel-resident (inputs already in HBM),
`e2e` goes through ckm_call_batch with pinned HOST buffers (H2D + kernels + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
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


def build_world(name, n_sigs, n_proteins, sd, rank, seed=12345):
    """Seeded synthetic world: prototypes -> signature k-mers -> (image bytes), plus this rank's proteins."""
    t0 = time.time()
    n_protos = max(64, -(-n_sigs // 293) + 8)
    protos = synth.make_prototypes(seed, n_protos, 300, sd)
    batch = synth.make_proteins_parallel(seed + 1 + rank, protos, n_proteins)
    log(f"[bench r{rank}] prototypes+proteins: {time.time() - t0:.1f}s  ({batch.n} proteins, {batch.residues.nbytes / 1e6:.0f} MB)")
    return protos, batch


def build_image(protos, n_sigs, rank):
    from close_kmers_b200 import api
    t0 = time.time()
    sig = synth.make_signatures(protos, n_sigs, dedupe=n_sigs <= 2_000_000)
    nb = synth.bucket_count(len(sig.keys))
    img = api.build_image(nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    log(f"[bench r{rank}] image: {len(sig.keys)} k-mers, {nb} buckets, {img.nbytes / 1e9:.2f} GB in {time.time() - t0:.1f}s")
    return sig, nb, img


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


class CpuEngine:
    """The reference's CPU path for this workload: oracle/_ref (the reference's own object code, one KmerGuts
    per thread sharing one mmapped image, like threadpool.cc:18-45) when present, else the plain-C port."""

    def __init__(self, kdir, img, threads):
        import cpu_checkers as cc
        cc.ensure_built()
        self.threads = threads
        self.kind = "reference" if (os.path.exists(cc.REF_SO) and kdir is not None) else "port"
        if self.kind == "reference":
            self.eng = cc.Ref().open(kdir, threads)
            self.eng.set_params()
            self.run = lambda b: self.eng.bench_calls(b, True)
        else:
            self.eng = cc.Oracle()
            self.eng.open_image(img) if img is not None else self.eng.open(kdir)
            self.run = lambda b: self.eng.bench_calls(b, threads, True)
        self.rate = None

    def sample(self, batch, target_s):
        """One bounded sample sized for ~target_s of wall time on all threads; returns the cpu_baseline dict."""
        def sub(n):
            n = min(n, batch.n)
            return synth.Batch(batch.residues[: int(batch.offsets[n])], batch.offsets[: n + 1])
        if self.rate is None:  # calibrate once on a small slice
            self.run(sub(min(batch.n, 50 * self.threads)))
            cal_n = min(batch.n, 500 * self.threads)
            self.rate = cal_n / max(self.run(sub(cal_n)), 1e-6)
        n = int(max(min(batch.n, 500 * self.threads), min(batch.n, self.rate * target_s)))
        b = sub(n)
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
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n_sigs, n_prot, sd, desc = WORKLOADS[args.workload]
    n_sigs = args.sigs or n_sigs
    n_prot = args.proteins or n_prot
    K, W = args.steps, max(args.warmup, 0)
    tag = os.environ.get("MASTER_PORT", str(os.getpid())) + f"_{args.workload}_{n_sigs}"
    config = {"workload": f"{args.workload}: {desc}", "signature_kmers": n_sigs, "proteins_per_gpu_per_step": n_prot,
              "flags": "WANT_BEST (process_aa_seq + find_best_call)", "params": "defaults (min_hits 5, max_gap 200)",
              "l2": "inputs larger than L2 (table, residues and hit regions each exceed 126 MB)" if args.workload != "c1"
                    else "table fits L2 (96 MB); parity config, not an HBM test", "seed": 12345}

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        protos, batch = build_world(args.workload, n_sigs, n_prot, sd, 0)
        sig, nb, img = build_image(protos, n_sigs, 0)
        kdir = image_dir(tag + "_ref", img.nbytes)
        if kdir:
            img.tofile(os.path.join(kdir, "kmer.table.mem_map"))
            synth.write_index_files(kdir, sig.n_functions, 0)
        T = cpu_threads()
        try:
            with stdout_to_stderr():
                eng = CpuEngine(kdir, img, T)
                vals = []
                for s in range(W + K):
                    r = eng.sample(batch, target_s=max(1.0, 40.0 / (W + K)))
                    if s >= W:
                        vals.append(r)
                eng.close()
        finally:
            if kdir:
                shutil.rmtree(kdir, ignore_errors=True)
        v = float(np.mean([r["value"] for r in vals]))
        ms = float(np.mean([r["seconds"] for r in vals])) * 1e3
        line = {"impl": "reference", "metric": "proteins/sec", "value": v, "unit": "proteins/s", "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 keys / f32 scores", "data": "synthetic", "config": config,
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

    protos, batch = build_world(args.workload, n_sigs, n_prot, sd, rank)
    # rank 0 builds the image once and publishes it as a reference-format kmer dir; every rank opens that
    # directory with ckm_open (the reference's own load path: mmap + validate, kmer_image.cc:41-108)
    kdir = image_dir(tag, 24 * synth.bucket_count(n_sigs))
    img = None
    if rank == 0 or kdir is None:
        sig, nb, img = build_image(protos, n_sigs, rank)
        if kdir:
            t0 = time.time()
            img.tofile(os.path.join(kdir, "kmer.table.mem_map"))
            synth.write_index_files(kdir, sig.n_functions, 0)
            log(f"[bench r{rank}] wrote {kdir} in {time.time() - t0:.1f}s")
    barrier()
    t0 = time.time()
    guts = api.KmerGuts(kmer_dir=kdir, device=local) if kdir else api.KmerGuts(image=img, device=local)
    log(f"[bench r{rank}] table resident in HBM: {guts.num_sigs} buckets x {guts.slot_bytes} B in {time.time() - t0:.1f}s")
    guts.set_default_parameters()
    L = api.lib()
    flags = api.WANT_BEST
    n = batch.n
    total = int(batch.offsets[-1])
    max_len = int(np.diff(batch.offsets.astype(np.int64)).max())

    # device-resident copy of the batch (+16 B slack) for the kernel-resident measurement
    d_res = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    d_res[:total] = torch.from_numpy(batch.residues).cuda()
    d_off = torch.from_numpy(batch.offsets.astype(np.int64)).cuda()
    # pinned host copy for the end-to-end measurement
    import ctypes as C
    hp_res, hp_off = C.c_void_p(), C.c_void_p()
    api._check(L.ckm_host_alloc(C.byref(hp_res), total + 64))
    api._check(L.ckm_host_alloc(C.byref(hp_off), (n + 1) * 8))
    C.memmove(hp_res.value, batch.residues.ctypes.data, total)
    C.memmove(hp_off.value, batch.offsets.ctypes.data, (n + 1) * 8)

    stream = torch.cuda.ExternalStream(guts.stream, device=torch.device("cuda", local))

    def step_resident():
        guts.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), n, total, max_len, flags)

    def step_e2e():
        return guts.call_batch_raw(hp_res.value, hp_off.value, n, flags)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(W, 3)):
        step_resident()
    guts.synchronize()
    # parity spot check on the live workload: first 2000 proteins against the C oracle (never in the timed region)
    if rank == 0:
        import cpu_checkers as cc
        cc.ensure_built()
        m = min(n, 2000)
        sb = synth.Batch(batch.residues[: int(batch.offsets[m])], batch.offsets[: m + 1])
        orc = cc.Oracle()
        orc.open(kdir) if kdir else orc.open_image(img)
        want = orc.call_batch(sb, cc.WANT_BEST)["best"]
        got = guts.process_aa_seq_batch(sb.residues, sb.offsets, flags)["best"]
        assert got.tobytes() == want.tobytes(), "bench: CUDA best calls differ from the oracle"
        orc.close()
        log(f"[bench] parity spot check ok ({m} proteins, {int((got['function_index'] >= 0).sum())} confident calls)")

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
    n_probes, n_hits, n_calls = guts.read_totals()
    chain = guts.chain_info
    plain_probe_ms = None
    if chain["entries"]:
        # A/B, outside the timed region: the same steps with plain hash probing (probe_kernel) instead of
        # hint_kernel + probe_hint_kernel -- identical results, every hit its own DRAM transaction
        guts.set_tuning(32)
        step_resident()
        guts.synchronize()
        guts.profile_enable(True)
        guts.profile_read()
        for _ in range(min(K, 5)):
            step_resident()
        guts.synchronize()
        pm, _, nbp = guts.profile_read()
        guts.profile_enable(False)
        plain_probe_ms = pm / max(nbp, 1)
        guts.set_tuning(0)

    # ---- end to end through the C ABI with HOST buffers (H2D + kernels + D2H in the timed region)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out = step_e2e()
    t_e2e = time.perf_counter() - t0
    assert out.n_probes == n_probes, (out.n_probes, n_probes)
    clocks = sampler.stop(t_clk0, time.time())  # both timed regions (device-resident steps, then end-to-end steps)

    ms_t = torch.tensor([ms_total, t_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([n_probes, n], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        probes_all, prot_all = float(cnt[0]), float(cnt[1])
    else:
        probes_all, prot_all = float(n_probes), float(n)
    ms_total, ms_e2e = float(ms_t[0]), float(ms_t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = 32.0 * n_probes + float(total)  # one 32 B sector per probe + 1 B per residue
        probe_ms_avg = probe_ms / max(nb_prof, 1)
        achieved = alg_bytes / (probe_ms_avg * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(
                args.workload + ("_chain" if chain["entries"] else ""))
        except (OSError, ValueError):
            pass
        # independent random 16 B reads over the resident table: best of a few occupancies
        cal_rate = max(guts.calibrate_gather(16, u, 64, b)[0] for u, b in ((1, 8), (4, 4), (4, 8)))
        value = prot_all * K / (ms_total * 1e-3)
        probe_name = "hint_kernel+probe_hint_kernel" if chain["entries"] else "probe_kernel"
        line = {
            "metric": "proteins/sec", "value": value, "unit": "proteins/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 keys / f32 scores", "data": "synthetic", "config": config,
            "probes_per_s": probes_all * K / (ms_total * 1e-3),
            "per_step": {"proteins": n, "residues": total, "probes": n_probes, "hits": n_hits, "calls": n_calls},
            "kernels_ms": {probe_name: probe_ms_avg, "scan_kernel": scan_ms / max(nb_prof, 1)},
            "roofline": {"bound": "hbm", "kernel": probe_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "gather_calibration": {"accesses_per_s": cal_rate, "as_32B_sectors_GBps": cal_rate * 32 / 1e9,
                                                "probe_kernel_probes_per_s": n_probes / (probe_ms_avg * 1e-3),
                                                "note": "ceiling of independent random reads over this table on this GPU "
                                                        "(DESIGN.md section 6); probes/s / accesses_per_s = fraction of it"}},
            "e2e": {"value": prot_all * K / (ms_e2e * 1e-3), "unit": "proteins/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": total + (n + 1) * 8, "d2h_bytes_per_step": n * 28 + 24},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "table": {"buckets": guts.num_sigs, "slot_bytes": guts.slot_bytes, "l2_fetch_granularity": guts.l2_fetch_granularity,
                      "occupancy_bitmap": guts.has_occupancy_bitmap,
                      "neighbour_copy": {"entries": chain["entries"], "chains": chain["chains"], "build_ms": chain["build_ms"],
                                         "hits_answered_from_copy": chain["hits_from_copy"],
                                         "plain_hash_probe_kernel_ms": plain_probe_ms}},
        }
        if not args.no_cpu_baseline and world == 1:
            with stdout_to_stderr():
                eng = CpuEngine(kdir, img, cpu_threads())
                cb = eng.sample(batch, target_s=8.0)
                eng.close()
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "probes_per_s")}
        emit(line)
    barrier()
    guts.close()
    L.ckm_host_free(hp_res)
    L.ckm_host_free(hp_off)
    if kdir and rank == 0:
        shutil.rmtree(kdir, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
