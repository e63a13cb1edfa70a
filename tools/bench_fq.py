"""C4-style measurement: fastq reads -> 6 frames -> fragments > 10 aa -> calling + family voting -> best frame, through
ckm_fq_batch with host buffers (end to end), beside the C oracle on a bounded sample.  python tools/bench_fq.py [n_reads]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from close_kmers_b200 import api, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
protos = synth.make_prototypes(4242, -(-n_sigs // 293) + 8, 300, 60.0)
sig = synth.make_signatures(protos, n_sigs, dedupe=True)  # family tables are keyed by k-mer: keys must be distinct
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
fam = synth.make_families(7, sig)
chunk = 250_000
parts = [synth.make_reads(100 + k, protos, min(chunk, n_reads - k * chunk)) for k in range(-(-n_reads // chunk))]
reads = synth.Batch(np.concatenate([p.residues for p in parts]), np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(150))
g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
g.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
g.fq_batch(reads.residues[: 150 * 100_000], reads.offsets[: 100_001])
t0 = time.perf_counter()
res = g.fq_batch(reads.residues, reads.offsets)   # first full-size call: grows every work buffer
dt_first = time.perf_counter() - t0
t0 = time.perf_counter()
res = g.fq_batch(reads.residues, reads.offsets)   # steady state, incl. numpy copies of the results
dt_py = time.perf_counter() - t0
# the C-ABI call alone: host (pageable) buffers in, results in the library's pinned buffers
import ctypes as C
o = api.FqOutC()
bases = np.ascontiguousarray(reads.residues, np.uint8)
offs = np.ascontiguousarray(reads.offsets, np.uint64)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    rc = api.lib().ckm_fq_batch(g._h, bases.ctypes.data, offs.ctypes.data, n_reads, C.byref(o))
    ts.append(time.perf_counter() - t0)
    assert rc == 0
dt = min(ts)
# ... and with the reads in page-locked memory (what a front end that owns its buffers would use)
pin = api.pinned_empty(bases.nbytes) if hasattr(api, "pinned_empty") else None
dt_pinned = None
if pin is not None:
    pin[:] = bases
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        rc = api.lib().ckm_fq_batch(g._h, pin.ctypes.data, offs.ctypes.data, n_reads, C.byref(o))
        ts.append(time.perf_counter() - t0)
    dt_pinned = min(ts)
out = dict(reads=n_reads, signature_kmers=len(sig.keys), fragments=int(res["n_fragments"]), probes=int(res["n_probes"]),
           reads_with_output=int((res["best_frame"] != 0).sum()), gpu_first_call_s=dt_first, gpu_e2e_s=dt, gpu_reads_per_s=n_reads / dt,
           python_wrapper_e2e_s=dt_py, gpu_e2e_pinned_input_s=dt_pinned,
           gpu_probes_per_s=int(res["n_probes"]) / dt)
import cpu_checkers as cc
cc.ensure_built()
orc = cc.Oracle().open_image(img)
orc.family_load(fam)
m = 20_000
sub = synth.Batch(reads.residues[: 150 * m], reads.offsets[: m + 1])
t0 = time.perf_counter()
want = orc.fq_batch(sub)
dtc = time.perf_counter() - t0
got = g.fq_batch(sub.residues, sub.offsets)
out.update(cpu_port_reads_per_s_1thread=m / dtc, parity_on_sample=bool(np.array_equal(got["best_frame"], want["best_frame"]) and
           np.array_equal(got["best_score"], want["best_score"]) and np.array_equal(got["matches"]["lfam"], want["matches"]["lfam"])))
print(json.dumps(out), flush=True)
