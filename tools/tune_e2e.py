"""End-to-end (host buffers in, best calls out) sweep over pipeline chunk sizes on the C2 workload (GPU box), through the ASCII
entry point and the packed one.  python tools/tune_e2e.py [n_proteins] [n_sigs]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from close_kmers_b200 import api, synth

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
protos = synth.make_prototypes(12345, max(64, -(-n_sigs // 293) + 8), 300, 60.0)
batch = synth.make_proteins_parallel(12346, protos, n_prot)
sig = synth.make_signatures(protos, n_sigs, dedupe=n_sigs <= 2_000_000)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
L = api.lib()
total = int(batch.offsets[-1])
hp_res, hp_off = C.c_void_p(), C.c_void_p()
api._check(L.ckm_host_alloc(C.byref(hp_res), total + 64))
api._check(L.ckm_host_alloc(C.byref(hp_off), (batch.n + 1) * 8))
C.memmove(hp_res.value, batch.residues.ctypes.data, total)
C.memmove(hp_off.value, batch.offsets.ctypes.data, (batch.n + 1) * 8)
pk, woff = api.pack_residues(batch.residues, batch.offsets)
hp_pk, hp_woff = C.c_void_p(), C.c_void_p()
api._check(L.ckm_host_alloc(C.byref(hp_pk), pk.nbytes + 64))
api._check(L.ckm_host_alloc(C.byref(hp_woff), (batch.n + 1) * 8))
C.memmove(hp_pk.value, pk.ctypes.data, pk.nbytes)
C.memmove(hp_woff.value, woff.ctypes.data, (batch.n + 1) * 8)
for chunk_kb, ramp, tail in ((49152, 12, 8), (65536, 16, 8), (65536, 8, 4), (98304, 12, 6), (98304, 24, 8), (131072, 16, 8), (131072, 32, 16), (32768, 8, 8), (49152, 6, 4), (196608, 24, 12)):
    os.environ["CKM_PIPELINE_CHUNK_KB"] = str(chunk_kb)
    os.environ["CKM_PIPELINE_RAMP_DIV"] = str(ramp)
    os.environ["CKM_PIPELINE_TAIL_DIV"] = str(tail)
    g = api.KmerGuts(image=img)
    for _ in range(2):
        g.call_batch_raw(hp_res.value, hp_off.value, batch.n, api.WANT_BEST)
    t0 = time.perf_counter()
    K = 8
    for _ in range(K):
        g.call_batch_raw(hp_res.value, hp_off.value, batch.n, api.WANT_BEST)
    dt = (time.perf_counter() - t0) / K
    for _ in range(2):
        g.call_batch_packed_raw(hp_pk.value, hp_woff.value, batch.n, api.WANT_BEST)
    t0 = time.perf_counter()
    for _ in range(K):
        g.call_batch_packed_raw(hp_pk.value, hp_woff.value, batch.n, api.WANT_BEST)
    dtp = (time.perf_counter() - t0) / K
    print(json.dumps(dict(chunk_kb=chunk_kb, ramp_div=ramp, tail_div=tail, e2e_ms=dt * 1e3, proteins_per_s=batch.n / dt,
                          packed_e2e_ms=dtp * 1e3, packed_proteins_per_s=batch.n / dtp)), flush=True)
    g.close()
