"""N4 measurement: the signature table built on the GPU (ckm_image_build_device / ckm_open_built) beside the sequential host
builder (ckm_image_build = the reference's insert_kmer loop).  python tools/bench_build.py [n_sigs]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from close_kmers_b200 import api, synth

n_sigs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
protos = synth.make_prototypes(4242, -(-n_sigs // 293) + 8, 300, 60.0)
sig = synth.make_signatures(protos, n_sigs)
nb = synth.bucket_count(len(sig.keys))
args = (nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
api.build_image_device(3769, sig.keys[:100], sig.fI[:100], sig.oI[:100], sig.avg[:100], sig.wt[:100])  # CUDA context up
t0 = time.perf_counter()
dev = api.build_image_device(*args)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
g = api.KmerGuts(built=args, function_names=synth.function_names(sig.n_functions))
t_open = time.perf_counter() - t0
g.close()
t0 = time.perf_counter()
host = api.build_image(*args)
t_host = time.perf_counter() - t0
print(json.dumps(dict(signature_kmers=len(sig.keys), buckets=nb, image_mb=dev.nbytes / 1e6, gpu_build_to_host_image_s=t_dev,
                      gpu_build_and_open_s=t_open, host_sequential_build_s=t_host, identical=bool(dev.tobytes() == host.tobytes()))))
