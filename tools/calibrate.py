"""Random-gather calibration sweep over the resident table (GPU box): python tools/calibrate.py [n_sigs]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from close_kmers_b200 import api, synth

n_sigs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nb = synth.bucket_count(n_sigs)
rng = np.random.default_rng(1)
# a table of the right SIZE is all the calibration needs: random distinct-ish keys, no prototypes
keys = rng.integers(0, 20**8, n_sigs, dtype=np.uint64)
img = api.build_image(nb, keys, np.zeros(n_sigs, np.int32), np.full(n_sigs, -1, np.int32), np.zeros(n_sigs, np.uint16),
                      np.ones(n_sigs, np.float32))
g = api.KmerGuts(image=img)
del img
out = []
print(json.dumps(dict(l2_fetch_granularity=g.l2_fetch_granularity)), flush=True)
for nbytes in (16, 32):
    for unroll in (1, 4, 8):
        for bps in (4, 8):
            rate, ms = g.calibrate_gather(nbytes, unroll, 64, bps)
            rec = dict(bytes=nbytes, unroll=unroll, blocks_per_sm=bps, accesses_per_s=rate, ms=ms,
                       sector_GBps=rate * 32 / 1e9, table_GB=g.num_sigs * g.slot_bytes / 1e9)
            out.append(rec)
            print(json.dumps(rec), flush=True)
