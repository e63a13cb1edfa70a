"""Probe-kernel tuning sweep on the C2 workload (GPU box): builds the world once, then times K1/K2 per setting.
python tools/tune_probe.py [n_proteins] [n_sigs]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from close_kmers_b200 import api, synth

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
protos = synth.make_prototypes(12345, max(64, -(-n_sigs // 293) + 8), 300, 60.0)
batch = synth.make_proteins_parallel(12346, protos, n_prot)
sig = synth.make_signatures(protos, n_sigs, dedupe=n_sigs <= 2_000_000)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
total = int(batch.offsets[-1])
max_len = int(np.diff(batch.offsets.astype(np.int64)).max())
d_res = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
d_res[:total] = torch.from_numpy(batch.residues).cuda()
d_off = torch.from_numpy(batch.offsets.astype(np.int64)).cuda()
for bitmap in ("1", "0"):
    os.environ["CKM_OCCUPANCY_BITMAP"] = bitmap
    for persist in ("0",):
        g = api.KmerGuts(image=img)
        for tuning in (0, 16):
            if bitmap == "0" and (tuning & 2):
                continue
            g.set_tuning(tuning)
            g.profile_enable(True)
            for _ in range(3):
                g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
            g.profile_read()
            for _ in range(5):
                g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
            p, s, nb = g.profile_read()
            probes = g.read_totals()[0]
            print(json.dumps(dict(bitmap=bitmap, no_persist=persist, tuning=tuning, probe_ms=p / nb, scan_ms=s / nb,
                                  gprobes_per_s=probes / (p / nb) / 1e6)), flush=True)
        g.close()
