"""probe_chain_kernel vs probe_kernel on a C2-shaped workload (GPU box): builds the world once, then times K1 per setting.
python tools/tune_chain.py [n_proteins] [n_sigs] [tuning values ...]
tuning (ckm_set_tuning, include/ckm.h): 0 = hint_kernel + probe_hint_kernel<128 threads, 7 blocks/SM, payload staged in shared
memory>; n << 16 selects another build of it (1 = 256x4 staged, 2 = 128x8 staged, 3 = 128x6 staged, 4 = 256x3 registers,
5 = 256x3 staged); 1 = without evict_first on chain/slot loads; 0x80000 = without L2 prefetches; 128 / 64 = the walking
probe_chain_kernel at 3 / 2 blocks per SM; 32 = plain hash probing (probe_kernel)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from close_kmers_b200 import api, synth

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
tunings = [int(x) for x in sys.argv[3:]] or [0, 128, 32]
protos = synth.make_prototypes(12345, max(64, -(-n_sigs // 293) + 8), 300, 60.0)
batch = synth.make_proteins_parallel(12346, protos, n_prot)
sig = synth.make_signatures(protos, n_sigs, dedupe=n_sigs <= 2_000_000)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
total = int(batch.offsets[-1])
max_len = int(np.diff(batch.offsets.astype(np.int64)).max())
d_res = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
d_res[:total] = torch.from_numpy(batch.residues).cuda()
d_off = torch.from_numpy(batch.offsets.astype(np.int64)).cuda()
os.environ.setdefault("CKM_CHAIN", "1")
g = api.KmerGuts(image=img)
print(json.dumps(dict(n_proteins=n_prot, n_sigs=n_sigs, buckets=g.num_sigs, chain=g.chain_info)), flush=True)
ref_best = None
for tuning in tunings:
    g.set_tuning(tuning)
    g.profile_enable(True)
    for _ in range(3):
        g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
    g.profile_read()
    for _ in range(5):
        g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
    p, s, nb = g.profile_read()
    probes, hits, calls = g.read_totals()
    print(json.dumps(dict(tuning=tuning, probe_ms=p / nb, scan_ms=s / nb, gprobes_per_s=probes / (p / nb) / 1e6, hits=hits, calls=calls,
                          hits_from_copy=g.chain_info["hits_from_copy"],
                          frac_of_6549GBps=(32.0 * probes + total) / (p / nb * 1e-3) / 1e9 / 6549.4)), flush=True)
g.close()
