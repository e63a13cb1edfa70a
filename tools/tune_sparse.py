"""K1 through the neighbour copy against plain hash probing on signature sets of decreasing density (GPU box): the dense world of
make_signatures (every window of a prototype a signature) and make_signatures_sparse at several `keep` fractions.  Reports, per
world, K1 with the copy pinned on (CKM_TUNE_NO_FALLBACK), with plain probing, and what the automatic fall-back settles on.
python tools/tune_sparse.py [n_proteins] [n_sigs] [keep ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from close_kmers_b200 import api, parallel, synth

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
keeps = [float(x) for x in sys.argv[3:]] or [1.0, 0.8, 0.6, 0.4, 0.25]
foreign = float(os.environ.get("FOREIGN", "0.1"))
jitter = int(os.environ.get("JITTER", "12"))
os.environ.setdefault("CKM_CHAIN", "1")
for keep in keeps:
    n_protos = max(64, int(-(-n_sigs // 293) / keep) + 8)
    protos = synth.make_prototypes(12345, n_protos, 300, 60.0)
    batch = synth.make_proteins_parallel(12346, protos, n_prot)
    sig = synth.make_signatures(protos, n_sigs, dedupe=False) if keep >= 1.0 and foreign == 0.1 else synth.make_signatures_sparse(protos, n_sigs, keep=keep, jitter=jitter, foreign=foreign, dedupe=False)
    img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    total = int(batch.offsets[-1])
    max_len = int(np.diff(batch.offsets.astype(np.int64)).max())
    d_res = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    d_res[:total] = torch.from_numpy(batch.residues).cuda()
    d_off = torch.from_numpy(batch.offsets.astype(np.int64)).cuda()
    g = api.KmerGuts(image=img)
    rec = dict(keep=keep, foreign=foreign, jitter=jitter, n_sigs=len(sig.keys), buckets=g.num_sigs, chains=g.chain_info["chains"], entries=g.chain_info["entries"])
    best = {}
    for name, tuning in (("copy", api.TUNE_NO_FALLBACK), ("plain", api.TUNE_NO_FALLBACK | api.TUNE_PLAIN_PROBE),
                         ("copy_unfused", api.TUNE_NO_FALLBACK | api.TUNE_UNFUSED),
                         ("plain_unfused", api.TUNE_NO_FALLBACK | api.TUNE_PLAIN_PROBE | api.TUNE_UNFUSED), ("auto", 0)):
        g.set_tuning(tuning)
        for _ in range(3):
            g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
            g.read_totals()  # the counters reach the host: the automatic fall-back judges the batch
        g.profile_enable(True)
        g.profile_read()
        for _ in range(5):
            g.call_batch_device(d_res.data_ptr(), d_off.data_ptr(), batch.n, total, max_len, api.WANT_BEST)
        p, s, nb = g.profile_read()
        g.profile_enable(False)
        probes, hits, calls = g.read_totals()
        rec[name + "_K1_ms"] = p / nb
        rec[name + "_K1K2_ms"] = (p + s) / nb
        if name == "copy":
            rec.update(probes=probes, hits=hits, hits_from_copy=g.chain_info["hits_from_copy"],
                       share_of_probes_from_copy=g.chain_info["hits_from_copy"] / max(probes, 1))
        if name == "auto":
            rec["auto_state"] = g.copy_state
        o = g.device_results()
        best[name] = parallel._alias(o.d_best, batch.n * 28, "|u1", torch, torch.device("cuda", 0)).cpu().numpy().tobytes()
    rec["identical_best_calls"] = len(set(best.values())) == 1
    print(json.dumps(rec), flush=True)
    g.close()
    del d_res, d_off
