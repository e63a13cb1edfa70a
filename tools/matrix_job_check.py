"""MatrixJob (close_kmers_b200/parallel.py) against one GPU doing the whole /add + /matrix, on the world bench.py uses for C5.
python tools/matrix_job_check.py [n_proteins]                      (one GPU: the job without collectives)
torchrun --nproc-per-node N tools/matrix_job_check.py [n_proteins]  (N ranks over NCCL)"""
import json
import os
import sys

sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
import torch
from close_kmers_b200 import api, parallel, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
protos = synth.make_prototypes(777, max(n // 10, 8), 300, 60.0)
sig = synth.make_signatures(protos, min(1_000_000, int(protos.offsets[-1]) - 8 * protos.n), dedupe=True)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
batch = synth.make_proteins_parallel(778, protos, n, mix=(0.9, 0.1, 0.0, 0.0))
eids = np.arange(batch.n, dtype=np.uint32)
g = api.KmerGuts(image=img, device=local)
job = parallel.MatrixJob(g, eids, batch, rank, world, device=torch.device("cuda", local))
for _ in range(3):
    merged, stats = job.run()
if rank == 0:
    g2 = api.KmerGuts(image=img, device=local)
    g2.postings_add(eids, batch.residues, batch.offsets)
    whole = api.merge_pairs(g2.matrix_rows(eids, batch.residues, batch.offsets))
    same_ctx = api.merge_pairs(g.matrix_rows(eids, batch.residues, batch.offsets))
    rec = dict(world=world, n=n, pairs=len(merged), pairs_one_gpu=len(whole), equals_one_gpu=bool(whole.tobytes() == merged.tobytes()),
               same_ctx_equals_one_gpu=bool(same_ctx.tobytes() == whole.tobytes()), postings=stats["postings"], postings_one_gpu=g2.postings_count,
               phase_ms=stats["phase_ms"])
    if not rec["equals_one_gpu"] and len(merged) == len(whole):
        bad = np.nonzero(merged != whole)[0]
        rec["first_differences"] = [[str(merged[k]), str(whole[k])] for k in bad[:4]]
    print(json.dumps(rec), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
