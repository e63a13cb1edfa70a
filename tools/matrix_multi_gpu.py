"""Row-block sharded /matrix on N GPUs with an NCCL gather of the COO tiles, checked against a single-GPU
run and the CPU oracle.  Launch: torchrun --nproc-per-node N tools/matrix_multi_gpu.py [n_proteins]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist
from close_kmers_b200 import api, parallel, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
# C5: n proteins drawn from n/10 prototypes (10 mutated copies each), /add-ed once, then one matrix request
protos = synth.make_prototypes(777, max(n // 10, 8), 300, 60.0)
sig = synth.make_signatures(protos, min(1_000_000, int(protos.offsets[-1]) - 8 * protos.n), dedupe=True)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
batch = synth.make_proteins_parallel(778, protos, n, mix=(0.9, 0.1, 0.0, 0.0))
eids = np.arange(batch.n, dtype=np.uint32)
g = api.KmerGuts(image=img, device=local)
t0 = time.perf_counter()
g.postings_add(eids, batch.residues, batch.offsets)
g.synchronize()
t_add = time.perf_counter() - t0
lengths = np.diff(batch.offsets.astype(np.int64))
for rep in range(2):  # first pass builds the postings index
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    merged = parallel.matrix_sharded(lambda a, b: g.matrix_rows(eids, batch.residues, batch.offsets, a, b), lengths, rank, world)
    torch.cuda.synchronize()
    t_mat = time.perf_counter() - t0
if rank == 0:
    whole = api.merge_pairs(g.matrix_rows(eids, batch.residues, batch.offsets))
    ok_single = merged.tobytes() == whole.tobytes()
    ok_oracle = None
    if n <= 5000:
        import cpu_checkers as cc
        cc.ensure_built()
        orc = cc.Oracle().open_image(img)
        orc.postings_new()
        orc.postings_add(eids, batch)
        ok_oracle = api.merge_pairs(orc.matrix_rows(eids, batch)).tobytes() == merged.tobytes()
    print(json.dumps(dict(n_gpus=world, proteins=n, postings=g.postings_count, pairs=len(merged), add_s=t_add, matrix_s=t_mat,
                          pairs_per_s=len(merged) / t_mat, equals_single_gpu=bool(ok_single), equals_oracle=ok_oracle,
                          row_blocks=parallel.shard_rows(lengths, world))), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
g.close()
