"""Two ranks over NCCL: the one path with an exchange step (/add + /matrix: postings and tiles all-gathered device to device,
close_kmers_b200/parallel.py MatrixJob) and the batch-sharded calling path (no collective).  Skipped on a one-GPU box; the
same host logic runs over gloo on CPU in tests/test_parallel_gloo.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    import workloads as wl
    from close_kmers_b200 import api, parallel, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    protos, sig, img = wl.small_world(seed=47, n_protos=300, n_sigs=80_000, otu_mode="minus1")
    batch = synth.make_proteins(9, protos, 3000, mix=(0.9, 0.1, 0.0, 0.0))
    eids = np.arange(batch.n, dtype=np.uint32)
    g = api.KmerGuts(image=img, device=rank)
    job = parallel.MatrixJob(g, eids, batch, rank, world, device=torch.device("cuda", rank))
    merged, stats = job.run()
    # calling, sharded by batch: every rank its slice, no collective; the slices concatenate to the whole batch's answer
    a, b = rank * batch.n // world, (rank + 1) * batch.n // world
    o = batch.offsets
    mine = g.process_aa_seq_batch(batch.residues[int(o[a]): int(o[b])], (o[a: b + 1] - o[a]).astype(np.uint64), api.WANT_BEST)["best"]
    np.save(os.path.join(out_dir, f"best{rank}.npy"), mine)
    dist.barrier()
    if rank == 0:
        g2 = api.KmerGuts(image=img, device=0)
        g2.postings_add(eids, batch.residues, batch.offsets)
        whole = api.merge_pairs(g2.matrix_rows(eids, batch.residues, batch.offsets))
        best = g2.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_BEST)["best"]
        parts = np.concatenate([np.load(os.path.join(out_dir, f"best{r}.npy")) for r in range(world)])
        np.save(os.path.join(out_dir, "ok.npy"), np.array([merged.tobytes() == whole.tobytes(), len(whole), parts.tobytes() == best.tobytes(),
                                                           stats["postings"]]))
        g2.close()
    dist.barrier()
    g.close()
    dist.destroy_process_group()


def test_matrix_job_and_sharded_calling_over_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ok, n, same_best, postings = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and n > 1000 and same_best == 1 and postings > 100_000
