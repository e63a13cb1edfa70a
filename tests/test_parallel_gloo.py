"""world_size-2 gloo test of the only path with an exchange step: row-block sharded /matrix with a COO tile gather.
The per-rank compute stand-in is the plain-C oracle (this is a CPU test of the host-side sharding / gather / merge)."""
import os
import socket

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, parallel, synth


def test_shard_rows_partitions_and_balances():
    rng = np.random.default_rng(0)
    work = rng.integers(50, 1200, 1000)
    for world in (1, 2, 3, 8):
        blocks = parallel.shard_rows(work, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == len(work)
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
        sums = [work[a:b].sum() for a, b in blocks]
        assert max(sums) <= work.sum() / world + work.max()
    assert parallel.shard_rows([], 4) == [(0, 0)] * 4
    assert parallel.shard_rows([5], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


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
    import torch.distributed as dist
    import cpu_checkers as cc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    protos, sig, img = wl.small_world(seed=41, n_protos=30, n_sigs=8_000, otu_mode="minus1")
    batch = synth.make_proteins(3, protos, 120, mix=(1.0, 0.0, 0.0, 0.0))
    eids = np.arange(batch.n, dtype=np.uint32)
    orc = cc.Oracle().open_image(img)
    orc.postings_new()
    orc.postings_add(eids, batch)
    lengths = np.diff(batch.offsets.astype(np.int64))
    merged = parallel.matrix_sharded(lambda a, b: orc.matrix_rows(eids, batch, a, b), lengths, rank, world)
    if rank == 0:
        whole = api.merge_pairs(orc.matrix_rows(eids, batch))
        np.save(os.path.join(out_dir, "ok.npy"), np.array([merged.tobytes() == whole.tobytes(), len(whole)]))
    dist.barrier()
    dist.destroy_process_group()


def test_matrix_row_blocks_gathered_over_gloo(checkers, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ok, n = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and n > 100


class _OracleMatrixOps:
    """The plain-C oracle behind the five calls MatrixJob makes (host tensors; the collectives run over gloo)."""

    def __init__(self, orc):
        import torch
        self.orc, self.torch = orc, torch
        orc.postings_new()

    def clear(self):
        self.orc.postings_new()

    def add(self, eids, residues, offsets):
        self.orc.postings_add(eids, synth.Batch(np.ascontiguousarray(residues), np.ascontiguousarray(offsets, np.uint64)))

    def export(self):
        keys, eids = self.orc.postings_export()
        return self.torch.from_numpy(keys.view(np.int64).copy()), self.torch.from_numpy(eids.view(np.int32).copy())

    def install(self, keys, pegs):
        self.orc.postings_import(keys.numpy().view(np.uint64), pegs.numpy().view(np.uint32))

    def rows(self, eids, residues, offsets, a, b):
        # ckm_matrix_rows_device leaves every row's entries ordered by partner id: the oracle's rows come merged the same way
        pairs = api.merge_pairs(self.orc.matrix_rows(eids, synth.Batch(residues, offsets), a, b))
        return self.torch.from_numpy(pairs.view(np.uint8).copy()), int(pairs["count"].sum())

    @property
    def postings_count(self):
        return len(self.orc.postings_export()[0])


def _job_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    import cpu_checkers as cc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    protos, sig, img = wl.small_world(seed=43, n_protos=40, n_sigs=10_000, otu_mode="minus1")
    batch = synth.make_proteins(5, protos, 150, mix=(0.9, 0.1, 0.0, 0.0))
    eids = np.arange(batch.n, dtype=np.uint32)
    orc = cc.Oracle().open_image(img)
    job = parallel.MatrixJob(None, eids, batch, rank, world, ops=_OracleMatrixOps(orc))
    merged, stats = job.run()
    if rank == 0:
        orc.postings_new()
        orc.postings_add(eids, batch)
        whole = api.merge_pairs(orc.matrix_rows(eids, batch))
        np.save(os.path.join(out_dir, "job.npy"), np.array([merged.tobytes() == whole.tobytes(), len(whole), stats["postings"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_matrix_job_sharded_add_and_exchange_over_gloo(checkers, tmp_path):
    """MatrixJob at world_size 2: hit extraction by protein block, all-gather of the postings, row blocks, all-gather of the
    tiles -- the host logic of the multi-GPU /matrix with the oracle as each rank's engine."""
    import torch.multiprocessing as mp
    mp.spawn(_job_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ok, n, postings = np.load(os.path.join(str(tmp_path), "job.npy"))
    assert ok == 1 and n > 100 and postings > 1000
