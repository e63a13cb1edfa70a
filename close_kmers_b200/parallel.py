"""Multi-GPU plumbing for the one stage with an exchange step: /matrix.

Calling, family voting and the fastq path shard by batch with no data-path collective (see bench.py).  The matrix is
sharded by ROW BLOCK: every rank holds the full postings index, computes rows [begin, end) of the strictly-lower-
triangular count matrix with ``KmerGuts.matrix_rows`` and the COO tiles are gathered with torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests) and merged in (eid_i, eid_j) order.
"""
from __future__ import annotations

import numpy as np

PAIR_DT = np.dtype([("eid_i", "<u4"), ("eid_j", "<u4"), ("count", "<u8")])


def shard_rows(work, world: int):
    """Contiguous row blocks with (nearly) equal total ``work`` (e.g. residues per protein: a row's cost is its hits
    times the postings they walk, which does not depend on the row index).  Returns [(begin, end)] * world."""
    work = np.asarray(work, dtype=np.float64)
    n = len(work)
    if n == 0:
        return [(0, 0)] * world
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def gather_pairs(local: np.ndarray, group=None, device=None) -> np.ndarray:
    """all_gather of variable-size COO tiles; returns the concatenation over ranks (rank order)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, PAIR_DT)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    nbytes = torch.tensor([local.nbytes], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, nbytes, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 16)
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if local.nbytes:
        buf[: local.nbytes] = torch.from_numpy(local.view(np.uint8).copy()).to(device)
    outs = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    parts = [o[:s].cpu().numpy().view(PAIR_DT) for o, s in zip(outs, sizes)]
    return np.concatenate(parts) if parts else local


def matrix_sharded(compute_rows, lengths, rank: int, world: int, group=None, device=None) -> np.ndarray:
    """Row-block sharded matrix: ``compute_rows(begin, end)`` returns this rank's unordered COO entries (e.g.
    ``lambda a, b: guts.matrix_rows(eids, residues, offsets, a, b)``); the tiles are gathered and merged."""
    from . import api
    begin, end = shard_rows(lengths, world)[rank]
    local = compute_rows(begin, end) if end > begin else np.zeros(0, PAIR_DT)
    return api.merge_pairs(gather_pairs(local, group, device))
