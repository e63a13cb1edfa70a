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


class _DeviceView:
    """A raw device pointer as something torch.as_tensor can alias without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _alias(ptr: int, count: int, typestr: str, torch, device):
    if count == 0 or not ptr:
        return torch.zeros(0, dtype={"<u8": torch.int64, "<u4": torch.int32, "|u1": torch.uint8}[typestr], device=device)
    t = torch.as_tensor(_DeviceView(ptr, count, {"<u8": "<i8", "<u4": "<i4", "|u1": "|u1"}[typestr]), device=device)
    return t


def _all_gather_var(local, world, dist, torch, group=None):
    """all_gather of 1-D tensors of different lengths (device tensors over NCCL, host tensors over gloo); returns the list of
    per-rank tensors (views of one buffer)."""
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x) for x in sizes]
    mx = max(max(sizes), 1)
    send = torch.zeros(mx, dtype=local.dtype, device=local.device)
    send[: local.numel()] = local
    recv = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        dist.all_gather(list(recv.view(world, mx).unbind(0)), send, group=group)
    return [recv[r * mx: r * mx + sizes[r]] for r in range(world)]


class GpuMatrixOps:
    """What MatrixJob needs from one rank's engine, on the GPU: the postings and the tile stay in HBM and are handed to the
    collective as torch tensors that alias the library's buffers (no copy)."""

    def __init__(self, guts, device):
        import torch
        self.g, self.torch = guts, torch
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def clear(self):
        self.g.postings_clear()

    def add(self, eids, residues, offsets):
        self.g.postings_add(eids, residues, offsets)

    def export(self):
        dk, de, n = self.g.postings_device()
        return _alias(dk, n, "<u8", self.torch, self.device), _alias(de, n, "<u4", self.torch, self.device)

    def install(self, keys, pegs):
        # the gathered tensors were produced on torch's streams (NCCL, torch.cat); the library copies them on the context's own
        # non-blocking stream, which does not wait for those: finish them first.  (The other direction needs nothing: every
        # library call returns with its stream synchronised.)
        self.torch.cuda.current_stream(self.device).synchronize()
        self.g.postings_import_device(keys.data_ptr(), pegs.data_ptr(), keys.numel())

    def rows(self, eids, residues, offsets, a, b):
        ptr, n_pairs, walked = self.g.matrix_rows_device(eids, residues, offsets, a, b)
        return _alias(ptr, n_pairs * 16, "|u1", self.torch, self.device), walked

    @property
    def postings_count(self):
        return self.g.postings_count


class MatrixJob:
    """/add + /matrix of one request over `world` GPUs (matrix_request.cc:78-95, 130-189):

      1. hit extraction sharded by PROTEIN BLOCK: rank r runs ckm_postings_add on its contiguous block only;
      2. the (k-mer, peg id) postings are exchanged DEVICE TO DEVICE (NCCL all-gather over NVLink on the postings buffers
         themselves; no host bounce) and every rank installs the full set (ckm_postings_import_device);
      3. every rank computes its ROW BLOCK of the strictly-lower-triangular count matrix (ckm_matrix_rows_device); blocks are
         balanced by the posting work of their rows -- hits x posting-list length, for which the residue count stands in
         (every row walks the whole list of each of its hits; the j < i rule filters entries, it does not shorten the walk);
      4. the COO tiles, each already in (eid_i, eid_j) order, are all-gathered device to device and copied back once.

    With world == 1 the same calls run without the collectives.  eids must be the request's ids; when they ascend in request
    order (the usual case: KmerPegMapping::encode_id numbers ids in order of first appearance, kmer.cc:272-286) the
    concatenated tiles are final, otherwise they go through api.merge_pairs.

    `ops` is the per-rank engine: GpuMatrixOps(guts) by default; the CPU test of the sharding / exchange logic
    (tests/test_parallel_gloo.py, gloo, world_size 2) passes the plain-C oracle behind the same five calls."""

    def __init__(self, guts, eids, batch, rank: int, world: int, device=None, group=None, ops=None):
        self.eids, self.batch, self.rank, self.world, self.group = np.ascontiguousarray(eids, np.uint32), batch, rank, world, group
        self.ops = ops if ops is not None else GpuMatrixOps(guts, device)
        lengths = np.diff(batch.offsets.astype(np.int64))
        self.blocks = shard_rows(lengths, world)
        self.ascending = bool(np.all(np.diff(self.eids.astype(np.int64)) > 0)) if len(self.eids) > 1 else True
        self._pinned = None

    def _sub(self, a, b):
        o = self.batch.offsets
        return self.eids[a:b], self.batch.residues[int(o[a]): int(o[b])], (o[a: b + 1] - o[a]).astype(np.uint64)

    def _gather_sizes(self, vals, ref, dist, torch):
        """all_gather of a few int64 per rank -> list of lists"""
        mine = torch.tensor(vals, dtype=torch.int64, device=ref.device)
        got = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(got, mine, group=self.group)
        return [[int(x) for x in g] for g in got]

    def _gather_bytes(self, local, sizes, dist, torch):
        """all_gather of one uint8 tensor per rank (sizes[r] bytes from rank r); returns the per-rank views"""
        mx = (max(max(sizes), 1) + 15) // 16 * 16
        send = torch.zeros(mx, dtype=torch.uint8, device=local.device)
        send[: local.numel()] = local
        recv = torch.empty(self.world * mx, dtype=torch.uint8, device=local.device)
        if local.is_cuda:
            dist.all_gather_into_tensor(recv, send, group=self.group)
        else:
            dist.all_gather(list(recv.view(self.world, mx).unbind(0)), send, group=self.group)
        return [recv[r * mx: r * mx + sizes[r]] for r in range(self.world)]

    def _to_host(self, parts, torch):
        """the tiles, one after the other, in host memory (page-locked and reused from call to call on a GPU)"""
        total = sum(int(p.numel()) for p in parts)
        if not parts or not parts[0].is_cuda:
            return torch.cat(parts).numpy().view(PAIR_DT) if parts else np.zeros(0, PAIR_DT)
        if self._pinned is None or self._pinned.numel() < total:
            self._pinned = torch.empty(max(total, 1 << 20), dtype=torch.uint8, pin_memory=True)
        o = 0
        for p in parts:
            self._pinned[o: o + p.numel()].copy_(p, non_blocking=True)
            o += int(p.numel())
        torch.cuda.current_stream(parts[0].device).synchronize()
        return self._pinned[:total].numpy().view(PAIR_DT).copy()

    def run(self):
        import time
        import torch
        from . import api
        ops, world = self.ops, self.world
        t = [time.perf_counter()]
        a, b = self.blocks[self.rank]
        ops.clear()
        if b > a:
            ops.add(*self._sub(a, b))
        t.append(time.perf_counter())
        if world > 1:
            import torch.distributed as dist
            k_local, p_local = ops.export()
            n_local = int(k_local.numel())
            # two collectives: the counts, then keys and peg ids of a rank as one run of bytes (8 n + 4 n)
            counts = [c[0] for c in self._gather_sizes([n_local], k_local, dist, torch)]
            packed = torch.cat([k_local.view(torch.uint8), p_local.view(torch.uint8)]) if n_local else torch.zeros(0, dtype=torch.uint8, device=k_local.device)
            got = self._gather_bytes(packed, [12 * c for c in counts], dist, torch)
            keys = [g[: 8 * c].view(torch.int64) for g, c in zip(got, counts)]
            pegs = [g[8 * c: 12 * c].view(torch.int32) for g, c in zip(got, counts)]
            ops.install(torch.cat(keys), torch.cat(pegs))
        t.append(time.perf_counter())
        tile, walked = ops.rows(self.eids, self.batch.residues, self.batch.offsets, a, b)
        t.append(time.perf_counter())
        if world > 1:
            info = self._gather_sizes([int(tile.numel()), int(walked)], tile, dist, torch)
            tiles = self._gather_bytes(tile, [i[0] for i in info], dist, torch)
            walked = sum(i[1] for i in info)
            # every rank holds every tile; the host copy is made where the response is written
            merged = self._to_host(tiles, torch) if self.rank == 0 else np.zeros(0, PAIR_DT)
        else:
            merged = self._to_host([tile], torch)
        t.append(time.perf_counter())
        if not self.ascending:
            merged = api.merge_pairs(merged)
        t.append(time.perf_counter())
        stats = {"postings": ops.postings_count, "walked": walked, "row_blocks": self.blocks,
                 "phase_ms": {"add_block": (t[1] - t[0]) * 1e3, "gather_postings": (t[2] - t[1]) * 1e3, "rows": (t[3] - t[2]) * 1e3,
                              "gather_tiles": (t[4] - t[3]) * 1e3, "merge": (t[5] - t[4]) * 1e3}}
        return merged, stats

    def close(self):
        pass
