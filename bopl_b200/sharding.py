"""Multi-GPU: candidates shard across the ranks of one box; only the best candidates travel.

One process per GPU (torchrun).  The candidate batch is split row-wise into contiguous blocks (global index =
block offset + local index); every rank scores its block and keeps its local top-k; ONE all-gather of k records
`(value, global index, x[d])` per rank (k*(2+d)*8 bytes, ~1.9 KB at k=20, d=10) follows and every rank runs the same
deterministic merge (best value first, ties -> lowest global index), so all ranks agree on the anchors without a
second collective.  There is no data-path collective: the MC / theta reductions are per candidate.
The reference has no counterpart (its only parallelism is a 4-process pool, uEI.py:91).

On GPUs the exchange is `bopl_topk_allgather` of the C-ABI (`TopkGatherer` below): pack kernel -> one `ncclAllGather` on the compute
stream -> merge kernel, no host synchronisation.  `allgather_topk` / `merge_topk` are the same logic over `torch.distributed` (gloo) in
NumPy — the form the CPU tests (world_size 2) exercise, and the statement of what the device merge must return.
"""
import ctypes

import numpy as np


def shard_bounds(N, world_size, rank):
    """Contiguous block [start, stop) of rank `rank`; block sizes differ by at most one row."""
    base, rem = divmod(int(N), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_topk(vals, idxs, k):
    """Deterministic k-way merge of per-rank records.  vals, idxs: (R, k_r) arrays (NaN value = empty / worst).
    Returns (vals (k,), idxs (k,), flat positions (k,)) ordered best first, ties -> lowest global index."""
    v = np.asarray(vals, dtype=np.float64).reshape(-1)
    i = np.asarray(idxs, dtype=np.int64).reshape(-1)
    key_v = np.where(np.isnan(v), -np.inf, v)
    nan_rank = np.isnan(v).astype(np.int8)
    order = np.lexsort((i, -key_v, nan_rank))          # last key is primary: non-NaN first, value desc, index asc
    order = order[:k]
    return v[order], i[order], order


def pack_records(val, idx, x, k):
    """(k, 2+d) float64 records; rows beyond len(val) are NaN-valued fillers that never win."""
    d = x.shape[1]
    rec = np.full((k, 2 + d), np.nan)
    rec[:, 1] = float(2 ** 53 - 1)
    kk = len(val)
    rec[:kk, 0] = val
    rec[:kk, 1] = idx.astype(np.float64)                 # exact below 2^53
    rec[:kk, 2:] = x
    return rec


def allgather_topk(val, idx, x, k, group=None, device=None):
    """All ranks receive the merged global top-k: (val (k,), idx (k,), x (k,d)).  NCCL when `device` is a CUDA device
    (records stay on the GPU), gloo otherwise."""
    import torch
    import torch.distributed as dist
    rec = torch.from_numpy(pack_records(np.asarray(val), np.asarray(idx), np.asarray(x), k))
    if device is not None:
        rec = rec.to(device)
    world = dist.get_world_size(group)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    allrec = torch.stack(out).cpu().numpy()              # (world, k, 2+d)
    v, i, pos = merge_topk(allrec[:, :, 0], allrec[:, :, 1].astype(np.int64), k)
    xs = allrec.reshape(-1, allrec.shape[2])[pos, 2:]
    keep = ~np.isnan(v)
    return v[keep], i[keep], xs[keep]


class TopkGatherer(object):
    """The library's own NCCL communicator over the ranks of `group` + the fused pack / all-gather / merge call.

    Rank 0 draws the NCCL unique id (`bopl_nccl_unique_id`), the 128 bytes travel through `torch.distributed` (whatever backend the
    group has), every rank joins with `bopl_nccl_comm_init`.  One per engine; `close()` destroys the communicator."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        from . import _lib
        self.engine, self.lib = engine, engine.lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        ident = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            _lib.check(self.lib.bopl_nccl_unique_id(ctypes.c_void_p(ident.ctypes.data)))
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        comm = ctypes.c_void_p()
        _lib.check(self.lib.bopl_nccl_comm_init(engine.h, self.world, self.rank, ctypes.c_void_p(ident.ctypes.data), ctypes.byref(comm)),
                   engine.h)
        self.comm = comm

    def close(self):
        if getattr(self, "comm", None):
            self.lib.bopl_nccl_comm_destroy(self.comm)
            self.comm = None

    def allgather_dev(self, val, idx, Xc, index_offset, k):
        """Device tensors in (this rank's `Engine.topk_dev` output and its candidate block), device tensors out: the global
        (val (k,), idx (k,), x (k,d)), identical on every rank.  Stream-ordered, no host synchronisation."""
        import torch
        eng = self.engine
        k_local = int(val.shape[0])
        val_all = eng.empty(k)
        idx_all = torch.empty(k, dtype=torch.int64, device=eng.device)
        x_all = eng.empty(k, eng.d)
        eng._check(self.lib.bopl_topk_allgather(eng.h, self.comm, val.data_ptr(), idx.data_ptr(), k_local, Xc.data_ptr(),
                                                int(index_offset), int(k), val_all.data_ptr(), idx_all.data_ptr(), x_all.data_ptr(),
                                                eng._stream()))
        return val_all, idx_all, x_all

    def allgather_host(self, val, idx, x, k):
        """NumPy records in (`uEI.top_candidates` output + the matching candidate rows), NumPy out; NaN-valued fillers (ranks with
        fewer than k candidates) are dropped."""
        eng = self.engine
        val = np.ascontiguousarray(val, dtype=np.float64)
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(len(val), eng.d)
        v, i, xs = np.empty(k), np.empty(k, dtype=np.int64), np.empty((k, eng.d))
        p = lambda a: ctypes.c_void_p(a.ctypes.data)
        eng._check(self.lib.bopl_topk_allgather_host(eng.h, self.comm, p(val), p(idx), len(val), p(x), int(k), p(v), p(i), p(xs)))
        keep = ~np.isnan(v)
        return v[keep], i[keep], xs[keep]


class ShardedAnchorSelector(object):
    """Scores this rank's block with `local_top(X_block, k, index_offset) -> (val, global idx)` and merges.

    `local_top` is `uEI.top_candidates` on a GPU box; the gloo tests inject a CPU scorer to exercise the
    partitioning + collective + merge logic without a device."""

    def __init__(self, local_top, group=None, device=None, gatherer=None):
        self.local_top = local_top
        self.group = group
        self.device = device
        self.gatherer = gatherer          # TopkGatherer: the exchange runs through the library's NCCL call instead of torch.distributed

    def select_from_full(self, X, k):
        """Every rank holds (or can generate) the full candidate matrix; it only scores its own block."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        s, e = shard_bounds(len(X), world, rank)
        return self.select_block(X[s:e], s, k)

    def select_block(self, X_block, index_offset, k):
        kk = min(k, len(X_block))
        if kk > 0:
            val, idx = self.local_top(X_block, kk, index_offset)
            x = X_block[np.asarray(idx) - index_offset]
        else:
            val, idx, x = np.empty(0), np.empty(0, dtype=np.int64), np.empty((0, X_block.shape[1]))
        if self.gatherer is not None:
            return self.gatherer.allgather_host(val, idx, x, k)
        return allgather_topk(val, idx, x, k, self.group, self.device)
