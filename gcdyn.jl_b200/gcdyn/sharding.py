"""Tree sharding for multi-GPU evaluation.

Trees are independent (`loglikelihood(model, trees, T)` is a plain sum, mbp.jl:258), so a forest is
split by trees across ranks — balanced by node count, because tree sizes are heavy-tailed — the small
parameter block is replicated, every rank produces its partial `[P]` sums, and ONE all-reduce of `[P]`
doubles combines them.  No other data-path collective exists.
"""
from __future__ import annotations

import numpy as np

from .flatten import FlatForest

__all__ = ["partition_trees", "subset_flat", "shard_flat", "allreduce_sum"]


def partition_trees(sizes, n_shards):
    """Longest-processing-time assignment of trees to shards; returns a list of sorted index arrays.
    Deterministic: ties broken by tree index, so every rank computes the same partition."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.lexsort((np.arange(len(sizes)), -sizes))
    load = np.zeros(n_shards, dtype=np.int64)
    bins = [[] for _ in range(n_shards)]
    for t in order:
        r = int(np.argmin(load))          # first minimum
        bins[r].append(int(t))
        load[r] += sizes[t] + 1           # +1: the per-tree root item
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def subset_flat(flat: FlatForest, tree_ids) -> FlatForest:
    """The sub-forest made of `tree_ids` (in the given order); positions inside trees are unchanged."""
    tree_ids = np.asarray(tree_ids, dtype=np.int64)
    lo, hi = flat.tree_off[tree_ids], flat.tree_off[tree_ids + 1]
    n = hi - lo
    tree_off = np.zeros(len(tree_ids) + 1, dtype=np.int64)
    np.cumsum(n, out=tree_off[1:])
    if len(tree_ids):
        node_sel = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]) if n.sum() else np.zeros(0, dtype=np.int64)
    else:
        node_sel = np.zeros(0, dtype=np.int64)
    c_lo, c_hi = flat.child_off[node_sel], flat.child_off[node_sel + 1]
    cn = c_hi - c_lo
    child_off = np.zeros(len(node_sel) + 1, dtype=np.int64)
    np.cumsum(cn, out=child_off[1:])
    child_sel = (np.concatenate([np.arange(a, b) for a, b in zip(c_lo, c_hi)])
                 if cn.sum() else np.zeros(0, dtype=np.int64))
    take = lambda a: np.ascontiguousarray(a[node_sel])
    return FlatForest(
        n_types=flat.n_types, tree_off=tree_off, event=take(flat.event), type_idx=take(flat.type_idx),
        btype_idx=take(flat.btype_idx), t_start=take(flat.t_start), t_end=take(flat.t_end),
        parent=take(flat.parent), child_off=child_off, child_idx=np.ascontiguousarray(flat.child_idx[child_sel]),
        live=take(flat.live), root_type_idx=np.ascontiguousarray(flat.root_type_idx[tree_ids]),
        root_time=np.ascontiguousarray(flat.root_time[tree_ids]),
        self_loop=np.ascontiguousarray(flat.self_loop[tree_ids]),
        event_counts=None if flat.event_counts is None else np.ascontiguousarray(flat.event_counts[tree_ids]))


def shard_flat(flat: FlatForest, n_shards):
    """Split a flattened forest into `n_shards` balanced sub-forests."""
    sizes = np.diff(flat.tree_off)
    return [subset_flat(flat, ids) for ids in partition_trees(sizes, n_shards)]


def allreduce_sum(tensor_or_array, group=None):
    """Sum partial per-pset log-likelihoods across ranks with one all-reduce (NCCL on GPUs, gloo on
    CPU).  Accepts a torch tensor (reduced in place, returned) or a numpy array (returned as numpy)."""
    import torch
    import torch.distributed as dist
    if isinstance(tensor_or_array, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(tensor_or_array, dtype=np.float64).copy())
        if dist.is_initialized() and dist.get_backend(group) == "nccl":
            t = t.cuda()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor_or_array, op=dist.ReduceOp.SUM, group=group)
    return tensor_or_array


class PeerAllReduce:
    """`gcdyn_comm_*`: all-reduce of the `[P]` per-pset sums with one kernel per rank over NVLink peer memory
    (P2P stores into every peer's window + sequence flags) instead of a library collective.  One process per
    GPU; the 64-byte CUDA IPC handles are exchanged once through `torch.distributed` (any backend).  Every
    rank ends with bit-identical sums.  Collective: every rank calls `allreduce` the same number of times."""

    def __init__(self, ctx, max_psets, group=None, row_bytes=0):
        """`row_bytes` > 0 also creates the ROW WINDOW of the trajectory sharding (`gcdyn_comm_rows_create`): sharded ODE
        evaluations of >= 512 parameter sets then integrate only this rank's slice of the sets and all-gather the
        coefficient rows over peer memory.  `row_window_bytes(P, K, degree)` sizes it."""
        import ctypes as C
        import torch.distributed as dist
        from .likelihood import _lib, _check
        self._lib, self._check, self.ctx = _lib(), _check, ctx
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        handle = C.create_string_buffer(64)
        h = C.c_void_p()
        _check(self._lib.gcdyn_comm_create(ctx._h, self.rank, self.world, int(max_psets), handle, C.byref(h)), ctx._h)
        self._h = h
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handle.raw, group=group)
            blob = C.create_string_buffer(b"".join(gathered), 64 * self.world)
            _check(self._lib.gcdyn_comm_connect(self._h, blob), ctx._h)
            dist.barrier(group)        # every window is mapped everywhere before the first store
        if row_bytes:
            rh = C.create_string_buffer(64)
            _check(self._lib.gcdyn_comm_rows_create(self._h, int(row_bytes), rh), ctx._h)
            if self.world > 1:
                gathered = [None] * self.world
                dist.all_gather_object(gathered, rh.raw, group=group)
                blob = C.create_string_buffer(b"".join(gathered), 64 * self.world)
                _check(self._lib.gcdyn_comm_rows_connect(self._h, blob), ctx._h)
                dist.barrier(group)

    @staticmethod
    def row_window_bytes(n_psets, n_types, degree=128):
        """Bytes `row_bytes` needs for batches of `n_psets` on a nodal grid of `degree` (128 = the largest the library
        uses): two halves of P rows of Jst = 2K+4 + (degree+1)·K + K² doubles (rounded up to 4) plus 32 bytes per set."""
        K = int(n_types)
        j1 = (2 * K + 4 + (int(degree) + 1) * K + 3) & ~3
        jst = (j1 + K * K + 3) & ~3
        return 2 * int(n_psets) * (jst * 8 + 32)

    def error(self):
        """`gcdyn_comm_error`: the communicator's error word, cleared on read (call after synchronising a resident
        evaluation): 0 none, 1 a peer did not arrive in time, 2 the ranks disagreed on the grid degree."""
        return int(self._lib.gcdyn_comm_error(self._h))

    def allreduce(self, d_inout, n_psets, stream=0):
        """Enqueue on `stream` (raw cudaStream_t, 0 = the context's stream): `d_inout` (raw device address of
        `[n_psets]` doubles) is summed over the ranks in place."""
        import ctypes as C
        self._check(self._lib.gcdyn_comm_allreduce(self._h, C.c_void_p(d_inout), int(n_psets), C.c_void_p(stream or None)),
                    self.ctx._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gcdyn_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
