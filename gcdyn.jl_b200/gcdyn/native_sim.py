"""Native forward simulator: `rand_tree`'s law (mbp.jl:318-461) executed by the C++ host code of
libgcdyn_b200 (OpenMP over trees), emitting the flattened forest directly — no `TreeNode` objects.
Used for large synthetic workloads (BASELINE config 5: 10⁴ trees of ~10³ nodes)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .flatten import FlatForest
from .likelihood import _check, _lib, _ptr, _p_f64, _p_i32, _p_i64, _p_u8, pack_params
from .multitype_branching_process import type_space_index
from .types import ArgumentError

__all__ = ["simulate_flat"]


def simulate_flat(model, present_time, init_type, n, *, seed=0, first_tree_index=0, reject_stubs=True,
                  min_leaves=0, max_rejections=500) -> FlatForest:
    """`n` trees from `rand_tree(model, present_time, init_type, n; reject_stubs, min_leaves, max_rejections)`'s
    law, already flattened.  Tree i uses the RNG stream (seed, first_tree_index + i): forests are
    reproducible piecewise and independent of the thread count."""
    lib = _lib()
    idx = type_space_index(model, init_type)
    if idx is None:
        raise ArgumentError("The type must be in the type space of the model.")
    pp = pack_params(model)
    ps = pp.struct()
    h = C.c_void_p()
    _check(lib.gcdyn_sim_forest(C.byref(ps), float(present_time), int(idx), int(n), C.c_uint64(int(seed)),
                                int(first_tree_index), int(bool(reject_stubs)), int(min_leaves), int(max_rejections),
                                C.byref(h)))
    try:
        nt, nn, nc, nr = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib.gcdyn_sim_sizes(h, C.byref(nt), C.byref(nn), C.byref(nc), C.byref(nr)))
        T, N, NC = nt.value, nn.value, nc.value
        a = dict(tree_off=np.zeros(T + 1, np.int64), event=np.zeros(N, np.uint8), type_idx=np.zeros(N, np.int32),
                 btype_idx=np.zeros(N, np.int32), t_start=np.zeros(N), t_end=np.zeros(N), parent=np.zeros(N, np.int32),
                 child_off=np.zeros(N + 1, np.int64), child_idx=np.zeros(NC, np.int32),
                 root_type_idx=np.zeros(T, np.int32), root_time=np.zeros(T))
        _check(lib.gcdyn_sim_export(h, _ptr(a["tree_off"], _p_i64), _ptr(a["event"], _p_u8), _ptr(a["type_idx"], _p_i32),
                                    _ptr(a["btype_idx"], _p_i32), _ptr(a["t_start"], _p_f64), _ptr(a["t_end"], _p_f64),
                                    _ptr(a["parent"], _p_i32), _ptr(a["child_off"], _p_i64), _ptr(a["child_idx"], _p_i32),
                                    _ptr(a["root_type_idx"], _p_i32), _ptr(a["root_time"], _p_f64)))
    finally:
        lib.gcdyn_sim_destroy(h)
    counts = np.zeros((T, 7), dtype=np.int64)
    if N:
        np.add.at(counts, (np.repeat(np.arange(T), np.diff(a["tree_off"])), a["event"]), 1)
    flat = FlatForest(n_types=len(model.type_space), live=np.ones(N, np.uint8), self_loop=np.zeros(T, np.uint8),
                      event_counts=counts, **a)
    flat.rejections = nr.value
    return flat
