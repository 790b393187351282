"""ctypes wrapper around oracle/refcpu.c (CPU oracle / CPU baseline — test infrastructure only).
Build: `gcc -O2 -fopenmp -fPIC -shared -std=c11 -o oracle/lib/librefcpu.so oracle/refcpu.c -lm`
(done by `__graft_entry__.build()` and by tests/conftest.py)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "lib", "librefcpu.so")
        lib = C.CDLL(path)
        lib.refcpu_loglik.restype = C.c_long
        lib.refcpu_max_threads.restype = C.c_int
        _LIB = lib
    return _LIB


def max_threads():
    return int(_lib().refcpu_max_threads())


def loglik(flat, par, present_time, rtol=1e-3, atol=1e-3, maxiters=100000, nthreads=0, return_nrhs=False):
    """Per-(pset, tree) log-likelihoods [P, T] by the reference algorithm as written.

    `flat` has the attributes of the flattened forest (tree_off, event, type_idx, btype_idx, t_start,
    t_end, child_off, child_idx, root_type_idx); `par` is a dict with lam [P,K], mu, delta, rho,
    sigma [P] and Gamma [P,K,K] or [K,K] (row-major, as numpy)."""
    lib = _lib()
    g = lambda name: getattr(flat, name) if hasattr(flat, name) else flat[name]
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    tree_off, event = c(g("tree_off"), np.int64), c(g("event"), np.uint8)
    type_idx, btype_idx = c(g("type_idx"), np.int32), c(g("btype_idx"), np.int32)
    t_start, t_end = c(g("t_start"), np.float64), c(g("t_end"), np.float64)
    child_off, child_idx = c(g("child_off"), np.int64), c(g("child_idx"), np.int32)
    root_type = c(g("root_type_idx"), np.int32)
    lam = c(par["lam"], np.float64)
    P, K = lam.shape
    G = c(par["Gamma"], np.float64)
    gstride = K * K if G.ndim == 3 and G.shape[0] == P and P > 1 else 0
    if G.ndim == 3 and gstride == 0:
        G = c(G[0], np.float64)
    mu, delta = c(par["mu"], np.float64), c(par["delta"], np.float64)
    rho, sigma = c(par["rho"], np.float64), c(par["sigma"], np.float64)
    Tn = len(tree_off) - 1
    out = np.empty((P, Tn), dtype=np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    nrhs = lib.refcpu_loglik(C.c_int64(Tn), C.c_int(K), vp(tree_off), vp(event), vp(type_idx), vp(btype_idx),
                             vp(t_start), vp(t_end), vp(child_off), vp(child_idx), vp(root_type), C.c_int(P),
                             vp(lam), vp(mu), vp(delta), vp(rho), vp(sigma), vp(G), C.c_int(gstride),
                             C.c_double(present_time), C.c_double(rtol), C.c_double(atol), C.c_long(int(maxiters)),
                             C.c_int(int(nthreads)), vp(out))
    return (out, int(nrhs)) if return_nrhs else out
