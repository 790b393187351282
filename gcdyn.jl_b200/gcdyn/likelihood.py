"""Likelihood entry points: flatten on the host, evaluate on the B200 through the C ABI.

Replaces `StatsAPI.loglikelihood(model, tree(s), present_time; …)` (mbp.jl:110-259) and the
alternates of extra_likelihoods.jl:10-151.  This module is the Python stand-in for the Julia
`ccall` wrapper: it only packs arrays and calls `libgcdyn_b200.so`
(`include/gcdyn_b200.h`).  There is NO CPU fallback — if the library or a CUDA device is
missing every call raises `GcdynError`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings

import numpy as np

from .types import (AbstractBranchingProcess, ArgumentError, ConstantBranchingProcess,
                    DiscreteBranchingProcess, DomainError, SigmoidalBranchingProcess, TreeNode)
from .flatten import FlatForest, flatten

__all__ = ["loglikelihood_and_gradient", "loglikelihood", "naive_loglikelihood", "stadler_appx_loglikelihood",
           "stadler_appx_unconditioned_loglikelihood", "Forest", "Context", "pack_params",
           "DeviceParams", "evaluate", "evaluate_resident", "GcdynError", "library_path"]

METHOD_CODE = {"ode": 0, "naive": 1, "stadler": 2, "stadler_unconditioned": 3}
STATUS_SELF_LOOP, STATUS_SOLVER, STATUS_DOMAIN = 1, 2, 4
FLAG_SWEEP_ONLY = 4
FLAG_TOTALS = 8

_ERRNAMES = {-1: "EINVAL", -2: "ECUDA", -3: "ENOMEM", -4: "EUNSUPPORTED", -5: "ENODEVICE"}


class GcdynError(RuntimeError):
    """The native library reported an error (or could not be loaded)."""

    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


# ----------------------------------------------------------------------------- ctypes mirror of the header

_p_f64 = C.POINTER(C.c_double)
_p_i32 = C.POINTER(C.c_int32)
_p_i64 = C.POINTER(C.c_int64)
_p_u8 = C.POINTER(C.c_uint8)


class ForestDesc(C.Structure):
    _fields_ = [("n_trees", C.c_int64), ("n_nodes", C.c_int64), ("n_types", C.c_int32), ("reserved", C.c_int32),
                ("tree_off", _p_i64), ("event", _p_u8), ("type_idx", _p_i32), ("btype_idx", _p_i32),
                ("t_start", _p_f64), ("t_end", _p_f64), ("parent", _p_i32), ("child_off", _p_i64),
                ("child_idx", _p_i32), ("live", _p_u8), ("root_type_idx", _p_i32), ("root_time", _p_f64),
                ("self_loop", _p_u8)]


class Params(C.Structure):
    _fields_ = [("n_psets", C.c_int32), ("n_types", C.c_int32), ("response", C.c_int32),
                ("gamma_pset_stride", C.c_int32),
                ("lambda_", _p_f64), ("xscale", _p_f64), ("xshift", _p_f64), ("yscale", _p_f64), ("yshift", _p_f64),
                ("type_space", _p_f64), ("mu", _p_f64), ("delta", _p_f64), ("rho", _p_f64), ("sigma", _p_f64),
                ("Gamma", _p_f64)]


class MhDesc(C.Structure):
    """`gcdyn_mh_desc` (include/gcdyn_b200.h)."""
    _fields_ = [("n_chains", C.c_int32), ("n_types", C.c_int32), ("type_space", _p_f64), ("Gamma", _p_f64),
                ("rho", C.c_double), ("sigma", C.c_double), ("theta0", _p_f64), ("lower", _p_f64), ("upper", _p_f64),
                ("step", C.c_double), ("seed", C.c_uint64)]


class Opts(C.Structure):
    _fields_ = [("steps", C.c_int32), ("order", C.c_int32), ("tol", C.c_double), ("flags", C.c_int32),
                ("max_cells", C.c_int32)]


_LIB = None


def library_path():
    """Where the in-tree build puts the shared library (override: $GCDYN_B200_LIB)."""
    env = os.environ.get("GCDYN_B200_LIB")
    if env:
        return env
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libgcdyn_b200.so")


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise GcdynError(f"{path} not found: build it with `python __graft_entry__.py build` "
                         "(there is no CPU fallback for the likelihood path)")
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise GcdynError(f"cannot load {path}: {e} (there is no CPU fallback)") from e
    vp = C.c_void_p
    sig = {
        "gcdyn_version": (C.c_int, []),
        "gcdyn_device_count": (C.c_int, []),
        "gcdyn_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "gcdyn_ctx_destroy": (None, [vp]),
        "gcdyn_last_error": (C.c_char_p, [vp]),
        "gcdyn_forest_create": (C.c_int, [vp, C.POINTER(ForestDesc), C.POINTER(vp)]),
        "gcdyn_forest_destroy": (None, [vp]),
        "gcdyn_forest_info": (C.c_int, [vp, _p_i64, _p_i64, _p_i64, _p_i32]),
        "gcdyn_forest_event_counts": (C.c_int, [vp, vp, _p_i64]),
        "gcdyn_loglik_batch": (C.c_int, [vp, vp, C.POINTER(Params), C.c_double, C.c_int, C.POINTER(Opts),
                                         _p_f64, _p_f64, _p_i32]),
        "gcdyn_params_upload": (C.c_int, [vp, C.POINTER(Params), C.POINTER(vp)]),
        "gcdyn_dparams_destroy": (None, [vp]),
        "gcdyn_loglik_batch_resident": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.POINTER(Opts),
                                                  vp, vp, vp, vp]),
        "gcdyn_loglik_batch_resident_sharded": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.POINTER(Opts), vp,
                                                          vp, vp, vp, vp]),
        "gcdyn_loglik_grad_batch": (C.c_int, [vp, vp, C.POINTER(Params), C.c_double, C.POINTER(Opts), _p_f64, _p_f64, _p_i32]),
        "gcdyn_last_diag": (C.c_int, [vp, _p_i32, _p_i32, _p_i32, _p_f64, C.c_int32]),
        "gcdyn_last_cells": (C.c_int, [vp, _p_i32, C.c_int32]),
        "gcdyn_probe_fp64_peak": (C.c_int, [vp, _p_f64]),
        "gcdyn_last_cheb_degree": (C.c_int, [vp, _p_i32, _p_i32, _p_i32]),
        "gcdyn_sim_forest": (C.c_int, [C.POINTER(Params), C.c_double, C.c_int32, C.c_int64, C.c_uint64, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]),
        "gcdyn_sim_sizes": (C.c_int, [vp, _p_i64, _p_i64, _p_i64, _p_i64]),
        "gcdyn_sim_export": (C.c_int, [vp, _p_i64, _p_u8, _p_i32, _p_i32, _p_f64, _p_f64, _p_i32, _p_i64, _p_i32,
                                       _p_i32, _p_f64]),
        "gcdyn_sim_destroy": (None, [vp]),
        "gcdyn_set_profiling": (C.c_int, [vp, C.c_int]),
        "gcdyn_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "gcdyn_comm_create": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp, C.POINTER(vp)]),
        "gcdyn_comm_connect": (C.c_int, [vp, vp]),
        "gcdyn_comm_rows_create": (C.c_int, [vp, C.c_int64, vp]),
        "gcdyn_comm_rows_connect": (C.c_int, [vp, vp]),
        "gcdyn_comm_error": (C.c_int, [vp]),
        "gcdyn_last_traj_sharded": (C.c_int, [vp]),
        "gcdyn_comm_allreduce": (C.c_int, [vp, vp, C.c_int32, vp]),
        "gcdyn_comm_destroy": (None, [vp]),
        "gcdyn_loglik_batch_sharded": (C.c_int, [vp, vp, C.POINTER(Params), C.c_double, C.c_int, C.POINTER(Opts), vp,
                                                 _p_f64, _p_f64, _p_i32]),
        "gcdyn_mh_create": (C.c_int, [vp, vp, C.POINTER(MhDesc), C.c_double, vp, C.POINTER(vp)]),
        "gcdyn_mh_run": (C.c_int, [vp, C.c_int64, _p_f64, _p_f64]),
        "gcdyn_mh_state": (C.c_int, [vp, _p_f64, _p_f64, _p_i64, _p_i64]),
        "gcdyn_mh_restore": (C.c_int, [vp, _p_f64, _p_f64, _p_i64, C.c_int64]),
        "gcdyn_mh_destroy": (None, [vp]),
        "gcdyn_mh_uniform": (C.c_double, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)       # AttributeError here = header and library out of sync
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


_NULL_F64, _NULL_I32 = _p_f64(), _p_i32()


def _check(code, ctx_handle=None):
    if code == 0:
        return
    msg = _lib().gcdyn_last_error(ctx_handle)
    msg = msg.decode("utf-8", "replace") if msg else ""
    raise GcdynError(f"libgcdyn_b200: {_ERRNAMES.get(code, code)}: {msg}", code)


# ----------------------------------------------------------------------------- handles

class Context:
    """`gcdyn_ctx`: one CUDA device, its stream and workspace."""

    _default = {}

    def __init__(self, device=0):
        lib = _lib()
        h = C.c_void_p()
        _check(lib.gcdyn_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @classmethod
    def default(cls, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("GCDYN_B200_USE_LOCAL_RANK") else 0
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def close(self):
        if getattr(self, "_h", None):
            _lib().gcdyn_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_diag(self, n_psets=0):
        s, o, nl = C.c_int32(), C.c_int32(), C.c_int32()
        err = np.zeros(max(n_psets, 1))
        _check(_lib().gcdyn_last_diag(self._h, C.byref(s), C.byref(o), C.byref(nl),
                                      _ptr(err, _p_f64) if n_psets else _p_f64(), int(n_psets)), self._h)
        out = dict(steps=s.value, order=o.value, n_launches=nl.value, err_indicator=err[:n_psets])
        if n_psets and o.value > 0:
            cells = np.zeros(n_psets, dtype=np.int32)
            _check(_lib().gcdyn_last_cells(self._h, _ptr(cells, _p_i32), int(n_psets)), self._h)
            out["cells"] = cells
        return out

    def last_cheb(self):
        """Chebyshev/GEMM stage of the last ODE evaluation (all zeros if it did not run)."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _check(_lib().gcdyn_last_cheb_degree(self._h, C.byref(a), C.byref(b), C.byref(c)), self._h)
        return dict(interp_degree=a.value, effective_degree=b.value, n_fallback=c.value)

    def set_profiling(self, on=True):
        _check(_lib().gcdyn_set_profiling(self._h, 1 if on else 0), self._h)

    def last_kernel_ms(self):
        """Durations of the stages of the last evaluation, CUDA-event timed (see gcdyn_last_kernel_ms)."""
        ms = (C.c_float * 6)()
        _check(_lib().gcdyn_last_kernel_ms(self._h, ms), self._h)
        return dict(zip(("setup", "trajectory", "tree_const", "gemm", "sweep", "finish"), (float(x) for x in ms)))

    def probe_fp64_peak(self):
        v = C.c_double()
        _check(_lib().gcdyn_probe_fp64_peak(self._h, C.byref(v)), self._h)
        return v.value


class Forest:
    """A flattened forest resident in HBM (`gcdyn_forest`).  Flatten + upload once per dataset,
    then evaluate many parameter sets against it."""

    def __init__(self, trees, type_space, present_time=None, method="ode", ctx=None, *,
                 discrete_lambda=False):
        self.ctx = ctx or Context.default()
        self.method = method
        self.present_time = present_time
        self.flat = trees if isinstance(trees, FlatForest) else flatten(
            trees, type_space, present_time, method, discrete_lambda=discrete_lambda)
        f = self.flat
        d = ForestDesc(
            n_trees=f.n_trees, n_nodes=f.n_nodes, n_types=f.n_types, reserved=0,
            tree_off=_ptr(f.tree_off, _p_i64), event=_ptr(f.event, _p_u8), type_idx=_ptr(f.type_idx, _p_i32),
            btype_idx=_ptr(f.btype_idx, _p_i32), t_start=_ptr(f.t_start, _p_f64), t_end=_ptr(f.t_end, _p_f64),
            parent=_ptr(f.parent, _p_i32), child_off=_ptr(f.child_off, _p_i64), child_idx=_ptr(f.child_idx, _p_i32),
            live=_ptr(f.live, _p_u8), root_type_idx=_ptr(f.root_type_idx, _p_i32),
            root_time=_ptr(f.root_time, _p_f64), self_loop=_ptr(f.self_loop, _p_u8))
        h = C.c_void_p()
        _check(_lib().gcdyn_forest_create(self.ctx._h, C.byref(d), C.byref(h)), self.ctx._h)
        self._h = h

    @property
    def n_trees(self):
        return self.flat.n_trees

    def info(self):
        a, b, c, k = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        _check(_lib().gcdyn_forest_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(k)))
        return dict(n_trees=a.value, n_nodes=b.value, n_items=c.value, n_types=k.value)

    def event_counts(self):
        """Per-tree event histogram [T,7], computed by a device kernel."""
        out = np.zeros((self.flat.n_trees, 7), dtype=np.int64)
        _check(_lib().gcdyn_forest_event_counts(self.ctx._h, self._h, _ptr(out, _p_i64)), self.ctx._h)
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib().gcdyn_forest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PackedParams:
    """Host arrays of a parameter batch + the ctypes struct pointing at them."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    _FIELDS = ("lam", "xscale", "xshift", "yscale", "yshift", "type_space", "mu", "delta", "rho", "sigma", "Gamma")

    def __setattr__(self, name, value):
        # attaching another array (or changing a size) invalidates the cached struct; contents may change freely
        d = self.__dict__
        d[name] = value
        d["_struct_cache"] = None

    def struct(self):
        """The `gcdyn_params` view of the arrays.  Building the ctypes pointers costs more than a small evaluation
        takes on the GPU, so the struct is cached until an attribute is assigned (`__setattr__`); the cache keeps the
        arrays it points at alive, and their contents may change freely."""
        st = self.__dict__.get("_struct_cache")
        if st is not None:
            return st[1]
        arrays = tuple(getattr(self, f) for f in self._FIELDS)
        st = Params(
            n_psets=self.n_psets, n_types=self.n_types, response=self.response,
            gamma_pset_stride=self.gamma_pset_stride,
            lambda_=_ptr(self.lam, _p_f64), xscale=_ptr(self.xscale, _p_f64), xshift=_ptr(self.xshift, _p_f64),
            yscale=_ptr(self.yscale, _p_f64), yshift=_ptr(self.yshift, _p_f64),
            type_space=_ptr(self.type_space, _p_f64), mu=_ptr(self.mu, _p_f64), delta=_ptr(self.delta, _p_f64),
            rho=_ptr(self.rho, _p_f64), sigma=_ptr(self.sigma, _p_f64), Gamma=_ptr(self.Gamma, _p_f64))
        self.__dict__["_struct_cache"] = (arrays, st)
        return st


def pack_params(models, response="table"):
    """Canonical arrays for a batch of models: λ per type, μ, δ, ρ, σ per pset,
    Γ column-major (shared when every model holds the same matrix).

    response = "table" evaluates λ on the host exactly as the reference does (mbp.jl:27-43);
    "sigmoid" ships the four response parameters and lets the device evaluate them
    (all models must then be `SigmoidalBranchingProcess`)."""
    if isinstance(models, AbstractBranchingProcess):
        models = [models]
    models = list(models)
    if not models:
        raise ArgumentError("need at least one model")
    ts = models[0].type_space
    K = len(ts)
    for m in models[1:]:
        if len(m.type_space) != K or not np.array_equal(m.type_space, ts):
            raise ArgumentError("all models of a batch must share one type_space")
    P = len(models)
    f = lambda name: np.ascontiguousarray([getattr(m, name) for m in models], dtype=np.float64)
    mu, delta, rho, sigma = f("μ"), f("δ"), f("ρ"), f("σ")
    G0 = models[0].Γ
    shared = all(m.Γ is G0 or np.array_equal(m.Γ, G0) for m in models[1:])
    if shared:
        Gamma = np.ascontiguousarray(G0.ravel(order="F"))
        stride = 0
    else:
        Gamma = np.ascontiguousarray(np.concatenate([m.Γ.ravel(order="F") for m in models]))
        stride = K * K
    lam = xs = xsh = ys = ysh = None
    if response == "sigmoid":
        if not all(isinstance(m, SigmoidalBranchingProcess) for m in models):
            raise ArgumentError("response='sigmoid' needs SigmoidalBranchingProcess models")
        xs, xsh, ys, ysh = f("λ_xscale"), f("λ_xshift"), f("λ_yscale"), f("λ_yshift")
        code = 2
    elif all(isinstance(m, ConstantBranchingProcess) for m in models):
        lam = f("λ")
        code = 1
    else:
        lam = np.ascontiguousarray(np.stack([m.birth_rates() for m in models]), dtype=np.float64)
        code = 0
    return PackedParams(n_psets=P, n_types=K, response=code, gamma_pset_stride=stride, lam=lam,
                        xscale=xs, xshift=xsh, yscale=ys, yshift=ysh,
                        type_space=np.ascontiguousarray(ts, dtype=np.float64),
                        mu=mu, delta=delta, rho=rho, sigma=sigma, Gamma=Gamma)


class DeviceParams:
    """A parameter batch resident in HBM (`gcdyn_dparams`)."""

    def __init__(self, params: "PackedParams", ctx=None):
        self.ctx = ctx or Context.default()
        self.n_psets, self.n_types = params.n_psets, params.n_types
        ps = params.struct()
        h = C.c_void_p()
        _check(_lib().gcdyn_params_upload(self.ctx._h, C.byref(ps), C.byref(h)), self.ctx._h)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib().gcdyn_dparams_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def evaluate_resident(forest, dparams, present_time, d_out_pset, d_out_tree_pset=0, d_status=0, stream=0,
                      method="ode", steps=0, order=0, flags=0, comm=None):
    """`gcdyn_loglik_batch_resident`: everything already in HBM; `d_*` are raw device addresses
    (e.g. `tensor.data_ptr()`), `stream` a raw cudaStream_t (e.g. `torch.cuda.current_stream().cuda_stream`).
    Enqueues and returns without synchronising."""
    op = Opts(steps=int(steps), order=int(order), tol=0.0, flags=int(flags), max_cells=0)
    if comm is not None:            # the forest is this rank's shard: d_out_pset receives the sums over all ranks
        _check(_lib().gcdyn_loglik_batch_resident_sharded(
            forest.ctx._h, forest._h, dparams._h, float(present_time if present_time is not None else 0.0),
            METHOD_CODE[method], C.byref(op), comm._h, C.c_void_p(d_out_pset), C.c_void_p(d_out_tree_pset or None),
            C.c_void_p(d_status or None), C.c_void_p(stream or None)), forest.ctx._h)
        return
    _check(_lib().gcdyn_loglik_batch_resident(
        forest.ctx._h, forest._h, dparams._h, float(present_time if present_time is not None else 0.0),
        METHOD_CODE[method], C.byref(op), C.c_void_p(d_out_pset), C.c_void_p(d_out_tree_pset or None),
        C.c_void_p(d_status or None), C.c_void_p(stream or None)), forest.ctx._h)


# ----------------------------------------------------------------------------- evaluation

def _opts(reltol=None, abstol=None, steps=0, order=0, flags=0, maxiters=None):
    tol = 0.0
    for t in (reltol, abstol):
        if t is not None and 0 < t < 1e-11:   # we are always at least as tight as requested
            tol = t if tol == 0.0 else min(tol, t)
    # the reference's `maxiters` bounds the solver's steps per solve; here it bounds the adaptive cells per trajectory
    cap = 0 if maxiters is None or not (maxiters < 2 ** 31 - 1) else max(1, int(maxiters))
    return Opts(steps=int(steps), order=int(order), tol=float(tol), flags=int(flags), max_cells=cap)


def evaluate(forest: Forest, params: PackedParams, present_time, method="ode", *, per_tree=False,
             want_status=True, reltol=None, abstol=None, steps=0, order=0, flags=0, comm=None, maxiters=None):
    """One `gcdyn_loglik_batch` call.  Returns (out_pset[P], out_tree_pset[P,T] | None, status[P,T] | None).
    `comm` (a `gcdyn.sharding.PeerAllReduce`): the forest is this rank's shard and `out_pset` is summed over the
    ranks inside the call (`gcdyn_loglik_batch_sharded`)."""
    lib = _lib()
    P, T = params.n_psets, forest.flat.n_trees
    if params.n_types != forest.flat.n_types:
        raise ArgumentError("model type_space and forest disagree on the number of types")
    _check_present_time(forest, present_time, method)
    out = np.empty(P, dtype=np.float64)
    tp = np.empty((P, T), dtype=np.float64) if per_tree else None
    st = np.zeros((P, T), dtype=np.int32) if (want_status or per_tree) else None
    ps = params.struct()
    op = _opts(reltol, abstol, steps, order, flags, maxiters)
    T = float(present_time if present_time is not None else 0.0)
    if comm is not None:
        _check(lib.gcdyn_loglik_batch_sharded(forest.ctx._h, forest._h, C.byref(ps), T, METHOD_CODE[method], C.byref(op),
                                              comm._h, _ptr(out, _p_f64), _ptr(tp, _p_f64), _ptr(st, _p_i32)), forest.ctx._h)
    else:
        rc = lib.gcdyn_loglik_batch(forest.ctx._h, forest._h, C.byref(ps), T, METHOD_CODE[method], C.byref(op),
                                    out.ctypes.data_as(_p_f64), _NULL_F64 if tp is None else tp.ctypes.data_as(_p_f64),
                                    _NULL_I32 if st is None else st.ctypes.data_as(_p_i32))
        if rc:
            _check(rc, forest.ctx._h)
    return out, tp, st


def loglikelihood_and_gradient(forest: Forest, params: PackedParams, present_time, *, order=0):
    """`gcdyn_loglik_grad_batch`: per-parameter-set log-likelihood [P] and its exact gradient [P, n_dirs] by forward
    sensitivities of the trajectory — the derivatives the reference gets from ForwardDiff through its `T<:Real` structs
    (types.jl:92, mbp.jl:118-122).  Directions: (xscale, xshift, yscale, yshift, μ, δ) for `pack_params(..., "sigmoid")`,
    (λ, μ, δ) for constant-rate models."""
    lib = _lib()
    P = params.n_psets
    if params.n_types != forest.flat.n_types:
        raise ArgumentError("model type_space and forest disagree on the number of types")
    _check_present_time(forest, present_time, "ode")
    out = np.empty(P, dtype=np.float64)
    grad = np.empty(P * 6, dtype=np.float64)
    nd = C.c_int32()
    ps = params.struct()
    op = _opts(None, None, 0, order, 0)
    _check(lib.gcdyn_loglik_grad_batch(forest.ctx._h, forest._h, C.byref(ps), float(present_time), C.byref(op),
                                       _ptr(out, _p_f64), _ptr(grad, _p_f64), C.byref(nd)), forest.ctx._h)
    return out, grad[:P * nd.value].reshape(P, nd.value).copy()


def _check_present_time(forest, present_time, method):
    """The flattener validated leaf times against the forest's own present_time (survivors must sit exactly there,
    mbp.jl:125-130); evaluating at another one would extrapolate silently where the reference raises."""
    if method == "naive" or forest.present_time is None or present_time is None:
        return
    if float(present_time) != float(forest.present_time):
        raise ArgumentError(f"present_time {present_time} differs from the {forest.present_time} this forest was flattened for: "
                            "Sampled survival leaf must be at present time")


def _report(status, method):
    """Re-issue the reference's `@warn`s / exceptions from the status words."""
    if status is None:
        return
    bits = int(np.bitwise_or.reduce(status, axis=None)) if status.size else 0
    if bits & STATUS_DOMAIN:
        raise DomainError("log or sqrt of a negative number while evaluating the likelihood "
                          "(Julia raises DomainError here)")
    if bits & STATUS_SELF_LOOP:
        warnings.warn("Self-loop encountered at a type change event in the tree. Density will evaluate to zero.")
    if bits & STATUS_SOLVER:
        warnings.warn("Extinction probabilities left [0, 1] or became non-finite while integrating. "
                      "Density will evaluate to zero.")


def _run(method, model, trees, present_time, **kw):
    single = isinstance(model, AbstractBranchingProcess)
    models = [model] if single else list(model)
    params = pack_params(models)
    if isinstance(trees, Forest):
        forest, own = trees, False
    else:
        forest = Forest(trees, models[0].type_space, present_time, method,
                        discrete_lambda=isinstance(models[0], DiscreteBranchingProcess))
        own = True
    try:
        out, _, st = evaluate(forest, params, present_time, method, **kw)
    finally:
        if own:
            forest.close()
    _report(st, method)
    return float(out[0]) if single else out


def loglikelihood(model, trees, present_time, *, reltol=1e-3, abstol=1e-3, maxiters=1e5, steps=0, order=0):
    """`loglikelihood(model, tree, present_time; reltol, abstol, maxiters)` and the
    `trees::Vector{TreeNode}` method (mbp.jl:110-117, 250-259).

    `model` may also be a list of models — the batched entry — returning one value per model.
    `trees` may be a `TreeNode`, a list of them, or a `Forest` already resident on the GPU.
    The result is the CONVERGED solution (error indicator ≤ 1e-13), which is at least as accurate
    as any `reltol`/`abstol` the reference accepts.  `maxiters` caps the adaptive cells of a trajectory (the
    integrator's counterpart of solver steps): a parameter set that needs more is a solver failure — a warning and
    `-Inf`, as the reference does when `sol.retcode != :Success` (mbp.jl:154-158)."""
    return _run("ode", model, trees, present_time, reltol=reltol, abstol=abstol, steps=steps, order=order, maxiters=maxiters)


def _alt(method, model, tree, present_time):
    if isinstance(tree, (list, tuple)):
        raise TypeError(f"{method}: the reference has no method for a vector of trees; sum over trees "
                        "or pass a Forest")
    return _run(method, model, tree, present_time)


def naive_loglikelihood(model, tree):
    """`gcdyn.naive_loglikelihood(model, tree)` (extra_likelihoods.jl:10-34)."""
    return _alt("naive", model, tree, None)


def stadler_appx_loglikelihood(model, tree, present_time):
    """`gcdyn.stadler_appx_loglikelihood(model, tree, present_time)` (extra_likelihoods.jl:45-102)."""
    return _alt("stadler", model, tree, present_time)


def stadler_appx_unconditioned_loglikelihood(model, tree, present_time):
    """`gcdyn.stadler_appx_unconditioned_loglikelihood(model, tree, present_time)` (extra_likelihoods.jl:115-151)."""
    return _alt("stadler_unconditioned", model, tree, present_time)
