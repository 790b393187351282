"""Shared helpers for the test-suite (golden loading, relative error)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def load_golden(name):
    """Return (flat: gcdyn.FlatForest, params: PackedParams, present_time, expected: dict)."""
    import gcdyn
    from gcdyn.likelihood import PackedParams
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    f = {k[5:]: z[k] for k in z.files if k.startswith("flat_")}
    par = {k[4:]: z[k] for k in z.files if k.startswith("par_")}
    N, T = len(f["event"]), len(f["tree_off"]) - 1
    K = len(par["type_space"])
    flat = gcdyn.FlatForest(
        n_types=K, tree_off=f["tree_off"], event=f["event"], type_idx=f["type_idx"], btype_idx=f["btype_idx"],
        t_start=f["t_start"], t_end=f["t_end"], parent=f["parent"], child_off=f["child_off"],
        child_idx=f["child_idx"], live=np.ones(N, dtype=np.uint8), root_type_idx=f["root_type_idx"],
        root_time=f["root_time"], self_loop=np.zeros(T, dtype=np.uint8), event_counts=z["event_counts"])
    P = par["lam"].shape[0]
    G = par["Gamma"]                       # [P,K,K] row-major numpy
    shared = all(np.array_equal(G[0], G[p]) for p in range(P))
    Gcol = (np.ascontiguousarray(G[0].ravel(order="F")) if shared
            else np.ascontiguousarray(np.concatenate([G[p].ravel(order="F") for p in range(P)])))
    params = PackedParams(n_psets=P, n_types=K, response=0, gamma_pset_stride=0 if shared else K * K,
                          lam=np.ascontiguousarray(par["lam"]), xscale=None, xshift=None, yscale=None, yshift=None,
                          type_space=np.ascontiguousarray(par["type_space"]),
                          mu=np.ascontiguousarray(par["mu"]), delta=np.ascontiguousarray(par["delta"]),
                          rho=np.ascontiguousarray(par["rho"]), sigma=np.ascontiguousarray(par["sigma"]), Gamma=Gcol)
    expected = {k: z[k] for k in z.files if not k.startswith(("flat_", "par_"))}
    return flat, params, float(z["present_time"]), expected, par
