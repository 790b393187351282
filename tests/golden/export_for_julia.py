"""Writes the inputs of the REFERENCE-generated fixtures: `tests/golden/ref_inputs/<name>.txt`.

    python tests/golden/export_for_julia.py            # here (CPU container)
    julia --project=/path/to/gcdyn.jl tests/golden/make_reference_fixtures.jl     # wherever Julia 1.10.2 + the pinned Manifest exist

The Julia script rebuilds every tree as `TreeNode`s, every parameter set as a `DiscreteBranchingProcess` (a λ table — the
canonical form of all three model types) and calls the REAL `loglikelihood(model, tree, T; reltol=1e-12, abstol=1e-12)`,
`naive_loglikelihood`, `stadler_appx_loglikelihood` and `stadler_appx_unconditioned_loglikelihood`; it writes
`tests/golden/ref_<name>.txt`, which `tests/test_reference_fixtures.py` consumes when present.  Julia is not installed in
this image, so no `ref_*.txt` exists yet: parity stays pinned to the oracle.

Cases: C1, C2 (BASELINE configs 1, 2 at full size), a 50-tree × 8-pset slice of C3, and the two likelihood set-ups of the
reference's own tests (test/runtests.jl:20-29 fully observed, :31-38 no extinction) on 50 trees each.

Text format (one record per line, space separated, floats as `repr` = shortest round-trip):
    gcdyn_flat 1 | name | present_time | K | type_space ×K | P | pset p μ δ ρ σ λ_1..λ_K | gamma_shared 0/1 |
    gamma [p] Γ row-major K² | trees T | tree t n_nodes root_type_idx(0-based) root_time |
    node lines in post-order: event_code(0-based index into EVENTS, types.jl:6) type_idx(0-based) time parent_pos(−1 = root)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "gcdyn.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

OUT = os.path.join(HERE, "ref_inputs")


def write_case(name, flat, par, T, path=None):
    """flat: gcdyn.FlatForest; par: dict(lam[P,K], mu, delta, rho, sigma, Gamma[P,K,K] row-major, type_space[K])."""
    P, K = par["lam"].shape
    G = par["Gamma"]
    shared = all(np.array_equal(G[0], G[p]) for p in range(P))
    r = repr
    lines = ["gcdyn_flat 1", f"name {name}", f"present_time {r(float(T))}", f"K {K}",
             "type_space " + " ".join(r(float(x)) for x in par["type_space"]), f"P {P}"]
    for p in range(P):
        lines.append(f"pset {p} " + " ".join(r(float(x)) for x in (par["mu"][p], par["delta"][p], par["rho"][p], par["sigma"][p]))
                     + " " + " ".join(r(float(x)) for x in par["lam"][p]))
    lines.append(f"gamma_shared {1 if shared else 0}")
    for p in range(1 if shared else P):
        lines.append(f"gamma {p} " + " ".join(r(float(x)) for x in G[p].reshape(-1)))
    lines.append(f"trees {flat.n_trees}")
    for t in range(flat.n_trees):
        lo, hi = int(flat.tree_off[t]), int(flat.tree_off[t + 1])
        lines.append(f"tree {t} {hi - lo} {int(flat.root_type_idx[t])} {r(float(flat.root_time[t]))}")
        for i in range(lo, hi):
            lines.append(f"{int(flat.event[i])} {int(flat.type_idx[i])} {r(float(flat.t_end[i]))} {int(flat.parent[i])}")
    path = path or os.path.join(OUT, name + ".txt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def read_case(path):
    """Inverse of write_case (used by the round-trip test and by the fixture consumer)."""
    import gcdyn
    it = iter(open(path).read().split("\n"))
    nxt = lambda: next(it).split()
    assert nxt() == ["gcdyn_flat", "1"]
    name = nxt()[1]
    T = float(nxt()[1])
    K = int(nxt()[1])
    ts = np.array([float(x) for x in nxt()[1:]])
    P = int(nxt()[1])
    mu, de, rho, sig, lam = np.zeros(P), np.zeros(P), np.zeros(P), np.zeros(P), np.zeros((P, K))
    for p in range(P):
        w = nxt()
        assert w[0] == "pset" and int(w[1]) == p
        v = [float(x) for x in w[2:]]
        mu[p], de[p], rho[p], sig[p] = v[:4]
        lam[p] = v[4:]
    shared = int(nxt()[1]) == 1
    G = np.zeros((P, K, K))
    for p in range(1 if shared else P):
        w = nxt()
        G[p] = np.array([float(x) for x in w[2:]]).reshape(K, K)
    if shared:
        G[:] = G[0]
    ntr = int(nxt()[1])
    tree_off = [0]
    ev, ty, te, par, rtype, rtime = [], [], [], [], [], []
    for t in range(ntr):
        w = nxt()
        assert w[0] == "tree" and int(w[1]) == t
        n = int(w[2])
        rtype.append(int(w[3]))
        rtime.append(float(w[4]))
        for _ in range(n):
            a = nxt()
            ev.append(int(a[0])); ty.append(int(a[1])); te.append(float(a[2])); par.append(int(a[3]))
        tree_off.append(tree_off[-1] + n)
    ev = np.array(ev, dtype=np.uint8); ty = np.array(ty, dtype=np.int32); te = np.array(te); par = np.array(par, dtype=np.int32)
    tree_off = np.array(tree_off, dtype=np.int64)
    N = len(ev)
    bt, ts_ = np.zeros(N, dtype=np.int32), np.zeros(N)
    child_cnt = np.zeros(N, dtype=np.int64)
    for t in range(ntr):
        lo = int(tree_off[t])
        for i in range(lo, int(tree_off[t + 1])):
            if par[i] < 0:
                bt[i], ts_[i] = rtype[t], rtime[t]
            else:
                bt[i], ts_[i] = ty[lo + par[i]], te[lo + par[i]]
                child_cnt[lo + par[i]] += 1
    child_off = np.zeros(N + 1, dtype=np.int64)
    child_off[1:] = np.cumsum(child_cnt)
    child_idx = np.zeros(int(child_off[-1]), dtype=np.int32)
    fill = child_off[:-1].copy()
    for t in range(ntr):
        lo = int(tree_off[t])
        for i in range(lo, int(tree_off[t + 1])):        # increasing post-order position = `children` vector order
            if par[i] >= 0:
                j = lo + par[i]
                child_idx[fill[j]] = i - lo
                fill[j] += 1
    flat = gcdyn.FlatForest(n_types=K, tree_off=tree_off, event=ev, type_idx=ty, btype_idx=bt, t_start=ts_, t_end=te,
                            parent=par, child_off=child_off, child_idx=child_idx, live=np.ones(N, dtype=np.uint8),
                            root_type_idx=np.array(rtype, dtype=np.int32), root_time=np.array(rtime),
                            self_loop=np.zeros(ntr, dtype=np.uint8), event_counts=None)
    pdict = dict(lam=lam, mu=mu, delta=de, rho=rho, sigma=sig, Gamma=G, type_space=ts)
    return name, flat, pdict, T


def cases():
    """(name, flat, par, T) for every reference fixture."""
    import gcdyn
    import workloads
    out = []
    for name, cfg, ntrees, npsets in (("c1", "C1", 10, 1), ("c2", "C2", 100, 1), ("c3_slice", "C3", 50, 8)):
        flat = workloads.flat_forest(cfg, 0, ntrees)
        models = workloads.jittered_models(cfg, npsets)
        par = workloads.par_dict(models)
        par["type_space"] = np.asarray(models[0].type_space, dtype=float)
        out.append((name, flat, par, workloads.CONFIGS[cfg]["T"]))
    # the reference's own likelihood tests (test/runtests.jl:20-29 and :31-38), same parameters, 50 trees each
    for name, model, T, seed in (("rt_fully_observed", gcdyn.ConstantBranchingProcess(2.5, 1.1, 1.1, 1, 1, [1, 2, 3]), 3.0, 20),
                                 ("rt_no_extinction", gcdyn.ConstantBranchingProcess(2.5, 1.1, 0, 1, 0, [1, 2]), 3.0, 31)):
        trees = gcdyn.rand_tree(model, T, 1, 50, rng=np.random.default_rng(seed))
        flat = gcdyn.flatten(trees, model.type_space, T)
        par = workloads.par_dict([model])
        par["type_space"] = np.asarray(model.type_space, dtype=float)
        out.append((name, flat, par, T))
    return out


if __name__ == "__main__":
    for name, flat, par, T in cases():
        p = write_case(name, flat, par, T)
        print(p, "trees", flat.n_trees, "nodes", flat.n_nodes, "psets", par["lam"].shape[0])
