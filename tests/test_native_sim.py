"""Native simulator (C++ host code of libgcdyn_b200): same law and same structural invariants as the host
mirror of `rand_tree` (mbp.jl:318-461); reproducible piecewise; its forests evaluate like any other."""
import math

import numpy as np
import pytest

import gcdyn


def test_expected_survivors_law():
    # runtests.jl:4-18: E[#sampled survivors] = exp(T(λ−μ)), reject_stubs = false, γ = 0, ρ = σ = 1
    lam, mu, T = 2.5, 1.1, 2
    m = gcdyn.ConstantBranchingProcess(lam, mu, 0, 1, 1, range(1, 4))
    f = gcdyn.simulate_flat(m, T, 1, 20000, reject_stubs=False, seed=2)
    # with reject_stubs = false a tree may be a bare root: it contributes a tree with zero node records
    assert f.n_trees == 20000
    assert f.event_counts[:, 5].mean() == pytest.approx(math.exp(T * (lam - mu)), rel=3e-2)


def test_structural_invariants_and_reproducibility():
    ts = np.linspace(-2, 2, 7)
    m = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    T = 3.0
    f = gcdyn.simulate_flat(m, T, ts[3], 300, seed=11)
    ev, x, y = f.event.astype(int), f.btype_idx, f.type_idx
    nchild = np.diff(f.child_off)
    assert set(np.unique(ev)) <= {1, 2, 4, 5}                              # births, sampled deaths, type changes, sampled survivors
    assert np.all(nchild[ev == 1] == 2) and np.all(x[ev == 1] == y[ev == 1])
    assert np.all(nchild[ev == 4] == 1) and np.all(x[ev == 4] != y[ev == 4])
    assert np.all(nchild[(ev == 2) | (ev == 5)] == 0)
    assert np.all(f.t_end[ev == 5] == T) and np.all(f.t_end[ev == 2] < T) and np.all(f.t_end >= f.t_start)
    assert np.all(f.root_type_idx == 3) and np.all(f.root_time == 0)
    for t in range(f.n_trees):
        lo, hi = f.tree_off[t], f.tree_off[t + 1]
        assert hi > lo                                                     # reject_stubs: every tree has nodes
        par = f.parent[lo:hi]
        assert np.all((par == -1) | (par > np.arange(hi - lo))) and np.sum(par == -1) == 1 and par[-1] == -1
        for i in range(lo, hi):                                            # CSR children ↔ parent, branch types ↔ parent types
            for c in f.child_idx[f.child_off[i]:f.child_off[i + 1]]:
                assert f.parent[lo + c] == i - lo and f.btype_idx[lo + c] == f.type_idx[i]
                assert f.t_start[lo + c] == f.t_end[i]
    # the flattener accepts it as a valid forest of the C ABI (same dtypes as gcdyn.flatten)
    ref = gcdyn.flatten(gcdyn.rand_tree(m, T, ts[3], 2, rng=np.random.default_rng(0)), ts, T)
    for k in ("tree_off", "event", "type_idx", "btype_idx", "t_start", "t_end", "parent", "child_off", "child_idx",
              "root_type_idx", "root_time", "live", "self_loop"):
        assert getattr(f, k).dtype == getattr(ref, k).dtype, k
    # same seed ⇒ same forest; forests are reproducible piecewise
    g = gcdyn.simulate_flat(m, T, ts[3], 300, seed=11)
    a = gcdyn.simulate_flat(m, T, ts[3], 100, seed=11)
    b = gcdyn.simulate_flat(m, T, ts[3], 200, seed=11, first_tree_index=100)
    assert np.array_equal(f.t_end, g.t_end) and np.array_equal(f.event, g.event)
    assert np.array_equal(f.t_end, np.concatenate([a.t_end, b.t_end]))
    assert np.array_equal(f.type_idx, np.concatenate([a.type_idx, b.type_idx]))
    with pytest.raises(gcdyn.GcdynError):                                  # max_rejections (mbp.jl:389-391)
        gcdyn.simulate_flat(gcdyn.ConstantBranchingProcess(0.01, 50.0, 0, 1, 0, [1, 2]), 5.0, 1, 3, max_rejections=3)


def test_same_law_as_the_python_mirror():
    """Two independent restatements of the same simulator: tree-size and event-mix statistics must agree."""
    m = gcdyn.ConstantBranchingProcess(2.0, 1.0, 1.0, 0.5, 0.2, [1, 2, 3])
    T, n = 3.0, 3000
    f = gcdyn.simulate_flat(m, T, 1, n, seed=5)
    trees = gcdyn.rand_tree(m, T, 1, n, rng=np.random.default_rng(5))
    p = gcdyn.flatten(trees, m.type_space, T)
    size_n, size_p = np.diff(f.tree_off), np.diff(p.tree_off)
    se = math.sqrt(size_n.var() / n + size_p.var() / n)
    assert abs(size_n.mean() - size_p.mean()) <= 4 * se
    frac_n = f.event_counts.sum(0) / f.n_nodes
    frac_p = p.event_counts.sum(0) / p.n_nodes
    assert np.max(np.abs(frac_n - frac_p)) <= 0.01
    # type-change destinations follow Γ[i,:]/(−Γ[i,i]) with zero self weight in both
    tc_n = np.bincount(f.type_idx[f.event == 4], minlength=3) / max(1, np.sum(f.event == 4))
    tc_p = np.bincount(p.type_idx[p.event == 4], minlength=3) / max(1, np.sum(p.event == 4))
    assert np.max(np.abs(tc_n - tc_p)) <= 0.02


@pytest.mark.gpu
def test_native_forest_evaluates_like_the_oracle(ctx):
    import oracle
    import workloads
    from gcdyn.likelihood import evaluate, pack_params
    models = workloads.jittered_models("C3", 8)
    base, init = workloads.base_model("C3")
    T = workloads.CONFIGS["C3"]["T"]
    flat = gcdyn.simulate_flat(base, T, init, 150, seed=21)
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    out, tp, st = evaluate(forest, pack_params(models), T, per_tree=True)
    par = workloads.par_dict(models)
    for p in (0, 7):
        want = oracle.loglik_flat_shared(flat, par, p, T)
        assert np.max(np.abs(tp[p] - want) / np.abs(want)) <= 1e-9
        assert abs(out[p] - want.sum()) <= 1e-10 * abs(want.sum())
