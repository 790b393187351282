"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path, called through the C ABI, against
the committed golden fixtures and against the oracle on seeded inputs.

Tolerances: the north star asks for ≤ 1e-8 relative on the summed log-likelihood; we hold the sum to
1e-10 and every single tree to 1e-9 relative (the golden values themselves carry ≈1e-13 from the
oracle's DOP853 tolerance).  Integer outputs (event counts, statuses) are compared exactly.
"""
import math
import warnings

import numpy as np
import pytest

import gcdyn
from gcdyn.likelihood import evaluate, pack_params
from helpers import load_golden, rel

pytestmark = pytest.mark.gpu

SUM_RTOL = 1e-10
TREE_RTOL = 1e-9


def _check(forest, params, T, method, want, **kw):
    out, tp, st = evaluate(forest, params, T, method, per_tree=True, **kw)
    want = np.asarray(want)
    P = want.shape[0]
    assert rel(tp[:P], want) <= TREE_RTOL, (method, rel(tp[:P], want))
    assert rel(out[:P], want.sum(axis=1)) <= SUM_RTOL
    # out_pset must be the sum of the rows it was reduced from
    assert rel(out, tp.sum(axis=1)) <= 1e-13
    assert not st.any()
    return out, tp


@pytest.mark.parametrize("name,keys", [
    ("g1_config1", ("ode_branch", "ode_shared", "ode_closed")),
    ("g2_config2", ("ode_branch", "ode_shared")),
    ("g3_fully_observed", ("ode_branch", "naive")),       # runtests.jl:27 tightened: ρ=σ=1 ⇒ ode == naive
    ("g4_no_extinction", ("ode_branch", "ode_closed", "stadler")),   # runtests.jl:37 tightened: γ=0 ⇒ ode == stadler
    ("g5_batched_discrete", ("ode_branch", "ode_shared")),
])
def test_ode_against_golden(ctx, name, keys):
    flat, params, T, exp, _ = load_golden(name)
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    for k in keys:
        _check(forest, params, T, "ode", exp[k])


@pytest.mark.parametrize("name", ["g1_config1", "g2_config2", "g3_fully_observed", "g4_no_extinction",
                                  "g5_batched_discrete"])
def test_alternates_against_golden(ctx, name):
    flat, params, T, exp, _ = load_golden(name)
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    for method, key in (("naive", "naive"), ("stadler", "stadler"), ("stadler_unconditioned", "stadler_uncond")):
        if key in exp:
            _check(forest, params, T, method, exp[key])


def test_event_counts_bit_exact(ctx):
    for name in ("g1_config1", "g2_config2", "g5_batched_discrete"):
        flat, params, T, exp, _ = load_golden(name)
        forest = gcdyn.Forest(flat, None, T, ctx=ctx)
        assert np.array_equal(forest.event_counts(), exp["event_counts"])
        info = forest.info()
        assert info["n_nodes"] == flat.n_nodes and info["n_trees"] == flat.n_trees


@pytest.mark.parametrize("order", [8, 12, 16, 20, 24])
def test_every_taylor_order_converges(ctx, order):
    flat, params, T, exp, _ = load_golden("g2_config2")
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    _check(forest, params, T, "ode", exp["ode_branch"], order=order)
    d = ctx.last_diag(params.n_psets)
    assert d["order"] == order and d["steps"] >= 2 and np.all(d["err_indicator"] <= 1e-11)


def test_steps_refinement_and_explicit_steps(ctx):
    flat, params, T, exp, _ = load_golden("g2_config2")
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    # far too few steps, no refinement allowed: the pset is reported as a solver failure (-Inf), not a wrong number
    out, tp, st = evaluate(forest, params, T, "ode", per_tree=True, steps=2, order=8)
    assert np.all(np.isneginf(out)) and np.all(st & 2)
    # explicit, sufficient steps reproduce the converged value
    _check(forest, params, T, "ode", exp["ode_branch"], steps=64, order=16)


def test_adaptive_steps_per_pset_and_global_table_path(ctx):
    """Psets with very different rates in one batch: each adapts its own step count inside traj_kernel.
    A table too large for shared memory (explicit steps=512) takes the global-memory lookup path."""
    import oracle
    rng = np.random.default_rng(15)
    ts = np.linspace(-2, 2, 8)
    base = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    trees = gcdyn.rand_tree(base, 3.0, ts[3], 10, rng=rng)
    models = [base,
              gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 12.0, 3.0, 6.0, 5.0, 0.5, 0.1, ts),     # fast rates
              gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 0.3, 0.05, 0.1, 0.1, 0.5, 0.1, ts),     # slow rates
              gcdyn.SigmoidalBranchingProcess(3.0, 1.0, 6.0, 0.5, 1.0, 2.0, 1.0, 0.02, ts)]     # ρ=1, σ≈0: hardest cells
    forest = gcdyn.Forest(trees, ts, 3.0, ctx=ctx)
    want = np.array([oracle.loglikelihood_shared(m, trees, 3.0) for m in models])
    out, _, st = evaluate(forest, pack_params(models), 3.0)
    assert not st.any() and rel(out, want) <= SUM_RTOL
    d = ctx.last_diag(len(models))
    assert np.all(d["err_indicator"] <= 1e-11)
    out512, _, st = evaluate(forest, pack_params(models), 3.0, steps=512)
    assert not st.any() and rel(out512, want) <= SUM_RTOL
    assert ctx.last_diag()["steps"] == 512


def _tile(params, reps):
    """The same parameter sets repeated `reps` times (batches of ≥ 8 psets take the Chebyshev/GEMM path)."""
    from gcdyn.likelihood import PackedParams
    rep = lambda a, per: None if a is None else np.ascontiguousarray(np.tile(a.reshape(params.n_psets, per), (reps, 1))).reshape(-1)
    K = params.n_types
    lam_per = K if params.response == 0 else 1
    return PackedParams(n_psets=params.n_psets * reps, n_types=K, response=params.response,
                        gamma_pset_stride=params.gamma_pset_stride, lam=rep(params.lam.reshape(-1), lam_per),
                        xscale=None, xshift=None, yscale=None, yshift=None, type_space=params.type_space,
                        mu=rep(params.mu, 1), delta=rep(params.delta, 1), rho=rep(params.rho, 1), sigma=rep(params.sigma, 1),
                        Gamma=params.Gamma if params.gamma_pset_stride == 0 else rep(params.Gamma, K * K))


@pytest.mark.parametrize("name,key", [("g1_config1", "ode_closed"), ("g2_config2", "ode_branch"),
                                      ("g3_fully_observed", "naive"), ("g4_no_extinction", "stadler"),
                                      ("g5_batched_discrete", "ode_shared")])
def test_chebyshev_gemm_path_against_golden_and_sweep(ctx, name, key):
    """P ≥ 8 takes the DMMA GEMM over per-tree Chebyshev moments; it must agree with the golden values and,
    to rounding, with the item sweep (flags = SWEEP_ONLY) on the same inputs."""
    flat, params, T, exp, _ = load_golden(name)
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    reps = -(-8 // params.n_psets)
    big = _tile(params, reps)
    out_g, tp_g, st_g = evaluate(forest, big, T, per_tree=True)
    cheb = ctx.last_cheb()
    assert cheb["interp_degree"] >= 8 and 4 <= cheb["effective_degree"] <= cheb["interp_degree"] and cheb["n_fallback"] == 0
    out_s, tp_s, st_s = evaluate(forest, big, T, per_tree=True, flags=4)
    assert ctx.last_cheb()["interp_degree"] == 0
    assert not st_g.any() and not st_s.any()
    want = np.tile(exp[key], (reps, 1)) if exp[key].shape[0] == params.n_psets else None
    if want is None:
        want = np.tile(exp[key], (1, 1))
    n = want.shape[0]
    assert rel(tp_g[:n], want) <= TREE_RTOL and rel(out_g[:n], want.sum(axis=1)) <= SUM_RTOL
    assert rel(tp_g, tp_s) <= 1e-10 and rel(out_g, out_s) <= 1e-12
    assert rel(out_g, tp_g.sum(axis=1)) <= 1e-13


def test_config5_like_k50_batched(ctx):
    """BASELINE config 5 in miniature: K = 50 types (two types per lane in traj_kernel, K² = 2500 type-change
    pairs folded into the per-tree constant), 24 jittered parameter sets, GEMM path vs oracle and vs sweep."""
    import oracle
    import workloads
    flat = workloads.flat_forest("C5s", 0, 40)
    models = workloads.jittered_models("C5s", 24)
    T = workloads.CONFIGS["C5s"]["T"]
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    params = pack_params(models)
    out, tp, st = evaluate(forest, params, T, per_tree=True)
    cheb = ctx.last_cheb()
    assert cheb["interp_degree"] > 0 and cheb["n_fallback"] == 0 and not st.any()
    par = workloads.par_dict(models)
    for p in (0, 11, 23):
        want = oracle.loglik_flat_shared(flat, par, p, T)
        assert rel(tp[p], want) <= TREE_RTOL and abs(out[p] - want.sum()) <= SUM_RTOL * abs(want.sum())
    out_s, tp_s, _ = evaluate(forest, params, T, per_tree=True, flags=4)
    assert rel(tp, tp_s) <= 1e-10 and rel(out, out_s) <= 1e-12


def test_gemm_path_falls_back_per_pset(ctx):
    """Psets the GEMM cannot take — a non-finite node-term log the forest needs (σ = 0 with sampled deaths),
    or a trajectory too sharp for degree 128 — are redone by the item sweep inside the same call."""
    import oracle
    rng = np.random.default_rng(16)
    ts = np.linspace(-2, 2, 6)
    mk = lambda ys, mu, sig: gcdyn.SigmoidalBranchingProcess(1.5, 0.0, ys, 0.5, mu, 1.0, 0.5, sig, ts)
    base = mk(2.5, 1.0, 0.1)
    trees = gcdyn.rand_tree(base, 3.0, ts[2], 12, rng=rng)
    models = [mk(2.5 * f, 1.0, 0.1) for f in np.linspace(0.8, 1.2, 8)]
    models[3] = mk(2.5, 1.0, 0.0)          # log σ = -Inf, forest has sampled deaths
    models[5] = mk(60.0, 30.0, 0.1)        # T·rate ≈ 300: needs more than degree 128
    forest = gcdyn.Forest(trees, ts, 3.0, ctx=ctx)
    out, tp, st = evaluate(forest, pack_params(models), 3.0, per_tree=True)
    assert ctx.last_cheb()["n_fallback"] >= 1
    flat = forest.flat
    has_death = flat.event_counts[:, 2] > 0
    assert np.all(np.isneginf(tp[3][has_death])) and np.all(np.isfinite(tp[3][~has_death]))
    for p in (0, 5, 7):
        want = np.array([oracle.loglikelihood_shared(models[p], t, 3.0) for t in trees])
        assert rel(tp[p], want) <= TREE_RTOL, p
    out_s, tp_s, _ = evaluate(forest, pack_params(models), 3.0, per_tree=True, flags=4)
    finite = np.isfinite(tp_s)
    assert np.array_equal(finite, np.isfinite(tp)) and rel(tp[finite], tp_s[finite]) <= 1e-10


def test_public_api_matches_oracle_on_simulated_trees(ctx):
    """The reference-style call `loglikelihood(model, trees, present_time)` on freshly simulated trees."""
    import oracle
    rng = np.random.default_rng(11)
    ts = np.linspace(-1.5, 1.5, 6)
    model = gcdyn.SigmoidalBranchingProcess(1.2, 0.1, 2.2, 0.4, 0.9, 0.8, 0.7, 0.25, ts)
    trees = gcdyn.rand_tree(model, 3.0, ts[2], 15, rng=rng)
    want = oracle.loglikelihood_shared(model, trees, 3.0)
    got = gcdyn.loglikelihood(model, trees, 3.0)
    assert abs(got - want) <= SUM_RTOL * abs(want)
    # single tree method, and reltol/abstol/maxiters keywords are accepted
    one = gcdyn.loglikelihood(model, trees[0], 3.0, reltol=1e-6, abstol=1e-6, maxiters=1e4)
    assert abs(one - oracle.loglikelihood_shared(model, trees[0], 3.0)) <= TREE_RTOL * abs(one)
    # maxiters caps the adaptive cells of a trajectory; exceeding it is the reference's `@warn` + -Inf (mbp.jl:154-158),
    # on the item-sweep path (one model) and on the contraction path (a batch) alike
    with pytest.warns(UserWarning, match="Density will evaluate to zero"):
        assert gcdyn.loglikelihood(model, trees, 3.0, maxiters=2) == -np.inf
    with pytest.warns(UserWarning, match="Density will evaluate to zero"):
        assert np.all(np.isneginf(gcdyn.loglikelihood([model] * 9, trees, 3.0, maxiters=2)))
    assert abs(gcdyn.loglikelihood(model, trees, 3.0, maxiters=10 ** 5) - want) <= SUM_RTOL * abs(want)
    # batched entry: a list of models gives one value per model
    models = [model, gcdyn.SigmoidalBranchingProcess(1.0, 0.0, 2.0, 0.5, 1.0, 1.0, 0.7, 0.25, ts)]
    got2 = gcdyn.loglikelihood(models, trees, 3.0)
    want2 = [oracle.loglikelihood_shared(m, trees, 3.0) for m in models]
    assert rel(got2, want2) <= SUM_RTOL
    # alternates through the public API
    t0 = trees[0]
    assert abs(gcdyn.naive_loglikelihood(model, t0) - oracle.naive_loglikelihood(model, t0)) <= 1e-10 * abs(oracle.naive_loglikelihood(model, t0))
    assert abs(gcdyn.stadler_appx_loglikelihood(model, t0, 3.0) - oracle.stadler_appx_loglikelihood(model, t0, 3.0)) <= 1e-10 * abs(oracle.stadler_appx_loglikelihood(model, t0, 3.0))
    assert abs(gcdyn.stadler_appx_unconditioned_loglikelihood(model, t0, 3.0) - oracle.stadler_appx_unconditioned_loglikelihood(model, t0, 3.0)) <= 1e-10 * abs(oracle.stadler_appx_unconditioned_loglikelihood(model, t0, 3.0))


def test_sigmoid_evaluated_on_device(ctx):
    import oracle
    rng = np.random.default_rng(12)
    ts = np.linspace(-2, 2, 10)
    models = [gcdyn.SigmoidalBranchingProcess(1.5 * a, 0.1 * b, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
              for a, b in ((1.0, 0.0), (0.8, 1.0), (1.3, -1.0))]
    trees = gcdyn.rand_tree(models[0], 3.0, ts[4], 8, rng=rng)
    forest = gcdyn.Forest(trees, ts, 3.0, ctx=ctx)
    out_t, _, _ = evaluate(forest, pack_params(models), 3.0)
    out_s, _, _ = evaluate(forest, pack_params(models, response="sigmoid"), 3.0)
    assert rel(out_s, out_t) <= 1e-13
    want = [oracle.loglikelihood_shared(m, trees, 3.0) for m in models]
    assert rel(out_s, want) <= SUM_RTOL


def test_edge_cases(ctx):
    import oracle
    N = gcdyn.TreeNode
    ts = [1.0, 2.0]
    T = 2.0
    model = gcdyn.ConstantBranchingProcess(1.5, 0.7, 0.9, 0.8, 0.4, ts)

    def tiny():   # root -> survivor: the smallest legal tree (one node record)
        r = N("root", 0, 1.0)
        gcdyn.attach(r, N("sampled_survival", T, 1.0))
        return r

    def small():  # root -> type_change(1→2) -> birth -> (death, survivor)
        r = N("root", 0, 1.0)
        tc = N("type_change", 0.5, 2.0); gcdyn.attach(r, tc)
        b = N("birth", 1.1, 2.0); gcdyn.attach(tc, b)
        gcdyn.attach(b, N("sampled_death", 1.6, 2.0))
        gcdyn.attach(b, N("sampled_survival", T, 2.0))
        return r
    for tr in (tiny(), small()):
        got = gcdyn.loglikelihood(model, tr, T)
        want = oracle.loglikelihood(model, tr, T)
        assert abs(got - want) <= 1e-11 * max(1.0, abs(want))
    # ragged forest: tiny trees mixed with a large one
    rng = np.random.default_rng(13)
    big = gcdyn.rand_tree(model, T, 1.0, rng=rng, min_leaves=20)
    forest_trees = [tiny(), big, small(), tiny()]
    assert abs(gcdyn.loglikelihood(model, forest_trees, T) - oracle.loglikelihood_shared(model, forest_trees, T)) <= 1e-10 * 50
    # self loop: `@warn` + -Inf (mbp.jl:181-185)
    loop = small()
    loop.children[0].type = 1.0
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert gcdyn.loglikelihood(model, loop, T) == -math.inf
        assert any("Self-loop" in str(x.message) for x in w)
    assert oracle.loglikelihood(model, loop, T) == -math.inf
    # σ = 0 with a sampled death in the tree: log 0 = -Inf flows through silently (mbp.jl:171)
    m0 = gcdyn.ConstantBranchingProcess(1.5, 0.7, 0.9, 0.8, 0.0, ts)
    assert gcdyn.loglikelihood(m0, small(), T) == -math.inf
    assert oracle.loglikelihood(m0, small(), T) == -math.inf
    # ... but not when no sampled death occurs (0 · log 0 must not poison the sum)
    assert math.isfinite(gcdyn.loglikelihood(m0, tiny(), T))
    # ρ = 0 with a sampled survivor
    mr = gcdyn.ConstantBranchingProcess(1.5, 0.7, 0.9, 0.0, 0.4, ts)
    assert gcdyn.loglikelihood(mr, tiny(), T) == -math.inf
    # negative birth rate on a birth branch: Julia's log throws DomainError
    mneg = gcdyn.DiscreteBranchingProcess([1.0, -0.5], 0.7, 0.9, 0.8, 0.4, ts)
    with pytest.raises(gcdyn.DomainError):
        gcdyn.loglikelihood(mneg, small(), T)
    # empty forest: the sum over no trees is 0
    forest = gcdyn.Forest([], ts, T, ctx=ctx)
    out, _, _ = evaluate(forest, pack_params(model), T)
    assert out[0] == 0.0


def test_many_types_uses_multi_type_lanes(ctx):
    """K = 40 (32 < K ≤ 64): two integrating warps, one type per thread; K = 70 (> 64): several types per lane."""
    import oracle
    rng = np.random.default_rng(14)
    K = 40
    ts = np.linspace(-2, 2, K)
    model = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    trees = gcdyn.rand_tree(model, 3.0, ts[K // 2], 6, rng=rng)
    got = gcdyn.loglikelihood(model, trees, 3.0)
    want = oracle.loglikelihood_shared(model, trees, 3.0)
    assert abs(got - want) <= SUM_RTOL * abs(want)
    K = 70
    ts = np.linspace(-2, 2, K)
    model = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    trees = gcdyn.rand_tree(model, 2.5, ts[K // 2], 4, rng=rng)
    got = gcdyn.loglikelihood(model, trees, 2.5)
    want = oracle.loglikelihood_shared(model, trees, 2.5)
    assert abs(got - want) <= SUM_RTOL * abs(want)


@pytest.mark.parametrize("K,order", [(33, 16), (48, 8), (64, 24), (65, 16), (32, 20), (96, 16), (128, 16), (128, 24)])
def test_type_count_boundaries_of_the_trajectory_kernel(ctx, K, order):
    """The kernel switches layout at K = 32 (one warp) / 33..64 (two integrating warps, register rows of width 40, 48,
    56, 64) / 65.. (types strided over lanes, δΓ in shared memory).  Each side of every boundary: batch of 9 psets
    through the GEMM path == the item sweep, one pset == the oracle, a failed pset does not stall the others.
    K = 128 = GCDYN_MAX_TYPES: the widest layout keeps two cells in flight so that δΓ (K² doubles) still fits beside the ring;
    the nodal stage does not fit there and the batch takes the exact sweep."""
    import oracle
    rng = np.random.default_rng(100 + K)
    ts = np.linspace(-2, 2, K)
    mk = lambda f: gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5 * f, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    models = [mk(f) for f in np.linspace(0.85, 1.15, 9)]
    trees = gcdyn.rand_tree(models[4], 2.5, ts[K // 2], 4, rng=rng)
    forest = gcdyn.Forest(trees, ts, 2.5, ctx=ctx)
    params = pack_params(models)
    out, tp, st = evaluate(forest, params, 2.5, per_tree=True, order=order)
    out_s, tp_s, st_s = evaluate(forest, params, 2.5, per_tree=True, order=order, flags=4)
    assert not st.any() and not st_s.any()
    assert rel(tp, tp_s) <= 1e-10
    want = oracle.loglikelihood_shared(models[2], trees, 2.5)
    assert abs(out[2] - want) <= SUM_RTOL * abs(want)
    params.mu = params.mu.copy()
    params.mu[6] = np.nan
    out_b, _, st_b = evaluate(forest, params, 2.5, per_tree=True, order=order)
    keep = np.arange(9) != 6
    # the other psets are untouched by the failure (to rounding: the nodal grid may have been re-sized between the calls)
    assert np.all(st_b[6] & 2) and np.isneginf(out_b[6]) and rel(out_b[keep], out[keep]) <= 1e-13


def test_shards_sum_to_the_whole(ctx):
    """Multi-GPU emulation: R tree-shards evaluated separately sum to the unsharded result."""
    flat, params, T, exp, _ = load_golden("g2_config2")
    whole = gcdyn.Forest(flat, None, T, ctx=ctx)
    out_w, tp_w, _ = evaluate(whole, params, T, per_tree=True)
    from gcdyn.sharding import shard_flat
    for R in (2, 3):
        parts = [gcdyn.Forest(s, None, T, ctx=ctx) for s in shard_flat(flat, R)]
        total = sum(evaluate(p, params, T)[0] for p in parts)
        assert rel(total, out_w) <= 1e-13


@pytest.mark.gpu
def test_failed_trajectory_in_a_gemm_batch_does_not_stall_the_other_warps(ctx):
    """A pset whose step controller gives up (NaN rate: every cell is rejected) must release the three consumer
    warps of its CTA (CELL_FAIL through the ring) and come back as a solver failure; its neighbours in the same
    GEMM batch are untouched."""
    flat, params, T, exp, _ = load_golden("g2_config2")
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    tiled = _tile(params, 9)
    good, _, st0 = evaluate(forest, tiled, T, per_tree=True)
    assert not st0.any()
    bad = _tile(params, 9)
    bad.mu = bad.mu.copy()
    bad.mu[4] = np.nan
    out, tp, st = evaluate(forest, bad, T, per_tree=True)
    assert np.all(st[4] & 2) and np.all(np.isneginf(tp[4])) and np.isneginf(out[4])
    keep = np.arange(9) != 4
    assert np.array_equal(out[keep], good[keep]) and not st[keep].any()
    # the same through the item sweep (no Chebyshev stage: the consumers only build the table)
    out_s, _, st_s = evaluate(forest, bad, T, per_tree=True, flags=4)
    assert np.all(st_s[4] & 2) and np.isneginf(out_s[4]) and rel(out_s[keep], good[keep]) <= 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("K", [12, 40, 70])
def test_per_pset_gamma_and_device_sigmoid_across_kernel_layouts(ctx, K):
    """Every pset with its OWN mutation matrix (K² type-change columns in the GEMM, Γ read per pset by the integrating
    and the table-building warps) and the sigmoid evaluated on the device, in each layout of the trajectory kernel."""
    import oracle
    rng = np.random.default_rng(200 + K)
    ts = np.linspace(-2, 2, K)

    def gamma(seed):
        g = np.random.default_rng(seed).uniform(0.2, 1.0, (K, K))
        np.fill_diagonal(g, 0.0)
        g /= g.sum(axis=1, keepdims=True)
        np.fill_diagonal(g, -1.0)
        return g

    models = [gcdyn.SigmoidalBranchingProcess(1.5, 0.1 * i, 2.5, 0.5, 1.0, 0.8 + 0.05 * i, gamma(i), 0.5, 0.1, ts)
              for i in range(9)]
    trees = gcdyn.rand_tree(models[0], 2.5, ts[K // 2], 5, rng=rng)
    forest = gcdyn.Forest(trees, ts, 2.5, ctx=ctx)
    p_tab, p_sig = pack_params(models), pack_params(models, response="sigmoid")
    assert p_tab.gamma_pset_stride == K * K
    out, tp, st = evaluate(forest, p_tab, 2.5, per_tree=True)
    out_sig, _, _ = evaluate(forest, p_sig, 2.5)
    out_sw, tp_sw, _ = evaluate(forest, p_tab, 2.5, per_tree=True, flags=4)
    assert not st.any()
    assert rel(out_sig, out) <= 1e-12 and rel(tp, tp_sw) <= 1e-10
    for i in (0, 8):
        want = oracle.loglikelihood_shared(models[i], trees, 2.5)
        assert abs(out[i] - want) <= SUM_RTOL * abs(want), i


def test_totals_path_equals_the_sum_of_the_per_tree_path(ctx):
    """GCDYN_FLAG_TOTALS: per-pset sums from the coefficient rows and the forest's column sums, no contraction and no
    per-tree values.  Same numbers as summing the per-tree path (to the order of summation), −Inf for a failed pset and
    for a forest with a self-loop, EINVAL when per-tree outputs are requested with it."""
    from gcdyn.likelihood import FLAG_TOTALS, GcdynError
    flat, params, T, exp, _ = load_golden("g2_config2")
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    big = _tile(params, 12)
    big.mu = big.mu.copy()
    big.mu[5] *= 1.3
    big.mu[7] = np.nan                                           # a failed trajectory
    full, tp, st = evaluate(forest, big, T, per_tree=True)
    tot, _, _ = evaluate(forest, big, T, want_status=False, flags=FLAG_TOTALS)
    keep = np.arange(12) != 7
    assert np.isneginf(tot[7]) and np.isneginf(full[7])
    assert rel(tot[keep], full[keep]) <= 1e-12
    assert abs(tot[0] - exp["ode_branch"].sum()) <= SUM_RTOL * abs(exp["ode_branch"].sum())
    with pytest.raises(GcdynError):
        evaluate(forest, big, T, per_tree=True, flags=FLAG_TOTALS)
    # psets the grid cannot take are still summed exactly (finish_kernel runs for them alone)
    ts = np.linspace(-2, 2, 6)
    rng = np.random.default_rng(16)
    mk = lambda ys, mu, sig: gcdyn.SigmoidalBranchingProcess(1.5, 0.0, ys, 0.5, mu, 1.0, 0.5, sig, ts)
    trees = gcdyn.rand_tree(mk(2.5, 1.0, 0.1), 3.0, ts[2], 12, rng=rng)
    models = [mk(2.5 * f, 1.0, 0.1) for f in np.linspace(0.8, 1.2, 8)]
    models[3] = mk(2.5, 1.0, 0.0)                                # log σ = −Inf where the forest has sampled deaths
    f2 = gcdyn.Forest(trees, ts, 3.0, ctx=ctx)
    full2, _, _ = evaluate(f2, pack_params(models), 3.0)
    tot2, _, _ = evaluate(f2, pack_params(models), 3.0, want_status=False, flags=FLAG_TOTALS)
    fin = np.isfinite(full2)
    assert ctx.last_cheb()["n_fallback"] >= 1 and not fin[3] and np.isneginf(tot2[3])
    assert np.array_equal(fin, np.isfinite(tot2)) and rel(tot2[fin], full2[fin]) <= 1e-12


def test_self_loop_followed_by_an_out_of_space_type_and_present_time_mismatch(ctx):
    """The reference returns -Inf at the self-loop (mbp.jl:181-185) and never looks at the nodes that follow it in
    post-order — which may carry a type outside the type space.  The forest must be accepted (that tree is -Inf, its
    neighbours are untouched).  Evaluating a forest at another present_time than it was flattened for is the reference's
    ArgumentError (mbp.jl:125-130), not a silent extrapolation."""
    import oracle
    T = 2.0
    ts = np.array([1.0, 2.0])
    model = gcdyn.ConstantBranchingProcess(1.5, 0.7, 0.9, 0.8, 0.4, ts)
    N = gcdyn.TreeNode
    r = N("root", 0, 1.0)
    b = N("birth", 0.4, 1.0); gcdyn.attach(r, b)
    tc1 = N("type_change", 0.8, 1.0); gcdyn.attach(b, tc1); gcdyn.attach(tc1, N("sampled_survival", T, 1.0))     # self-loop
    tc2 = N("type_change", 0.9, 9.0); gcdyn.attach(b, tc2); gcdyn.attach(tc2, N("sampled_survival", T, 9.0))     # type ∉ type_space
    good = gcdyn.rand_tree(model, T, 1.0, rng=np.random.default_rng(3))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert gcdyn.loglikelihood(model, r, T) == -math.inf
        assert any("Self-loop" in str(x.message) for x in w)
    forest = gcdyn.Forest([good, r, good], ts, T, ctx=ctx)
    for reps in (1, 9):                                     # item sweep and contraction paths
        out, tp, st = evaluate(forest, pack_params([model] * reps), T, per_tree=True)
        want = oracle.loglikelihood_shared(model, good, T)
        assert np.all(np.isneginf(tp[:, 1])) and np.all(st[:, 1] & 1) and np.all(np.isneginf(out))
        assert rel(tp[:, 0], np.full(reps, want)) <= TREE_RTOL and np.array_equal(tp[:, 0], tp[:, 2])
    with pytest.raises(gcdyn.ArgumentError):
        evaluate(forest, pack_params(model), T + 0.5)
