"""GPU parity at the sizes BASELINE.json states: config 3 (1000 trees × 256 parameter sets,
K = 20), config 4 (64-chain Metropolis on the 1000-tree forest) and config 5 scaled to ≥ 1000 trees × 64 parameter sets
at K = 50 — all through the C ABI, against the oracle on the same seeded inputs.  Tolerances as in test_gpu_parity.py:
1e-10 relative on a per-pset sum (north star: 1e-8), 1e-9 per tree."""
import numpy as np
import pytest

import gcdyn
from gcdyn.likelihood import evaluate, pack_params
from helpers import rel

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

SUM_RTOL = 1e-10
TREE_RTOL = 1e-9

_SHARED = {}          # inherited by the forked oracle workers (the flattened forest is large; parameters travel as arguments)


def _oracle_pset_sum(args):
    import oracle
    par, p, T = args
    return float(oracle.loglik_flat_shared(_SHARED["flat"], par, p, T).sum())


@pytest.fixture(scope="module")
def c3(ctx):
    import workloads
    flat = workloads.flat_forest("C3", 0, 1000)
    models = workloads.jittered_models("C3", 256)
    T = workloads.CONFIGS["C3"]["T"]
    return dict(flat=flat, models=models, T=T, forest=gcdyn.Forest(flat, None, T, ctx=ctx), par=workloads.par_dict(models))


def test_config3_at_stated_size_32_psets_against_the_oracle(ctx, c3):
    import oracle
    out, tp, st = evaluate(c3["forest"], pack_params(c3["models"]), c3["T"], per_tree=True)
    assert out.shape == (256,) and tp.shape == (256, 1000) and not st.any()
    assert ctx.last_cheb()["interp_degree"] > 0 and ctx.last_cheb()["n_fallback"] == 0      # the DMMA contraction ran
    assert rel(out, tp.sum(axis=1)) <= 1e-13
    worst_sum = worst_tree = 0.0
    for p in range(0, 256, 8):                                                                # 32 psets spread over the batch
        want = oracle.loglik_flat_shared(c3["flat"], c3["par"], p, c3["T"])
        worst_tree = max(worst_tree, rel(tp[p], want))
        worst_sum = max(worst_sum, abs(out[p] - want.sum()) / abs(want.sum()))
    assert worst_sum <= SUM_RTOL and worst_tree <= TREE_RTOL, (worst_sum, worst_tree)
    # the item sweep (exact sums from the Taylor table, no Chebyshev stage) agrees with the contraction on every pset
    out_s, _, st_s = evaluate(c3["forest"], pack_params(c3["models"]), c3["T"], flags=4)
    assert not st_s.any() and rel(out, out_s) <= 1e-12


def test_config3_results_do_not_depend_on_call_history(ctx, c3):
    """After the first (calibrating) evaluation the grid degree belongs to the forest handle: repeating a call, or
    evaluating a sub-batch, reproduces the same bits."""
    params = pack_params(c3["models"])
    a, _, _ = evaluate(c3["forest"], params, c3["T"], want_status=False)
    b, _, _ = evaluate(c3["forest"], params, c3["T"], want_status=False)
    assert np.array_equal(a, b)
    sub, _, _ = evaluate(c3["forest"], pack_params(c3["models"][:64]), c3["T"], want_status=False)
    assert np.array_equal(sub, a[:64])


def test_config4_device_chain_on_the_1000_tree_forest_equals_the_oracle_driven_chain(ctx, c3):
    """64 chains, 20 iterations on the config-3 forest: the device-resident sampler (proposal, batched likelihood,
    accept/reject on the GPU) walks exactly the chain the host loop walks when it is fed the ORACLE's likelihoods and
    the same counter-based random numbers."""
    import oracle
    import workloads
    from gcdyn.mcmc import CounterRNG, device_metropolis, model_of, random_walk_metropolis, theta_of
    base, _ = workloads.base_model("C3")
    T = c3["T"]
    rng = np.random.default_rng(4)
    th0 = theta_of(base) + 0.05 * rng.standard_normal((64, 6))
    lo, up = theta_of(base) - 1.5, theta_of(base) + 1.5
    iters = 20
    dev = device_metropolis(c3["forest"], base, T, th0, iters, step=0.02, seed=4, lower=lo, upper=up)

    # the oracle costs ≈0.4 s per parameter set at this size: 21 batches of 64 go through a pool of forked CPU workers
    import multiprocessing as mp
    import os
    _SHARED["flat"] = c3["flat"]
    with mp.get_context("fork").Pool(min(16, len(os.sched_getaffinity(0)))) as pool:
        def oracle_ll(models):
            par = workloads.par_dict(models)
            return np.array(pool.map(_oracle_pset_sum, [(par, p, T) for p in range(len(models))]))
        ora = random_walk_metropolis(None, base, T, th0, iters, step=0.02, seed=4, lower=lo, upper=up,
                                     loglik_fn=oracle_ll, rng="counter")
    assert np.array_equal(dev.accepted, ora.accepted)
    # same decisions at every step; the states agree to the last bits of the device's vs NumPy's Box–Muller (log, cospi)
    assert np.max(np.abs(dev.chain - ora.chain)) <= 1e-12
    assert rel(dev.loglik, ora.loglik) <= SUM_RTOL
    assert 0 < dev.accepted.sum() < 64 * iters


def test_config5_scaled_1000_trees_64_psets_k50(ctx):
    """Config 5 scaled down in trees and psets only: K = 50 types (two integrating warps per pset), trees of 500-2000
    nodes from the native simulator, 1000 trees × 64 parameter sets; every pset's sum and 8 psets' per-tree values
    against the oracle."""
    import oracle
    import workloads
    from gcdyn.sharding import subset_flat
    K, T = 50, 5.0
    ts = np.linspace(-2.0, 2.0, K)
    base = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    pool = gcdyn.simulate_flat(base, T, ts[K // 2 - 1], 6000, seed=5)
    sizes = np.diff(pool.tree_off)
    ok = np.nonzero((sizes >= 500) & (sizes <= 2000))[0][:1000]
    assert len(ok) == 1000
    flat = subset_flat(pool, ok)
    rng = np.random.default_rng([5, 777])
    models = []
    for _ in range(64):
        f = np.exp(rng.normal(0.0, 0.2, 5))
        models.append(gcdyn.SigmoidalBranchingProcess(base.λ_xscale * f[0], base.λ_xshift + rng.normal(0.0, 0.2),
                                                      base.λ_yscale * f[1], base.λ_yshift * f[2], base.μ * f[3],
                                                      base.δ * f[4], base.Γ, base.ρ, base.σ, ts))
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    out, tp, st = evaluate(forest, pack_params(models), T, per_tree=True)
    assert not st.any() and ctx.last_cheb()["interp_degree"] > 0
    par = workloads.par_dict(models)
    for p in range(0, 64, 8):
        want = oracle.loglik_flat_shared(flat, par, p, T)
        assert rel(tp[p], want) <= TREE_RTOL, p
        assert abs(out[p] - want.sum()) <= SUM_RTOL * abs(want.sum()), p
