"""Random-walk Metropolis driver (BASELINE config 4): host logic on CPU with a stand-in likelihood, and on
the GPU the same seeded chain must come out whether the likelihood is evaluated by the library or by the oracle."""
import numpy as np
import pytest

import gcdyn
from gcdyn.mcmc import model_of, random_walk_metropolis, theta_of


def _template():
    ts = np.linspace(-1.5, 1.5, 6)
    return gcdyn.SigmoidalBranchingProcess(1.2, 0.1, 2.2, 0.4, 0.9, 0.8, 0.7, 0.25, ts)


def test_theta_round_trip_and_box_prior():
    tpl = _template()
    th = theta_of(tpl)
    m = model_of(th, tpl)
    assert m.λ_xscale == pytest.approx(tpl.λ_xscale) and m.δ == pytest.approx(tpl.δ) and m.Γ is tpl.Γ
    # stand-in target: independent standard normals on θ − θ0; box prior cuts the first coordinate at θ0 ± 0.1
    target = lambda models: np.array([-0.5 * np.sum((theta_of(mm) - th) ** 2) / 0.04 for mm in models])
    lower, upper = th - np.array([0.1, 9, 9, 9, 9, 9]), th + np.array([0.1, 9, 9, 9, 9, 9])
    res = random_walk_metropolis(None, tpl, 3.0, np.tile(th, (16, 1)), 400, step=0.15, lower=lower, upper=upper,
                                 seed=5, loglik_fn=target)
    assert res.chain.shape == (401, 16, 6) and np.all(res.chain[..., 0] >= lower[0]) and np.all(res.chain[..., 0] <= upper[0])
    assert 0.05 < res.acceptance_rate.mean() < 0.9
    assert np.abs(res.chain[100:, :, 1].mean() - th[1]) < 0.05          # unconstrained coordinate centred on θ0
    again = random_walk_metropolis(None, tpl, 3.0, np.tile(th, (16, 1)), 400, step=0.15, lower=lower, upper=upper,
                                   seed=5, loglik_fn=target)
    assert np.array_equal(res.chain, again.chain)


@pytest.mark.gpu
def test_gpu_chain_equals_oracle_chain(ctx):
    import oracle
    rng = np.random.default_rng(31)
    tpl = _template()
    T = 3.0
    trees = gcdyn.rand_tree(tpl, T, tpl.type_space[2], 8, rng=rng)
    forest = gcdyn.Forest(trees, tpl.type_space, T, ctx=ctx)
    th0 = theta_of(tpl) + 0.05 * rng.standard_normal((4, 6))
    kw = dict(step=0.05, seed=9)
    gpu = random_walk_metropolis(forest, tpl, T, th0, 12, **kw)
    ora = random_walk_metropolis(None, tpl, T, th0, 12, loglik_fn=lambda ms: [oracle.loglikelihood_shared(m, trees, T) for m in ms], **kw)
    assert np.array_equal(gpu.accepted, ora.accepted)
    assert np.array_equal(gpu.chain, ora.chain)
    assert np.max(np.abs(gpu.loglik - ora.loglik) / np.abs(ora.loglik)) <= 1e-10
