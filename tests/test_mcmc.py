"""Random-walk Metropolis driver (BASELINE config 4): host logic on CPU with a stand-in likelihood, and on
the GPU the same seeded chain must come out whether the likelihood is evaluated by the library or by the oracle."""
import numpy as np
import pytest

import gcdyn
from gcdyn.likelihood import pack_params
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


def test_counter_rng_is_a_sane_standard_normal_stream():
    from gcdyn.mcmc import CounterRNG
    r = CounterRNG(11)
    z = np.concatenate([r.normals(it, 32, 6).ravel() for it in range(1, 400)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    u = np.concatenate([np.exp(r.log_uniform(it, 32, 6)) for it in range(1, 400)])
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    assert np.array_equal(r.normals(7, 5, 6), CounterRNG(11).normals(7, 5, 6))          # stateless
    assert not np.array_equal(r.normals(7, 5, 6), CounterRNG(12).normals(7, 5, 6))
    # proposals of a chain do not depend on how many chains run beside it (multi-GPU ranks draw the same numbers)
    assert np.array_equal(r.normals(7, 5, 6), r.normals(7, 9, 6)[:5])


@pytest.mark.gpu
def test_device_resident_chain_equals_host_driven_and_oracle_chains(ctx):
    """`gcdyn_mh_*` (proposal, likelihood, accept/reject all on the device) against the host-driven sampler replaying
    the same counter-based stream — once with the GPU likelihood (16 chains: GEMM path), once with the oracle."""
    import oracle
    from gcdyn.mcmc import device_metropolis
    rng = np.random.default_rng(33)
    tpl = _template()
    T = 3.0
    trees = gcdyn.rand_tree(tpl, T, tpl.type_space[2], 8, rng=rng)
    forest = gcdyn.Forest(trees, tpl.type_space, T, ctx=ctx)
    th = theta_of(tpl)
    lower, upper = th - 0.6, th + 0.6
    lower[1], upper[1] = th[1] - 0.05, th[1] + 0.05          # a tight side of the box: some proposals fall outside
    for C, iters, use_oracle in ((16, 40, False), (4, 10, True)):
        th0 = np.clip(th + 0.03 * rng.standard_normal((C, 6)), lower, upper)
        kw = dict(step=0.04, seed=21, lower=lower, upper=upper)
        dev = device_metropolis(forest, tpl, T, th0, iters, **kw)
        fn = (lambda ms: [oracle.loglikelihood_shared(m, trees, T) for m in ms]) if use_oracle else None
        host = random_walk_metropolis(None if use_oracle else forest, tpl, T, th0, iters, rng="counter", loglik_fn=fn, **kw)
        assert dev.chain.shape == host.chain.shape == (iters + 1, C, 6)
        assert np.array_equal(dev.accepted, host.accepted)
        assert 0 < dev.accepted.sum() < C * iters
        assert np.max(np.abs(dev.chain - host.chain)) <= 1e-12
        assert np.max(np.abs(dev.loglik - host.loglik) / np.abs(host.loglik)) <= 1e-10
        assert np.all(dev.chain[..., 1] >= lower[1]) and np.all(dev.chain[..., 1] <= upper[1])


@pytest.mark.gpu
def test_batched_finite_difference_gradient_matches_high_precision_oracle_gradient(ctx):
    """13 parameter sets in one batch → log-likelihood + gradient; checked against Richardson-extrapolated central
    differences of the oracle (scipy DOP853 at 1e-13), parameter by parameter."""
    import oracle
    from gcdyn.mcmc import loglikelihood_gradient
    rng = np.random.default_rng(41)
    tpl = _template()
    T = 3.0
    trees = gcdyn.rand_tree(tpl, T, tpl.type_space[2], 6, rng=rng)
    forest = gcdyn.Forest(trees, tpl.type_space, T, ctx=ctx)
    th = theta_of(tpl)
    ll, grad = loglikelihood_gradient(forest, tpl, T, th)
    f = lambda t: oracle.loglikelihood_shared(model_of(t, tpl), trees, T)
    assert abs(ll - f(th)) <= 1e-10 * abs(ll)
    for k in range(6):
        e = np.zeros(6); e[k] = 1.0
        d = lambda h: (f(th + h * e) - f(th - h * e)) / (2 * h)
        want = (4.0 * d(1e-3) - d(2e-3)) / 3.0                    # O(h^4)
        assert abs(grad[k] - want) <= 2e-6 * max(1.0, abs(want)), (k, grad[k], want)


@pytest.mark.gpu
def test_device_chain_resumes_from_a_checkpoint(ctx):
    """30 iterations in one go == 12 iterations, checkpoint, new sampler, restore, 18 more (same counter-based stream)."""
    from gcdyn.mcmc import device_metropolis
    rng = np.random.default_rng(35)
    tpl = _template()
    T = 3.0
    trees = gcdyn.rand_tree(tpl, T, tpl.type_space[2], 8, rng=rng)
    forest = gcdyn.Forest(trees, tpl.type_space, T, ctx=ctx)
    th = theta_of(tpl)
    kw = dict(step=0.04, seed=23, lower=th - 0.6, upper=th + 0.6)
    th0 = th + 0.03 * rng.standard_normal((8, 6))
    whole = device_metropolis(forest, tpl, T, th0, 30, **kw)
    first = device_metropolis(forest, tpl, T, th0, 12, **kw)
    rest = device_metropolis(forest, tpl, T, th0, 18, checkpoint=first.checkpoint, **kw)
    assert first.checkpoint[3] == 12 and rest.checkpoint[3] == 30
    assert np.array_equal(first.chain, whole.chain[:13])
    assert np.array_equal(rest.chain[1:], whole.chain[13:]) and np.array_equal(rest.loglik[1:], whole.loglik[13:])
    assert np.array_equal(rest.accepted, whole.accepted)


def test_counter_rng_mirror_is_bit_identical_to_the_library():
    """`CounterRNG.uniform` (NumPy) against `gcdyn_mh_uniform` (the kernels' own hash, compiled for the host): same bits."""
    from gcdyn.likelihood import _lib
    from gcdyn.mcmc import CounterRNG
    lib = _lib()
    rng = np.random.default_rng(0)
    for seed in (0, 1, 21, 2**63 + 12345, 2**64 - 1):
        r = CounterRNG(seed)
        its = rng.integers(0, 2**40, 50)
        chains = rng.integers(0, 70000, 50)
        slots = rng.integers(0, 13, 50)
        for it, c, s in zip(its, chains, slots):
            want = lib.gcdyn_mh_uniform(int(seed) & (2**64 - 1), int(it), int(c), int(s))
            got = float(r.uniform(int(it), np.array([c]), np.array([s]))[0])
            assert got == want and 0.0 < got < 1.0


@pytest.mark.gpu
def test_exact_gradient_by_forward_sensitivities(ctx):
    """`gcdyn_loglik_grad_batch`: log-likelihood and its gradient with respect to the autodiff-able fields
    of the model, from the sensitivity trajectories.  Checked against Richardson-extrapolated central differences of the
    ORACLE (O(h⁴) truncation; its own DOP853 noise at 1e-13 of |LL| limits the reference to ≈1e-8 relative), against the
    finite differences of the GPU path, and the value against the plain evaluation."""
    import oracle
    from gcdyn.mcmc import loglikelihood_gradient, loglikelihood_gradient_exact
    from gcdyn.likelihood import FLAG_TOTALS
    rng = np.random.default_rng(41)
    tpl = _template()
    T = 3.0
    trees = gcdyn.rand_tree(tpl, T, tpl.type_space[2], 10, rng=rng)
    forest = gcdyn.Forest(trees, tpl.type_space, T, ctx=ctx)
    th = theta_of(tpl)
    pts = np.stack([th, th + 0.1 * rng.standard_normal(6), th - 0.1 * rng.standard_normal(6)])
    ll, grad = loglikelihood_gradient_exact(forest, tpl, T, pts)
    assert ll.shape == (3,) and grad.shape == (3, 6) and np.all(np.isfinite(grad))
    f = lambda t: oracle.loglikelihood_shared(model_of(t, tpl), trees, T)
    for i, t0 in enumerate(pts):
        assert abs(ll[i] - f(t0)) <= 1e-10 * abs(ll[i])
        for k in range(6):
            e = np.zeros(6); e[k] = 1.0
            d = lambda h: (f(t0 + h * e) - f(t0 - h * e)) / (2 * h)
            want = (4.0 * d(1e-3) - d(2e-3)) / 3.0                    # O(h^4)
            assert abs(grad[i, k] - want) <= 2e-7 * max(1.0, abs(want)), (i, k, grad[i, k], want)
    # consistency with the batched central differences of the GPU path (≈1e-7) and with the constant-rate directions
    _, fd = loglikelihood_gradient(forest, tpl, T, th)
    assert np.max(np.abs(fd - grad[0]) / np.maximum(1.0, np.abs(grad[0]))) <= 2e-6
    cm = gcdyn.ConstantBranchingProcess(1.7, 0.9, 1.1, 0.6, 0.3, tpl.type_space)
    ctrees = gcdyn.rand_tree(cm, T, tpl.type_space[1], 8, rng=rng)
    cf = gcdyn.Forest(ctrees, tpl.type_space, T, ctx=ctx)
    from gcdyn.likelihood import loglikelihood_and_gradient
    llc, gc = loglikelihood_and_gradient(cf, pack_params([cm] * 2), T)
    assert gc.shape == (2, 3)
    fc = lambda lam, mu, de: oracle.loglikelihood_shared(
        gcdyn.ConstantBranchingProcess(lam, mu, de, cm.Γ, cm.ρ, cm.σ, cm.type_space), ctrees, T)
    base = (1.7, 0.9, 1.0)
    assert abs(llc[0] - fc(*base)) <= 1e-10 * abs(llc[0])
    for k in range(3):
        def d(h):
            a = list(base); b = list(base); a[k] += h; b[k] -= h
            return (fc(*a) - fc(*b)) / (2 * h)
        want = (4.0 * d(1e-3) - d(2e-3)) / 3.0
        assert abs(gc[0, k] - want) <= 2e-7 * max(1.0, abs(want)), (k, gc[0, k], want)
