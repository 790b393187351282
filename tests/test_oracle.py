"""Pins the ORACLE (oracle/oracle.py, oracle/refcpu.c).  The reference has no golden vectors and cannot
be run here, so the oracle is pinned to the reference's own test relations (runtests.jl:27,28,37)
tightened from rtol 1e-1 to ≤1e-10, to closed forms built from extra_likelihoods.jl in 50-digit mpmath,
to an independent arbitrary-precision Taylor integration, and to the committed fixtures."""
import math

import mpmath
import numpy as np
import pytest

import gcdyn
import oracle
from helpers import load_golden, rel


def _trees(model, T, init, n, seed):
    return gcdyn.rand_tree(model, T, init, n, rng=np.random.default_rng(seed))


def test_three_integrations_agree():
    ts = np.linspace(-1, 1, 4)
    m = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.0, 0.5, 1.0, 1.0, 0.5, 0.3, ts)
    for t in _trees(m, 2.5, ts[1], 4, 21):
        a = oracle.loglikelihood(m, t, 2.5)
        b = oracle.loglikelihood_shared(m, t, 2.5)
        assert abs(a - b) <= 1e-11 * max(1.0, abs(a))


def test_equal_rates_closed_form_in_mpmath():
    # SURVEY §3.4 Corollary A, from extra_likelihoods.jl:52-61,82-97
    mpmath.mp.dps = 50
    for (rho, sig) in ((1, 1), (0.5, 0.3), (1, 0)):
        m = gcdyn.ConstantBranchingProcess(1.8, 1.0, 1.0, rho, sig, [1, 2])
        for t in _trees(m, 3.0, 1, 3, 22):
            want = float(oracle.loglikelihood_equal_rates(m, t, 3.0, mp=mpmath.mp))
            assert abs(oracle.loglikelihood(m, t, 3.0) - want) <= 1e-11 * max(1.0, abs(want))
            assert abs(oracle.loglikelihood_equal_rates(m, t, 3.0) - want) <= 1e-12 * max(1.0, abs(want))


def test_trajectory_against_arbitrary_precision_taylor():
    """p(T) and ∫p from mpmath's Taylor-series ODE solver (30 digits) vs the scipy trajectory."""
    mpmath.mp.dps = 30
    ts = np.array([-1.0, 0.5, 2.0])
    m = oracle.canonical(gcdyn.SigmoidalBranchingProcess(1.5, 0.2, 2.0, 0.4, 1.1, 0.9, 0.6, 0.2, ts))
    K = m.K
    lam = [mpmath.mpf(float(v)) for v in m.lam]
    G = [[mpmath.mpf(float(m.delta * m.Gamma[i, j])) for j in range(K)] for i in range(K)]
    mu, sig, rho = mpmath.mpf(m.mu), mpmath.mpf(m.sigma), mpmath.mpf(m.rho)

    def f(t, u):
        p = u[:K]
        dp = [-(lam[i] + mu) * p[i] + mu * (1 - sig) + lam[i] * p[i] ** 2 + sum(G[i][j] * p[j] for j in range(K))
              for i in range(K)]
        return dp + list(p)
    sol = mpmath.odefun(f, 0, [1 - rho] * K + [0] * K)
    want = np.array([float(v) for v in sol(2.0)])
    got = oracle.trajectory(m, 2.0).sol(2.0)
    assert rel(got, want) <= 1e-12


def test_reference_test_relations_tightened():
    # runtests.jl:20-29: ρ = σ = 1 ⇒ naive == loglikelihood (exactly: p ≡ 0), and naive ≈ stadler only at 1e-1
    m = gcdyn.ConstantBranchingProcess(2.5, 1.1, 1.1, 1, 1, range(1, 4))
    trees = _trees(m, 3, 1, 6, 23)
    naive = sum(oracle.naive_loglikelihood(m, t) for t in trees)
    ode = oracle.loglikelihood(m, trees, 3)
    appx = sum(oracle.stadler_appx_loglikelihood(m, t, 3) for t in trees)
    assert abs(naive - ode) <= 1e-10 * abs(naive)
    assert abs(naive - appx) <= 1e-1 * abs(naive)
    # runtests.jl:31-38: γ = 0, ρ = 1, σ = 0 ⇒ stadler == loglikelihood
    m = gcdyn.ConstantBranchingProcess(2.5, 1.1, 0, 1, 0, range(1, 3))
    trees = _trees(m, 3, 1, 6, 24)
    appx = sum(oracle.stadler_appx_loglikelihood(m, t, 3) for t in trees)
    assert abs(appx - oracle.loglikelihood(m, trees, 3)) <= 1e-10 * abs(appx)
    # the same fully-observed identity with type-dependent birth rates (never exercised by the reference's tests)
    ts = np.array([-1.0, 0.0, 0.5, 2.0])
    m = gcdyn.SigmoidalBranchingProcess(2.0, 0.2, 1.5, 0.4, 1.1, 1.1, 1, 1, ts)
    trees = _trees(m, 2.5, ts[1], 6, 25)
    assert abs(sum(oracle.naive_loglikelihood(m, t) for t in trees) - oracle.loglikelihood(m, trees, 2.5)) <= 1e-10 * 100


def test_golden_fixtures_are_self_consistent():
    for name in ("g1_config1", "g2_config2", "g3_fully_observed", "g4_no_extinction", "g5_batched_discrete"):
        flat, params, T, exp, par = load_golden(name)
        if "ode_shared" in exp:
            P = exp["ode_branch"].shape[0]
            assert rel(exp["ode_shared"][:P], exp["ode_branch"]) <= 1e-10
        if "ode_closed" in exp:
            assert rel(exp["ode_branch"], exp["ode_closed"]) <= 1e-10
    _, _, _, exp, _ = load_golden("g3_fully_observed")
    assert rel(exp["ode_branch"], exp["naive"]) <= 1e-10
    _, _, _, exp, _ = load_golden("g4_no_extinction")
    assert rel(exp["ode_branch"], exp["stadler"]) <= 1e-10


def test_hand_built_tree_by_hand():
    """root(type 1) → birth at 1 → two survivors at T = 2, ρ = σ = 1: LL = log λ − Λ·1 + 2(log ρ − Λ·1)."""
    r = oracle.Node("root", 0, 1.0)
    b = r.add(oracle.Node("birth", 1.0, 1.0))
    b.add(oracle.Node("sampled_survival", 2.0, 1.0))
    b.add(oracle.Node("sampled_survival", 2.0, 1.0))
    m = oracle.Params([1.5, 1.5], 0.7, 1.0, [[-0.9, 0.9], [0.9, -0.9]], 1.0, 1.0, [1.0, 2.0])
    Lam = 1.5 + 0.7 + 0.9
    want = math.log(1.5) - Lam * 1.0 + 2 * (0.0 - Lam * 1.0)
    assert abs(oracle.loglikelihood(m, r, 2.0) - want) <= 1e-12
    assert abs(oracle.naive_loglikelihood(m, r) - want) <= 1e-12


def test_refcpu_c_restatement_matches_python_oracle():
    import refcpu
    for name in ("g2_config2", "g5_batched_discrete", "g1_config1"):
        flat, params, T, exp, par = load_golden(name)
        P = exp["ode_branch"].shape[0]
        got = refcpu.loglik(flat, par, T, rtol=1e-12, atol=1e-12)[:P]
        assert rel(got, exp["ode_branch"]) <= 1e-9
        assert rel(got.sum(axis=1), exp["ode_branch"].sum(axis=1)) <= 1e-10
        # default tolerances of the reference (reltol = abstol = 1e-3) land within a few 1e-3 of converged
        loose = refcpu.loglik(flat, par, T, rtol=1e-3, atol=1e-3)[:P]
        assert rel(loose.sum(axis=1), exp["ode_branch"].sum(axis=1)) <= 2e-2
