"""Generates tests/golden/*.npz.  Run here (CPU container) with:  python tests/golden/make_golden.py

The reference (Julia) cannot be executed in this environment and ships no golden vectors
(test/runtests.jl holds three Monte-Carlo checks at rtol 1e-1), so these fixtures come from the
oracle: the per-branch recursion written as mbp.jl:110-248 with scipy DOP853 at 1e-13, the
closed form for equal birth rates in 50-digit mpmath where it applies, and the alternates of
extra_likelihoods.jl.  Each file stores the flattened forest (from oracle.flatten, not from the
product), the canonical parameter arrays and the expected values.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "gcdyn.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import mpmath  # noqa: E402
import gcdyn  # noqa: E402  (only the simulator, to draw trees)
import oracle  # noqa: E402


def canon_arrays(models):
    ps = [oracle.canonical(m) for m in models]
    return dict(lam=np.stack([p.lam for p in ps]), mu=np.array([p.mu for p in ps]),
                delta=np.array([p.delta for p in ps]), rho=np.array([p.rho for p in ps]),
                sigma=np.array([p.sigma for p in ps]),
                Gamma=np.stack([p.Gamma for p in ps]),   # row-major [P,K,K] as numpy; loaders convert to column-major
                type_space=ps[0].type_space)


def save(name, trees, models, T, extra):
    flat = oracle.flatten(trees, models[0].type_space)
    out = {f"flat_{k}": v for k, v in flat.items()}
    out.update({f"par_{k}": v for k, v in canon_arrays(models).items()})
    out["present_time"] = np.float64(T)
    out["event_counts"] = oracle.event_counts(trees)
    out.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "trees", len(trees), "nodes", len(flat["event"]), {k: np.shape(v) for k, v in extra.items()})


def per_tree(fn, models, trees, *a):
    return np.array([[fn(m, t, *a) for t in trees] for m in models])


def main():
    mpmath.mp.dps = 50
    # --- G1 = BASELINE config 1: 10 trees, 2 types, constant rates, symmetric mutation, T=5, ρ=1, σ=0
    rng = np.random.default_rng(1)
    m = gcdyn.ConstantBranchingProcess(1.8, 1.0, 1.0, 1, 0, [1, 2])
    T = 5.0
    trees = gcdyn.rand_tree(m, T, 1, 10, rng=rng)
    save("g1_config1", trees, [m], T, dict(
        ode_branch=per_tree(oracle.loglikelihood, [m], trees, T),
        ode_shared=per_tree(oracle.loglikelihood_shared, [m], trees, T),
        ode_closed=np.array([[float(oracle.loglikelihood_equal_rates(m, t, T, mp=mpmath.mp)) for t in trees]]),
        stadler=per_tree(oracle.stadler_appx_loglikelihood, [m], trees, T),
        stadler_uncond=per_tree(oracle.stadler_appx_unconditioned_loglikelihood, [m], trees, T),
        naive=per_tree(oracle.naive_loglikelihood, [m], trees)))

    # --- G2 = BASELINE config 2, reduced to 24 trees: K=10 sigmoid response, ρ=0.5, σ=0.1, T=4
    rng = np.random.default_rng(2)
    ts = np.linspace(-2, 2, 10)
    m = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    T = 4.0
    trees = gcdyn.rand_tree(m, T, ts[4], 24, rng=rng)
    save("g2_config2", trees, [m], T, dict(
        ode_branch=per_tree(oracle.loglikelihood, [m], trees, T),
        ode_shared=per_tree(oracle.loglikelihood_shared, [m], trees, T),
        stadler=per_tree(oracle.stadler_appx_loglikelihood, [m], trees, T),
        stadler_uncond=per_tree(oracle.stadler_appx_unconditioned_loglikelihood, [m], trees, T),
        naive=per_tree(oracle.naive_loglikelihood, [m], trees)))

    # --- G3: ρ = σ = 1 (fully observed): loglikelihood == naive_loglikelihood (runtests.jl:27), K=4 sigmoid
    rng = np.random.default_rng(3)
    ts = np.array([-1.0, 0.0, 0.5, 2.0])
    m = gcdyn.SigmoidalBranchingProcess(2.0, 0.2, 1.5, 0.4, 1.1, 1.1, 1, 1, ts)
    T = 2.5
    trees = gcdyn.rand_tree(m, T, ts[1], 12, rng=rng)
    save("g3_fully_observed", trees, [m], T, dict(
        ode_branch=per_tree(oracle.loglikelihood, [m], trees, T),
        naive=per_tree(oracle.naive_loglikelihood, [m], trees)))

    # --- G4: runtests.jl:31-38 — γ = 0, ρ = 1, σ = 0: stadler_appx == loglikelihood exactly
    rng = np.random.default_rng(4)
    m = gcdyn.ConstantBranchingProcess(2.5, 1.1, 0, 1, 0, [1, 2])
    T = 3.0
    trees = gcdyn.rand_tree(m, T, 1, 12, rng=rng)
    save("g4_no_extinction", trees, [m], T, dict(
        ode_branch=per_tree(oracle.loglikelihood, [m], trees, T),
        ode_closed=np.array([[float(oracle.loglikelihood_equal_rates(m, t, T, mp=mpmath.mp)) for t in trees]]),
        stadler=per_tree(oracle.stadler_appx_loglikelihood, [m], trees, T)))

    # --- G5: batched, DiscreteBranchingProcess, non-uniform Γ, δ ≠ 1, 5 parameter sets with their own Γ, ρ, σ
    rng = np.random.default_rng(5)
    ts = np.array([1.0, 2.0, 3.0])
    G0 = np.array([[-1.0, 0.7, 0.3], [0.2, -0.5, 0.3], [0.6, 0.6, -1.2]])
    base = gcdyn.DiscreteBranchingProcess([1.2, 2.0, 2.6], 0.9, 1.3, G0, 0.6, 0.3, ts)
    T = 3.0
    trees = gcdyn.rand_tree(base, T, 2.0, 10, rng=rng)
    models = [base]
    for _ in range(4):
        G = G0 * rng.uniform(0.5, 1.5, size=(3, 3))
        np.fill_diagonal(G, 0.0)
        np.fill_diagonal(G, -G.sum(axis=1))
        models.append(gcdyn.DiscreteBranchingProcess(rng.uniform(0.8, 3.0, 3), rng.uniform(0.5, 1.5),
                                                     rng.uniform(0.5, 2.0), G, rng.uniform(0.2, 1.0),
                                                     rng.uniform(0.1, 1.0), ts))
    save("g5_batched_discrete", trees, models, T, dict(
        ode_branch=per_tree(oracle.loglikelihood, models[:2], trees, T),
        ode_shared=per_tree(oracle.loglikelihood_shared, models, trees, T),
        stadler=per_tree(oracle.stadler_appx_loglikelihood, models, trees, T),
        naive=per_tree(oracle.naive_loglikelihood, models, trees)))


if __name__ == "__main__":
    main()
