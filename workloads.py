"""Concrete synthetic workloads for BASELINE.json's configs (SURVEY.md §8d fixes rates and seeds, since the
reference publishes none).  Trees come from the host mirror of the reference's own simulator
(`rand_tree`, mbp.jl:318-461); every tree has its own RNG stream, so a forest does not depend on how
many ranks generate it."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if os.path.join(ROOT, "gcdyn.jl_b200") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "gcdyn.jl_b200"))

import gcdyn  # noqa: E402

CONFIGS = {
    # name: K, trees, psets, present_time, rho, sigma, seed
    "C1": dict(K=2, trees=10, psets=1, T=5.0, seed=1),
    "C2": dict(K=10, trees=100, psets=1, T=4.0, seed=2),
    "C3": dict(K=20, trees=1000, psets=256, T=4.0, seed=3),
    "C5s": dict(K=50, trees=200, psets=64, T=4.0, seed=5),   # small stand-in for config 5 (parity tests)
    # config 5 at scale: 10^4 trees of ~10^3 nodes from the NATIVE simulator (gcdyn_sim_forest), kept when 500 <= nodes <= 2000
    "C5": dict(K=50, trees=10000, psets=1024, T=5.0, seed=5, native=True),
}
CONFIGS["C4"] = dict(CONFIGS["C3"], psets=64)               # config 4: 64 Metropolis chains on the config-3 forest


def base_model(name):
    c = CONFIGS[name]
    if name == "C1":
        return gcdyn.ConstantBranchingProcess(1.8, 1.0, 1.0, 1, 0, [1, 2]), 1
    ts = np.linspace(-2.0, 2.0, c["K"])
    init = ts[c["K"] // 2 - 1]            # Julia's type_space[K÷2] (1-based)
    return gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts), init


def jittered_models(name, n_psets=None):
    """P parameter sets = base × independent LogNormal(0, 0.2) on (xscale, yscale, yshift, μ, δ) and
    Normal(0, 0.2) on xshift; Γ shared, ρ and σ fixed."""
    c = CONFIGS[name]
    P = n_psets or c["psets"]
    name = "C3" if name == "C4" else name
    base, _ = base_model(name)
    if P == 1 or name == "C1":
        return [base]
    rng = np.random.default_rng([c["seed"], 777])
    out = []
    for _ in range(P):
        f = np.exp(rng.normal(0.0, 0.2, 5))
        out.append(gcdyn.SigmoidalBranchingProcess(
            base.λ_xscale * f[0], base.λ_xshift + rng.normal(0.0, 0.2), base.λ_yscale * f[1],
            base.λ_yshift * f[2], base.μ * f[3], base.δ * f[4], base.Γ, base.ρ, base.σ, base.type_space))
    return out


def trees(name, first=0, count=None):
    """Trees [first, first+count) of the workload's (conceptually unbounded) tree sequence."""
    c = CONFIGS[name]
    count = c["trees"] if count is None else count
    model, init = base_model(name)
    out = []
    for i in range(first, first + count):
        rng = np.random.default_rng([c["seed"], i])
        out.append(gcdyn.rand_tree(model, c["T"], init, rng=rng))
    return out


def flat_forest(name, first=0, count=None):
    c = CONFIGS[name]
    if c.get("native"):
        # one simulator stream per shard: shard index = first // count (strong scaling: the shards differ with N; the
        # workload is the distribution, not one fixed forest — a 10^7-node forest from the host simulator would take hours)
        from gcdyn.sharding import subset_flat
        count = c["trees"] if count is None else count
        model, init = base_model(name)
        pool = gcdyn.simulate_flat(model, c["T"], init, 6 * count, seed=c["seed"] + 1000 * (first // max(count, 1)))
        sizes = np.diff(pool.tree_off)
        ok = np.nonzero((sizes >= 500) & (sizes <= 2000))[0][:count]
        if len(ok) < count:
            raise RuntimeError(f"simulator produced only {len(ok)} trees of 500..2000 nodes out of {6 * count}")
        return subset_flat(pool, ok)
    name = "C3" if name == "C4" else name
    model, _ = base_model(name)
    return gcdyn.flatten(trees(name, first, count), model.type_space, CONFIGS[name]["T"])


def par_dict(models):
    """Canonical arrays as plain numpy (what the oracle takes)."""
    return dict(lam=np.stack([m.birth_rates() for m in models]), mu=np.array([m.μ for m in models]),
                delta=np.array([m.δ for m in models]), rho=np.array([m.ρ for m in models]),
                sigma=np.array([m.σ for m in models]), Gamma=np.stack([m.Γ for m in models]))
