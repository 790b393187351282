"""Batched random-walk Metropolis over branching-process parameters (BASELINE.json config 4).

The reference ships no sampler (its analyses moved to another repository, README.md:3); this is the
consumer the batched likelihood was built for: C chains advance in lock-step, every iteration evaluates
all C proposals in ONE batched likelihood call (P = C parameter sets) against a forest that stays
resident in HBM, trees may be sharded over ranks (one all-reduce of C doubles per iteration), and the
accept/reject step runs on the host with a seeded NumPy stream, identically on every rank.

Parameters sampled: log xscale, xshift, log yscale, log yshift, log μ, log δ of a
`SigmoidalBranchingProcess`; Γ, ρ, σ and the type space are fixed.  Flat prior on a box.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .types import SigmoidalBranchingProcess
from .likelihood import DeviceParams, Forest, evaluate, pack_params

__all__ = ["MetropolisResult", "random_walk_metropolis", "theta_of", "model_of"]

NAMES = ("log_xscale", "xshift", "log_yscale", "log_yshift", "log_mu", "log_delta")


def theta_of(model: SigmoidalBranchingProcess) -> np.ndarray:
    return np.array([np.log(model.λ_xscale), model.λ_xshift, np.log(model.λ_yscale), np.log(model.λ_yshift),
                     np.log(model.μ), np.log(model.δ)])


def model_of(theta, template: SigmoidalBranchingProcess) -> SigmoidalBranchingProcess:
    return SigmoidalBranchingProcess(float(np.exp(theta[0])), float(theta[1]), float(np.exp(theta[2])),
                                     float(np.exp(theta[3])), float(np.exp(theta[4])), float(np.exp(theta[5])),
                                     template.Γ, template.ρ, template.σ, template.type_space)


@dataclass
class MetropolisResult:
    chain: np.ndarray            # [iterations+1, C, 6]
    loglik: np.ndarray           # [iterations+1, C]
    accepted: np.ndarray         # [C] number of accepted proposals
    names: tuple = field(default=NAMES)

    @property
    def acceptance_rate(self):
        return self.accepted / max(1, self.chain.shape[0] - 1)


def random_walk_metropolis(forest: Forest, template: SigmoidalBranchingProcess, present_time, theta0,
                           iterations, *, step=0.05, lower=None, upper=None, seed=0, loglik_fn=None,
                           allreduce=None) -> MetropolisResult:
    """Run C = len(theta0) chains for `iterations` steps.

    `loglik_fn(models) -> [C]` defaults to the GPU batched likelihood on `forest`; tests pass the oracle here
    to check that both produce the same chain from the same seed.  `allreduce(array) -> array` sums the
    per-rank partial log-likelihoods when the forest is a shard (see `gcdyn.sharding.allreduce_sum`)."""
    theta = np.array(theta0, dtype=np.float64)
    C, D = theta.shape
    lower = np.full(D, -np.inf) if lower is None else np.asarray(lower, dtype=np.float64)
    upper = np.full(D, np.inf) if upper is None else np.asarray(upper, dtype=np.float64)
    rng = np.random.default_rng(seed)

    gamma_col = np.ascontiguousarray(template.Γ.ravel(order="F"))
    ts = np.ascontiguousarray(template.type_space)

    def pack_theta(th):
        """Canonical arrays straight from θ (vectorised `sigmoid`, mbp.jl:4-7) — no per-chain model objects."""
        from .likelihood import PackedParams
        xs, xsh, ys, ysh = np.exp(th[:, 0]), th[:, 1], np.exp(th[:, 2]), np.exp(th[:, 3])
        lam = ys[:, None] * (1.0 / (1.0 + np.exp(-(xs[:, None] * (ts[None, :] - xsh[:, None]))))) + ysh[:, None]
        n = len(th)
        return PackedParams(n_psets=n, n_types=len(ts), response=0, gamma_pset_stride=0,
                            lam=np.ascontiguousarray(lam), xscale=None, xshift=None, yscale=None, yshift=None,
                            type_space=ts, mu=np.ascontiguousarray(np.exp(th[:, 4])),
                            delta=np.ascontiguousarray(np.exp(th[:, 5])), rho=np.full(n, template.ρ),
                            sigma=np.full(n, template.σ), Gamma=gamma_col)

    def loglik(th):
        if loglik_fn is not None:
            ll = np.asarray(loglik_fn([model_of(t, template) for t in th]), dtype=np.float64)
        else:
            ll, _, _ = evaluate(forest, pack_theta(th), present_time, want_status=False)
        return allreduce(ll) if allreduce is not None else ll

    ll = loglik(theta)
    chain = np.empty((iterations + 1, C, D))
    lls = np.empty((iterations + 1, C))
    chain[0], lls[0] = theta, ll
    accepted = np.zeros(C, dtype=np.int64)
    for it in range(1, iterations + 1):
        prop = theta + step * rng.standard_normal((C, D))
        inside = np.all((prop >= lower) & (prop <= upper), axis=1)
        # proposals outside the prior box are rejected without looking at their likelihood, but they are
        # still evaluated (clipped) so that every iteration is one batch of exactly C parameter sets
        ll_prop = loglik(np.clip(prop, np.maximum(lower, -50.0), np.minimum(upper, 50.0)))
        logu = np.log(rng.random(C))
        with np.errstate(invalid="ignore"):
            accept = inside & (logu < ll_prop - ll)
        theta = np.where(accept[:, None], prop, theta)
        ll = np.where(accept, ll_prop, ll)
        accepted += accept
        chain[it], lls[it] = theta, ll
    return MetropolisResult(chain=chain, loglik=lls, accepted=accepted)
