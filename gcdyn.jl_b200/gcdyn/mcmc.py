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

__all__ = ["MetropolisResult", "random_walk_metropolis", "device_metropolis", "CounterRNG", "loglikelihood_gradient", "loglikelihood_gradient_exact", "theta_of", "model_of"]

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


class CounterRNG:
    """The counter-based stream of `csrc/metropolis.cuh`: uniform(seed, iteration, chain, slot) is a splitmix64
    hash, normals are Box–Muller on slots (2j, 2j+1), the accept draw is slot 12.  Pure NumPy, so the host-driven
    sampler below can replay exactly what the device-resident one draws."""

    M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

    def __init__(self, seed):
        self.seed = np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)

    @staticmethod
    def _mix(z):
        with np.errstate(over="ignore"):
            z = (z + np.uint64(0x9E3779B97F4A7C15))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, it, chain, slot):
        chain = np.asarray(chain, dtype=np.uint64)
        slot = np.asarray(slot, dtype=np.uint64)
        with np.errstate(over="ignore"):
            h = self._mix(self._mix(self._mix(self.seed) ^ np.uint64(it)) ^ (chain * np.uint64(16) + slot))
        return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)

    def normals(self, it, n_chains, dim):
        c = np.arange(n_chains)[:, None]
        j = np.arange(dim)[None, :]
        u1, u2 = self.uniform(it, c, 2 * j), self.uniform(it, c, 2 * j + 1)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)

    def log_uniform(self, it, n_chains, dim):
        return np.log(self.uniform(it, np.arange(n_chains), np.full(n_chains, 2 * dim)))


def device_metropolis(forest: Forest, template: SigmoidalBranchingProcess, present_time, theta0, iterations, *,
                      step=0.05, lower, upper, seed=0, comm=None, history=True, checkpoint=None) -> MetropolisResult:
    """`gcdyn_mh_*`: the same sampler as `random_walk_metropolis(..., rng="counter")`, but proposals, likelihood,
    all-reduce (`comm`: a `gcdyn.sharding.PeerAllReduce` when the forest is a shard) and accept/reject all run on
    the device; the host only enqueues.  `upper[2:]` must be finite.  `checkpoint` = the `.checkpoint` of an earlier
    result: the chains continue from there with the random numbers of the following iterations, so a split run
    reproduces an uninterrupted one."""
    import ctypes as C
    from .likelihood import MhDesc, _lib, _check, _ptr, _p_f64, _p_i64
    lib = _lib()
    theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
    n, D = theta0.shape
    assert D == 6
    ts = np.ascontiguousarray(template.type_space, dtype=np.float64)
    gam = np.ascontiguousarray(np.asarray(template.Γ, dtype=np.float64).ravel(order="F"))
    lo = np.ascontiguousarray(lower, dtype=np.float64)
    up = np.ascontiguousarray(upper, dtype=np.float64)
    desc = MhDesc(n, len(ts), _ptr(ts, _p_f64), _ptr(gam, _p_f64), float(template.ρ), float(template.σ),
                  _ptr(theta0, _p_f64), _ptr(lo, _p_f64), _ptr(up, _p_f64), float(step), int(seed) & 0xFFFFFFFFFFFFFFFF)
    h = C.c_void_p()
    ctxh = forest.ctx._h
    _check(lib.gcdyn_mh_create(ctxh, forest._h, C.byref(desc), float(present_time),
                               comm._h if comm is not None else None, C.byref(h)), ctxh)
    try:
        th = np.empty((n, 6)); ll = np.empty(n); acc = np.zeros(n, dtype=np.int64); it = C.c_int64()
        if checkpoint is not None:
            cth, cll, cacc, cit = checkpoint
            _check(lib.gcdyn_mh_restore(h, _ptr(np.ascontiguousarray(cth, dtype=np.float64), _p_f64),
                                        _ptr(np.ascontiguousarray(cll, dtype=np.float64), _p_f64),
                                        _ptr(np.ascontiguousarray(cacc, dtype=np.int64), _p_i64), int(cit)), ctxh)
        _check(lib.gcdyn_mh_run(h, 0, None, None), ctxh)          # initial log-likelihood (kept when restored)
        _check(lib.gcdyn_mh_state(h, _ptr(th, _p_f64), _ptr(ll, _p_f64), _ptr(acc, _p_i64), C.byref(it)), ctxh)
        if history:
            chain = np.empty((iterations + 1, n, 6)); lls = np.empty((iterations + 1, n))
            chain[0], lls[0] = th, ll
            _check(lib.gcdyn_mh_run(h, int(iterations), _ptr(chain[1:], _p_f64), _ptr(lls[1:], _p_f64)), ctxh)
        else:
            _check(lib.gcdyn_mh_run(h, int(iterations), None, None), ctxh)
        _check(lib.gcdyn_mh_state(h, _ptr(th, _p_f64), _ptr(ll, _p_f64), _ptr(acc, _p_i64), C.byref(it)), ctxh)
        if not history:
            chain, lls = th[None].copy(), ll[None].copy()
    finally:
        lib.gcdyn_mh_destroy(h)
    res = MetropolisResult(chain=chain, loglik=lls, accepted=acc)
    res.checkpoint = (th.copy(), ll.copy(), acc.copy(), int(it.value))
    return res


def random_walk_metropolis(forest: Forest, template: SigmoidalBranchingProcess, present_time, theta0,
                           iterations, *, step=0.05, lower=None, upper=None, seed=0, loglik_fn=None,
                           allreduce=None, rng="numpy") -> MetropolisResult:
    """Run C = len(theta0) chains for `iterations` steps.

    `loglik_fn(models) -> [C]` defaults to the GPU batched likelihood on `forest`; tests pass the oracle here
    to check that both produce the same chain from the same seed.  `allreduce(array) -> array` sums the
    per-rank partial log-likelihoods when the forest is a shard (see `gcdyn.sharding.allreduce_sum`).
    `rng="counter"` draws from `CounterRNG`, the stream of the device-resident sampler (`device_metropolis`)."""
    theta = np.array(theta0, dtype=np.float64)
    C, D = theta.shape
    lower = np.full(D, -np.inf) if lower is None else np.asarray(lower, dtype=np.float64)
    upper = np.full(D, np.inf) if upper is None else np.asarray(upper, dtype=np.float64)
    counter = CounterRNG(seed) if rng == "counter" else None
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
        prop = theta + step * (counter.normals(it, C, D) if counter else rng.standard_normal((C, D)))
        inside = np.all((prop >= lower) & (prop <= upper), axis=1)
        # proposals outside the prior box are rejected without looking at their likelihood, but they are
        # still evaluated (clipped) so that every iteration is one batch of exactly C parameter sets
        ll_prop = loglik(np.clip(prop, np.maximum(lower, -50.0), np.minimum(upper, 50.0)))
        logu = counter.log_uniform(it, C, D) if counter else np.log(rng.random(C))
        with np.errstate(invalid="ignore"):
            accept = inside & (logu < ll_prop - ll)
        theta = np.where(accept[:, None], prop, theta)
        ll = np.where(accept, ll_prop, ll)
        accepted += accept
        chain[it], lls[it] = theta, ll
    return MetropolisResult(chain=chain, loglik=lls, accepted=accepted)


def loglikelihood_gradient_exact(forest: Forest, template: SigmoidalBranchingProcess, present_time, thetas):
    """Log-likelihoods [C] and EXACT gradients [C, 6] w.r.t. θ = (log xscale, xshift, log yscale, log yshift, log μ, log δ)
    for a batch of C points: `gcdyn_loglik_grad_batch` integrates the forward sensitivities of the trajectory for the raw
    parameters (xscale, xshift, yscale, yshift, μ, δ) — what ForwardDiff does for the reference through its `T<:Real`
    structs (types.jl:92, mbp.jl:118-122) — and the chain rule through exp() maps them to θ."""
    from .likelihood import PackedParams, loglikelihood_and_gradient
    th = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    n = len(th)
    ts = np.ascontiguousarray(template.type_space)
    raw = np.stack([np.exp(th[:, 0]), th[:, 1], np.exp(th[:, 2]), np.exp(th[:, 3]), np.exp(th[:, 4]), np.exp(th[:, 5])], axis=1)
    c = lambda k: np.ascontiguousarray(raw[:, k])
    params = PackedParams(n_psets=n, n_types=len(ts), response=2, gamma_pset_stride=0, lam=None, xscale=c(0), xshift=c(1),
                          yscale=c(2), yshift=c(3), type_space=ts, mu=c(4), delta=c(5), rho=np.full(n, float(template.ρ)),
                          sigma=np.full(n, float(template.σ)),
                          Gamma=np.ascontiguousarray(np.asarray(template.Γ, dtype=np.float64).ravel(order="F")))
    ll, g = loglikelihood_and_gradient(forest, params, present_time)
    scale = np.ones_like(raw)
    for k in (0, 2, 3, 4, 5):
        scale[:, k] = raw[:, k]                  # d/d log x = x d/dx
    return ll, g * scale


def loglikelihood_gradient(forest: Forest, template: SigmoidalBranchingProcess, present_time, theta, *, rel_step=1e-4,
                           comm=None):
    """Log-likelihood and its gradient w.r.t. θ = (log xscale, xshift, log yscale, log yshift, log μ, log δ)
    (first form).  The reference is parametric in `T<:Real` so that ForwardDiff can push dual
    numbers through `loglikelihood` (types.jl:92, mbp.jl:118); behind a Float64 C ABI the same quantity comes from
    ONE batched evaluation of 13 parameter sets — θ and θ ± h·e_k — i.e. central differences riding on the batch
    dimension (they take the tensor-core path like any batch of ≥ 8).  With the likelihood accurate to ~1e-13 relative
    and h = rel_step·max(1, |θ_k|), the gradient is good to ~1e-7 relative (truncation h²·f‴/6 against round-off
    ε|f|/h); forward sensitivities of the (p, F) trajectory remain the exact alternative and are not built.
    Returns (loglik, grad[6])."""
    theta = np.asarray(theta, dtype=np.float64)
    D = len(theta)
    h = rel_step * np.maximum(1.0, np.abs(theta))
    pts = np.tile(theta, (2 * D + 1, 1))
    for k in range(D):
        pts[1 + 2 * k, k] += h[k]
        pts[2 + 2 * k, k] -= h[k]
    ts = np.ascontiguousarray(template.type_space)
    from .likelihood import PackedParams
    xs, xsh, ys, ysh = np.exp(pts[:, 0]), pts[:, 1], np.exp(pts[:, 2]), np.exp(pts[:, 3])
    lam = ys[:, None] * (1.0 / (1.0 + np.exp(-(xs[:, None] * (ts[None, :] - xsh[:, None]))))) + ysh[:, None]
    n = len(pts)
    params = PackedParams(n_psets=n, n_types=len(ts), response=0, gamma_pset_stride=0, lam=np.ascontiguousarray(lam),
                          xscale=None, xshift=None, yscale=None, yshift=None, type_space=ts,
                          mu=np.ascontiguousarray(np.exp(pts[:, 4])), delta=np.ascontiguousarray(np.exp(pts[:, 5])),
                          rho=np.full(n, template.ρ), sigma=np.full(n, template.σ),
                          Gamma=np.ascontiguousarray(np.asarray(template.Γ, dtype=np.float64).ravel(order="F")))
    ll, _, _ = evaluate(forest, params, present_time, want_status=False, comm=comm)
    grad = (ll[1::2] - ll[2::2]) / (2.0 * h)
    return float(ll[0]), grad
