"""CPU ORACLE — test infrastructure only.

A restatement, in numpy/scipy, of gcdyn.jl's multitype branching-process tree
log-likelihood and its alternates.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; the
product (`gcdyn.jl_b200/`) never does and has no CPU fallback.

PARITY UNPINNED BY THE REFERENCE: gcdyn.jl ships no golden vectors (its three tests
are Monte-Carlo checks at rtol = 1e-1, test/runtests.jl:17,27,28,37) and neither
Julia nor OrdinaryDiffEq.jl (pinned 6.72.0, Manifest.toml:1067-1071) exists in this
environment, so the reference itself cannot be run to make fixtures.  The oracle is
pinned instead to (a) the reference's own test relations tightened to 1e-10
(`naive == ode` when ρ = σ = 1, runtests.jl:27; `stadler == ode` when γ = 0,
runtests.jl:37), (b) the closed form for equal birth rates that follows from
extra_likelihoods.jl:52-61,82-97, evaluated in 50-digit mpmath, and (c) agreement of
three independent integrations (per-branch recursion as written, shared trajectory,
mpmath Taylor series).  See tests/test_oracle.py and tests/golden/.

The ODE arithmetic of the reference lives in OrdinaryDiffEq.jl's `Tsit5()`; parity is
defined against the CONVERGED solution ("the reference's own solver, run at tight
tolerances"), so any integrator converged far below 1e-8 is an equivalent oracle:
here scipy's DOP853 at rtol = atol = 1e-13.

Trees are duck-typed: any object with `.event` (str), `.time`, `.type`,
`.children` (list) and `.up`.  Models are duck-typed on the reference's field names.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.integrate import solve_ivp

EVENTS = ("root", "birth", "sampled_death", "unsampled_death", "type_change",
          "sampled_survival", "unsampled_survival")   # types.jl:6


class Node:
    """Minimal stand-in for `TreeNode` (types.jl:23-37) for hand-built oracle trees."""

    def __init__(self, event, time, type):
        assert event in EVENTS
        self.event, self.time, self.type = event, float(time), type
        self.children, self.up = [], None

    def add(self, child):          # attach!, treenode.jl:10-15
        self.children.append(child)
        child.up = self
        return child


# ------------------------------------------------------------------ rates (mbp.jl:4-97)

class Params:
    """Canonical parameter set: what every model type reduces to on the likelihood path."""

    def __init__(self, lam, mu, delta, Gamma, rho, sigma, type_space):
        self.lam = np.asarray(lam, dtype=np.float64).copy()
        self.mu, self.delta = float(mu), float(delta)
        self.Gamma = np.asarray(Gamma, dtype=np.float64).copy()
        self.rho, self.sigma = float(rho), float(sigma)
        self.type_space = np.asarray(type_space, dtype=np.float64).copy()
        self.K = len(self.type_space)

    def index(self, t):            # type_space_index, mbp.jl:16-18 (0-based here)
        hits = np.nonzero(self.type_space == t)[0]
        if len(hits) == 0:
            raise ValueError("The type must be in the type space of the model.")
        return int(hits[0])

    def gamma_out(self, i):        # γ(model, type), mbp.jl:67-75
        return self.delta * -self.Gamma[i, i]

    def gamma_to(self, i, j):      # γ(model, from, to), mbp.jl:84-97
        return 0.0 if i == j else self.delta * self.Gamma[i, j]


def canonical(model) -> Params:
    """Reduce a reference-style model object to `Params` (λ evaluated per type:
    Constant mbp.jl:27-29, Discrete :31-39, Sigmoidal :41-43 with `sigmoid` :4-7)."""
    if isinstance(model, Params):
        return model
    ts = np.asarray(model.type_space, dtype=np.float64)
    if hasattr(model, "λ_xscale"):
        lam = np.array([model.λ_yscale * (1.0 / (1.0 + math.exp(-(model.λ_xscale * (x - model.λ_xshift)))))
                        + model.λ_yshift for x in ts])
    elif np.ndim(model.λ) == 0:
        lam = np.full(len(ts), float(model.λ))
    else:
        lam = np.asarray(model.λ, dtype=np.float64)
    return Params(lam, model.μ, model.δ, model.Γ, model.ρ, model.σ, ts)


# ------------------------------------------------------------------ traversals (treenode.jl:159-215)

def postorder(root):
    out = []

    def rec(n):
        for c in n.children:
            rec(c)
        out.append(n)
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    rec(root)
    return out


def leaves(root):
    return [n for n in postorder(root) if len(n.children) == 0]   # same left-to-right order


# ------------------------------------------------------------------ ODE right-hand sides

def dp_dt(p, m: Params):
    """mbp.jl:266-278."""
    return -(m.lam + m.mu) * p + m.mu * (1.0 - m.sigma) + m.lam * p * p + m.delta * (m.Gamma @ p)


def _solve(f, y0, span, rtol, atol):
    if span[1] == span[0]:
        return np.array(y0, dtype=np.float64)
    sol = solve_ivp(f, span, y0, method="DOP853", rtol=rtol, atol=atol)
    if not sol.success:
        raise RuntimeError(sol.message)
    return sol.y[:, -1]


def loglikelihood(model, tree, present_time, rtol=1e-13, atol=1e-13):
    """Per-branch recursion exactly as written in mbp.jl:110-248 (pass A, pass B, pass C)."""
    if isinstance(tree, (list, tuple)):           # mbp.jl:250-259
        return sum(loglikelihood(model, t, present_time, rtol, atol) for t in tree)
    m = canonical(model)
    K = m.K
    p_start, p_end, logq_start, logq_end = {}, {}, {}, {}
    f_p = lambda t, p: dp_dt(p, m)
    # pass A, mbp.jl:124-164
    for leaf in leaves(tree):
        if leaf.event == "sampled_survival":
            if leaf.time != present_time:
                raise ValueError("Sampled survival events must all be at the present time.")
            p_end[id(leaf)] = np.full(K, 1.0 - m.rho)
        elif leaf.event == "sampled_death":
            if leaf.time == present_time:
                raise ValueError("Sampled death events must not be at the present time.")
            p_end[id(leaf)] = _solve(f_p, np.full(K, 1.0 - m.rho), (0.0, present_time - leaf.time), rtol, atol)
        else:
            raise ValueError("Leaf event must be either `:sampled_survival` or `:sampled_death`.")
    # pass B, mbp.jl:166-218
    for ev in postorder(tree.children[0]):
        x = m.index(ev.up.type)
        if ev.event == "sampled_survival":
            logq_end[id(ev)] = _log(m.rho)
        elif ev.event == "sampled_death":
            logq_end[id(ev)] = _log(m.mu) + _log(m.sigma)
        elif ev.event == "birth":
            p_end[id(ev)] = p_start[id(ev.children[0])]
            logq_end[id(ev)] = _log(m.lam[x]) + logq_start[id(ev.children[0])] + logq_start[id(ev.children[1])]
        elif ev.event == "type_change":
            if ev.type == ev.up.type:
                return -math.inf
            p_end[id(ev)] = p_start[id(ev.children[0])]
            logq_end[id(ev)] = _log(m.gamma_to(x, m.index(ev.type))) + logq_start[id(ev.children[0])]
        else:
            raise ValueError(f"Unknown event type {ev.event}")
        lam_x, Lam_x = m.lam[x], m.lam[x] + m.mu + m.gamma_out(x)

        def f_pq(t, u):            # dp_logq_dt!, mbp.jl:285-300
            du = np.empty_like(u)
            du[:-1] = dp_dt(u[:-1], m)
            du[-1] = -Lam_x + 2.0 * lam_x * u[x]
            return du
        lq0 = logq_end[id(ev)]
        if not math.isfinite(lq0):
            # log 0 = -Inf flows through the solve untouched (dlogq does not depend on logq)
            u = _solve(f_pq, np.append(p_end[id(ev)], 0.0), (ev.up.time, ev.time), rtol, atol)
            u[-1] = lq0
        else:
            u = _solve(f_pq, np.append(p_end[id(ev)], lq0), (ev.up.time, ev.time), rtol, atol)
        p_start[id(ev)] = u[:-1]
        logq_start[id(ev)] = u[-1]
    result = logq_start[id(tree.children[0])]
    # pass C, mbp.jl:220-247
    p = _solve(f_p, np.full(K, 1.0 - m.rho), (0.0, float(present_time)), rtol, atol)
    result -= _log(1.0 - p[m.index(tree.type)])
    return result


def _log(x):
    if x == 0:
        return -math.inf
    if x < 0:
        raise ValueError("DomainError: log of a negative number")
    return math.log(x)


# ------------------------------------------------------------------ shared-trajectory form

def trajectory(model, t_max, rtol=1e-13, atol=1e-13):
    """Dense solution of u = (p, I), I_i(τ) = ∫_0^τ p_i, on [0, t_max]."""
    m = canonical(model)
    K = m.K

    def f(t, u):
        return np.concatenate([dp_dt(u[:K], m), u[:K]])
    u0 = np.concatenate([np.full(K, 1.0 - m.rho), np.zeros(K)])
    sol = solve_ivp(f, (0.0, float(t_max)), u0, method="DOP853", rtol=rtol, atol=atol, dense_output=True)
    assert sol.success
    return sol


def loglikelihood_shared(model, tree, present_time, rtol=1e-13, atol=1e-13, sol=None):
    """LL = Σ_v [node_term − Λ_x Δt + 2 λ_x (I_x(T−t_up) − I_x(T−t_v))] − log(1 − p_r(T))."""
    m = canonical(model)
    K = m.K
    T = float(present_time)
    if sol is None:
        sol = trajectory(m, T, rtol, atol)
    if isinstance(tree, (list, tuple)):
        return sum(loglikelihood_shared(m, t, T, rtol, atol, sol) for t in tree)
    total = 0.0
    for v in postorder(tree.children[0]):
        x = m.index(v.up.type)
        if v.event == "sampled_survival":
            term = _log(m.rho)
        elif v.event == "sampled_death":
            term = _log(m.mu) + _log(m.sigma)
        elif v.event == "birth":
            term = _log(m.lam[x])
        elif v.event == "type_change":
            if v.type == v.up.type:
                return -math.inf
            term = _log(m.gamma_to(x, m.index(v.type)))
        else:
            raise ValueError(v.event)
        Lam = m.lam[x] + m.mu + m.gamma_out(x)
        Ia = sol.sol(T - v.up.time)[K + x]
        Ib = sol.sol(T - v.time)[K + x]
        total += term - Lam * (v.time - v.up.time) + 2.0 * m.lam[x] * (Ia - Ib)
    pT = sol.sol(T)[m.index(tree.type)]
    return total - _log(1.0 - pT)


# ------------------------------------------------------------------ closed form, equal birth rates (SURVEY §3.4 Corollary A)

def loglikelihood_equal_rates(model, tree, present_time, mp=None):
    """Exact LL when all λ_i are equal (every ConstantBranchingProcess): all p_i coincide, the
    mutation term of dp_dt! vanishes (rows of Γ sum to 0, types.jl:139), and p, ∫p take the
    single-type closed forms of extra_likelihoods.jl:52-61,82-97 with γ = 0 inside Λ.
    `mp` = an mpmath context for arbitrary precision, else float64."""
    m = canonical(model)
    assert np.all(m.lam == m.lam[0])
    if isinstance(tree, (list, tuple)):
        return sum(loglikelihood_equal_rates(m, t, present_time, mp) for t in tree)
    if mp is None:
        F, sqrt, exp, log = float, math.sqrt, math.exp, math.log
    else:
        F, sqrt, exp, log = mp.mpf, mp.sqrt, mp.exp, mp.log
    lam, mu, rho, sig, T = F(m.lam[0]), F(m.mu), F(m.rho), F(m.sigma), F(present_time)
    Lam = lam + mu
    c = sqrt(Lam * Lam - 4 * mu * (1 - sig) * lam)
    xx = (-Lam - c) / 2
    yy = (-Lam + c) / 2

    def h(t):
        return (yy + lam * (1 - rho)) * exp(-c * t) - xx - lam * (1 - rho)

    def lg(v):
        if v == 0:
            return F("-inf")
        return log(v)
    total = F(0)
    for v in postorder(tree.children[0]):
        x = m.index(v.up.type)
        if v.event == "sampled_survival":
            term = lg(rho)
        elif v.event == "sampled_death":
            term = lg(mu) + lg(sig)
        elif v.event == "birth":
            term = lg(lam)
        elif v.event == "type_change":
            if v.type == v.up.type:
                return -math.inf
            term = lg(F(m.delta) * F(m.Gamma[x, m.index(v.type)]))
        else:
            raise ValueError(v.event)
        gx = F(m.delta) * -F(m.Gamma[x, x])
        ts, te = T - F(v.up.time), T - F(v.time)
        total += term + c * (te - ts) + 2 * (log(h(te)) - log(h(ts))) - gx * (ts - te)
    p = -1 / lam * ((yy + lam * (1 - rho)) * xx * exp(-c * T) - yy * (xx + lam * (1 - rho))) / (
        (yy + lam * (1 - rho)) * exp(-c * T) - (xx + lam * (1 - rho)))
    m.index(tree.type)
    return total - log(1 - p)


# ------------------------------------------------------------------ alternates (extra_likelihoods.jl)

def _exp_logpdf(theta, x):
    # Distributions.jl 0.25.107 (Manifest.toml:323-327): logpdf(Exponential(θ), x) = -(log θ + x/θ) on x ≥ 0
    return -(math.log(theta) + x / theta) if x >= 0 else -math.inf


def _exp_logccdf_as_written(theta, x):
    # the reference writes log(1 - cdf(d, x)); cdf(Exponential) = -expm1(-max(x,0)/θ)
    z = max(x, 0.0) / theta
    return _log(1.0 - (-math.expm1(-z)))


def naive_loglikelihood(model, tree):
    """extra_likelihoods.jl:10-34."""
    m = canonical(model)
    result = 0.0
    for node in postorder(tree.children[0]):
        x = m.index(node.up.type)
        lam, mu, gam = m.lam[x], m.mu, m.gamma_out(x)
        Lam = lam + mu + gam
        dt = node.time - node.up.time
        if node.event == "birth":
            result += _exp_logpdf(1 / Lam, dt) + _log(lam / Lam)
        elif node.event == "sampled_death":
            result += _exp_logpdf(1 / Lam, dt) + _log(mu / Lam)
        elif node.event == "type_change":
            rate = 0.0 if node.type == node.up.type else m.gamma_to(x, m.index(node.type))
            result += _exp_logpdf(1 / Lam, dt) + _log(rate) - _log(Lam)
        elif node.event == "sampled_survival":
            result += _exp_logccdf_as_written(1 / Lam, dt) + _log(m.rho)
        elif node.event == "unsampled_survival":
            result += _exp_logccdf_as_written(1 / Lam, dt) + _log(1 - m.rho)
        else:
            raise ValueError(f"Tree contains incompatible event {node.event}")
    return result


def _stadler_sum(m: Params, tree, present_time):
    # extra_likelihoods.jl:49-77 (= :118-148)
    result = 0.0
    rho, sig = m.rho, m.sigma
    for node in postorder(tree.children[0]):
        x = m.index(node.up.type)
        lam, mu, gam = m.lam[x], m.mu, m.gamma_out(x)
        Lam = lam + mu + gam
        c = math.sqrt(Lam ** 2 - 4 * mu * (1 - sig) * lam)
        xx = (-Lam - c) / 2
        yy = (-Lam + c) / 2
        helper = lambda t: (yy + lam * (1 - rho)) * math.exp(-c * t) - xx - lam * (1 - rho)
        t_s = present_time - node.up.time
        t_e = present_time - node.time
        result += c * (t_e - t_s) + 2 * (_log(helper(t_e)) - _log(helper(t_s)))
        if node.event == "birth":
            result += _log(lam)
        elif node.event == "sampled_death":
            result += _log(sig) + _log(mu)
        elif node.event == "type_change":
            rate = 0.0 if node.type == node.up.type else m.gamma_to(x, m.index(node.type))
            result += _log(rate)
        elif node.event == "sampled_survival":
            result += _log(rho)
        else:
            raise ValueError(f"Tree contains incompatible event {node.event}")
    return result


def stadler_appx_unconditioned_loglikelihood(model, tree, present_time):
    """extra_likelihoods.jl:115-151."""
    return _stadler_sum(canonical(model), tree, present_time)


def stadler_appx_loglikelihood(model, tree, present_time):
    """extra_likelihoods.jl:45-102."""
    m = canonical(model)
    result = _stadler_sum(m, tree, present_time)
    r = m.index(tree.type)
    lam, mu, gam = m.lam[r], m.mu, m.gamma_out(r)
    rho, sig = m.rho, m.sigma
    Lam = lam + mu + gam
    root_time = present_time - 0
    c = math.sqrt(Lam ** 2 - 4 * mu * (1 - sig) * lam)
    xx = (-Lam - c) / 2
    yy = (-Lam + c) / 2
    p = (-1 / lam
         * ((yy + lam * (1 - rho)) * xx * math.exp(-c * root_time) - yy * (xx + lam * (1 - rho)))
         / ((yy + lam * (1 - rho)) * math.exp(-c * root_time) - (xx + lam * (1 - rho))))
    return result - _log(1 - p)


# ------------------------------------------------------------------ flattener restatement (integer parity)

def flatten(trees, type_space):
    """Independent restatement of the post-order struct-of-arrays the product's flattener must
    reproduce bit-for-bit: nodes = PostOrderTraversal(tree.children[1]) (treenode.jl:159-175,
    mbp.jl:166); type index = findfirst(==(type), type_space) − 1 (mbp.jl:17), −1 if absent."""
    ts = np.asarray(type_space, dtype=np.float64)

    def idx(t):
        hits = np.nonzero(ts == t)[0]
        return int(hits[0]) if len(hits) else -1
    out = dict(tree_off=[0], event=[], type_idx=[], btype_idx=[], t_start=[], t_end=[], parent=[],
               child_off=[0], child_idx=[], root_type_idx=[], root_time=[])
    for tree in trees:
        nodes = postorder(tree.children[0])
        pos = {id(n): i for i, n in enumerate(nodes)}
        for n in nodes:
            out["event"].append(EVENTS.index(n.event))
            out["type_idx"].append(idx(n.type))
            out["btype_idx"].append(idx(n.up.type))
            out["t_start"].append(n.up.time)
            out["t_end"].append(n.time)
            out["parent"].append(pos[id(n.up)] if id(n.up) in pos else -1)
            out["child_idx"].extend(pos[id(c)] for c in n.children)
            out["child_off"].append(len(out["child_idx"]))
        out["tree_off"].append(len(out["event"]))
        out["root_type_idx"].append(idx(tree.type))
        out["root_time"].append(tree.time)
    dt = dict(tree_off=np.int64, event=np.uint8, type_idx=np.int32, btype_idx=np.int32, t_start=np.float64,
              t_end=np.float64, parent=np.int32, child_off=np.int64, child_idx=np.int32,
              root_type_idx=np.int32, root_time=np.float64)
    return {k: np.asarray(v, dtype=dt[k]) for k, v in out.items()}


def event_counts(trees):
    """Per-tree histogram of event codes over PostOrderTraversal(tree.children[1])."""
    out = np.zeros((len(trees), len(EVENTS)), dtype=np.int64)
    for t, tree in enumerate(trees):
        for n in postorder(tree.children[0]):
            out[t, EVENTS.index(n.event)] += 1
    return out


# ------------------------------------------------------------------ vectorised shared-trajectory form on flat arrays

def loglik_flat_shared(flat, par, p, present_time, rtol=1e-13, atol=1e-13):
    """Per-tree log-likelihoods [T] for parameter set `p` of `par` (dict: lam [P,K], mu, delta, rho, sigma [P],
    Gamma [P,K,K] or [K,K] row-major) on a flattened forest, by the shared-trajectory formula of
    `loglikelihood_shared` evaluated for all nodes at once.  Used for parity checks at sizes where the
    per-branch recursion is too slow.  Ignores `live`/`self_loop` (valid, simulator-made trees only)."""
    g = lambda name: np.asarray(getattr(flat, name) if hasattr(flat, name) else flat[name])
    G = np.asarray(par["Gamma"], dtype=np.float64)
    G = G[p] if G.ndim == 3 else G
    m = Params(par["lam"][p], par["mu"][p], par["delta"][p], G, par["rho"][p], par["sigma"][p],
               np.arange(len(par["lam"][p])))
    K, T = m.K, float(present_time)
    sol = trajectory(m, T, rtol, atol)
    ev, x, y = g("event").astype(np.int64), g("btype_idx").astype(np.int64), g("type_idx").astype(np.int64)
    ts, te = g("t_start"), g("t_end")
    tree_off = g("tree_off")
    n = len(ev)

    def I_at(tau, typ):
        out = np.empty(len(tau))
        for lo in range(0, len(tau), 65536):
            sl = slice(lo, lo + 65536)
            out[sl] = sol.sol(tau[sl])[K + typ[sl], np.arange(len(tau[sl]))]
        return out
    Ia, Ib = I_at(T - ts, x), I_at(T - te, x)
    Lam = m.lam + m.mu + m.delta * -np.diag(m.Gamma)
    with np.errstate(divide="ignore"):
        loglam = np.log(m.lam)
        logG = np.log(m.delta * m.Gamma)
        term = np.select([ev == 5, ev == 2, ev == 1, ev == 4],
                         [np.full(n, _log(m.rho)), np.full(n, _log(m.mu) + _log(m.sigma)), loglam[x],
                          logG[x, np.maximum(y, 0)]], default=np.nan)
    contrib = term - Lam[x] * (te - ts) + 2.0 * m.lam[x] * (Ia - Ib)
    sums = np.add.reduceat(contrib, tree_off[:-1]) if n else np.zeros(len(tree_off) - 1)
    pT = sol.sol(T)[:K]
    return sums - np.log1p(-pT)[g("root_type_idx")]


def alt_flat(flat, par, p, present_time, method):
    """Per-tree values [T] of the alternates (extra_likelihoods.jl:10-34, :45-102, :115-151) for parameter set `p` on a
    flattened forest — the per-node formulas of `naive_loglikelihood` / `_stadler_sum` above applied node by node.
    method ∈ {"naive", "stadler", "stadler_uncond"(= "stadler_unconditioned")}."""
    g = lambda name: np.asarray(getattr(flat, name) if hasattr(flat, name) else flat[name])
    G = np.asarray(par["Gamma"], dtype=np.float64)
    G = G[p] if G.ndim == 3 else G
    m = Params(par["lam"][p], par["mu"][p], par["delta"][p], G, par["rho"][p], par["sigma"][p],
               np.arange(len(par["lam"][p])))
    ev, xs, ys = g("event"), g("btype_idx"), g("type_idx")
    ts, te, off = g("t_start"), g("t_end"), g("tree_off")
    names = EVENTS
    out = np.zeros(len(off) - 1)
    T = present_time
    for t in range(len(off) - 1):
        result = 0.0
        for i in range(int(off[t]), int(off[t + 1])):
            x, y, e = int(xs[i]), int(ys[i]), names[int(ev[i])]
            lam, mu, gam = m.lam[x], m.mu, m.gamma_out(x)
            Lam = lam + mu + gam
            rate = 0.0 if y == x else (m.gamma_to(x, y) if y >= 0 else float("nan"))
            if method == "naive":
                dt = te[i] - ts[i]
                if e == "birth":
                    result += _exp_logpdf(1 / Lam, dt) + _log(lam / Lam)
                elif e == "sampled_death":
                    result += _exp_logpdf(1 / Lam, dt) + _log(mu / Lam)
                elif e == "type_change":
                    result += _exp_logpdf(1 / Lam, dt) + _log(rate) - _log(Lam)
                elif e == "sampled_survival":
                    result += _exp_logccdf_as_written(1 / Lam, dt) + _log(m.rho)
                elif e == "unsampled_survival":
                    result += _exp_logccdf_as_written(1 / Lam, dt) + _log(1 - m.rho)
                else:
                    raise ValueError(f"Tree contains incompatible event {e}")
            else:
                c = math.sqrt(Lam ** 2 - 4 * mu * (1 - m.sigma) * lam)
                xx, yy = (-Lam - c) / 2, (-Lam + c) / 2
                helper = lambda tt: (yy + lam * (1 - m.rho)) * math.exp(-c * tt) - xx - lam * (1 - m.rho)
                t_s, t_e = T - ts[i], T - te[i]
                result += c * (t_e - t_s) + 2 * (_log(helper(t_e)) - _log(helper(t_s)))
                if e == "birth":
                    result += _log(lam)
                elif e == "sampled_death":
                    result += _log(m.sigma) + _log(mu)
                elif e == "type_change":
                    result += _log(rate)
                elif e == "sampled_survival":
                    result += _log(m.rho)
                else:
                    raise ValueError(f"Tree contains incompatible event {e}")
        if method == "stadler":
            r = int(g("root_type_idx")[t])
            lam, mu, gam = m.lam[r], m.mu, m.gamma_out(r)
            Lam = lam + mu + gam
            c = math.sqrt(Lam ** 2 - 4 * mu * (1 - m.sigma) * lam)
            xx, yy = (-Lam - c) / 2, (-Lam + c) / 2
            rt = T - 0
            pr = (-1 / lam * ((yy + lam * (1 - m.rho)) * xx * math.exp(-c * rt) - yy * (xx + lam * (1 - m.rho)))
                  / ((yy + lam * (1 - m.rho)) * math.exp(-c * rt) - (xx + lam * (1 - m.rho))))
            result -= _log(1 - pr)
        out[t] = result
    return out
