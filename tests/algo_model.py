"""numpy MODEL of the GPU algorithm — test infrastructure, not a product path.

It restates, step for step, what the CUDA kernels in `gcdyn.jl_b200/csrc/` do
(Taylor-series trajectory, per-cell polynomial table of F_x, lookup items, per-tree
sums, moment contraction) so the numerical design (order M, steps S) can be
validated against the oracle on a machine without a GPU.  It is never imported by
the product.
"""
from __future__ import annotations

import numpy as np

EV_ROOT, EV_BIRTH, EV_SDEATH, EV_UDEATH, EV_TC, EV_SSURV, EV_USURV = range(7)


def taylor_table(lam, mu, delta, Gamma, rho, sigma, T, S, M):
    """Integrate p' = -(λ+μ)p + μ(1-σ) + λp² + δΓp from p(0)=1-ρ over [0,T] in S Taylor steps of
    order M.  Returns (Fc[S, M+2, K], p_T[K], err) where on cell s, with θ = (τ-τ_s)/h ∈ [0,1],
        F_x(τ) = Σ_n Fc[s,n,x] θ^n,   F_x' = -Λ_x + 2 λ_x p_x,  F_x(0) = 0,
    and err = max over steps of |ĉ_M| + |ĉ_{M-1}| (h-scaled last Taylor coefficients of p)."""
    lam = np.asarray(lam, float)
    K = len(lam)
    G = delta * np.asarray(Gamma, float)
    a = lam + mu
    b = mu * (1.0 - sigma)
    Lam = lam + mu - np.diag(G)
    h = T / S
    Fc = np.zeros((S, M + 2, K))
    c = np.zeros((M + 1, K))
    p = np.full(K, 1.0 - rho)
    F = np.zeros(K)
    err = 0.0
    bad = False
    for s in range(S):
        c[0] = p
        for n in range(M):
            conv = np.zeros(K)
            for m in range(n + 1):
                conv += c[m] * c[n - m]
            rhs = -a * c[n] + lam * conv + G @ c[n]
            if n == 0:
                rhs = rhs + b
            c[n + 1] = rhs * (h / (n + 1))
        Fc[s, 0] = F
        Fc[s, 1] = h * (2.0 * lam * c[0] - Lam)
        for n in range(1, M + 1):
            Fc[s, n + 1] = h * 2.0 * lam * c[n] / (n + 1)
        err = max(err, np.max(np.abs(c[M]) + np.abs(c[M - 1])))
        p = c.sum(axis=0)
        F = Fc[s].sum(axis=0)
        if not np.all(np.isfinite(p)) or np.any(p < -1e-9) or np.any(p > 1 + 1e-9):
            bad = True
    return Fc, p, err, bad


def auto_steps(present_time, lam, mu, delta, Gamma, M=16):
    """Mirror of the host rule in csrc/gcdyn_b200.cu: h · max_i(|λ_i| + |μ| + 2|γ_i|) ≤ target(M)."""
    target = {8: 0.055, 12: 0.20, 16: 0.36, 20: 0.50, 24: 0.62}[M]
    g = np.abs(delta * np.diag(np.asarray(Gamma, float)))
    rate = np.max(np.abs(lam) + abs(mu) + 2.0 * g)
    return int(min(8192, max(4, np.ceil(present_time * rate / target))))


def build_items(flat, present_time):
    """Lookup items of the regrouped sum.  Per tree, for each live node v with
    branch type x = btype(v), own type y = type(v), time t:
        birth            -> (t, x, -1), (t, y, +nlive_children)   [merged to (t, x, +1) when y == x]
        type_change      -> (t, x, -1), (t, y, +1)
        sampled_death    -> (t, x, -1)
        sampled_survival -> nothing (τ = 0, F(0) = 0)
    plus one root item (root_time, root_type, +1).
    Returns dict of arrays: item_off[T+1], time, type, w, and per-tree node-term bookkeeping."""
    T = flat.n_trees
    K = flat.n_types
    item_off = [0]
    tm, ty, w = [], [], []
    nb = np.zeros((T, K), dtype=np.int64)      # births by branch type  -> log λ_x
    ntc = np.zeros((T, K * K), dtype=np.int64)  # type changes x->y       -> log(δ Γ[x,y])
    nd = np.zeros(T, dtype=np.int64)
    ns = np.zeros(T, dtype=np.int64)
    for t in range(T):
        lo, hi = int(flat.tree_off[t]), int(flat.tree_off[t + 1])
        for i in range(lo, hi):
            if not flat.live[i]:
                continue
            ev, x, y, te = int(flat.event[i]), int(flat.btype_idx[i]), int(flat.type_idx[i]), float(flat.t_end[i])
            if ev == EV_BIRTH:
                nb[t, x] += 1
                if x == y:
                    tm.append(te); ty.append(x); w.append(1.0)
                else:
                    tm.append(te); ty.append(x); w.append(-1.0)
                    tm.append(te); ty.append(y); w.append(2.0)
            elif ev == EV_TC:
                ntc[t, x * K + y] += 1
                tm.append(te); ty.append(x); w.append(-1.0)
                tm.append(te); ty.append(y); w.append(1.0)
            elif ev == EV_SDEATH:
                nd[t] += 1
                tm.append(te); ty.append(x); w.append(-1.0)
            elif ev == EV_SSURV:
                ns[t] += 1
        tm.append(float(flat.root_time[t])); ty.append(int(flat.root_type_idx[t])); w.append(1.0)
        item_off.append(len(tm))
    return dict(item_off=np.asarray(item_off), time=np.asarray(tm), type=np.asarray(ty, dtype=np.int64),
                w=np.asarray(w), nb=nb, ntc=ntc, nd=nd, ns=ns)


def _xlogy_count(n, logv):
    """n * logv with 0 * (-inf) = 0."""
    n = np.asarray(n, dtype=float)
    with np.errstate(invalid="ignore"):
        out = n * logv
    return np.where(n == 0, 0.0, out)


def loglik_per_tree(flat, items, lam, mu, delta, Gamma, rho, sigma, present_time, S, M):
    """Per-tree log-likelihoods [T] for ONE parameter set, the way the sweep kernel forms them."""
    K = flat.n_types
    T = float(present_time)
    Fc, pT, err, bad = taylor_table(lam, mu, delta, Gamma, rho, sigma, T, S, M)
    with np.errstate(divide="ignore", invalid="ignore"):
        loglam = np.log(np.asarray(lam, float))
        logG = np.log(delta * np.asarray(Gamma, float)).reshape(-1)   # row-major [x*K+y]
        logmusig = np.log(mu) + np.log(sigma)
        logrho = np.log(rho)
        logcond = np.log1p(-pT)
    tau = T - items["time"]
    u = tau * (S / T)
    s = np.minimum(u.astype(np.int64), S - 1)
    th = u - s
    acc = np.zeros(len(tau))
    for n in range(M + 1, -1, -1):
        acc = acc * th + Fc[s, n, items["type"]]
    contrib = items["w"] * acc
    out = np.add.reduceat(contrib, items["item_off"][:-1])
    out = out + _xlogy_count(items["nb"], loglam).sum(axis=1)
    out = out + _xlogy_count(items["ntc"], logG).sum(axis=1)
    out = out + _xlogy_count(items["nd"], logmusig) + _xlogy_count(items["ns"], logrho)
    out = out - logcond[flat.root_type_idx]
    out = np.where(flat.self_loop != 0, -np.inf, out)
    if bad:
        out[:] = -np.inf
    return out, err


def moments(items, present_time, S, M, K):
    """Pset-independent sufficient statistics of the forest: Mom[s, n, x] = Σ_items w θ^n."""
    T = float(present_time)
    tau = T - items["time"]
    u = tau * (S / T)
    s = np.minimum(u.astype(np.int64), S - 1)
    th = u - s
    Mom = np.zeros((S, M + 2, K))
    pw = items["w"].copy()
    for n in range(M + 2):
        np.add.at(Mom, (s, n, items["type"]), pw)
        pw = pw * th
    return Mom


# ------------------------------------------------------------------ Chebyshev-moment (GEMM) formulation

def cheb_tree_moments(flat, items, present_time, D):
    """Pset-independent per-tree statistics: Mt[t, d, x] = Σ_{items of tree t with type x} w · T_d(u),
    u = 2τ/T − 1, d = 0..D (Chebyshev polynomials by the three-term recurrence)."""
    T = float(present_time)
    nT, K = flat.n_trees, flat.n_types
    u = 2.0 * (T - items["time"]) / T - 1.0
    tree_of = np.repeat(np.arange(nT), np.diff(items["item_off"]))
    Mt = np.zeros((nT, D + 1, K))
    Tm, Tc = np.ones_like(u), u.copy()
    for d in range(D + 1):
        cur = Tm if d == 0 else Tc
        np.add.at(Mt, (tree_of, d, items["type"]), items["w"] * cur)
        if d >= 1:
            Tm, Tc = Tc, 2.0 * u * Tc - Tm
    return Mt


def cheb_coeffs_from_table(Fc, present_time, D):
    """Chebyshev interpolation coefficients a[d, x] of F_x on [0, T] from its piecewise Taylor table:
    values at the D+1 Chebyshev–Gauss–Lobatto nodes, then a DCT-I.  Returns (a, tail) where
    tail = max_x Σ of the last three |a|."""
    S, L, K = Fc.shape
    T = float(present_time)
    k = np.arange(D + 1)
    tau = T * (1.0 + np.cos(np.pi * k / D)) / 2.0
    uu = tau * (S / T)
    s = np.minimum(uu.astype(np.int64), S - 1)
    th = uu - s
    V = np.zeros((D + 1, K))
    for n in range(L - 1, -1, -1):
        V = V * th[:, None] + Fc[s, n, :]
    wk = np.ones(D + 1); wk[0] = wk[-1] = 0.5
    a = np.zeros((D + 1, K))
    for d in range(D + 1):
        a[d] = (2.0 / D) * ((wk * np.cos(np.pi * k * d / D))[:, None] * V).sum(axis=0)
    a[0] *= 0.5
    a[-1] *= 0.5
    tail = np.abs(a[-3:]).sum(axis=0).max()
    return a, tail


def loglik_per_tree_cheb(flat, items, Mt, lam, mu, delta, Gamma, rho, sigma, present_time, S, M, D):
    """Per-tree log-likelihoods as the GEMM path forms them: Σ_{d,x} a[d,x]·Mt[t,d,x] + count terms."""
    Fc, pT, err, bad = taylor_table(lam, mu, delta, Gamma, rho, sigma, present_time, S, M)
    a, tail = cheb_coeffs_from_table(Fc, present_time, D)
    with np.errstate(divide="ignore", invalid="ignore"):
        loglam = np.log(np.asarray(lam, float))
        logG = np.log(delta * np.asarray(Gamma, float)).reshape(-1)
        logmusig = np.log(mu) + np.log(sigma)
        logrho = np.log(rho)
        logcond = np.log1p(-pT)
    out = np.einsum("tdx,dx->t", Mt[:, :D + 1, :], a)
    out = out + _xlogy_count(items["nb"], loglam).sum(axis=1) + _xlogy_count(items["ntc"], logG).sum(axis=1)
    out = out + _xlogy_count(items["nd"], logmusig) + _xlogy_count(items["ns"], logrho)
    out = out - logcond[flat.root_type_idx]
    out = np.where(flat.self_loop != 0, -np.inf, out)
    return out, tail
