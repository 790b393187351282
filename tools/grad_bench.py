"""Exact gradients (gcdyn_loglik_grad_batch) at config-3 size: cost next to a plain evaluation, accuracy against the batched
central differences of the plain path.  `python tools/grad_bench.py`"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gcdyn.jl_b200")]
import numpy as np
import torch
import gcdyn, workloads
from gcdyn.likelihood import evaluate, pack_params, loglikelihood_and_gradient, FLAG_TOTALS

P = 256
flat = workloads.flat_forest("C3")
models = workloads.jittered_models("C3", P)
T = workloads.CONFIGS["C3"]["T"]
ctx = gcdyn.Context(0)
forest = gcdyn.Forest(flat, None, T, ctx=ctx)
ps = pack_params(models, response="sigmoid")
for _ in range(3):
    ll, g = loglikelihood_and_gradient(forest, ps, T)
    tot, _, _ = evaluate(forest, ps, T, want_status=False, flags=FLAG_TOTALS)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    ll, g = loglikelihood_and_gradient(forest, ps, T)
t_grad = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for _ in range(n):
    tot, _, _ = evaluate(forest, ps, T, want_status=False, flags=FLAG_TOTALS)
t_tot = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for _ in range(n):
    full, _, _ = evaluate(forest, ps, T, want_status=False)
t_full = (time.perf_counter() - t0) / n
# central differences of the plain path for 4 parameter sets
names = ("xscale", "xshift", "yscale", "yshift", "mu", "delta")
worst = 0.0
for p in (0, 85, 170, 255):
    m = models[p]
    base = np.array([m.λ_xscale, m.λ_xshift, m.λ_yscale, m.λ_yshift, m.μ, m.δ])
    pts = []
    for k in range(6):
        for sgn in (+1, -1):
            for hh in (1e-3, 2e-3):
                q = base.copy(); q[k] += sgn * hh * max(1.0, abs(base[k])); pts.append(q)
    ms = [gcdyn.SigmoidalBranchingProcess(q[0], q[1], q[2], q[3], q[4], q[5], m.Γ, m.ρ, m.σ, m.type_space) for q in pts]
    v, _, _ = evaluate(forest, pack_params(ms), T, want_status=False)
    v = v.reshape(6, 2, 2)
    for k in range(6):
        h1 = 1e-3 * max(1.0, abs(base[k]))
        d1 = (v[k, 0, 0] - v[k, 1, 0]) / (2 * h1)
        d2 = (v[k, 0, 1] - v[k, 1, 1]) / (4 * h1)
        want = (4 * d1 - d2) / 3
        worst = max(worst, abs(g[p, k] - want) / max(1.0, abs(want)))
print(json.dumps({"config": "C3: 1000 trees x 256 parameter sets, K=20, 6 directions (sigmoid response)",
                  "ms_gradient_call": 1e3 * t_grad, "ms_totals_call": 1e3 * t_tot, "ms_full_call": 1e3 * t_full,
                  "gradient_over_totals": t_grad / t_tot, "gradient_over_full": t_grad / t_full,
                  "max_rel_err_vs_richardson_differences": worst,
                  "ll_vs_totals_max_rel": float(np.max(np.abs(ll - tot) / np.abs(tot)))}))
