"""Timing of the K > 64 layout of the trajectory kernel (one integrating warp, several types per lane, δΓ in shared memory) next
to the K <= 64 layouts: `python tools/large_k_bench.py` prints one JSON line per K (256 parameter sets, 300 simulated trees,
present time 3).  The tests cover these sizes for parity (tests/test_gpu_parity.py); this is the missing timing."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gcdyn.jl_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import gcdyn, oracle, workloads
from gcdyn.likelihood import DeviceParams, evaluate_resident, pack_params

P, T, NT = 256, 3.0, 300
ctx = gcdyn.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K in ([int(a) for a in sys.argv[1:]] or [20, 50, 64, 65, 96, 128]):
    ts = np.linspace(-2.0, 2.0, K)
    base = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    flat = gcdyn.simulate_flat(base, T, ts[K // 2 - 1], NT, seed=40 + K)
    rng = np.random.default_rng([K, 777])
    models = []
    for _ in range(P):
        f = np.exp(rng.normal(0.0, 0.2, 5))
        models.append(gcdyn.SigmoidalBranchingProcess(base.λ_xscale * f[0], base.λ_xshift + rng.normal(0.0, 0.2), base.λ_yscale * f[1],
                                                      base.λ_yshift * f[2], base.μ * f[3], base.δ * f[4], base.Γ, base.ρ, base.σ, ts))
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    dpar = DeviceParams(pack_params(models), ctx)
    Tn = flat.n_trees
    out = torch.zeros(P, dtype=torch.float64, device="cuda")
    out_tree = torch.zeros(P * Tn, dtype=torch.float64, device="cuda")
    step = lambda: evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        flush.zero_(); a.record(stream); step(); b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    ctx.set_profiling(True)
    acc = {}
    for _ in range(5):
        flush.zero_(); step()
        for k, v in ctx.last_kernel_ms().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    ctx.set_profiling(False)
    torch.cuda.synchronize()
    host = out_tree.cpu().numpy().reshape(P, Tn)
    want = oracle.loglik_flat_shared(flat, workloads.par_dict(models[:1]), 0, T)
    rel = float(abs(host[0].sum() - want.sum()) / abs(want.sum()))
    print(json.dumps({"K": K, "trees": Tn, "nodes": int(flat.n_nodes), "psets": P, "ms_per_step": ms,
                      "kernel_ms": {k: acc[k] for k in ("trajectory", "gemm", "sweep", "finish")}, "chebyshev": ctx.last_cheb(),
                      "cells_max": ctx.last_diag(P)["steps"], "rel_err_sum_pset0_vs_oracle": rel}), flush=True)
    forest.close()
