import os, sys, time
sys.path[:0] = ["/root/repo", "/root/repo/gcdyn.jl_b200"]
import numpy as np, torch, gcdyn, workloads
from gcdyn.likelihood import evaluate, pack_params
P = 256
flat = workloads.flat_forest("C3"); models = workloads.jittered_models("C3", P)
params = pack_params(models); ctx = gcdyn.Context(0); forest = gcdyn.Forest(flat, None, 4.0, ctx=ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    evaluate(forest, params, 4.0, want_status=False)
for mode in ("flush", "noflush"):
    tot = 0.0
    for i in range(30):
        if mode == "flush": flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        evaluate(forest, params, 4.0, want_status=False)
        tot += time.perf_counter() - t0
    print(mode, "python-side e2e us per call", 1e6 * tot / 30, flush=True)
