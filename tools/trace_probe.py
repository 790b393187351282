"""Phase timing of traj_kernel (clock64 stamps, -DGCDYN_TRAJ_TRACE build): GCDYN_B200_LIB=tools/_trace/libgcdyn_trace.so python tools/trace_probe.py"""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/gcdyn.jl_b200"]
import numpy as np, torch, gcdyn, workloads
from gcdyn.likelihood import evaluate, pack_params
flat = workloads.flat_forest("C3"); models = workloads.jittered_models("C3", 256)
params = pack_params(models); ctx = gcdyn.Context(0); forest = gcdyn.Forest(flat, None, 4.0, ctx=ctx)
for _ in range(3):
    evaluate(forest, params, 4.0, want_status=False)
    torch.cuda.synchronize()
    print("----", flush=True)
