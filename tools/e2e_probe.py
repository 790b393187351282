import sys, time, ctypes as C
sys.path[:0]=["/root/repo","/root/repo/gcdyn.jl_b200"]
import numpy as np, torch, gcdyn, workloads
from gcdyn.likelihood import evaluate, pack_params, _lib, Opts, _ptr, _p_f64, _p_i32
flat = workloads.flat_forest("C3"); models = workloads.jittered_models("C3", 256)
params = pack_params(models); ctx = gcdyn.Context(0); forest = gcdyn.Forest(flat, None, 4.0, ctx=ctx)
for _ in range(5): evaluate(forest, params, 4.0, want_status=False)
N=300
t0=time.perf_counter()
for _ in range(N): evaluate(forest, params, 4.0, want_status=False)
t1=time.perf_counter(); print("python evaluate(): %.1f us/call" % (1e6*(t1-t0)/N))
lib=_lib(); ps=params.struct(); op=Opts(0,0,0.0,0,0); out=np.empty(256)
t0=time.perf_counter()
for _ in range(N): lib.gcdyn_loglik_batch(forest.ctx._h, forest._h, C.byref(ps), 4.0, 0, C.byref(op), _ptr(out,_p_f64), None, None)
t1=time.perf_counter(); print("raw ctypes gcdyn_loglik_batch: %.1f us/call" % (1e6*(t1-t0)/N))
