"""(c_wait: per cell that waits for a ring slot, i.e. cells 7.., c_wait0: loop back-edge of the first six cells.)
Phase timing of traj_kernel (clock64 stamps per pset, dumped through gcdyn_trace_dump).  Needs the trace build:
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp -shared -lgomp \
       -DGCDYN_TRAJ_TRACE -o tools/_trace/libgcdyn_trace.so gcdyn.jl_b200/csrc/gcdyn_b200.cu
  GCDYN_B200_LIB=tools/_trace/libgcdyn_trace.so python tools/trace_probe.py
The stamps perturb the kernel (≈1.5× slower); use the numbers relative to each other."""
import os, sys, ctypes as C
sys.path[:0] = ["/root/repo", "/root/repo/gcdyn.jl_b200"]
import numpy as np, torch, gcdyn, workloads
from gcdyn.likelihood import evaluate, pack_params, _lib
P = 256
flat = workloads.flat_forest("C3"); models = workloads.jittered_models("C3", P)
params = pack_params(models); ctx = gcdyn.Context(0); forest = gcdyn.Forest(flat, None, 4.0, ctx=ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    if not os.environ.get("NOFLUSH"): flush.zero_()
    evaluate(forest, params, 4.0, want_status=False, order=int(os.environ.get("ORDER", "0")))
    torch.cuda.synchronize()
lib = _lib()
lib.gcdyn_trace_dump.argtypes = [C.POINTER(C.c_longlong)]
buf = np.zeros(1024 * 24, dtype=np.int64)
print("rc", lib.gcdyn_trace_dump(buf.ctypes.data_as(C.POINTER(C.c_longlong))))
t = buf.reshape(1024, 24)[:P]
g0 = t[:, 14].min()
rows = []
for p in range(P):
    r = t[p]
    rows.append(dict(p=p, sm=int(r[16]), S=int(r[18]), start_ns=int(r[14] - g0), end_ns=int(r[15] - g0), setup=int(r[2] - r[0]),
                     integ=int(r[3] - r[2]), per_cell=int((r[3] - r[2]) / max(1, r[18])), cons_after=int(r[6] - r[3]),
                     sync=int(r[4] - max(r[3], r[6])), finish=int(r[5] - r[4]), total=int(r[5] - r[0]),
                     c_wait=int(r[7] / max(1, r[18] - 6)), c_wait0=int(r[10] / min(6, max(1, r[18]))), c_stage=int(r[8] / max(1, r[18])), c_epi=int(r[9] / max(1, r[18]))))
rows.sort(key=lambda r: -r["end_ns"])
print("kernel span ns:", rows[0]["end_ns"], " start spread ns:", max(r["start_ns"] for r in rows))
for r in rows[:6] + rows[-3:]:
    print(r)
for key in ("setup", "integ", "per_cell", "cons_after", "sync", "finish", "total", "S", "c_wait", "c_wait0", "c_stage", "c_epi"):
    v = np.array([r[key] for r in rows]); print(f"{key:10s} mean {v.mean():9.0f} min {v.min():8d} max {v.max():8d}")
print("cheb", ctx.last_cheb(), "diag", {k: v for k, v in ctx.last_diag(4).items() if k != "err_indicator"})
import collections
by_sm = collections.defaultdict(list)
for r in rows: by_sm[r["sm"]].append(r)
alone = [r["c_stage"] for r in rows if len(by_sm[r["sm"]]) == 1]; shared = [r["c_stage"] for r in rows if len(by_sm[r["sm"]]) > 1]
print("c_stage alone on its SM: n=%d mean %.0f | sharing an SM: n=%d mean %.0f" % (len(alone), np.mean(alone) if alone else 0, len(shared), np.mean(shared) if shared else 0))
print("CTAs per SM histogram", collections.Counter(len(v) for v in by_sm.values()))
