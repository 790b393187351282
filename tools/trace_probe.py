"""Phase timing of traj_kernel (clock64 stamps per pset, dumped through gcdyn_trace_dump).  Needs the trace build:
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp -shared -lgomp \
       -DGCDYN_TRAJ_TRACE -o tools/_trace/libgcdyn_trace.so gcdyn.jl_b200/csrc/gcdyn_b200.cu
  GCDYN_B200_LIB=tools/_trace/libgcdyn_trace.so python tools/trace_probe.py
The stamps perturb the kernel (≈1.5× slower); use the numbers relative to each other."""
import sys, ctypes as C
sys.path[:0] = ["/root/repo", "/root/repo/gcdyn.jl_b200"]
import numpy as np, torch, gcdyn, workloads
from gcdyn.likelihood import evaluate, pack_params, _lib
P = 256
flat = workloads.flat_forest("C3"); models = workloads.jittered_models("C3", P)
params = pack_params(models); ctx = gcdyn.Context(0); forest = gcdyn.Forest(flat, None, 4.0, ctx=ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    flush.zero_()
    evaluate(forest, params, 4.0, want_status=False)
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
    rows.append(dict(p=p, sm=int(r[16]), pw=int(r[17]), S=int(r[18]), att=int(r[19]), wid0=int(r[20]),
                     start_ns=int(r[14] - g0), end_ns=int(r[15] - g0), setup=int(r[2] - r[0]), integ=int(r[3] - r[2]),
                     per_cell=int((r[3] - r[2]) / max(1, r[19])), wait=int(r[4] - r[3]), fold=int(r[7] - r[6]), dct=int(r[8] - r[7]),
                     deff=int(r[9] - r[8]), write=int(r[10] - r[9]), later_cells=(int(r[1] / max(1, r[18] - 1)), int(r[5] / max(1, r[18] - 1)), int(r[11] / max(1, r[18] - 1)))))
import collections
by_sm = collections.defaultdict(list)
for r in rows: by_sm[r["sm"]].append(r)
print("SMs used", len(by_sm), "CTAs/SM histogram", collections.Counter(len(v) for v in by_sm.values()))
same = sum(1 for v in by_sm.values() if len(v) > 1 and len({(r["wid0"] + r["pw"]) & 3 for r in v}) < len(v))
print("SMs whose CTAs integrate on the same sub-partition:", same)
rows.sort(key=lambda r: -r["end_ns"])
print("kernel span ns:", rows[0]["end_ns"], " start spread ns:", max(r["start_ns"] for r in rows))
for r in rows[:6] + rows[-3:]:
    print(r)
for key in ("setup", "integ", "per_cell", "wait", "fold", "dct", "deff", "write"):
    v = np.array([r[key] for r in rows]); print(f"{key:9s} mean {v.mean():9.0f} min {v.min():8d} max {v.max():8d}")
alone = [r["per_cell"] for r in rows if len(by_sm[r["sm"]]) == 1]; shared = [r["per_cell"] for r in rows if len(by_sm[r["sm"]]) > 1]
print("per_cell alone", np.mean(alone) if alone else None, "shared", np.mean(shared) if shared else None)
