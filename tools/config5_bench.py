"""BASELINE config 5 at scale: 10^4 trees of ~10^3 nodes (K = 50 types), 1024 parameter sets.
One GPU: `python tools/config5_bench.py [n_trees] [P]`.  Several GPUs (torchrun, one rank per GPU):
`... tools/config5_bench.py n_trees P strong|weak` — strong: the n_trees are split over the ranks, weak: n_trees per
rank; every rank simulates its own trees (seed keyed by rank); the [P] sums are combined by the peer all-reduce kernel.
Trees from the native simulator, kept when 500 <= nodes <= 2000.  Prints one JSON line with the
per-kernel breakdown (CUDA events inside the library), the GEMM's FP64 rate and a parity check against the oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gcdyn.jl_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import gcdyn, workloads, oracle
from gcdyn.likelihood import DeviceParams, evaluate_resident, pack_params
from gcdyn.sharding import subset_flat

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mode = sys.argv[3] if len(sys.argv) > 3 else "strong"
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
n_trees = n_total // world if mode == "strong" else n_total
K, T = 50, 5.0
ts = np.linspace(-2.0, 2.0, K)
base = gcdyn.SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
t0 = time.time()
pool = gcdyn.simulate_flat(base, T, ts[K // 2 - 1], 6 * n_trees, seed=5 + 1000 * rank)
sizes = np.diff(pool.tree_off)
ok = np.nonzero((sizes >= 500) & (sizes <= 2000))[0][:n_trees]
flat = subset_flat(pool, ok)
t_sim = time.time() - t0
rng = np.random.default_rng([5, 777])
models = []
for _ in range(P):
    f = np.exp(rng.normal(0.0, 0.2, 5))
    models.append(gcdyn.SigmoidalBranchingProcess(base.λ_xscale * f[0], base.λ_xshift + rng.normal(0.0, 0.2), base.λ_yscale * f[1],
                                                  base.λ_yshift * f[2], base.μ * f[3], base.δ * f[4], base.Γ, base.ρ, base.σ, ts))
params = pack_params(models)
ctx = gcdyn.Context(local)
forest = gcdyn.Forest(flat, None, T, ctx=ctx)
dpar = DeviceParams(params, ctx)
Tn = flat.n_trees
out = torch.zeros(P, dtype=torch.float64, device="cuda")
out_tree = torch.zeros(P * Tn, dtype=torch.float64, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
peer = None
if world > 1:
    from gcdyn.sharding import PeerAllReduce
    # the row window of the trajectory sharding (C5_TRAJ_SHARD=0: every rank integrates all P parameter sets)
    rows = PeerAllReduce.row_window_bytes(P, K) if os.environ.get("C5_TRAJ_SHARD", "1") != "0" else 0
    peer = PeerAllReduce(ctx, P, row_bytes=rows)


def step():
    # N > 1: the all-reduce of the [P] sums is fused into finish_kernel (gcdyn_loglik_batch_resident_sharded)
    evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream, comm=peer)


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
for _ in range(3):
    step()
sync_all()
steps = 10
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
# (N > 1: a peer barrier between the flush and the timed step, so that the ranks' flush kernels do not skew the step; C5_NO_ALIGN=1 = off)
align = torch.zeros(1, dtype=torch.float64, device="cuda") if (peer is not None and not os.environ.get("C5_NO_ALIGN")) else None
for a, b in ev:
    flush.zero_()
    if align is not None:
        peer.allreduce(align.data_ptr(), 1, stream.cuda_stream)
    a.record(stream); step(); b.record(stream)
sync_all()
ms = sum(a.elapsed_time(b) for a, b in ev) / steps
if world > 1:                                   # device time of the slowest rank
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
from gcdyn.likelihood import _lib
traj_sharded = bool(_lib().gcdyn_last_traj_sharded(ctx._h))
ctx.set_profiling(True)
acc_sh = {}
if world > 1:                                   # the sharded step's own breakdown ("trajectory" includes the row exchange)
    for _ in range(5):
        flush.zero_()
        step()
        for k, v in ctx.last_kernel_ms().items():
            acc_sh[k] = acc_sh.get(k, 0.0) + v / 5
    sync_all()
    assert peer.error() == 0
acc = {}
evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream)   # (allocates its own row buffer)
sync_all()
for _ in range(5):
    flush.zero_()
    evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream)
    for k, v in ctx.last_kernel_ms().items():
        acc[k] = acc.get(k, 0.0) + v / 5
sync_all()
cheb = ctx.last_cheb()
diag = ctx.last_diag(P)
J = 2 * K + 4 + (cheb["interp_degree"] + 1) * K
gemm_flop = 2.0 * Tn * P * J
peak = ctx.probe_fp64_peak()
host = out_tree.cpu().numpy().reshape(P, Tn)
par = workloads.par_dict(models[:1])
sub = subset_flat(flat, np.arange(min(Tn, 500)))
want = oracle.loglik_flat_shared(sub, par, 0, T)
rel_tree = float(np.max(np.abs(host[0, :len(want)] - want) / np.abs(want)))
# ... and a parameter set of the LAST rank's slice, from the sharded step (its coefficient row came over NVLink)
step(); sync_all()
host_sh = out_tree.cpu().numpy().reshape(P, Tn)
want_last = oracle.loglik_flat_shared(subset_flat(flat, np.arange(min(Tn, 100))), workloads.par_dict(models[P - 1:]), 0, T)
rel_last = float(np.max(np.abs(host_sh[P - 1, :len(want_last)] - want_last) / np.abs(want_last)))
same_as_unsharded = bool(np.array_equal(host_sh, host))
rel_sum = float(abs(host[0, :len(want)].sum() - want.sum()) / abs(want.sum()))
if rank == 0:
  print(json.dumps({"config": f"C5: {Tn} trees per GPU ({flat.n_nodes} nodes, {forest.info()['n_items']} items), K={K}, P={P}, T={T}, "
                            f"{world} GPU(s), {mode} scaling, {Tn * world} trees in total",
                  "sim_seconds": t_sim, "ms_per_step": ms, "tree_pset_evals_per_s": Tn * world * P / (ms * 1e-3),
                  "trajectory_sharded": traj_sharded, "aligned_start": align is not None, "kernel_ms_sharded_step": acc_sh or None,
                  "kernel_ms": acc, "chebyshev": cheb, "taylor_cells_max": diag["steps"],
                  "gemm": {"J": J, "gflop": gemm_flop / 1e9, "tflops": gemm_flop / (acc["gemm"] * 1e-3) / 1e12,
                           "frac_of_measured_dfma_peak": gemm_flop / (acc["gemm"] * 1e-3) / 1e12 / peak, "dfma_peak_tflops": peak},
                  "parity_first_500_trees_pset0": {"max_rel_err_tree": rel_tree, "rel_err_sum": rel_sum},
                  "parity_first_100_trees_last_pset_sharded_step": rel_last,
                  "sharded_step_per_tree_values_bit_identical_to_unsharded": same_as_unsharded}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
