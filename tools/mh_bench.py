"""Config 4 timing: 64-chain random-walk Metropolis on the config-3 forest (1000 trees per GPU, K=20).
Run as `python tools/mh_bench.py [iterations]` (1 GPU) or under torchrun with trees sharded per rank."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gcdyn.jl_b200")]
import numpy as np
import torch
import gcdyn, workloads
from gcdyn.mcmc import random_walk_metropolis, device_metropolis, theta_of
from gcdyn.sharding import allreduce_sum, PeerAllReduce

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
base, _ = workloads.base_model("C3")
T = workloads.CONFIGS["C3"]["T"]
flat = workloads.flat_forest("C3", rank * 1000, 1000)
ctx = gcdyn.Context(local)
forest = gcdyn.Forest(flat, None, T, ctx=ctx)
rng = np.random.default_rng(4)
th0 = theta_of(base) + 0.05 * rng.standard_normal((64, 6))
box = 1.5
ar = (lambda a: allreduce_sum(a)) if world > 1 else None
random_walk_metropolis(forest, base, T, th0, 5, step=0.02, seed=4, lower=theta_of(base) - box, upper=theta_of(base) + box, allreduce=ar)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = random_walk_metropolis(forest, base, T, th0, iters, step=0.02, seed=4, lower=theta_of(base) - box, upper=theta_of(base) + box, allreduce=ar)
dt = time.perf_counter() - t0
# device-resident sampler: proposals, likelihood, peer all-reduce and accept/reject without host round trips
comm = PeerAllReduce(ctx, 64) if world > 1 else None
lo, up = theta_of(base) - box, theta_of(base) + box
device_metropolis(forest, base, T, th0, 5, step=0.02, seed=4, lower=lo, upper=up, comm=comm, history=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
dres = device_metropolis(forest, base, T, th0, iters, step=0.02, seed=4, lower=lo, upper=up, comm=comm, history=False)
ddt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": "C4 device-resident (gcdyn_mh_*): 64 chains x %d iterations, %d trees on %d GPU(s)" % (iters, 1000 * world, world),
                      "iterations_per_s": iters / ddt, "ms_per_iteration": 1e3 * ddt / iters,
                      "tree_pset_evals_per_s": iters * 64 * 1000 * world / ddt,
                      "acceptance_rate_mean": float(dres.accepted.mean() / iters),
                      "loglik_last": float(dres.loglik[-1].mean())}))
if rank == 0:
    print(json.dumps({"config": "C4: 64 chains x %d iterations, %d trees on %d GPU(s)" % (iters, 1000 * world, world),
                      "iterations_per_s": iters / dt, "ms_per_iteration": 1e3 * dt / iters,
                      "tree_pset_evals_per_s": iters * 64 * 1000 * world / dt,
                      "acceptance_rate_mean": float(res.acceptance_rate.mean()),
                      "loglik_first_last": [float(res.loglik[0].mean()), float(res.loglik[-1].mean())]}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
