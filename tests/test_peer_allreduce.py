"""Peer-memory all-reduce of the [P] per-pset sums (gcdyn_comm_*).
World 1 runs on any GPU box; the 2-rank case needs two GPUs with P2P access and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path[:0] = [os.path.join("@ROOT@", "gcdyn.jl_b200")]
import numpy as np, torch, torch.distributed as dist
import gcdyn
from gcdyn.sharding import PeerAllReduce
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = gcdyn.Context(local)
P = 300
peer = PeerAllReduce(ctx, P)
stream = torch.cuda.Stream()
rng = np.random.default_rng(5)
base = rng.standard_normal((world, 40, P))          # every rank draws the same numbers
for it in range(40):                                 # many calls: sequence numbers and the two parities
    x = torch.from_numpy(base[rank, it].copy()).cuda()
    torch.cuda.synchronize()
    peer.allreduce(x.data_ptr(), P, stream.cuda_stream)
    stream.synchronize()
    want = np.zeros(P)
    for r in range(world):
        want += base[r, it]                          # rank order, like the kernel
    got = x.cpu().numpy()
    assert np.array_equal(got, want), (rank, it, np.abs(got - want).max())
    gathered = [None] * world
    dist.all_gather_object(gathered, got.tobytes())
    assert all(g == gathered[0] for g in gathered)   # bit-identical on every rank
# the host-buffer entry point for a sharded forest: partial sums -> peer all-reduce -> pinned host memory
from gcdyn.likelihood import evaluate, pack_params
ts = np.linspace(-1.0, 1.0, 5)
models = [gcdyn.SigmoidalBranchingProcess(1.2, 0.0, 2.0 + 0.1 * i, 0.5, 1.0, 1.0, 0.6, 0.2, ts) for i in range(9)]
trees = gcdyn.rand_tree(models[0], 2.5, ts[2], 8, rng=np.random.default_rng(3))      # same forest on every rank
whole = gcdyn.Forest(trees, ts, 2.5, ctx=ctx)
shard = gcdyn.Forest(trees[rank::world], ts, 2.5, ctx=ctx)
full, _, _ = evaluate(whole, pack_params(models), 2.5)
for _ in range(3):
    got, _, st = evaluate(shard, pack_params(models), 2.5, comm=peer)
    assert not st.any() and np.max(np.abs(got - full) / np.abs(full)) <= 1e-12, (rank, got, full)
    gathered = [None] * world
    dist.all_gather_object(gathered, got.tobytes())
    assert all(g == gathered[0] for g in gathered)
# sharded device-resident Metropolis (proposal, likelihood, peer all-reduce, accept/reject inside one captured CUDA graph per
# iteration): every rank walks the chain of the unsharded sampler
from gcdyn.mcmc import device_metropolis, theta_of
tpl = models[0]
th = theta_of(tpl)
th0 = th + 0.03 * np.random.default_rng(7).standard_normal((6, 6))
kw = dict(step=0.04, seed=11, lower=th - 0.6, upper=th + 0.6)
ref = device_metropolis(whole, tpl, 2.5, th0, 15, **kw)
sh = device_metropolis(shard, tpl, 2.5, th0, 15, comm=peer, **kw)
assert np.array_equal(sh.accepted, ref.accepted), (rank, sh.accepted, ref.accepted)
assert np.max(np.abs(sh.chain - ref.chain)) <= 1e-12 and np.max(np.abs(sh.loglik - ref.loglik) / np.abs(ref.loglik)) <= 1e-12
gathered = [None] * world
dist.all_gather_object(gathered, sh.chain.tobytes() + sh.loglik.tobytes())
assert all(g == gathered[0] for g in gathered)       # bit-identical chains on every rank
peer.close()
dist.barrier()
sys.stdout.write("rank%d-ok\n" % rank)
"""

# A peer that never arrives: the waiting rank's kernel gives up after GCDYN_PEER_TIMEOUT_MS, the call returns an error
# and NaN sums instead of hanging inside cudaStreamSynchronize.
TIMEOUT_WORKER = r"""
import os, sys, time
os.environ["GCDYN_PEER_TIMEOUT_MS"] = "400"
sys.path[:0] = [os.path.join("@ROOT@", "gcdyn.jl_b200")]
import numpy as np, torch, torch.distributed as dist
import gcdyn
from gcdyn.sharding import PeerAllReduce
from gcdyn.likelihood import evaluate, pack_params, GcdynError
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = gcdyn.Context(local)
peer = PeerAllReduce(ctx, 16)
ts = np.linspace(-1.0, 1.0, 5)
models = [gcdyn.SigmoidalBranchingProcess(1.2, 0.0, 2.0 + 0.1 * i, 0.5, 1.0, 1.0, 0.6, 0.2, ts) for i in range(9)]
trees = gcdyn.rand_tree(models[0], 2.5, ts[2], 8, rng=np.random.default_rng(3))
shard = gcdyn.Forest(trees[rank::world], ts, 2.5, ctx=ctx)
if rank == 0:
    t0 = time.time()
    try:
        evaluate(shard, pack_params(models), 2.5, comm=peer)
        raise SystemExit("the call returned although rank 1 never arrived")
    except GcdynError as e:
        assert "timed out" in str(e), str(e)
    assert time.time() - t0 < 30.0
dist.barrier()
sys.stdout.write("rank%d-ok\n" % rank)
os._exit(0)                                          # the windows are out of step now: skip the orderly teardown
"""

# Trajectory sharding: with a row window every rank integrates a slice of the parameter sets and the coefficient rows are
# all-gathered over peer memory; values must equal the unsplit evaluation, including a parameter set that the calibrated grid
# does not resolve (its exact fallback needs the Taylor table on EVERY rank: the second, "missing" trajectory launch).
ROWS_WORKER = r"""
import os, sys
os.environ["GCDYN_TRAJ_SHARD_MIN_PSETS"] = "8"
sys.path[:0] = [os.path.join("@ROOT@", "gcdyn.jl_b200")]
import numpy as np, torch, torch.distributed as dist
import gcdyn
from gcdyn.sharding import PeerAllReduce
from gcdyn.likelihood import evaluate, pack_params, _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = gcdyn.Context(local)
P, K = 27, 5                                         # an odd count: uneven slices
peer = PeerAllReduce(ctx, P, row_bytes=PeerAllReduce.row_window_bytes(P, K))
plain = PeerAllReduce(ctx, P)                        # no row window: every rank integrates everything
ts = np.linspace(-1.0, 1.0, K)
mk = lambda i, ys=2.0: gcdyn.SigmoidalBranchingProcess(1.2, 0.0, ys + 0.02 * i, 0.5, 1.0, 1.0, 0.6, 0.2, ts)
models = [mk(i) for i in range(P)]
trees = gcdyn.rand_tree(models[0], 2.5, ts[2], 40, rng=np.random.default_rng(3))     # same forest on every rank
whole = gcdyn.Forest(trees, ts, 2.5, ctx=ctx)
shard = gcdyn.Forest(trees[rank::world], ts, 2.5, ctx=ctx)
shard2 = gcdyn.Forest(trees[rank::world], ts, 2.5, ctx=ctx)
def check(ms, expect_fallback):
    full, _, _ = evaluate(whole, pack_params(ms), 2.5)
    for _ in range(3):                               # both halves of the row window
        got, tp, st = evaluate(shard, pack_params(ms), 2.5, comm=peer, per_tree=True)
        assert _lib().gcdyn_last_traj_sharded(ctx._h) == (1 if world > 1 else 0)
        d = ctx.last_diag(P)
        assert (d["cells"] > 0).all(), d["cells"]    # the metadata of the other ranks' parameter sets arrived
        assert (ctx.last_cheb()["n_fallback"] > 0) == expect_fallback, ctx.last_cheb()
        assert not st.any() and np.max(np.abs(got - full) / np.abs(full)) <= 1e-12, (rank, got, full)
        ref, tpr, _ = evaluate(shard2, pack_params(ms), 2.5, comm=plain, per_tree=True)
        assert _lib().gcdyn_last_traj_sharded(ctx._h) == 0
        assert np.array_equal(tp, tpr) and np.array_equal(got, ref)       # bit-identical to the unsplit sharded evaluation
        gathered = [None] * world
        dist.all_gather_object(gathered, got.tobytes())
        assert all(g == gathered[0] for g in gathered)
check(models, False)
# two much faster parameter sets, one in each rank's slice: the grid calibrated above does not resolve them
stiff = list(models)
stiff[2] = mk(2, 12.0)
stiff[P - 3] = mk(P - 3, 12.5)
check(stiff, True)
assert peer.error() == 0
peer.close(); plain.close()
dist.barrier()
sys.stdout.write("rank%d-ok\n" % rank)
"""


def _run(world, worker=None):
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write((worker or WORKER).replace("@ROOT@", ROOT))
        path = f.name
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), path]
    try:
        r = subprocess.run(cmd, env=dict(os.environ), capture_output=True, text=True, timeout=100)
    finally:
        os.unlink(path)
    # every assertion lives in the workers; torchrun exits non-zero if any rank fails
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("-ok") == world, r.stdout


@pytest.mark.gpu
def test_peer_allreduce_single_rank():
    _run(1)


@pytest.mark.gpu
def test_peer_allreduce_two_ranks_bit_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2)


@pytest.mark.gpu
def test_peer_allreduce_wait_is_bounded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, TIMEOUT_WORKER)


@pytest.mark.gpu
def test_trajectory_sharding_matches_the_unsplit_evaluation():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, ROWS_WORKER)
