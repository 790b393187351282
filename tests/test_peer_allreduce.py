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
sys.path[:0] = [os.path.join(%(root)r, "gcdyn.jl_b200")]
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
peer.close()
dist.barrier()
sys.stdout.write("rank%d-ok\n" % rank)
"""


def _run(world):
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(WORKER % {"root": ROOT})
        path = f.name
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), path]
    try:
        r = subprocess.run(cmd, env=dict(os.environ), capture_output=True, text=True, timeout=300)
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
