"""World-size-2 test of the multi-rank host logic on CPU (gloo): deterministic tree partition, per-rank
partial [P] sums, ONE all-reduce.  The per-tree values stand in for what each rank's GPU would produce
(they are the golden per-tree log-likelihoods), so this checks the plumbing, not the kernels."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path[:0] = [r"{root}", r"{root}/gcdyn.jl_b200", r"{root}/tests"]
    from gcdyn.sharding import partition_trees, subset_flat, allreduce_sum
    from helpers import load_golden
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    flat, params, T, exp, par = load_golden("g5_batched_discrete")
    per_tree = exp["ode_shared"]                      # [P, T] stand-in for this rank's GPU output
    parts = partition_trees(np.diff(flat.tree_off), world)
    mine = parts[rank]
    shard = subset_flat(flat, mine)                   # what this rank would upload
    assert shard.n_trees == len(mine) and shard.n_nodes == int(np.diff(flat.tree_off)[mine].sum())
    partial = per_tree[:, mine].sum(axis=1)
    total = allreduce_sum(partial)
    want = per_tree.sum(axis=1)
    assert np.allclose(total, want, rtol=1e-13, atol=0), (total, want)
    # every rank computed the same partition, and together they cover every tree exactly once
    sizes = [None, None]
    dist.all_gather_object(sizes, mine.tolist())
    assert sorted(sizes[0] + sizes[1]) == list(range(flat.n_trees))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_sharded_sum_with_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=_free_port()))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
