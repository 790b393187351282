"""The benchmark's synthetic workloads and the host-side sizing of the trajectory sharding (CPU only)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcdyn.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import workloads  # noqa: E402
from gcdyn.sharding import PeerAllReduce  # noqa: E402


def test_config5_forest_comes_from_the_native_simulator_and_is_reproducible():
    # config 5: trees of 500..2000 nodes, K = 50; one simulator stream per shard
    a = workloads.flat_forest("C5", 0, 6)
    b = workloads.flat_forest("C5", 0, 6)
    c = workloads.flat_forest("C5", 6, 6)          # the next shard of six trees
    sizes = np.diff(a.tree_off)
    assert a.n_trees == 6 and a.n_types == 50 and np.all((sizes >= 500) & (sizes <= 2000))
    assert np.array_equal(a.t_end, b.t_end) and np.array_equal(a.event, b.event)
    assert c.n_trees == 6 and not (c.n_nodes == a.n_nodes and np.array_equal(c.t_end, a.t_end))
    models = workloads.jittered_models("C5", 5)
    assert len(models) == 5 and len(models[0].type_space) == 50


def test_config4_is_the_config3_forest_with_64_chains():
    assert workloads.CONFIGS["C4"]["psets"] == 64 and workloads.CONFIGS["C4"]["trees"] == workloads.CONFIGS["C3"]["trees"]
    a, b = workloads.flat_forest("C4", 0, 3), workloads.flat_forest("C3", 0, 3)
    assert np.array_equal(a.t_end, b.t_end) and np.array_equal(a.tree_off, b.tree_off)


def test_row_window_size_follows_the_library_row_length():
    # Jst as gcdyn_b200.cu:ensure_moments computes it: C1 = 2K+4, J1 = round4(C1 + (D+1)K), Jst = round4(J1 + K^2)
    for K, D, P in ((20, 34, 256), (50, 42, 1024), (5, 128, 27), (128, 128, 512)):
        j1 = (2 * K + 4 + (D + 1) * K + 3) // 4 * 4
        jst = (j1 + K * K + 3) // 4 * 4
        assert PeerAllReduce.row_window_bytes(P, K, D) == 2 * P * (jst * 8 + 32)
    # the default covers the largest degree the library uses
    assert PeerAllReduce.row_window_bytes(64, 20) >= PeerAllReduce.row_window_bytes(64, 20, 96)


def test_bench_command_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--scaling"):
        assert flag in out.stdout


def test_reference_arm_prints_one_line_in_the_bench_schema():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line, same metric/unit/config keys,
    `impl: reference`, a `cpu_baseline` that describes the run and an `e2e` that repeats the line's own value."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--trees", "64", "--psets", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "tree_loglik_evals_per_sec" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and "workload" in d["config"]
