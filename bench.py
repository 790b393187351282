#!/usr/bin/env python
"""Benchmark of the hot path: batched multitype branching-process tree log-likelihood.

    python bench.py --gpus N --steps K --warmup W            # this build, one rank per GPU (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference algorithm on host cores

Default workload at every N: BASELINE.json config 3 — 1000 trees per GPU × 256 parameter sets, K = 20 types
(weak scaling: trees are sharded, the [P] partial sums are combined by ONE all-reduce fused into the last kernel).
`--workload C5` = config 5 at scale (10^4 trees of ~10^3 nodes from the native simulator × 1024 parameter sets, K = 50),
`--workload C4` = config 4 (64 device-resident Metropolis chains on the config-3 forest; a step = 100 iterations);
`--scaling strong` keeps the TOTAL tree count fixed and splits it over the ranks (with >= 512 parameter sets the
trajectories are sharded too: each rank integrates P/N parameter sets and pushes the coefficient rows to its peers).  A step is one
batched evaluation of all P parameter sets against the rank's trees, producing the per-(tree,pset)
matrix and the [P] sums.  `value` = (tree, pset) evaluations per second over the whole job, with the
forest and the parameters resident in HBM; `e2e` is the same through the host-buffer C-ABI call
(`gcdyn_loglik_batch`: parameters copied H2D from pinned staging, results D2H) every step.

The real Julia reference cannot run here (no Julia in the image); the CPU arm times `oracle/refcpu.c`,
a C restatement of the reference algorithm as written (one adaptive RK5(4) solve per node, reference
default tolerances), OpenMP over (tree, pset) pairs on all host cores.  It prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gcdyn.jl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOAD = "C3"
METRIC = "tree_loglik_evals_per_sec"
UNIT = "(tree,pset) evals/s"


def config_dict(name, trees_per_gpu, psets, n_gpus, extra=None):
    import workloads
    c = workloads.CONFIGS[name]
    what = {"C3": "BASELINE config 3", "C4": "BASELINE config 4 (Metropolis chains = parameter sets, config-3 forest)",
            "C5": "BASELINE config 5 (trees of 500..2000 nodes, native simulator)"}.get(name, name)
    d = {"workload": f"{what}: {trees_per_gpu} trees/GPU x {psets} parameter sets, K={c['K']} "
                     f"sigmoid birth response, rho=0.5, sigma=0.1, present_time={c['T']}, simulator seed {c['seed']}",
         "trees_per_gpu": trees_per_gpu, "trees_total": trees_per_gpu * n_gpus, "psets": psets, "K": c["K"],
         "present_time": c["T"], "parallelism": f"trees sharded over {n_gpus} GPU(s), one all-reduce of [P] doubles"
         if n_gpus > 1 else "single GPU",
         "l2": "L2 flushed (256 MiB write) before every timed step; inputs (forest+table < 126 MB) would otherwise stay in L2"}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------- CPU arm

def host_threads():
    """Host threads the CPU arm may use: the process's affinity set (torchrun pins OMP_NUM_THREADS to 1, which
    would otherwise silently turn the reference arm into a single-thread run)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(flat, par, T, n_trees, n_psets, nthreads=None):
    """Time oracle/refcpu.c (reference algorithm as written, reltol = abstol = 1e-3 like the reference's
    defaults) on the first `n_trees` trees × first `n_psets` parameter sets.  Returns (pairs/s, seconds)."""
    import refcpu
    from gcdyn.sharding import subset_flat
    sub = subset_flat(flat, np.arange(n_trees))
    sp = {k: (v[:n_psets] if k != "Gamma" else v[:n_psets]) for k, v in par.items()}
    t0 = time.perf_counter()
    refcpu.loglik(sub, sp, T, rtol=1e-3, atol=1e-3, nthreads=nthreads or host_threads())
    dt = time.perf_counter() - t0
    return n_trees * n_psets / dt, dt


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def cpu_baseline(flat, par, T, budget_s=12.0):
    cores = host_threads()
    n_trees = min(flat.n_trees, 64)
    n_psets = min(par["lam"].shape[0], 4)
    rate, dt = cpu_reference_rate(flat, par, T, n_trees, n_psets)
    # grow the sample towards the budget (bounded by the full workload)
    P, Tn = par["lam"].shape[0], flat.n_trees
    while dt < budget_s / 3 and (n_trees < Tn or n_psets < P):
        grow = min(8.0, max(2.0, 0.6 * budget_s / max(dt, 1e-3)))
        if n_trees < Tn:
            n_trees = int(min(Tn, n_trees * grow))
        else:
            n_psets = int(min(P, max(n_psets + 1, n_psets * grow)))
        rate, dt = cpu_reference_rate(flat, par, T, n_trees, n_psets)
    # the reference itself is serial (mbp.jl:258): one thread on a smaller sample of the same workload
    n1 = max(8, int(min(Tn, n_trees * n_psets / max(cores, 1) / 4)))
    rate1, dt1 = cpu_reference_rate(flat, par, T, min(n1, Tn), 1, nthreads=1)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": rate1,
            "sample": f"first {n_trees} trees x first {n_psets} parameter sets of the same workload, "
                      f"{dt:.1f} s on {cores} threads of {cpu_model()}, oracle/refcpu.c (C restatement of mbp.jl:110-259 as "
                      f"written: one adaptive RK5(4) solve per node, reltol=abstol=1e-3), OpenMP over (tree,pset) pairs; "
                      f"single_thread_value: first {min(n1, Tn)} trees x 1 parameter set, {dt1:.1f} s on one thread "
                      f"(the reference is serial); the real Julia reference cannot run in this image "
                      f"(bench/ref_julia.jl times it for readers who have Julia, unexecuted here)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import workloads
    c = workloads.CONFIGS[WORKLOAD]
    # (config 5: the CPU arm only ever touches a bounded sample, so it only generates one)
    flat = workloads.flat_forest(WORKLOAD, 0, min(args.trees, 256) if c.get("native") else args.trees)
    par = workloads.par_dict(workloads.jittered_models(WORKLOAD, args.psets))
    cores = host_threads()
    # bounded sample per step: sized from a probe so that (warmup + steps) stay within a few minutes
    n_trees, n_psets = min(flat.n_trees, 64), min(args.psets, 4)
    rate, dt = cpu_reference_rate(flat, par, c["T"], n_trees, n_psets)
    target = 6.0
    if dt < target / 2:
        n_trees = int(min(flat.n_trees, max(n_trees, n_trees * target / max(dt, 1e-3))))
        if n_trees == flat.n_trees:
            n_psets = int(min(args.psets, max(n_psets, rate * target / n_trees)))
    for _ in range(args.warmup):
        cpu_reference_rate(flat, par, c["T"], min(n_trees, 32), 1)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_reference_rate(flat, par, c["T"], n_trees, n_psets)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_trees * n_psets / (ms * 1e-3)
    sample = (f"each step = first {n_trees} trees x first {n_psets} parameter sets of the workload "
              f"(bounded sample), oracle/refcpu.c at reltol=abstol=1e-3, {cores} OpenMP threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(WORKLOAD, args.trees, args.psets, args.gpus,
                                  {"note": "CPU arm: C restatement of the reference algorithm (Julia unavailable in this image)"}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm

def run_gpu(args):
    import torch
    import torch.distributed as dist
    import gcdyn
    import workloads
    from gcdyn.likelihood import DeviceParams, evaluate, evaluate_resident, pack_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    c = workloads.CONFIGS[WORKLOAD]
    T, K, P, Tn = c["T"], c["K"], args.psets, args.trees
    if args.scaling == "strong":
        Tn = args.trees // world                         # --trees is the TOTAL under strong scaling
    native = bool(c.get("native"))

    # ---- inputs: this rank's shard of the forest, the replicated parameter batch
    flat = workloads.flat_forest(WORKLOAD, rank * Tn, Tn)
    models = workloads.jittered_models(WORKLOAD, P)
    params = pack_params(models)
    ctx = gcdyn.Context(local)
    torch.cuda.synchronize()
    t_c0 = time.perf_counter()
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)       # validate, index, upload (gcdyn_forest_create)
    t_c1 = time.perf_counter()
    evaluate(forest, params, T, steps=args.taylor_steps, order=args.order, want_status=False)   # moments, calibration, first evaluation
    t_c2 = time.perf_counter()
    cold = {"forest_create_ms": 1e3 * (t_c1 - t_c0), "first_evaluation_ms": 1e3 * (t_c2 - t_c1),
            "note": "once per dataset: gcdyn_forest_create (validate, derive lookup items, upload), then the first gcdyn_loglik_batch "
                    "(nodal Chebyshev moments built twice: generous grid, probe, one synchronisation, calibrated grid); "
                    "flattening the pointer trees is host code of the wrapper and not included"}
    dpar = DeviceParams(params, ctx)
    info = forest.info()
    out = torch.zeros(P, dtype=torch.float64, device="cuda")
    out_tree = torch.zeros(P * Tn, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    # a non-default stream: its handle is what the library launches on, so these CUDA events bracket the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    # multi-GPU: the [P] partial sums are combined by one peer-memory kernel per rank (gcdyn_comm_*, NVLink P2P
    # stores + sequence flags); NCCL stays as the fallback when the windows cannot be mapped (no P2P between the GPUs)
    peer, allreduce_kind = None, "none"
    if world > 1:
        from gcdyn.sharding import PeerAllReduce
        ok = 1
        try:
            # >= 512 parameter sets: a row window, so that every rank integrates only its slice of them
            peer = PeerAllReduce(ctx, P, row_bytes=PeerAllReduce.row_window_bytes(P, K) if P >= 512 else 0)
        except Exception as e:          # noqa: BLE001 - any failure on any rank -> everybody uses NCCL
            print(f"[bench] rank {rank}: peer all-reduce unavailable ({e}); using NCCL", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            peer = None
        allreduce_kind = "peer memory, fused into finish_kernel (gcdyn_loglik_batch_resident_sharded)" if peer else "nccl"

    def step_resident():
        # N > 1: gcdyn_loglik_batch_resident_sharded — the all-reduce of the [P] sums is fused into finish_kernel
        evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream,
                          steps=args.taylor_steps, order=args.order, comm=peer)
        if peer is None and world > 1:
            dist.all_reduce(out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the ranks' 75 us flush kernels differ by a few us, and a step that contains a rendezvous (the fused all-reduce)
    # would charge that difference to the path.  One peer barrier (the library's one-CTA all-reduce on a dummy double) after
    # the flush and BEFORE the start event lets every rank's timed step begin together; what remains inside the timed
    # region is the skew of the path's own kernels.  (--no-align restores the old behaviour.)
    align = torch.zeros(1, dtype=torch.float64, device="cuda") if (peer is not None and not args.no_align) else None

    def flush_and_align():
        flush.zero_()
        if align is not None:
            peer.allreduce(align.data_ptr(), 1, stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        flush_and_align()
        step_resident()
    barrier()
    diag = ctx.last_diag(P)
    launches_per_step = diag["n_launches"]

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream, L2 flushed between
    sampler = ClockSampler(local) if rank == 0 else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush_and_align()
        starts[i].record(stream)
        step_resident()
        ends[i].record(stream)
    barrier()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    result = out.cpu().numpy().copy()           # all-reduced [P] sums of the last timed step
    # keep the GPU busy a little longer so the clock sampler sees load even for very short runs; local work
    # only — a time-bounded loop must not contain collectives (ranks would issue different counts)
    t_end = time.perf_counter() + 1.0
    while time.perf_counter() < t_end:
        for _ in range(20):
            evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream,
                              steps=args.taylor_steps, order=args.order)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    tm = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_per_step = float(tm.item()) / args.steps
    value = world * Tn * P / (ms_per_step * 1e-3)

    # ---- e2e: the host-buffer C-ABI call (parameters H2D, results D2H inside the timed region)
    for _ in range(3):
        evaluate(forest, params, T, steps=args.taylor_steps, order=args.order, want_status=False, comm=peer)
    barrier()
    e2e_s = 0.0
    for i in range(args.steps):
        flush_and_align()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o, _, _ = evaluate(forest, params, T, steps=args.taylor_steps, order=args.order, want_status=False, comm=peer)
        if world > 1 and peer is None:
            t = torch.from_numpy(o).cuda()
            dist.all_reduce(t)
            o = t.cpu().numpy()
        e2e_s += time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Tn * P / (float(te.item()) / args.steps)
    n_par_doubles = sum(int(np.size(a)) for a in (params.lam, params.mu, params.delta, params.rho, params.sigma, params.Gamma)
                        if a is not None)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_par_doubles,
           "d2h_bytes_per_step": P * 8, "ms_per_step": 1e3 * float(te.item()) / args.steps,
           "api": ("gcdyn_loglik_batch_sharded" if peer is not None else "gcdyn_loglik_batch") +
                  " (host buffers; forest handle created once per dataset, as a sampler would)"}
    if world > 1:
        assert np.allclose(o, result, rtol=1e-12, atol=0)
    # every rank's own step WITHOUT the rendezvous (same kernels, no peer traffic): the sharded step cannot be faster than the
    # slowest of these, whatever the collective costs — the GPUs of one box differ by several per cent
    per_rank_ms = None
    if world > 1:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream,
                              steps=args.taylor_steps, order=args.order)
            b.record(stream)
        torch.cuda.synchronize()
        mine_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / len(evs)], dtype=torch.float64, device="cuda")
        allms = [torch.zeros_like(mine_ms) for _ in range(world)]
        dist.all_gather(allms, mine_ms)
        per_rank_ms = [float(x.item()) for x in allms]
    # the fused all-reduce against a library collective over the ranks' own per-tree values
    allreduce_check = None
    traj_sharded = False
    if world > 1:
        from gcdyn.likelihood import _lib
        step_resident()
        torch.cuda.synchronize()
        traj_sharded = bool(_lib().gcdyn_last_traj_sharded(ctx._h))
        mine = out_tree.view(P, Tn).sum(dim=1)
        dist.all_reduce(mine)
        got = out.cpu().numpy()
        allreduce_check = float(np.max(np.abs(got - mine.cpu().numpy()) / np.abs(got)))
        assert allreduce_check <= 1e-11, allreduce_check
        assert peer is None or peer.error() == 0

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel durations (CUDA events inside the library, on the launching stream), roofline
    ctx.set_profiling(True)
    acc = {}
    nprof = 10
    for _ in range(nprof):
        flush.zero_()
        evaluate_resident(forest, dpar, T, out.data_ptr(), out_tree.data_ptr(), 0, stream.cuda_stream,
                          steps=args.taylor_steps, order=args.order)
        for k, v in ctx.last_kernel_ms().items():
            acc[k] = acc.get(k, 0.0) + v / nprof
    ctx.set_profiling(False)
    S, M = diag["steps"], diag["order"]
    n_items = info["n_items"]
    fp64_peak = ctx.probe_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # algorithmic FLOPs per launch
    cheb = ctx.last_cheb()
    C = 2 * K + 4          # count columns (shared Γ: type changes are a per-tree constant, not K² columns)
    Dn = cheb["interp_degree"]
    J = C + (Dn + 1) * K if Dn else 0
    S_mean = float(np.mean(diag["cells"])) if "cells" in diag else float(S)
    gemm_path = bool(Dn)
    # trajectory: per cell M stages of (K-term dot + seed + convolution) and the F rows; node values of the grid
    flops = {"trajectory": P * S_mean * (M * (2.0 * K * K + 4.0 * K) + M * (M + 1.0) * K / 2 + 8.0 * (M + 2) * K) +
                           (P * K * (Dn + 1) * 2.0 * (M + 2) if gemm_path else 0.0),
             "gemm": 2.0 * Tn * P * J,
             "sweep": 0.0 if gemm_path else n_items * P * (2.0 * (M + 1) + 8.0),
             "finish": (4.0 * Tn * P + n_items * cheb["n_fallback"] * (2.0 * (M + 1) + 8.0)) if gemm_path else 1.0 * Tn * P}
    ran = ("trajectory", "gemm", "finish") if gemm_path else ("trajectory", "sweep", "finish")
    dom = max(ran, key=lambda k: acc.get(k, 0.0))
    achieved = flops[dom] / (acc[dom] * 1e-3) / 1e12
    per_kernel = {k: {"ms": acc[k], "algorithmic_gflop": flops[k] / 1e9,
                      "tflops": flops[k] / (acc[k] * 1e-3) / 1e12,
                      "frac_of_fp64_peak": flops[k] / (acc[k] * 1e-3) / 1e12 / fp64_peak}
                  for k in ran}
    # algorithmic bytes of the GEMM operands (read once) and its output
    gemm_bytes = (Tn * J + P * J + Tn * P) * 8.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_dram_traffic.json" if WORKLOAD == "C3" else f"r2_dram_traffic_{WORKLOAD}.json")))
        traffic = tr.get(f"{dom}_kernel", {}).get("dram_bytes_per_launch")
    except Exception:
        tr = {}
    # "tensor": the contraction runs on the FP64 tensor pipe (DMMA); "fp64": the trajectory kernel's DFMA chain — both are
    # measured against the same live FP64 FMA peak (B200's DMMA and DFMA rates are equal)
    roofline = {"kernel": f"{dom}_kernel", "bound": "tensor" if dom == "gemm" else "fp64", "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "peak_source": "measured live: register-resident DFMA loop on all SMs (gcdyn_probe_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "traffic": traffic, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of this "
                "command (profiles/r2_dram_traffic.json); L2 is flushed before every step, so the kernels' inputs come from HBM",
                "traffic_per_kernel": {k: v.get("dram_bytes_per_launch") for k, v in tr.items() if isinstance(v, dict)} if tr else None,
                "mean_cells_per_pset": S_mean, "max_cells_per_pset": S,
                "kernel_ms": {k: acc[k] for k in ran},
                "kernel_share": {k: acc[k] / sum(acc[r] for r in ran) for k in ran},
                "algorithmic_flops_per_launch": flops[dom], "per_kernel": per_kernel, "chebyshev": cheb,
                "hbm": {"achieved_gbs": gemm_bytes / (acc["gemm"] * 1e-3) / 1e9 if acc.get("gemm") else None,
                        "peak_gbs": peaks.get("hbm_gbs"),
                        "note": "GEMM operand + output bytes; the working set is L2-resident, HBM is not the bound"}}

    # ---- parity beside the number: oracle (shared-trajectory form, scipy DOP853 1e-13) on 16 parameter sets spread over
    #      the batch, per-pset sums of this rank's shard; on N > 1 also the ALL-REDUCED sums of 4 parameter sets against the
    #      oracle summed over every rank's shard (rank 0 regenerates the other shards: every tree has its own RNG stream)
    import oracle
    par = workloads.par_dict(models)
    host_tree = out_tree.cpu().numpy().reshape(P, Tn)
    check = sorted(set(int(round(x)) for x in np.linspace(0, P - 1, min(16, P))))
    rels = []
    n_par = min(Tn, 64) if native else Tn            # config 5: ~10^3 nodes per tree, the oracle takes ~0.1 s per (tree, pset)
    from gcdyn.sharding import subset_flat
    flat_par = subset_flat(flat, np.arange(n_par)) if n_par < Tn else flat
    for p in check:
        want = oracle.loglik_flat_shared(flat_par, par, p, T)
        rels.append(abs(host_tree[p, :n_par].sum() - want.sum()) / abs(want.sum()))
    parity = {"psets_checked": len(check), "psets": check, "trees_checked": n_par, "max_rel_err_sum": float(max(rels)), "gate": 1e-8}
    if allreduce_check is not None:
        parity["fused_allreduce_vs_nccl_of_per_tree_values"] = allreduce_check
    if world > 1 and not native:
        chk = check[:: max(1, len(check) // 4)][:4]
        tot = {p: 0.0 for p in chk}
        for r in range(world):
            fr = flat if r == 0 else workloads.flat_forest(WORKLOAD, r * Tn, Tn)
            for p in chk:
                tot[p] += float(oracle.loglik_flat_shared(fr, par, p, T).sum())
        parity["allreduced_psets_checked"] = len(chk)
        parity["allreduced_max_rel_err_sum"] = float(max(abs(result[p] - tot[p]) / abs(tot[p]) for p in chk))
        assert parity["allreduced_max_rel_err_sum"] <= 1e-8, parity
    assert parity["max_rel_err_sum"] <= 1e-8, parity

    cpu = cpu_baseline(flat, par, T) if world == 1 else None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(WORKLOAD, Tn, P, world, {
                "nodes_per_gpu": info["n_nodes"], "lookup_items_per_gpu": n_items, "taylor_steps": S,
                "taylor_order": M, "max_err_indicator": float(np.max(diag["err_indicator"])),
                "allreduce": allreduce_kind, "trajectories_sharded": traj_sharded,
                "per_rank_step_ms_without_rendezvous": per_rank_ms,
                "aligned_start": "peer barrier between the L2 flush and each timed step (outside the timed region)" if align is not None else None}),
            "clocks": clocks, "e2e": e2e, "cold": cold,
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "parity": parity}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_mh(args):
    """Config 4: C device-resident random-walk Metropolis chains (gcdyn_mh_*) on the config-3 forest.  A step is
    MH_ITERS iterations of all chains = one gcdyn_mh_run call (it synchronises; the library owns its stream, so the step is
    timed by the host clock around the call — >= 100 iterations per call keep the ~20 us call overhead below 1 %).
    `value`: chains keep their history on the device (no per-iteration output); `e2e`: the call that also returns every
    iteration's parameters and log-likelihoods to host memory."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import gcdyn
    import workloads
    import oracle
    from gcdyn.likelihood import MhDesc, _lib, _check, _ptr, _p_f64, _p_i64
    from gcdyn.mcmc import theta_of, model_of
    MH_ITERS = 100
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    c = workloads.CONFIGS["C4"]
    T, K, n_chains = c["T"], c["K"], args.psets
    Tn = args.trees // world if args.scaling == "strong" else args.trees
    flat = workloads.flat_forest("C4", rank * Tn, Tn)
    base, _ = workloads.base_model("C3")
    ctx = gcdyn.Context(local)
    forest = gcdyn.Forest(flat, None, T, ctx=ctx)
    comm = None
    if world > 1:
        from gcdyn.sharding import PeerAllReduce
        comm = PeerAllReduce(ctx, n_chains)
    lib = _lib()
    th = theta_of(base)
    th0 = np.ascontiguousarray(th + 0.05 * np.random.default_rng(4).standard_normal((n_chains, 6)))
    lo, up = np.ascontiguousarray(th - 1.5), np.ascontiguousarray(th + 1.5)
    ts = np.ascontiguousarray(base.type_space, dtype=np.float64)
    gam = np.ascontiguousarray(np.asarray(base.Γ, dtype=np.float64).ravel(order="F"))
    desc = MhDesc(n_chains, len(ts), _ptr(ts, _p_f64), _ptr(gam, _p_f64), float(base.ρ), float(base.σ), _ptr(th0, _p_f64),
                  _ptr(lo, _p_f64), _ptr(up, _p_f64), 0.02, 4)
    h = C.c_void_p()
    _check(lib.gcdyn_mh_create(ctx._h, forest._h, C.byref(desc), float(T), comm._h if comm else None, C.byref(h)), ctx._h)
    run = lambda n, a=None, b=None: _check(lib.gcdyn_mh_run(h, int(n), a, b), ctx._h)
    run(0)
    for _ in range(max(args.warmup, 3)):
        run(MH_ITERS)
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(MH_ITERS)
    dt = time.perf_counter() - t0
    chain = np.empty((MH_ITERS, n_chains, 6)); lls = np.empty((MH_ITERS, n_chains))
    for _ in range(3):
        run(MH_ITERS, _ptr(chain, _p_f64), _ptr(lls, _p_f64))
    t1 = time.perf_counter()
    for _ in range(args.steps):
        run(MH_ITERS, _ptr(chain, _p_f64), _ptr(lls, _p_f64))
    dte = time.perf_counter() - t1
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([dt, dte], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dte = float(tt[0]), float(tt[1])
    state_th = np.empty((n_chains, 6)); state_ll = np.empty(n_chains); acc = np.zeros(n_chains, dtype=np.int64); it = C.c_int64()
    _check(lib.gcdyn_mh_state(h, _ptr(state_th, _p_f64), _ptr(state_ll, _p_f64), _ptr(acc, _p_i64), C.byref(it)), ctx._h)
    lib.gcdyn_mh_destroy(h)
    if rank != 0:
        dist.barrier(); dist.destroy_process_group()
        return
    # parity beside the number: the sampler's own log-likelihood of four chains' final states against the oracle (all shards)
    chk = [0, n_chains // 3, 2 * n_chains // 3, n_chains - 1]
    models = [model_of(state_th[cix], base) for cix in chk]
    par = workloads.par_dict(models)
    rel = 0.0
    tot = np.zeros(len(chk))
    for r in range(world):
        fr = flat if r == 0 else workloads.flat_forest("C4", r * Tn, Tn)
        for j in range(len(chk)):
            tot[j] += float(oracle.loglik_flat_shared(fr, par, j, T).sum())
    rel = float(np.max(np.abs(state_ll[chk] - tot) / np.abs(tot)))
    assert rel <= 1e-8, rel
    n_it = args.steps * MH_ITERS
    ms_step = 1e3 * dt / args.steps
    value = world * Tn * n_chains * n_it / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict("C4", Tn, n_chains, world, {
                "step": f"{MH_ITERS} Metropolis iterations of {n_chains} chains (one gcdyn_mh_run call, CUDA graphs of 16 fused iterations)",
                "ms_per_iteration": ms_step / MH_ITERS, "iterations_per_s": n_it / dt,
                "acceptance_rate_mean": float(acc.mean() / max(1, it.value)),
                "timing": "host clock around the synchronising call (the library owns the stream); no L2 flush: a sampler's working "
                          "set stays in L2 between iterations by design"}),
            "clocks": clocks,
            "e2e": {"value": world * Tn * n_chains * n_it / dte, "unit": UNIT, "ms_per_step": 1e3 * dte / args.steps,
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": MH_ITERS * n_chains * 7 * 8,
                    "api": "gcdyn_mh_run with host history buffers (every iteration's parameters and log-likelihoods copied back)"},
            "gpu_launches": args.steps * MH_ITERS * 2,
            "roofline": None, "parity": {"chains_checked": len(chk), "max_rel_err_loglik_vs_oracle": rel, "gate": 1e-8}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gcdyn_b200", choices=["gcdyn_b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C4", "C5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--trees", type=int, default=None, help="trees per GPU (weak) / in total (strong); default: the workload's")
    ap.add_argument("--psets", type=int, default=None)
    ap.add_argument("--no-align", action="store_true", help="N > 1: no peer barrier between the L2 flush and the timed step")
    ap.add_argument("--order", type=int, default=0)
    ap.add_argument("--taylor-steps", type=int, default=0)
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    import workloads
    args.trees = args.trees or workloads.CONFIGS[WORKLOAD]["trees"]
    args.psets = args.psets or workloads.CONFIGS[WORKLOAD]["psets"]
    if args.impl != "reference" and WORKLOAD == "C4":
        return run_mh(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
