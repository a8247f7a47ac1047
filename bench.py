#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: PageRank GTEPS on synthetic RMAT (edge factor 16), 20 iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scale S]          our arm  (B200, CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     the reference's CPU path

One "step" = one `execute(20)` of the PageRank vertex program (the reference's "Execute time" window,
src/vp/vertex_program.hpp:416-437: the iteration loop only) on the RMAT graph.  GTEPS = nnz * 20 / t
(SURVEY.md §8d).  N = 1 runs scale 26 (BASELINE.json configs[2] at one B200); N > 1 keeps the same graph
and shards it with the reference's 2D tile grid (strong scaling, as configs[2] is quoted).

Prints ONE JSON line.  `value`: inputs resident in HBM, device-timed (CUDA events on the engine stream,
max over ranks).  `e2e`: the same K steps through the public API with HOST buffers — per step the initial
vertex states go host->device from pinned memory and the final states come back device->host, copies
inside the timed region.  `roofline`: the dominant kernel (the SpMV pass) timed live with CUDA events.
`cpu_baseline`: the unmodified reference (oracle/_ref, fork+shm MPI stand-in) on a bounded RMAT sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"          # NCCL's version banner goes to stdout, which carries exactly one JSON line

ITERS = 20
METRIC = "pagerank_gteps"
UNIT = "GTEPS"


def workload(scale):
    return f"PageRank {ITERS} iters on synthetic RMAT scale-{scale} (ef=16)"


def measured_peak_gbs():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
def pick_np(limit=16):
    n = os.cpu_count() or 1
    p = 1
    while p * 2 <= min(n, limit):
        p *= 2
    return p


def reference_sample(scale, repeats, np_ranks):
    """Times the unmodified reference's PageRank on an RMAT sample: returns (GTEPS, [execute seconds], nnz)."""
    from oracle import oracle as O
    if not O.ref_available():
        raise RuntimeError("oracle/_ref/ref_pr is missing (build it in the container that has /root/reference)")
    d = tempfile.mkdtemp(prefix="gtbench_")
    path = os.path.join(d, f"rmat{scale}.bin")
    nnz = O.write_rmat(path, scale)
    env = dict(os.environ)
    env.pop("GT_MPI_NP", None)
    if np_ranks > 1:
        env["GT_MPI_NP"] = str(np_ranks)
    out = subprocess.run([os.path.join(O.REF_DIR, "ref_pr"), path, str(1 << scale), str(ITERS), "--repeat", str(repeats)],
                         capture_output=True, text=True, env=env, check=True).stdout
    os.remove(path)
    secs = [float(l.split()[2]) for l in out.splitlines() if l.startswith("Execute time:")][1:]   # [0] is the Deg pass
    assert len(secs) == repeats, out[-2000:]
    return nnz, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    np_ranks = pick_np()
    scale = args.cpu_scale
    nnz, secs = reference_sample(scale, args.warmup + args.steps, np_ranks)
    timed = secs[args.warmup:]
    t = sum(timed)
    value = nnz * ITERS * len(timed) / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / len(timed), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload(args.scale), "parallelism": f"mpi-np{np_ranks} (fork+shm stand-in)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": np_ranks, "kind": "reference",
                         "sample": f"unmodified reference pr (oracle/_ref), RMAT scale-{scale} ef=16 ({nnz} edges), {ITERS} iterations per step, "
                                   f"np={np_ranks} of {os.cpu_count()} host cores, reference 'Execute time' window"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import numpy as np
    from graphtap_b200 import capi, engine as E

    E.Env.quiet = True                     # stdout carries exactly one JSON line
    E.Env.init()
    rank, nranks = E.Env.rank, E.Env.nranks
    dist = E.Env._dist
    scale = args.scale
    L = capi.lib()

    t_build = time.time()
    G = E.Graph(weighted=False)
    G.load_rmat(scale, directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=True,
                compression_type=E._TCSC_CF_)                                   # src/apps/pr.cpp:26-32
    gi = G.info()
    nnz = gi.nnz_global
    D = E.Deg_Program(G, True, False, False, E._COL_)                           # pr.cpp:40-43
    D.execute(1)
    P = E.PR_Program(G, True, False, False, E._ROW_)
    if args.pr_layout is not None:
        P.set("pr_layout", args.pr_layout)
    P.initialize(D)
    E.Env.barrier()
    t_build = time.time() - t_build

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: resident inputs, device-timed ----------------------------------------------------------
    sampler = None
    step_ms = []
    launches = 0
    for s in range(args.warmup + args.steps):
        P.initialize(D)                                   # untimed: rank = alpha, degrees (pr.cpp:47)
        E.Env.barrier()
        if s == args.warmup and rank == 0:
            sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
            sampler.start()
        P.execute(ITERS)
        tm = P.timing()
        if s >= args.warmup:
            step_ms.append(tm.execute_ms)
            launches += tm.kernel_launches
    E.Env.barrier()
    total_s = max_over_ranks(sum(step_ms)) * 1e-3
    clocks = sampler.summary() if sampler else None
    value = nnz * ITERS * args.steps / total_s / 1e9

    # ---- per-phase and dominant-kernel timing (CUDA events on the engine stream) -------------------------------
    phases = []
    for ph in range(3):
        capi.check(L.gt_program_run_phase(P.handle, ph))              # warm
        reps = 5
        capi.check(L.gt_ctx_timer_begin(E.Env.ctx))
        for _ in range(reps):
            capi.check(L.gt_program_run_phase(P.handle, ph))
        ms = C.c_double()
        capi.check(L.gt_ctx_timer_end(E.Env.ctx, C.byref(ms)))
        phases.append(ms.value / reps)
    th = gi.layout.tile_height
    # algorithmic bytes of the SpMV pass over this rank's tiles: IA 4 B/edge + JA + one read of
    # each tile's x segment (8 B per non-empty column) + one write of each y segment (8 B per non-empty row)
    kb = 0
    for k in range(gi.ntiles_local):
        tv = capi.TileView()
        capi.check(L.gt_graph_tile_view(G.handle, k, C.byref(tv)))
        if tv.nnz:
            kb += 4 * tv.nnz + 4 * (tv.nnzcols + 1) + 8 * tv.nnzcols
    for slot in range(gi.layout.rank_nrowgrps):
        n = C.c_uint32()
        capi.check(L.gt_graph_rowgrp_maps(G.handle, slot, None, None, C.byref(n)))
        kb += 8 * n.value
    peak, peak_src = measured_peak_gbs()
    achieved = kb / (phases[1] * 1e-3) / 1e9
    tm = P.timing()
    iter_bytes = tm.bytes_algorithmic / max(1, tm.iterations)
    iter_ms = sum(step_ms) / len(step_ms) / ITERS
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of the SpMV kernel, one ncu --set full capture per launch
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"pagerank_rmat{scale}_p{nranks}"
        if key in t:
            traffic = t[key]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "combine phase = y zero-fill + SpMV over the local tiles", "kernel_ms": phases[1], "kernel_algorithmic_bytes": kb,
                "peak_source": peak_src,
                "iteration_algorithmic_bytes": iter_bytes, "iteration_ms": iter_ms,
                "iteration_frac": iter_bytes / (iter_ms * 1e-3) / 1e9 / peak,
                "phases_ms": {"scatter_gather": phases[0], "combine": phases[1], "apply": phases[2]}}

    # ---- e2e: host buffers, copies inside the timed region ------------------------------------------------------
    sb = 16 * th
    pin_in, pin_out = C.c_void_p(), C.c_void_p()
    capi.check(L.gt_host_alloc_pinned(sb, C.byref(pin_in)))
    capi.check(L.gt_host_alloc_pinned(sb, C.byref(pin_out)))
    P.initialize(D)
    capi.check(L.gt_program_state_to_host(P.handle, pin_in, sb))        # the initial states, on the host
    e2e_ms = []
    for s in range(min(2, args.warmup) + args.steps):
        E.Env.barrier()
        t0 = time.perf_counter()
        capi.check(L.gt_program_state_from_host(P.handle, pin_in, sb))  # H2D from pinned memory
        P.set("iteration", 0)
        capi.check(L.gt_program_execute(P.handle, ITERS, None))
        capi.check(L.gt_program_state_to_host(P.handle, pin_out, sb))   # D2H (synchronises)
        dt = (time.perf_counter() - t0) * 1e3
        if s >= min(2, args.warmup):
            e2e_ms.append(dt)
    e2e_total = max_over_ranks(sum(e2e_ms)) * 1e-3
    e2e = {"value": nnz * ITERS * args.steps / e2e_total / 1e9, "unit": UNIT, "h2d_bytes_per_step": sb * nranks, "d2h_bytes_per_step": sb * nranks,
           "ms_per_step": 1e3 * e2e_total / args.steps}
    out_states = np.frombuffer((C.c_char * sb).from_address(pin_out.value), dtype=E.PR_STATE)
    rank_sum = float(out_states["rank"].sum())

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------------------
    cpu = None
    if nranks == 1 and not args.no_cpu_baseline:
        try:
            np_ranks = pick_np()
            cn, secs = reference_sample(args.cpu_scale, 2, np_ranks)
            cpu = {"value": cn * ITERS / secs[-1] / 1e9, "unit": UNIT, "cores": np_ranks, "kind": "reference",
                   "sample": f"unmodified reference pr (oracle/_ref), RMAT scale-{args.cpu_scale} ef=16 ({cn} edges), {ITERS} iterations, "
                             f"np={np_ranks} of {os.cpu_count()} host cores, reference 'Execute time' window"}
        except Exception as ex:                                                   # report, never fake
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(scale), "nnz": int(nnz), "vertices": 1 << scale, "iterations_per_step": ITERS,
                       "parallelism": f"2dt-p{nranks}", "l2": "inputs (>= 4 GB of IA per pass) exceed the 126 MB L2, no flush needed",
                       "build_seconds": round(t_build, 2), "rank_sum_check": rank_sum},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    P.free(); D.free(); G.free()
    E.Env.barrier()
    E.Env.finalize()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=26)
    ap.add_argument("--cpu-scale", type=int, default=22, help="RMAT scale of the bounded CPU sample")
    ap.add_argument("--pr-layout", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
