#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: PageRank GTEPS on synthetic RMAT scale-26 (edge factor 16), 20 iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scale S]          our arm  (B200, CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     the reference's CPU path

One "step" = one `execute(20)` of the PageRank vertex program (the reference's "Execute time" window,
src/vp/vertex_program.hpp:416-437: the iteration loop only) on the RMAT graph.  GTEPS = nnz * iterations / t
(SURVEY.md §8d).  N = 1 runs scale 26 (BASELINE.json configs[2] at one B200); N > 1 keeps the same graph
and shards it with the reference's 2D tile grid (strong scaling, as configs[2] is quoted).

Prints ONE JSON line.
  value          inputs resident in HBM, device-timed (CUDA events on the engine stream, max over ranks)
  e2e            the same K steps through the public API with HOST buffers — per step the initial vertex states go
                 host->device from pinned memory and the final states come back device->host, copies inside the timed
                 region
  roofline       the dominant kernel (the SpMV pass) timed live with CUDA events
  config.rank_sum_global / rank_sqsum_global
                 sum of rank and rank^2 over ALL vertices (all-reduced over the ranks): the same at every N to 1e-9
  multi_gpu_parity (N > 1)
                 before timing, every rank runs PR / BFS / CC / SSSP on the reference's own fixture and on a seeded
                 RMAT-12 graph and compares its owned segment per vertex with the committed dumps of the UNMODIFIED
                 reference (tests/golden/*.npz, made by tests/golden/make_golden.py) — nothing under oracle/ is touched
  other_configs  BASELINE.json configs[1], [3], [4] device-timed beside the headline: BFS RMAT-22 (N = 1), SSSP weighted
                 RMAT-25 (every N), CC RMAT-27 (N = 8), each with its algorithmic bytes, roofline fraction and checksum
  cpu_baseline   (N = 1) the unmodified reference (oracle/_ref, fork+shm MPI stand-in) on a bounded RMAT sample

`--impl reference` times the reference's own CPU implementation (oracle/_ref/ref_pr = src/apps/pr.cpp compiled
unmodified) with all the host threads it can use on the SAME graph — RMAT scale-26, generated from the same seed
and written to a temporary file — if the host has the memory and disk for it (~28 GB RSS, 8 GiB file), else on the
largest scale that fits, and `config.workload` names the scale that really ran.  Each step is a bounded sample of
the workload: REF_SAMPLE_ITERS of the 20 iterations (PageRank does the same work every iteration), so that the
whole run ends within a few minutes.
"""
import argparse
import json
import os
import shutil
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
REF_SAMPLE_ITERS = 5
METRIC = "pagerank_gteps"
UNIT = "GTEPS"
GOLDEN = os.path.join(ROOT, "tests", "golden")


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


# ---- the reference's CPU path ----------------------------------------------------------------------------------
def pick_np(limit=16):
    n = os.cpu_count() or 1
    p = 1
    while p * 2 <= min(n, limit):
        p *= 2
    return p


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def pick_reference_scale(want):
    """Largest scale <= want whose edge file (8 B x 16 x 2^s, page-cached while it is read) and resident set (measured:
    21 B per edge summed over the ranks at scale 26, np = 8) fit this host with a 1.7x margin."""
    tmp_free = shutil.disk_usage(tempfile.gettempdir()).free / 1e9
    mem = mem_available_gb()
    s = want
    while s > 16:
        edges = 16 << s
        if edges * 8 / 1e9 * 1.25 <= tmp_free and edges * 24 / 1e9 * 1.5 + edges * 8 / 1e9 <= mem:
            return s
        s -= 2
    return s


def reference_sample(scale, repeats, np_ranks, iters=ITERS):
    """Times the unmodified reference's PageRank on RMAT scale `scale`: returns (nnz, [execute seconds of `iters`
    iterations] * repeats)."""
    from oracle import oracle as O
    if not O.ref_available():
        raise RuntimeError("oracle/_ref/ref_pr is missing (build it in the container that has /root/reference)")
    d = tempfile.mkdtemp(prefix="gtbench_")
    path = os.path.join(d, f"rmat{scale}.bin")
    try:
        nnz = O.write_rmat(path, scale)
        env = dict(os.environ)
        env.pop("GT_MPI_NP", None)
        if np_ranks > 1:
            env["GT_MPI_NP"] = str(np_ranks)
        out = subprocess.run([os.path.join(O.REF_DIR, "ref_pr"), path, str(1 << scale), str(iters), "--repeat", str(repeats)],
                             capture_output=True, text=True, env=env, check=True).stdout
    finally:
        shutil.rmtree(d, ignore_errors=True)
    secs = [float(l.split()[2]) for l in out.splitlines() if l.startswith("Execute time:")][1:]   # [0] is the Deg pass
    assert len(secs) == repeats, out[-2000:]
    return nnz, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    np_ranks = pick_np()
    scale = pick_reference_scale(args.scale) if args.cpu_scale is None else args.cpu_scale
    iters = min(ITERS, REF_SAMPLE_ITERS) if scale >= 24 else ITERS
    nnz, secs = reference_sample(scale, args.warmup + args.steps, np_ranks, iters)
    timed = secs[args.warmup:]
    t = sum(timed)
    value = nnz * iters * len(timed) / t / 1e9
    sample = (f"unmodified reference pr (oracle/_ref = src/apps/pr.cpp, _TCSC_CF_), RMAT scale-{scale} ef=16 ({nnz} edges, same generator and seed as "
              f"the GPU arm), each step = {iters} of the {ITERS} iterations of one execute() on that graph (every PageRank iteration does the same "
              f"work), np={np_ranks} of {os.cpu_count()} host cores, reference 'Execute time' window")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / len(timed), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload(scale), "nnz": int(nnz), "vertices": 1 << scale, "iterations_per_step": iters,
                   "parallelism": f"mpi-np{np_ranks} (fork+shm stand-in)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": np_ranks, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- our arm -------------------------------------------------------------------------------------------------------
class Reducer:
    """max / min / sum over the ranks (torch.distributed is plumbing only)."""

    def __init__(self, dist):
        self.dist = dist

    def _red(self, x, op):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def max(self, x): return self._red(x, "MAX")
    def min(self, x): return self._red(x, "MIN")
    def sum(self, x): return self._red(x, "SUM")


def golden_parity(E, red):
    """PR / BFS / CC / SSSP on the reference fixture and a seeded RMAT-12 graph at this N, every rank's owned segment
    against the committed dumps of the unmodified reference (tests/golden/*.npz).  Returns the JSON block."""
    import numpy as np
    from graphtap_b200.rmat import rmat_edges
    ok, max_rel, ncases = True, 0.0, 0
    fw = np.fromfile(os.path.join(GOLDEN, "rmat10_1024_w.bin"), dtype="<u4").reshape(-1, 3)
    cases = [("fixture", fw, 1024, np.load(os.path.join(GOLDEN, "fixture.npz"))),
             ("rmat12_seed12", rmat_edges(12, seed=12, weighted=True), 4096, np.load(os.path.join(GOLDEN, "rmat12_seed12.npz")))]
    for name, tw, n, gold in cases:
        tu = tw[:, :2].copy()

        def loader(tri):
            def load(G, **fl):
                ct = fl.pop("compression_type")
                G.load_triples(tri, n, compression_type=ct, **fl)
            return load

        for app in ("pr", "bfs", "cc", "sssp"):
            tri = tw if app == "sssp" else tu
            if app == "pr":
                G, V = E.run_pr(loader(tri), ITERS)
            elif app == "bfs":
                G, V = E.run_bfs(loader(tri), 0)
            elif app == "cc":
                G, V = E.run_cc(loader(tri))
            else:
                G, V = E.run_sssp(loader(tri), 0)
            mine = V.V
            lay = G.info().layout
            lo = lay.owned_segment * lay.tile_height
            hi = min(lo + lay.tile_height, n + 1)
            ref = gold[f"{app}_np1_V"][lo:hi]
            m = mine[: max(0, hi - lo)]
            good = V.iteration == int(gold[f"{app}_np1_meta"][0])
            if len(ref):
                if app == "pr":
                    rel = float((np.abs(m["rank"] - ref["rank"]) / np.abs(ref["rank"])).max())
                    max_rel = max(max_rel, rel)
                    good &= rel <= 1e-6 and bool((m["degree"] == ref["degree"]).all())
                else:
                    for f in m.dtype.names:
                        good &= bool((m[f] == (ref[f] if ref.dtype.names else ref)).all())
            cs = V.checksum(quiet=True)
            if app != "pr":                      # PageRank's truncating checksum depends on the partition (K11)
                good &= cs == (int(gold[f"{app}_np1_meta"][1]), int(gold[f"{app}_np1_meta"][2]))
            V.free(); G.free()
            ok &= bool(good)
            ncases += 1
    return {"p": E.Env.nranks, "pass": bool(red.min(1.0 if ok else 0.0) == 1.0), "max_rel": red.max(max_rel), "cases": ncases,
            "against": "tests/golden/{fixture,rmat12_seed12}.npz = per-vertex dumps of the unmodified reference (np=1); integer apps bit-exact, PageRank <= 1e-6"}


def hbm_used_gb():
    """Device memory in use on this rank's GPU (nvidia-smi's view: every allocation incl. the CUDA contexts); None if unavailable."""
    try:
        idx = os.environ.get("LOCAL_RANK", "0")
        out = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits", "-i", idx],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return float(out) * 1.048576e-3          # MiB -> GB
    except Exception:
        return -1.0


def other_config(E, red, app, scale, peak, runs=3):
    """One of BASELINE.json configs[1,3,4]: device-timed execute() (the reference's "Execute time" window) of a
    non-stationary program on synthetic RMAT, the mean of `runs` runs after one warm-up run."""
    weighted = app == "sssp"
    G = E.Graph(weighted=weighted)
    t0 = time.time()
    if app == "bfs":      # src/apps/bfs.cpp:26-33
        G.load_rmat(scale, directed=False, transpose=False, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)
    elif app == "cc":     # src/apps/cc.cpp:25-32
        G.load_rmat(scale, directed=False, transpose=False, self_loops=True, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)
    else:                 # src/apps/sssp.cpp:26-40
        G.load_rmat(scale, directed=True, transpose=True, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)
    E.Env.barrier()
    build_s = time.time() - t0
    ms, tm, it, cs = [], None, 0, None
    for r in range(runs + 1):
        V = {"bfs": E.BFS_Program, "cc": E.CC_Program, "sssp": E.SSSP_Program}[app](G, False, app != "bfs", app == "bfs", E._ROW_)
        V.root = 0
        E.Env.quiet = True
        it = V.execute()
        tm = V.timing()
        if r:
            ms.append(tm.execute_ms)
        if r == runs:
            cs = V.checksum(quiet=True)
            used = red.max(hbm_used_gb())          # graph + program resident: the footprint of the config per GPU
        V.free()
        E.Env.barrier()
    nnz = G.info().nnz_global
    G.free()
    t = red.max(sum(ms) / len(ms)) * 1e-3
    bytes_all = red.sum(float(tm.bytes_algorithmic))
    n = E.Env.nranks
    name = {"bfs": "BFS from root 0", "cc": "Connected Components", "sssp": "SSSP from root 0 on weighted"}[app]
    return {"workload": f"{name} synthetic RMAT scale-{scale} (ef=16)", "n_gpus": n, "execute_ms": t * 1e3, "iterations": int(it),
            "sparse_iterations": int(tm.sparse_iterations), "nnz": int(nnz), "gteps": nnz / t / 1e9,
            "bytes_algorithmic": int(bytes_all), "achieved_gbs": bytes_all / t / 1e9, "frac": bytes_all / t / 1e9 / (peak * n),
            "checksum": [int(cs[0]), int(cs[1])], "build_seconds": round(build_s, 2), "kernel_launches": int(tm.kernel_launches),
            "hbm_used_gb_per_gpu": round(used, 2) if used >= 0 else None}


def run_ours(args):
    import ctypes as C
    import numpy as np
    from graphtap_b200 import capi, engine as E

    E.Env.quiet = True                     # stdout carries exactly one JSON line
    E.Env.init()
    rank, nranks = E.Env.rank, E.Env.nranks
    red = Reducer(E.Env._dist)
    scale = args.scale
    L = capi.lib()
    peak, peak_src = measured_peak_gbs()

    parity = golden_parity(E, red) if nranks > 1 and not args.no_parity else None

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
    hbm_used = red.max(hbm_used_gb())                     # tiles + pull layout + vectors + windows: everything the path keeps resident

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
    total_s = red.max(sum(step_ms)) * 1e-3
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
    # algorithmic bytes of the timed SpMV pass over this rank's tiles: IA 4 B/edge + the column pointer and
    # one read of x per non-empty column of a tile (12 B) + one write of y per non-empty row (8 B), counted by the library
    # for exactly what that pass runs: on the _TCSC_CF_ graph of pr.cpp a middle iteration is the REG x REG list only
    # (98.4 % of the entries at this scale; the reference skips the same entries, vertex_program.hpp:1264-1281)
    kb = int(P.timing().combine_bytes)
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
    e2e_ms, parts = [], [0.0, 0.0, 0.0]
    for s in range(min(2, args.warmup) + args.steps):
        E.Env.barrier()
        t0 = time.perf_counter()
        capi.check(L.gt_program_state_from_host(P.handle, pin_in, sb))  # H2D from pinned memory (synchronises)
        t1 = time.perf_counter()
        P.set("iteration", 0)
        capi.check(L.gt_program_execute(P.handle, ITERS, None))
        t2 = time.perf_counter()
        capi.check(L.gt_program_state_to_host(P.handle, pin_out, sb))   # D2H (synchronises)
        t3 = time.perf_counter()
        if s >= min(2, args.warmup):
            e2e_ms.append((t3 - t0) * 1e3)
            for k, d in enumerate((t1 - t0, t2 - t1, t3 - t2)):
                parts[k] += d * 1e3 / args.steps
    e2e_total = red.max(sum(e2e_ms)) * 1e-3
    # where a step's wall time goes on the slowest rank of each part (host clock; every call ends with the stream drained)
    e2e = {"value": nnz * ITERS * args.steps / e2e_total / 1e9, "unit": UNIT, "h2d_bytes_per_step": sb * nranks, "d2h_bytes_per_step": sb * nranks,
           "ms_per_step": 1e3 * e2e_total / args.steps,
           "parts_ms": {"state_from_host": red.max(parts[0]), "execute": red.max(parts[1]), "state_to_host": red.max(parts[2])}}
    out_states = np.frombuffer((C.c_char * sb).from_address(pin_out.value), dtype=E.PR_STATE)
    # Vertex_Program::checksum's vertex range (vid < nrows, :1936); the sums are the same numbers at every N up to f64
    # summation order, so N = 1, 2, 4, 8 can be compared directly
    nvalid = max(0, min(th, gi.layout.nrows - gi.layout.owned_segment * th))
    r = out_states["rank"][:nvalid].astype(np.float64)
    rank_sum = red.sum(float(r.sum()))
    rank_sqsum = red.sum(float((r * r).sum()))

    # ---- the other BASELINE configs, device-timed beside the headline ------------------------------------------------
    P.free(); D.free(); G.free()
    E.Env.barrier()
    others = []
    if not args.no_other_configs:
        try:
            if nranks == 1:
                others.append(other_config(E, red, "bfs", 22 if scale >= 22 else scale, peak))
            others.append(other_config(E, red, "sssp", 25 if scale >= 26 else max(10, scale - 1), peak))
            if nranks == 8:
                others.append(other_config(E, red, "cc", 27 if scale >= 26 else scale + 1, peak))
        except Exception as ex:                                                   # report, never fake
            others.append({"error": str(ex)})

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------------------
    cpu = None
    if nranks == 1 and not args.no_cpu_baseline:
        try:
            np_ranks = pick_np()
            cs = 22 if args.cpu_scale is None else args.cpu_scale
            cn, secs = reference_sample(cs, 2, np_ranks)
            cpu = {"value": cn * ITERS / secs[-1] / 1e9, "unit": UNIT, "cores": np_ranks, "kind": "reference",
                   "sample": f"unmodified reference pr (oracle/_ref), RMAT scale-{cs} ef=16 ({cn} edges), {ITERS} iterations, "
                             f"np={np_ranks} of {os.cpu_count()} host cores, reference 'Execute time' window "
                             f"(bounded sample; `bench.py --impl reference` runs the scale-{scale} graph itself)"}
        except Exception as ex:                                                   # report, never fake
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(scale), "nnz": int(nnz), "vertices": 1 << scale, "iterations_per_step": ITERS,
                       "parallelism": f"2dt-p{nranks}", "l2": "inputs (>= 4 GB of IA per pass) exceed the 126 MB L2, no flush needed",
                       "build_seconds": round(t_build, 2), "rank_sum_global": rank_sum, "rank_sqsum_global": rank_sqsum,
                       "hbm_used_gb_per_gpu": round(hbm_used, 2) if hbm_used >= 0 else None},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "other_configs": others,
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        print(json.dumps(line))
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
    ap.add_argument("--cpu-scale", type=int, default=None, help="RMAT scale of the CPU runs (default: the headline scale for --impl reference "
                                                                "if the host can hold it, 22 for the cpu_baseline leg)")
    ap.add_argument("--pr-layout", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
