"""Sweep of the pull layout's virtual-row length (GT_PULL_VROW) at launch sizes that match one GPU's share at
p = 2, 4, 8 (a kernel launch of 2^25 .. 2^28 entries instead of 2^30): one process, one graph build per point.

    python tools/sweep_vrow.py [scales ...] [KNOB=V[,KNOB=V...] ...]     default scales 21 22 23 24, built-in knob list

Prints device-timed PageRank execute(20) and the combine phase per iteration."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphtap_b200 import capi, engine as E  # noqa: E402


def point(scale, env):
    for k in ("GT_PULL_VROW", "GT_PULL_CTAS", "GT_PULL_THREADS", "GT_PULL_BAND", "GT_PULL_L2HINT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    L = capi.lib()
    G = E.Graph(weighted=False)
    G.load_rmat(scale, directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=True, compression_type=E._TCSC_CF_)
    nnz = G.info().nnz_global
    D = E.Deg_Program(G, True, False, False, E._COL_)
    D.execute(1)
    P = E.PR_Program(G, True, False, False, E._ROW_)
    best = 1e30
    for s in range(6):
        P.initialize(D)
        P.execute(20)
        if s >= 2:
            best = min(best, P.timing().execute_ms)
    capi.check(L.gt_program_run_phase(P.handle, 0))
    capi.check(L.gt_program_run_phase(P.handle, 1))
    capi.check(L.gt_ctx_timer_begin(E.Env.ctx))
    for _ in range(10):
        capi.check(L.gt_program_run_phase(P.handle, 1))
    ms = C.c_double()
    capi.check(L.gt_ctx_timer_end(E.Env.ctx, C.byref(ms)))
    combine = ms.value / 10
    print(f"scale {scale} {' '.join(f'{k}={v}' for k, v in env.items()) or 'default':40s} execute(20) {best:8.3f} ms  {nnz * 20 / best / 1e6:7.1f} GTEPS  "
          f"combine {combine * 1e3:8.1f} us  {nnz / combine / 1e6:6.1f} G entries/s", flush=True)
    P.free(); D.free(); G.free()


def main():
    scales = [int(a) for a in sys.argv[1:] if "=" not in a and a != "default"] or [21, 22, 23, 24]
    envs = [dict(kv.split("=") for kv in a.split(",")) if a != "default" else {} for a in sys.argv[1:] if "=" in a or a == "default"]
    if not envs:
        envs = [{}, {"GT_PULL_VROW": "256"}, {"GT_PULL_VROW": "128"}, {"GT_PULL_VROW": "64"}, {"GT_PULL_VROW": "32"},
                {"GT_PULL_VROW": "64", "GT_PULL_CTAS": "1"}, {"GT_PULL_VROW": "64", "GT_PULL_THREADS": "512", "GT_PULL_CTAS": "4"}]
    E.Env.quiet = True
    E.Env.init()
    for scale in scales:
        for env in envs:
            point(scale, env)
    E.Env.finalize()


if __name__ == "__main__":
    main()
