"""Runs one of BASELINE.json's non-PageRank configs on synthetic RMAT and prints a JSON line:

    python tools/run_config.py bfs  --scale 22          # config #2   (1 GPU)
    python tools/run_config.py sssp --scale 25          # config #4   (1/2/4/8 GPUs under torchrun)
    python tools/run_config.py cc   --scale 27          # config #5   (8 GPUs under torchrun)

Reports the reference's "Execute time" window (device-timed), iterations, how many of them took the
frontier SpMSpV, traversed-edge rate (stored entries * iterations / time, an upper bound on real work for
the sparse iterations) and a result checksum (Vertex_Program::checksum)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphtap_b200 import engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("app", choices=["bfs", "cc", "sssp"])
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--root", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    E.Env.init()
    t0 = time.time()
    G = E.Graph(weighted=(a.app == "sssp"))
    if a.app == "bfs":
        G.load_rmat(a.scale, directed=False, transpose=False, self_loops=False, parallel_edges=False)
        mk = lambda: E.BFS_Program(G, False, False, True, E._ROW_)
    elif a.app == "cc":
        G.load_rmat(a.scale, directed=False, transpose=False, self_loops=True, parallel_edges=False)
        mk = lambda: E.CC_Program(G, False, True, False, E._ROW_)
    else:
        G.load_rmat(a.scale, directed=True, transpose=True, self_loops=False, parallel_edges=False)
        mk = lambda: E.SSSP_Program(G, False, True, False, E._ROW_)
    E.Env.barrier()
    build_s = time.time() - t0
    gi = G.info()
    times, it, tm, cs = [], 0, None, None
    for _ in range(a.repeat):
        V = mk()
        V.root = a.root
        E.Env.barrier()
        it = V.execute()
        tm = V.timing()
        times.append(tm.execute_ms)
        cs = V.checksum(quiet=True)
        V.free()
    best = min(times)
    if E.Env.rank == 0:
        print(json.dumps({"app": a.app, "scale": a.scale, "n_gpus": E.Env.nranks, "nnz_stored": int(gi.nnz_global), "iterations": it,
                          "sparse_iterations": tm.sparse_iterations, "execute_ms": best, "execute_ms_all": times,
                          "gteps_upper": gi.nnz_global * it / best / 1e6, "kernel_launches": int(tm.kernel_launches),
                          "checksum": cs, "build_seconds": round(build_s, 2)}))
    G.free()
    E.Env.barrier()
    E.Env.finalize()


if __name__ == "__main__":
    main()
