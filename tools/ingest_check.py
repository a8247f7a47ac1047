"""Partitioned ingest at config scale, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 8 ... tools/ingest_check.py cc --scale 27          # BASELINE config #5's graph
    python -m torch.distributed.run --nproc-per-node 8 ... tools/ingest_check.py pr --scale 22 --file   # + the byte-range split of a file

Builds the graph twice — every rank scanning the whole record stream (gt_graph_build_rmat) and every rank generating /
reading only its 1/p share with the entries routed to the tile owners (gt_graph_build_rmat_partitioned /
gt_graph_build_partitioned) — and compares every array of every local tile and segment map by CRC.  Prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphtap_b200 import engine as E  # noqa: E402
from graphtap_b200.rmat import rmat_edges  # noqa: E402

FLAGS = {  # the drivers' Graph::load arguments
    "pr": dict(directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=True, compression_type=E._TCSC_CF_),
    "bfs": dict(directed=False, transpose=False, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_),
    "cc": dict(directed=False, transpose=False, self_loops=True, acyclic=False, parallel_edges=False, compression_type=E._TCSC_),
    "sssp": dict(directed=True, transpose=True, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_),
}


def fingerprint(G):
    """CRC of every tile array and segment map of this rank, one tile at a time (the arrays are GBs at config scale)."""
    crc, gi = 0, G.info()
    rows, cols = set(), set()
    for k in range(gi.ntiles_local):
        t = G.tile(k)
        rows.add(t["row_slot"]); cols.add(t["col_slot"])
        for f in ("JA", "IA", "A", "JC", "IR"):
            if t[f] is not None:
                crc = zlib.crc32(t[f].tobytes(), crc)
        del t
    for s in sorted(rows):
        for a in G.rowgrp_maps(s)[:2]:
            crc = zlib.crc32(a.tobytes(), crc)
    for s in sorted(cols):
        for a in G.colgrp_maps(s)[:2]:
            crc = zlib.crc32(a.tobytes(), crc)
    return crc, int(gi.nnz_local), int(gi.nnz_global), int(gi.nedges_input)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("app", choices=sorted(FLAGS))
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--file", action="store_true", help="also write the records to a binary file and let every rank read its byte range of it")
    a = ap.parse_args()
    E.Env.init()
    rank, p = E.Env.rank, E.Env.nranks
    weighted = a.app == "sssp"
    fl = FLAGS[a.app]
    out = {"app": a.app, "scale": a.scale, "n_gpus": p}

    def timed(build):
        E.Env.barrier()
        t0 = time.time()
        G = build()
        E.Env.barrier()
        return G, time.time() - t0

    Gg, out["build_global_s"] = timed(lambda: E.Graph(weighted=weighted).load_rmat(a.scale, **fl))
    ref = fingerprint(Gg)
    Gg.free()
    Gp, out["build_partitioned_s"] = timed(lambda: E.Graph(weighted=weighted).load_rmat(a.scale, partitioned=True, **fl))
    same = fingerprint(Gp) == ref
    Gp.free()
    if a.file:
        path = os.path.join(tempfile.gettempdir(), f"gt_ingest_{os.environ.get('MASTER_PORT', '0')}_{a.scale}.bin")
        if rank == 0:
            rmat_edges(a.scale, seed=a.scale, weighted=weighted).astype("<u4").tofile(path)
        Gf, out["build_from_file_shares_s"] = timed(lambda: E.Graph(weighted=weighted).load(path, 1 << a.scale, 1 << a.scale, **fl))
        same_file = fingerprint(Gf) == ref
        Gf.free()
        E.Env.barrier()
        if rank == 0:
            out["file_bytes"] = os.path.getsize(path)
            os.unlink(path)
    # every rank must agree
    import torch
    flags = torch.tensor([int(same), int(same_file) if a.file else 1], device="cuda")
    if E.Env._dist is not None:
        E.Env._dist.all_reduce(flags, op=E.Env._dist.ReduceOp.MIN)
    out["nnz_global"], out["records"] = ref[2], ref[3]
    out["partitioned_identical_on_every_rank"] = bool(flags[0].item())
    if a.file:
        out["file_shares_identical_on_every_rank"] = bool(flags[1].item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    E.Env.finalize()
    sys.exit(0 if bool(flags.min().item()) else 1)


if __name__ == "__main__":
    main()
