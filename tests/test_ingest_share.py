"""Partitioned ingest, host side: every rank reads its share of a binary edge file exactly as Graph::parread_binary
splits it (src/mat/graph.hpp:317-323: equal whole-record shares, the last rank takes the remainder), and the shares
tile the file without gap or overlap.  The device side (routing to tile owners, bit-identical graph) is checked on
GPUs by tools/multi_gpu_check.py / tests/test_gpu_multi.py."""
import os

import numpy as np
import pytest

from graphtap_b200 import engine as E

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_split(filesize, rec, rank, nranks):
    # the reference's arithmetic, restated from src/mat/graph.hpp:317-323
    share = (filesize // nranks) // rec * rec
    offset = share * rank
    endpos = filesize if rank == nranks - 1 else offset + share
    return offset, endpos


@pytest.mark.parametrize("rec", [8, 12])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 7, 8, 16])
def test_share_of_file_is_the_reference_split(rec, nranks):
    for nrec in (0, 1, 5, 16, 1000, 16384, 16385, 99991):
        filesize = nrec * rec
        prev_end = 0
        for r in range(nranks):
            off, end = E.share_of_file(filesize, rec, r, nranks)
            assert (off, end) == _ref_split(filesize, rec, r, nranks)
            assert off == prev_end and off % rec == 0 and end % rec == 0
            prev_end = end
        assert prev_end == filesize


@pytest.mark.parametrize("name,weighted", [("rmat10_1024.bin", False), ("rmat10_1024_w.bin", True)])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_shares_of_the_fixture_concatenate_to_the_file(name, weighted, nranks):
    path = os.path.join(G, name)
    whole = E.read_edge_list(path, weighted)
    parts = [E.read_edge_list_share(path, weighted, r, nranks) for r in range(nranks)]
    assert all(p.shape[1] == (3 if weighted else 2) for p in parts)
    assert (np.concatenate(parts) == whole).all()
    assert max(len(p) for p in parts[:-1] or parts) <= len(parts[-1]) or nranks == 1


def test_text_shares_concatenate(tmp_path):
    whole = E.read_edge_list(os.path.join(G, "rmat10_1024.bin"), False)[:1001]
    path = tmp_path / "g.txt"
    path.write_text("# header\n% another\n" + "".join(f"{r} {c}\n" for r, c in whole))
    for nranks in (1, 3, 4):
        parts = [E.read_edge_list_share(str(path), False, r, nranks) for r in range(nranks)]
        assert (np.concatenate(parts) == whole).all()


# ---- the exchange plan of gt_graph_build_partitioned (host arithmetic of the product library) ------------------------------
import ctypes as C
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plan(capi, counts, rank):
    p = counts.shape[0]
    arr = lambda: (C.c_uint64 * p)()
    so, ro, mo = arr(), arr(), arr()
    ns, nr, mx = C.c_uint64(), C.c_uint64(), C.c_uint64()
    flat = np.ascontiguousarray(counts, dtype=np.uint64)
    capi.check(capi.lib().gt_ingest_route_plan(p, rank, flat.ctypes.data_as(C.POINTER(C.c_uint64)), so, ro, mo, C.byref(ns), C.byref(nr), C.byref(mx)))
    return list(so), list(ro), list(mo), ns.value, nr.value, mx.value


@pytest.mark.parametrize("p", [1, 2, 3, 4, 8])
def test_route_plan_blocks_tile_every_buffer(p):
    from graphtap_b200 import capi
    rng = np.random.default_rng(p)
    for trial in range(8):
        counts = rng.integers(0, 50, size=(p, p)).astype(np.uint64)
        counts[rng.integers(0, p), :] = 0                     # a rank with an empty share
        if trial % 2:
            counts[:, rng.integers(0, p)] = 0                 # a rank that owns nothing
        plans = [_plan(capi, counts, r) for r in range(p)]
        for r, (so, ro, mo, ns, nr, mx) in enumerate(plans):
            assert ns == counts[r].sum() and nr == counts[:, r].sum() and mx == counts.sum(axis=0).max()
            assert so == list(np.concatenate([[0], np.cumsum(counts[r])[:-1]]))          # send buffer: destination order
            assert ro == list(np.concatenate([[0], np.cumsum(counts[:, r])[:-1]]))       # receive buffer: sender order
            for q in range(p):                                # my block lands in q's buffer exactly where q expects it
                assert mo[q] == plans[q][1][r]


def _route_worker(rank, world, port, out_dir):
    """One rank of a partitioned ingest with gloo in place of the peer window: flags on the share (graph.hpp:337-356),
    owner of every entry from the product's tile->rank table, the product's route plan, blocks written at the SENDER's
    remote offsets into the receiver's buffer."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from graphtap_b200 import capi, engine as E
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1024
    path = os.path.join(G, "rmat10_1024.bin")
    share = E.read_edge_list_share(path, False, rank, world).astype(np.int64)
    r, c = share[:, 0], share[:, 1]
    keep = r != c                                             # bfs.cpp's flags: undirected, no self-loops
    r, c = r[keep], c[keep]
    rr, cc = np.concatenate([r, c]), np.concatenate([c, r])   # mirrored
    cnt = C.c_uint32()
    capi.check(capi.lib().gt_layout_table(n, world, rank, capi.GT_LT_TILE_RANK, None, 0, C.byref(cnt)))
    tab = (C.c_int32 * cnt.value)()
    capi.check(capi.lib().gt_layout_table(n, world, rank, capi.GT_LT_TILE_RANK, tab, cnt.value, C.byref(cnt)))
    tile_rank = np.array(list(tab)).reshape(world, world)
    th = (n + 1) // world + 1
    dest = tile_rank[rr // th, cc // th]
    mine = np.bincount(dest, minlength=world).astype(np.int64)
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, torch.from_numpy(mine))
    counts = np.stack([t.numpy() for t in allc]).astype(np.uint64)
    so, ro, mo, ns, nr, mx = _plan(capi, counts, rank)
    order = np.argsort(dest, kind="stable")
    sendbuf = np.stack([rr[order], cc[order]], axis=1)        # grouped by destination, as k_route<true> leaves it
    recvbuf = np.full((nr, 2), -1, dtype=np.int64)
    reqs, stash = [], []
    for q in range(world):                                    # every block travels with the offset the SENDER computed
        blk = sendbuf[so[q]: so[q] + int(counts[rank, q])]
        if q == rank:
            recvbuf[mo[q]: mo[q] + len(blk)] = blk
        else:
            reqs.append(dist.isend(torch.tensor([mo[q]], dtype=torch.int64), q, tag=1))
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(blk)), q, tag=2))
    for q in range(world):
        if q == rank:
            continue
        off = torch.zeros(1, dtype=torch.int64)
        dist.recv(off, q, tag=1)
        blk = torch.zeros((int(counts[q, rank]), 2), dtype=torch.int64)
        dist.recv(blk, q, tag=2)
        assert off.item() == ro[q]                            # ... and it is the offset the RECEIVER's plan has for that sender
        recvbuf[off.item(): off.item() + blk.shape[0]] = blk.numpy()
    for w in reqs:
        w.wait()
    np.save(os.path.join(out_dir, f"recv{rank}.npy"), recvbuf)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_ingest_plan_over_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    from graphtap_b200 import capi, engine as E
    mp.spawn(_route_worker, args=(world, 29650 + world, str(tmp_path)), nprocs=world, join=True)
    # what every rank must hold: the entries of the WHOLE mirrored, loop-free list that fall in its tiles
    n = 1024
    whole = E.read_edge_list(os.path.join(G, "rmat10_1024.bin"), False).astype(np.int64)
    r, c = whole[:, 0], whole[:, 1]
    keep = r != c
    rr, cc = np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]])
    th = (n + 1) // world + 1
    cnt = C.c_uint32()
    capi.check(capi.lib().gt_layout_table(n, world, 0, capi.GT_LT_TILE_RANK, None, 0, C.byref(cnt)))
    tab = (C.c_int32 * cnt.value)()
    capi.check(capi.lib().gt_layout_table(n, world, 0, capi.GT_LT_TILE_RANK, tab, cnt.value, C.byref(cnt)))
    owner = np.array(list(tab)).reshape(world, world)[rr // th, cc // th]
    total = 0
    for k in range(world):
        got = np.load(tmp_path / f"recv{k}.npy")
        assert (got >= 0).all()                               # no gap in the receive buffer
        want = np.stack([rr[owner == k], cc[owner == k]], axis=1)
        key = lambda a: np.sort(a[:, 0] * (1 << 32) + a[:, 1])
        assert (key(got) == key(want)).all()                  # same multiset of entries (the owner sorts them anyway)
        total += len(got)
    assert total == len(rr)
