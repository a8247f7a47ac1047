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
