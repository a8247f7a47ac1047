"""The counter-based RMAT generator: host (numpy) properties, and host == device bit for bit."""
import ctypes as C

import numpy as np
import pytest

from graphtap_b200.rmat import permute_labels, rmat_edges


def test_permutation_is_a_bijection():
    for s in (1, 5, 10, 16):
        v = permute_labels(np.arange(1 << s), s, seed=s)
        assert len(np.unique(v)) == 1 << s and v.max() < (1 << s)


def test_stream_is_counter_based_and_seeded():
    a = rmat_edges(12, 5000, seed=7, weighted=True)
    b = np.concatenate([rmat_edges(12, 2000, seed=7, weighted=True), rmat_edges(12, 3000, seed=7, weighted=True, first_edge=2000)])
    assert (a == b).all()
    assert not (a == rmat_edges(12, 5000, seed=8, weighted=True)).all()
    assert a[:, 2].min() >= 1 and a[:, 2].max() <= 128                 # src/misc/converter.cpp:81
    assert a[:, :2].max() < 4096


def test_graph500_skew_and_root():
    e = rmat_edges(14)
    n = 1 << 14
    out = np.bincount(e[:, 0], minlength=n)
    assert 0.5 < (out > 0).mean() < 0.8              # ~67 % at scale 14, falling to ~52 % at scale 20
    assert out[0] > 0 and np.bincount(e[:, 1], minlength=n)[0] > 0     # root 0 is not isolated
    assert out.max() > 200 * out.mean()              # heavy tail


@pytest.mark.gpu
def test_device_generator_matches_host():
    from graphtap_b200 import capi
    from graphtap_b200.engine import Env
    Env.init()
    for scale, w in ((10, 0), (13, 1), (20, 0)):
        n = 1 << 15
        first = 12345
        host = rmat_edges(scale, n, seed=scale + 3, weighted=bool(w), first_edge=first)
        dev = C.c_void_p()
        capi.check(capi.lib().gt_dev_alloc(Env.ctx, host.nbytes, C.byref(dev)))
        capi.check(capi.lib().gt_rmat_generate(Env.ctx, scale, first, n, scale + 3, w, dev))
        out = np.empty_like(host)
        capi.check(capi.lib().gt_dev_download(Env.ctx, out.ctypes.data_as(C.c_void_p), dev, host.nbytes))
        capi.check(capi.lib().gt_dev_free(Env.ctx, dev))
        assert (out == host).all()
