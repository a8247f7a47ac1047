"""The C-ABI library loads without a GPU, exports every symbol include/graphtap_b200.h declares, and
refuses to compute without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from graphtap_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "graphtap_b200.h")).read()
    return re.findall(r"^GT_API [^;(]*?\b(gt_[a-z0-9_]+)\(", hdr, flags=re.M)


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 30
    assert sorted(syms) == sorted(capi.PROTOTYPES)


def test_every_declared_symbol_is_exported():
    l = capi.lib()
    for s in declared_symbols():
        assert hasattr(l, s), s
    assert l.gt_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    st = capi.lib().gt_ctx_create(0, 0, 1, None, C.byref(ctx))
    assert st == capi.GT_ERR_NO_DEVICE
    assert b"no CPU fallback" in capi.lib().gt_last_error()
    with pytest.raises(capi.GraphTapError):
        capi.check(st)


def test_product_does_not_import_oracle():
    """The product path must not route through the CPU oracle (parity would be void)."""
    pkg = os.path.join(ROOT, "graphtap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "gt_oracle" not in txt and "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_edge_list_readers(tmp_path):
    """Host-side ingest of the reference's two file formats (src/mat/graph.hpp:194-372): binary records and
    text with '#'/'%' headers; both must give the same records."""
    import numpy as np
    from graphtap_b200.engine import read_edge_list
    rng = np.random.default_rng(3)
    tri = rng.integers(0, 1000, size=(200, 3), dtype=np.uint32)
    tri[:, 2] = tri[:, 2] % 128 + 1
    for weighted in (False, True):
        t = tri if weighted else tri[:, :2].copy()
        b = tmp_path / f"g{int(weighted)}.bin"
        t.tofile(b)
        txt = tmp_path / f"g{int(weighted)}.txt"
        with open(txt, "w") as f:
            f.write("# comment\n% another\n\n")
            for r in t:
                f.write(" ".join(str(int(x)) for x in r) + "\n")
            f.write("\n9 9 9\n")                     # after the first empty line: ignored, as in the reference
        assert (read_edge_list(str(b), weighted) == t).all()
        assert (read_edge_list(str(txt), weighted) == t).all()
    bad = tmp_path / "bad.txt"
    bad.write_text("1 2 3 4\n")
    with pytest.raises(capi.GraphTapError):
        read_edge_list(str(bad), False)


def test_every_environment_knob_is_documented():
    """The library reads its tuning knobs with getenv(); README.md lists every one of them."""
    import glob
    import re
    knobs = set()
    for path in glob.glob(os.path.join(ROOT, "graphtap_b200", "csrc", "*.c*")):
        knobs |= set(re.findall(r'getenv\("([A-Z0-9_]+)"\)', open(path).read()))
    assert knobs, "no getenv() found: the pattern is stale"
    readme = open(os.path.join(ROOT, "README.md")).read()
    missing = sorted(k for k in knobs if f"`{k}`" not in readme)
    assert not missing, f"undocumented environment knobs: {missing}"
