"""The source-compatible C++ shim (include/graphtap/graphtap.hpp): apps/gt_apps.cpp makes the reference
drivers' calls and must print the reference's own checksum lines on the reference's own fixtures."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "apps", "gt_apps")
G = os.path.join(ROOT, "tests", "golden")
EXPECT = {  # Iterations / Value checksum / Reachable vertices
    "pr": ("rmat10_1024.bin", "20", (20, 70, 1025)),
    "pr_until_converged": ("rmat10_1024.bin", None, (12, 51, 1025)),        # _TCSC_CF_ convergence mode, SURVEY.md §8c "quirky path"
    "bfs": ("rmat10_1024.bin", "0", (4, 1912, 887)),
    "cc": ("rmat10_1024.bin", None, (4, 69590, 1025)),
    "sssp": ("rmat10_1024_w.bin", "0", (7, 53366, 471)),
}


@pytest.mark.parametrize("defs", [[], ["-DTIMING"]])
def test_shim_compiles_against_the_c_abi(defs):
    subprocess.run(["g++", "-std=c++14", "-fsyntax-only"] + defs + ["-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "apps", "gt_apps.cpp")], check=True)


@pytest.mark.gpu
@pytest.mark.parametrize("app", sorted(EXPECT))
def test_cpp_driver_prints_reference_checksums(app):
    f, arg, (it, cs, reach) = EXPECT[app]
    cmd = [EXE, app.split("_")[0], os.path.join(G, f), "1024"] + ([arg] if arg else [])
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    get = lambda k: int([l for l in out.splitlines() if l.startswith(k)][-1].split()[-1])
    assert (get("Iterations:"), get("Value checksum:"), get("Reachable vertices:")) == (it, cs, reach)
    if app == "bfs":
        assert "vertex[1]:Parent=317,Hops=2" in out and "vertex[5]:Parent=866,Hops=3" in out
    if app == "pr":
        assert "vertex[4]:Rank=1.238176,Degree=2" in out and "vertex[5]:Rank=0.150000,Degree=0" in out
    if app == "sssp":
        assert "vertex[2]:Distance=INF" in out and "vertex[9]:Distance=116" in out


@pytest.mark.gpu
@pytest.mark.parametrize("app", ["pr", "sssp"])
def test_timing_build_prints_the_reference_report(app):
    """-DTIMING (src/vp/vertex_program.hpp:2134-2152): the per-phase report and the one-line `TIMING init sg_sum sg_avg sg_std
    cb_sum cb_avg cb_std ap_sum ap_avg ap_std execute` record, plus one `Iteration:` line per iteration (:431)."""
    f, arg, (it, cs, reach) = EXPECT[app]
    out = subprocess.run([EXE + "_timing", app, os.path.join(G, f), "1024", arg], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    get = lambda k: int([l for l in lines if l.startswith(k)][-1].split()[-1])
    assert (get("Iterations:"), get("Value checksum:"), get("Reachable vertices:")) == (it, cs, reach)      # timing does not change results
    its = [int(l.split()[-1]) for l in lines if l.startswith("Iteration: ")]
    assert its[-it:] == list(range(1, it + 1))
    rec = [l for l in lines if l.startswith("TIMING ")][-1].split()[1:]
    assert len(rec) == 11
    v = [float(x) for x in rec]
    assert all(x >= 0 for x in v) and v[10] > 0
    for k in (1, 4, 7):                      # sum = avg * iterations for each of the three phases
        assert abs(v[k] - v[k + 1] * it) <= 1e-3 * max(v[k], 1e-9) + 1e-6
    for head in ("Init           time:", "Scatter_gather time (sum: avg +/- std_dev):", "Combine        time (sum: avg +/- std_dev):",
                 "Apply          time (sum: avg +/- std_dev):", "Execute        time:"):
        assert any(l.startswith(head) for l in lines), head
