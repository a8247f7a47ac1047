"""First end-to-end check on a B200: every app on the reference fixture and on seeded RMAT graphs,
compared per vertex with the unmodified reference (oracle/_ref binaries, which travel with gpurun)."""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphtap_b200 import engine as E
from graphtap_b200.rmat import rmat_edges

REF = os.path.join(ROOT, "oracle", "_ref")

def ref_run(app, path, n, arg, dtype):
    with tempfile.TemporaryDirectory() as d:
        cmd = [os.path.join(REF, "ref_" + app), path, str(n)] + ([str(arg)] if arg is not None else []) + ["--dump", os.path.join(d, "o")]
        out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
        V = np.fromfile(os.path.join(d, "o.r0.V.bin"), dtype=dtype)
        it = int([l for l in out.splitlines() if l.startswith("Iterations:")][-1].split()[1])
        return V, it, out

def loader_for(tri):
    def load(G, **fl):
        ct = fl.pop("compression_type")
        G.load_triples(tri, n_vertices, compression_type=ct, **fl)
    return load

ok = True
cases = [("fixture", os.path.join(ROOT, "tests/golden/rmat10_1024.bin"), os.path.join(ROOT, "tests/golden/rmat10_1024_w.bin"), 1024)]
tmp = tempfile.mkdtemp()
for scale in (12, 16):
    pu, pw = os.path.join(tmp, f"r{scale}.bin"), os.path.join(tmp, f"r{scale}_w.bin")
    rmat_edges(scale, weighted=True).tofile(pw)
    rmat_edges(scale, weighted=True)[:, :2].copy().tofile(pu)
    cases.append((f"rmat{scale}", pu, pw, 1 << scale))
for name, pu, pw, n_vertices in cases:
    tu = np.fromfile(pu, dtype="<u4").reshape(-1, 2)
    tw = np.fromfile(pw, dtype="<u4").reshape(-1, 3)
    t0 = time.time()
    G, VR = E.run_pr(loader_for(tu), 20)
    mine = VR.V; it = VR.iteration; VR.free(); G.free()
    ref, rit, _ = ref_run("pr", pu, n_vertices, 20, E.PR_STATE)
    n = n_vertices + 1
    rel = np.abs(mine["rank"][:n] - ref["rank"][:n]) / np.abs(ref["rank"][:n])
    good = rel.max() < 1e-6 and (mine["degree"][:n] == ref["degree"][:n]).all() and it == rit
    print(f"{name} PR   max rel err {rel.max():.3e} degrees {(mine['degree'][:n] == ref['degree'][:n]).all()} iters {it}/{rit} -> {'OK' if good else 'FAIL'}  ({time.time()-t0:.1f}s)")
    ok &= good
    for app, runner, tri, dt, path, field in (("bfs", E.run_bfs, tu, E.BFS_STATE, pu, None), ("cc", E.run_cc, tu, E.CC_STATE, pu, None), ("sssp", E.run_sssp, tw, E.SSSP_STATE, pw, None)):
        G, V = (runner(loader_for(tri), 0) if app != "cc" else runner(loader_for(tri)))
        mine = V.V; it = V.iteration; tm = V.timing(); V.free(); G.free()
        ref, rit, _ = ref_run(app, path, n_vertices, 0 if app != "cc" else None, dt)
        good = it == rit
        for f in dt.names:
            good &= bool((mine[f][:n] == ref[f][:n]).all())
        print(f"{name} {app.upper():4s} bit-exact {good} iters {it}/{rit} sparse_iters {tm.sparse_iterations} launches {tm.kernel_launches}")
        ok &= good
print("SMOKE", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
