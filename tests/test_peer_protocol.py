"""Model check of the PageRank peer-window exchange (graphtap_b200/csrc/gt_engine.cu pull_scatter_gather /
pull_combine / pull_apply + gt_peer.cu): the same per-rank sequence of stream operations, executed by a randomised
scheduler that may run any enabled operation of any rank next, with every buffer access checked.

What is modelled: per rank a main stream and a side stream (in-order queues), CUDA events between them, puts as
non-atomic copies (begin ... end; the destination is garbage in between) followed by the arrival counter, polling
kernels, SpMV / applicator kernels as begin ... end intervals during which their inputs must not change.  The process
grid is the reference's: ranks sharing a column group exchange x, ranks sharing a row group exchange partial y
(tests/test_layout.py pins those groups to the reference; here any R x C grid is enough).

Claims checked: with TWO buffers per window, alternating by epoch parity, no interleaving produces a
read of stale / half-written data or a write under a reader, and nothing deadlocks — although there is no barrier
anywhere in the loop.  With ONE buffer the same scheduler finds a violation, so the check has teeth."""
import random

import pytest


class Hazard(Exception):
    pass


class Cell:
    """One chunk of a window buffer: the version it holds, whether a copy into it is in flight, how many kernels read it."""
    __slots__ = ("version", "writing", "readers")

    def __init__(self):
        self.version, self.writing, self.readers = 0, False, 0


class Rank:
    def __init__(self):
        self.main, self.side = [], []             # in-order queues of (name, ready(), run())
        self.events = {}                          # event name -> number of completed records


class Sim:
    """R x C process grid.  Column group of rank (r, c) = all ranks with the same c (size R) — they share x chunks;
    row group = all ranks with the same r (size C) — they share y.  (Names follow the data, not the reference's MPI
    communicators; only the group structure matters for the protocol.)"""

    def __init__(self, R, C, iters, nbuf, seed, nbuf_y=None):
        self.R, self.C, self.iters, self.nbuf = R, C, iters, nbuf
        self.nbuf_y = nbuf if nbuf_y is None else nbuf_y
        self.rng = random.Random(seed)
        self.ranks = {(r, c): Rank() for r in range(R) for c in range(C)}
        # x window of a rank: [parity][member of its column group]; y window: [parity][member of its row group]
        self.xwin = {k: [[Cell() for _ in range(R)] for _ in range(nbuf)] for k in self.ranks}
        self.xflag = {k: [0] * R for k in self.ranks}
        self.ywin = {k: [[Cell() for _ in range(C)] for _ in range(self.nbuf_y)] for k in self.ranks}
        self.yflag = {k: [0] * C for k in self.ranks}
        self.ylocal = {k: [Cell() for _ in range(C)] for k in self.ranks}        # Yh: one chunk per row-group member
        self.done_iters = {k: 0 for k in self.ranks}

    # ---- primitive operations ------------------------------------------------------------------------------
    @staticmethod
    def write_begin(cell, what):
        if cell.readers:
            raise Hazard(f"write under a reader: {what}")
        if cell.writing:
            raise Hazard(f"two writers: {what}")
        cell.writing = True

    @staticmethod
    def write_end(cell, version):
        cell.writing = False
        cell.version = version

    @staticmethod
    def read_begin(cell, version, what):
        if cell.writing:
            raise Hazard(f"read of a half-written buffer: {what}")
        if cell.version != version:
            raise Hazard(f"read of version {cell.version}, expected {version}: {what}")
        cell.readers += 1

    @staticmethod
    def read_end(cell):
        cell.readers -= 1

    # ---- the per-rank program, enqueued exactly in the engine's order ------------------------------------------------
    def enqueue_iteration(self, key, k):
        """Iteration k (1-based).  x(k) is produced by applicator k-1 (or the messenger for k = 1)."""
        rk = self.ranks[key]
        r, c = key
        R, C, nb = self.R, self.C, self.nbuf
        always = lambda: True

        def ev_record(q, name):
            q.append((f"record {name}", always, lambda: rk.events.__setitem__(name, rk.events.get(name, 0) + 1)))

        def ev_wait(q, name, count):
            q.append((f"wait {name}", lambda: rk.events.get(name, 0) >= count, lambda: None))

        def kernel(q, name, reads, writes, version_written=None):
            """reads: [(cell, version)], writes: [cell]; modelled as two queue entries (begin, end)."""
            def begin():
                for cell, v in reads:
                    self.read_begin(cell, v, f"{key} {name}")
                for cell in writes:
                    self.write_begin(cell, f"{key} {name}")

            def end():
                for cell, _ in reads:
                    self.read_end(cell)
                for cell in writes:
                    self.write_end(cell, version_written)
            q.append((f"{name} begin", always, begin))
            q.append((f"{name} end", always, end))

        def put(q, name, src_cell, src_version, dst_cell, version, flags, idx):
            def begin():
                self.read_begin(src_cell, src_version, f"{key} {name} (source)")
                self.write_begin(dst_cell, f"{key} {name}")

            def end():
                self.read_end(src_cell)
                self.write_end(dst_cell, version)
            q.append((f"{name} begin", always, begin))
            q.append((f"{name} end", always, end))
            q.append((f"{name} flag", always, lambda: flags.__setitem__(idx, version)))     # same stream: lands after the payload

        par = k % nb
        ypar = k % self.nbuf_y
        xw, yw = self.xwin[key], self.ywin[key]
        if k == 1:                                   # messenger writes the own chunk of parity 1
            kernel(rk.main, "messenger", [], [xw[par][r]], 1)
        # --- pull_scatter_gather: x puts on the side stream
        ev_record(rk.main, f"x{k}")
        ev_wait(rk.side, f"x{k}", 1)
        for j in range(1, R):
            q = ((r + j) % R, c)
            put(rk.side, f"xput{k}->{q}", xw[par][r], k, self.xwin[q][par][r], k, self.xflag[q], r)
        # --- pull_combine
        if k > 1:
            ev_wait(rk.main, f"yput{k - 1}", 1)      # the previous y puts have read Yh
        yl = self.ylocal[key]
        kernel(rk.main, "memset", [], list(yl), ("y", k, 0))
        kernel(rk.main, "spmv own part", [(xw[par][r], k)], list(yl), ("y", k, 1))
        rk.main.append(("wait x", lambda: all(self.xflag[key][m] >= k for m in range(R) if m != r), lambda: None))
        others = [m for m in range(C) if m != c]
        kernel(rk.main, "spmv rest, follower segments", [(xw[par][m], k) for m in range(R)], [yl[m] for m in others], ("y", k, 2))
        ev_record(rk.main, f"b{k}")
        ev_wait(rk.side, f"b{k}", 1)
        for m in others:
            q = (r, m)
            put(rk.side, f"yput{k}->{q}", yl[m], ("y", k, 2), self.ywin[q][ypar][c], k, self.yflag[q], c)
        ev_record(rk.side, f"yput{k}")
        kernel(rk.main, "spmv rest, owned segment", [(xw[par][m], k) for m in range(R)], [yl[c]], ("y", k, 2))
        rk.main.append(("wait y", lambda: all(self.yflag[key][m] >= k for m in others), lambda: None))
        # --- pull_apply: reads own partial + the followers' partials, writes x(k+1) into the other parity
        nxt = (k + 1) % nb
        kernel(rk.main, "applicator", [(yl[c], ("y", k, 2))] + [(yw[ypar][m], k) for m in others], [xw[nxt][r]], k + 1)
        rk.main.append(("iteration done", always, lambda: self.done_iters.__setitem__(key, k)))

    def run(self):
        for key in self.ranks:
            for k in range(1, self.iters + 1):
                self.enqueue_iteration(key, k)
        queues = [q for rk in self.ranks.values() for q in (rk.main, rk.side)]
        heads = [0] * len(queues)
        while True:
            ready = [i for i, q in enumerate(queues) if heads[i] < len(q) and q[heads[i]][1]()]
            if not ready:
                if all(heads[i] == len(q) for i, q in enumerate(queues)):
                    return
                stuck = [(i, queues[i][heads[i]][0]) for i in range(len(queues)) if heads[i] < len(queues[i])]
                raise Hazard(f"deadlock: {stuck[:6]}")
            # a biased scheduler: sometimes let one stream run far ahead, which is where buffer reuse bites
            i = self.rng.choice(ready)
            burst = self.rng.choice((1, 1, 1, 3, 10, 40))
            for _ in range(burst):
                if heads[i] < len(queues[i]) and queues[i][heads[i]][1]():
                    queues[i][heads[i]][2]()
                    heads[i] += 1
                else:
                    break


@pytest.mark.parametrize("grid", [(2, 1), (2, 2), (4, 2), (2, 4), (4, 4)])
def test_two_buffers_are_enough(grid):
    R, C = grid
    for seed in range(60):
        sim = Sim(R, C, iters=6, nbuf=2, seed=seed)
        sim.run()
        assert all(v == 6 for v in sim.done_iters.values())


@pytest.mark.parametrize("nbuf_x,nbuf_y", [(1, 2), (2, 1), (1, 1)])
def test_one_buffer_is_not(nbuf_x, nbuf_y):
    """The checker must be able to fail: with a single x buffer, or a single y buffer, some interleaving reuses it too
    early (x: the applicator overwrites a chunk a put or a peer's SpMV still reads; y: a row-group peer that is one
    iteration ahead overwrites the partial the leader's applicator is reading)."""
    found = 0
    for seed in range(300):
        try:
            Sim(2, 2, iters=6, nbuf=nbuf_x, seed=seed, nbuf_y=nbuf_y).run()
        except Hazard:
            found += 1
    assert found > 0


# ---- non-stationary programs: single-buffered frontier window, ordered by the per-iteration all-reduce -----------------
def run_frontier_model(R, iters, seed, allreduce=True):
    """One column group of R ranks (graphtap_b200/csrc/gt_engine.cu scatter_gather / combine for BFS, CC, SSSP): per
    iteration every rank stores its frontier into the peers' windows with a kernel on its only stream, polls the
    arrival counters, reads all segments in combine, applies, and joins the world all-reduce of the convergence count.
    The window has ONE buffer; the claim is that the all-reduce makes that safe."""
    rng = random.Random(seed)
    win = {r: [Cell() for _ in range(R)] for r in range(R)}          # win[rank][segment]
    flag = {r: [0] * R for r in range(R)}
    arrived = [0] * (iters + 2)
    queues = {r: [] for r in range(R)}
    always = lambda: True
    for r in range(R):
        q = queues[r]
        for k in range(1, iters + 1):
            def put_begin(r=r, k=k):
                for m in range(R):
                    Sim.write_begin(win[m][r], f"rank {r} frontier put {k} -> {m}")

            def put_end(r=r, k=k):
                for m in range(R):
                    Sim.write_end(win[m][r], k)
                    flag[m][r] = k
            q.append(("put begin", always, put_begin))
            q.append(("put end", always, put_end))
            q.append(("wait", lambda r=r, k=k: all(flag[r][m] >= k for m in range(R)), lambda: None))
            q.append(("combine begin", always, lambda r=r, k=k: [Sim.read_begin(win[r][m], k, f"rank {r} combine {k}") for m in range(R)]))
            q.append(("combine end", always, lambda r=r: [Sim.read_end(win[r][m]) for m in range(R)]))
            if allreduce:
                q.append(("all-reduce arrive", always, lambda k=k: arrived.__setitem__(k, arrived[k] + 1)))
                q.append(("all-reduce done", lambda k=k: arrived[k] == R, lambda: None))
    heads = {r: 0 for r in range(R)}
    while True:
        ready = [r for r in range(R) if heads[r] < len(queues[r]) and queues[r][heads[r]][1]()]
        if not ready:
            if all(heads[r] == len(queues[r]) for r in range(R)):
                return
            raise Hazard("deadlock")
        r = rng.choice(ready)
        for _ in range(rng.choice((1, 1, 2, 8, 30))):
            if heads[r] < len(queues[r]) and queues[r][heads[r]][1]():
                queues[r][heads[r]][2]()
                heads[r] += 1
            else:
                break


@pytest.mark.parametrize("R", [2, 4])
def test_frontier_window_single_buffer_is_ordered_by_the_allreduce(R):
    for seed in range(100):
        run_frontier_model(R, iters=6, seed=seed, allreduce=True)
    found = 0
    for seed in range(100):
        try:
            run_frontier_model(R, iters=6, seed=seed, allreduce=False)
        except Hazard:
            found += 1
    assert found > 0                     # without it a fast rank overwrites a segment a slow one is still reading
