"""Synthetic RMAT edge lists in the reference's headerless binary format (host-side, numpy).

The reference ships no generator; SURVEY.md §8(d) fixes the convention: Graph500 parameters
(a, b, c, d) = (0.57, 0.19, 0.19, 0.05), edge factor 16, random vertex-label permutation, uint32 ids
in [0, 2^scale), records ``{u32 src, u32 dst[, u32 w]}`` (reference ``src/ds/triple.hpp:9-18,40-49``,
read by ``src/mat/graph.hpp:307-372``), weights i.i.d. uniform in [1, 128] (the reference's own
convention, ``src/misc/converter.cpp:81``).

The generator is COUNTER-BASED: edge ``e`` depends only on ``(seed, e)`` through splitmix64 and
integer thresholds, so this numpy version and the CUDA version in ``csrc/rmat_gen.cu`` produce
bit-identical edge lists (tests/test_rmat.py checks that on the GPU), and any rank can generate any
slice of the list without coordination.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
_MASK32 = U64(0xFFFFFFFF)

# integer quadrant thresholds on a uniform u32: [0,A) -> a, [A,AB) -> b, [AB,ABC) -> c, else d
_T_A = int(0.57 * 2**32)
_T_AB = int((0.57 + 0.19) * 2**32)
_T_ABC = int((0.57 + 0.19 + 0.19) * 2**32)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x + U64(0x9E3779B97F4A7C15)).astype(U64)
        z = ((z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)).astype(U64)
        z = ((z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)).astype(U64)
        return (z ^ (z >> U64(31))).astype(U64)


def _scalar_splitmix64(x: int) -> int:
    return int(splitmix64(np.array([x & 0xFFFFFFFFFFFFFFFF], dtype=U64))[0])


def permute_labels(v: np.ndarray, scale: int, seed: int) -> np.ndarray:
    """Bijection on [0, 2^scale): three rounds of (odd multiply, add, xor-shift), all mod 2^scale."""
    mask = U64((1 << scale) - 1)
    v = v.astype(U64)
    k = _scalar_splitmix64(seed * 0x632BE59BD9B4E019 + 0x1234567)
    sh = U64(max(1, scale // 2))
    with np.errstate(over="ignore"):
        for r in range(3):
            k = _scalar_splitmix64(k + r)
            mul = U64((k | 1) & 0xFFFFFFFFFFFFFFFF)
            add = U64((k >> 17) & 0xFFFFFFFFFFFFFFFF)
            v = ((v * mul + add) & mask).astype(U64)
            v = (v ^ (v >> sh)).astype(U64)
    return v


def root_pre_image(scale: int) -> int:
    """Pre-permutation id that is relabelled to vertex 0: popcount scale//4, i.e. a vertex of typical
    edge-endpoint popularity (mean popcount of an RMAT endpoint is 0.24*scale), so that the BASELINE
    configs' ``root 0`` is neither isolated nor the biggest hub."""
    return (1 << max(1, scale // 4)) - 1


def rmat_edges(scale: int, nedges: int | None = None, seed: int | None = None, weighted: bool = False,
               first_edge: int = 0, permute: bool = True) -> np.ndarray:
    """Return ``nedges`` records starting at edge index ``first_edge`` as a (nedges, 2|3) uint32 array
    (columns src, dst[, w]); ``tofile()`` of it is exactly the reference's binary input format."""
    if nedges is None:
        nedges = 16 << scale
    if seed is None:
        seed = scale
    base = _scalar_splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5)
    e = np.arange(first_edge, first_edge + nedges, dtype=U64)
    src = np.zeros(nedges, dtype=U64)
    dst = np.zeros(nedges, dtype=U64)
    with np.errstate(over="ignore"):
        ctr = (U64(base) + e * U64(32)).astype(U64)
        for lvl in range(scale):
            if lvl % 2 == 0:
                r = splitmix64((ctr + U64(lvl // 2)).astype(U64))
                u = r & _MASK32
            else:
                u = r >> U64(32)
            sbit = (u >= U64(_T_AB)).astype(U64)                      # c or d  -> src bit 1
            dbit = (((u >= U64(_T_A)) & (u < U64(_T_AB))) | (u >= U64(_T_ABC))).astype(U64)  # b or d
            src = (src << U64(1)) | sbit
            dst = (dst << U64(1)) | dbit
        if permute:
            zero = permute_labels(np.array([root_pre_image(scale)], dtype=U64), scale, seed)[0]
            src = permute_labels(src, scale, seed) ^ zero
            dst = permute_labels(dst, scale, seed) ^ zero
        cols = [src.astype(np.uint32), dst.astype(np.uint32)]
        if weighted:
            w = splitmix64((ctr + U64(31)).astype(U64))
            cols.append(((w >> U64(33)) % U64(128) + U64(1)).astype(np.uint32))
    return np.ascontiguousarray(np.stack(cols, axis=1))


def write_binary(path: str, scale: int, seed: int | None = None, weighted: bool = False,
                 nedges: int | None = None, chunk: int = 1 << 22) -> int:
    """Stream the edge list to ``path`` in the reference's format; returns the number of records."""
    if nedges is None:
        nedges = 16 << scale
    with open(path, "wb") as f:
        done = 0
        while done < nedges:
            n = min(chunk, nedges - done)
            rmat_edges(scale, n, seed, weighted, first_edge=done).tofile(f)
            done += n
    return nedges
