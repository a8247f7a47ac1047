// gt_build.cu — edge list -> 2DT tiles in TCSC, entirely on the device.
//
// Replaces, for this rank (reference paths relative to the GraphTap repo):
//   Graph::parread_binary     per-edge flags: drop self-loops, acyclic, transpose, mirror   src/mat/graph.hpp:337-356
//   Matrix::distribute        every rank sees the GLOBAL list and keeps its own tiles        src/mat/matrix.hpp:692-810
//   Matrix::init_tiles        sort by (col,row) [weighted: see below], adjacent dedup        src/mat/matrix.hpp:537-560, src/ds/triple.hpp:78-98
//   Matrix::filter_vertices   group-wide non-empty bitmaps I/J + prefix maps IV/JV           src/mat/matrix.hpp:860-1122
//   TCSC_BASE::populate       JA / IA / A / JC / IR                                          src/ds/compressed_column.hpp:370-417
//
// Because each rank scans the whole edge list it can mark EVERY row and column of the matrix, so the
// row-group / column-group OR-reduce + broadcast of the bitmaps (matrix.hpp:973-1083) needs no
// communication: the maps come out identical on all ranks of a group by construction.
//
// Weighted tiles: the reference sorts by (col, weight) with an unstable std::sort and then drops
// ADJACENT equal (row,col) pairs, so which heavier duplicates survive is unspecified; the lightest
// copy of every (row,col) always survives.  Here the order is (col, row, weight) and only the
// lightest copy is kept, which gives identical min-plus results (min is idempotent) with a
// deterministic layout.  Unweighted tiles are bit-identical to the reference's.
#include "gt_graph.h"
#include "gt_peer.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <memory>

namespace gt {

// ---------------------------------------------------------------------------------------------
// counter-based RMAT stream; must stay bit-identical to graphtap_b200/rmat.py
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct RmatParams {
    uint64_t base, mul[3], add[3], mask, zero;
    uint32_t sh, scale;
    int weighted, enabled;
};

__host__ __device__ __forceinline__ uint64_t rmat_permute(const RmatParams& P, uint64_t v) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
        v = (v * P.mul[r] + P.add[r]) & P.mask;
        v ^= v >> P.sh;
    }
    return v;
}

static RmatParams make_rmat_params(uint32_t scale, uint64_t seed, int weighted) {
    RmatParams P{};
    P.enabled = 1;
    P.scale = scale;
    P.weighted = weighted;
    P.mask = (scale >= 64) ? ~0ull : ((1ull << scale) - 1);
    P.sh = std::max(1u, scale / 2);
    uint64_t k = splitmix64(seed * 0x632BE59BD9B4E019ull + 0x1234567ull);
    for (int r = 0; r < 3; r++) {
        k = splitmix64(k + (uint64_t) r);
        P.mul[r] = k | 1ull;
        P.add[r] = k >> 17;
    }
    P.base = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull);
    P.zero = 0;
    uint64_t root_pre = (1ull << std::max(1u, scale / 4)) - 1;
    P.zero = rmat_permute(P, root_pre);
    return P;
}

__device__ __forceinline__ void rmat_edge(const RmatParams& P, uint64_t e, uint32_t& src, uint32_t& dst, uint32_t& w) {
    const uint64_t ctr = P.base + e * 32ull;
    uint64_t s = 0, d = 0, r = 0;
    for (uint32_t lvl = 0; lvl < P.scale; lvl++) {
        uint32_t u;
        if ((lvl & 1) == 0) { r = splitmix64(ctr + (lvl >> 1)); u = (uint32_t) r; }
        else u = (uint32_t) (r >> 32);
        uint32_t sbit = u >= 3264175144u;
        uint32_t dbit = (u >= 2448131358u && u < 3264175144u) || (u >= 4080218931u);
        s = (s << 1) | sbit;
        d = (d << 1) | dbit;
    }
    src = (uint32_t) (rmat_permute(P, s) ^ P.zero);
    dst = (uint32_t) (rmat_permute(P, d) ^ P.zero);
    w = P.weighted ? (uint32_t) ((splitmix64(ctr + 31) >> 33) % 128ull) + 1u : 1u;
}

__global__ void k_rmat_generate(RmatParams P, uint64_t first, uint64_t n, uint32_t* out) {
    const int stride = P.weighted ? 3 : 2;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        uint32_t s, d, w;
        rmat_edge(P, first + i, s, d, w);
        out[i * stride] = s;
        out[i * stride + 1] = d;
        if (P.weighted) out[i * stride + 2] = w;
    }
}

// ---------------------------------------------------------------------------------------------
// ingest: flags -> (row, col, w) entries -> bitmaps, ownership filter, sort keys
// ---------------------------------------------------------------------------------------------
struct IngestParams {
    uint32_t th, p;
    int self_loops, acyclic, transpose, directed, weighted;
    int bw;                          // bits of an in-tile coordinate
    const int32_t* tile_local;       // [p*p] local tile index (row order) or -1
    uint8_t* I_all;                  // [p*th] any entry in this row    (count pass only)
    uint8_t* J_all;                  // [p*th] any entry in this column (count pass only)
    uint32_t* rdeg_all;              // [p*th] entries in the vertex's row    (count pass only; raw records, before dedup)
    uint32_t* cdeg_all;              // [p*th] entries in the vertex's column (count pass only)
    unsigned long long* counters;    // [0] owned entries, [1] append cursor, [2] out-of-range records
    uint64_t* keys;                  // fill pass
    uint32_t* wts;                   // fill pass, weighted
    // partitioned ingest (every rank holds a share of the records): route every entry to the rank that owns its tile
    const int32_t* tile_rank;        // [p*p] owner of every tile (after the leader swap)
    unsigned long long* dest_count;  // [nranks] entries per destination (count pass) / running cursors (fill pass)
    const unsigned long long* dest_off;   // [nranks] first slot of every destination in the send buffer
    uint32_t* sendbuf;               // {row, col[, w]} records grouped by destination
    int no_marks;                    // the records are already-routed entries: marks and degrees came from the all-reduce
};

template <bool FILL>
__device__ __forceinline__ void ingest_emit(const IngestParams& Q, uint32_t r, uint32_t c, uint32_t w, bool valid,
                                            unsigned long long& local_count) {
    int t = -1;
    uint32_t rg = 0, cg = 0;
    if (valid) {
        rg = r / Q.th;
        cg = c / Q.th;
        t = Q.tile_local[rg * Q.p + cg];
        if (!FILL && !Q.no_marks) { Q.I_all[r] = 1; Q.J_all[c] = 1; atomicAdd(Q.rdeg_all + r, 1u); atomicAdd(Q.cdeg_all + c, 1u); }
    }
    const bool own = valid && t >= 0;
    if (!FILL) {
        local_count += own ? 1 : 0;
    } else {
        // warp-aggregated append (all 32 lanes reach this point together)
        const unsigned ballot = __ballot_sync(0xffffffffu, own);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&Q.counters[1], (unsigned long long) __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (own) {
                const unsigned long long pos = base + __popc(ballot & ((1u << lane) - 1));
                Q.keys[pos] = ((uint64_t) t << (2 * Q.bw)) | ((uint64_t) (c - cg * Q.th) << Q.bw) | (uint64_t) (r - rg * Q.th);
                if (Q.weighted) Q.wts[pos] = w;
            }
        }
    }
}

// Partitioned ingest (every rank holds a SHARE of the records, as Graph::parread_binary reads 1/p of the file,
// src/mat/graph.hpp:307-335): apply the per-edge flags (:337-356) to this rank's share and route every resulting entry
// to the rank that owns its tile (Matrix::distribute, src/mat/matrix.hpp:692-810).  Count pass: entries per destination +
// this share's row / column marks and degrees (all-reduced over the world afterwards, where the reference OR-reduces
// its bitmaps along the groups, :973-1083).  Fill pass: {row, col[, w]} records into the send buffer, grouped by
// destination (the order inside a group is irrelevant: the owner sorts).  Counters are bumped once per warp and
// destination (__match_any), not once per entry.
template <bool FILL>
__device__ __forceinline__ void route_emit(const IngestParams& Q, uint32_t r, uint32_t c, uint32_t w, bool valid) {
    const int dest = valid ? Q.tile_rank[(r / Q.th) * Q.p + (c / Q.th)] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, dest);
    if (!valid) return;
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(Q.dest_count + dest, (unsigned long long) __popc(peers));
    if (!FILL) {
        Q.I_all[r] = 1; Q.J_all[c] = 1; atomicAdd(Q.rdeg_all + r, 1u); atomicAdd(Q.cdeg_all + c, 1u);
    } else {
        base = __shfl_sync(peers, base, leader);
        const unsigned long long pos = Q.dest_off[dest] + base + __popc(peers & ((1u << lane) - 1));
        const int rec = Q.weighted ? 3 : 2;
        Q.sendbuf[pos * rec] = r; Q.sendbuf[pos * rec + 1] = c;
        if (Q.weighted) Q.sendbuf[pos * rec + 2] = w;
    }
}
template <bool FILL>
__global__ void __launch_bounds__(256) k_route(IngestParams Q, RmatParams G, const uint32_t* triples, uint64_t first, uint64_t n) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) / 32 * 32;
    const uint32_t limit = Q.p * Q.th;
    unsigned long long bad = 0;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool valid = i < n;
        uint32_t r = 0, c = 0, w = 1;
        if (valid) {
            if (G.enabled) rmat_edge(G, first + i, r, c, w);
            else if (Q.weighted) { r = triples[i * 3]; c = triples[i * 3 + 1]; w = triples[i * 3 + 2]; }
            else { const uint2 rc = reinterpret_cast<const uint2*>(triples)[i]; r = rc.x; c = rc.y; }
            if (r >= limit || c >= limit) { bad++; valid = false; }
        }
        if (valid && r == c && !Q.self_loops) valid = false;                 // graph.hpp:339-342
        if (valid && Q.acyclic && c < r) { uint32_t x = r; r = c; c = x; }   // :344-347
        if (valid && Q.transpose) { uint32_t x = r; r = c; c = x; }          // :349-350
        route_emit<FILL>(Q, r, c, w, valid);                                 // :352
        if (!Q.directed) route_emit<FILL>(Q, c, r, w, valid);                // :354-357
    }
    if (!FILL && bad) atomicAdd(&Q.counters[2], bad);
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_ingest(IngestParams Q, RmatParams G, const uint32_t* triples, uint64_t first, uint64_t n) {
    unsigned long long local_count = 0, bad = 0;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) / 32 * 32;
    const uint32_t limit = Q.p * Q.th;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool valid = i < n;
        uint32_t r = 0, c = 0, w = 1;
        if (valid) {
            if (G.enabled) rmat_edge(G, first + i, r, c, w);
            else if (Q.weighted) { r = triples[i * 3]; c = triples[i * 3 + 1]; w = triples[i * 3 + 2]; }
            else { const uint2 rc = reinterpret_cast<const uint2*>(triples)[i]; r = rc.x; c = rc.y; }
            if (r >= limit || c >= limit) { bad++; valid = false; }
        }
        if (valid && r == c && !Q.self_loops) valid = false;          // graph.hpp:339-342
        if (valid && Q.acyclic && c < r) { uint32_t x = r; r = c; c = x; }   // :344-347
        if (valid && Q.transpose) { uint32_t x = r; r = c; c = x; }          // :349-350
        ingest_emit<FILL>(Q, r, c, w, valid, local_count);                   // :352
        if (!Q.directed) ingest_emit<FILL>(Q, c, r, w, valid, local_count);  // :354-357
    }
    if (!FILL) {
        typedef cub::BlockReduce<unsigned long long, 256> BR;
        __shared__ typename BR::TempStorage tmp;
        unsigned long long tot = BR(tmp).Sum(local_count);
        __syncthreads();
        unsigned long long totbad = BR(tmp).Sum(bad);
        if (threadIdx.x == 0) {
            if (tot) atomicAdd(&Q.counters[0], tot);
            if (totbad) atomicAdd(&Q.counters[2], totbad);
        }
    }
}

__global__ void k_u8_to_u32(const uint8_t* in, uint32_t* out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) out[i] = in[i];
}

// segment slot maps from the global bitmap + its exclusive scan
__global__ void k_seg_maps(const uint8_t* bits_all, const uint32_t* scan_all, uint32_t seg_base, uint32_t th,
                           uint8_t* bits, uint32_t* prefix, uint32_t* ids) {
    const uint32_t s0 = scan_all[seg_base];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        const uint8_t b = bits_all[seg_base + i];
        bits[i] = b;
        const uint32_t k = scan_all[seg_base + i] - s0;
        prefix[i] = b ? k : 0;                    // matrix.hpp:1031-1040
        if (b) ids[k] = i;                        // compressed_column.hpp:399-416 (JC / IR)
    }
}

// hot order: key = (class, ~degree, local id) for vertices with a non-empty row or column, all-ones otherwise.
// class (only on _TCSC_CF_ graphs, else 0 for everyone): 0 regular, 1 source row, 2 sink column.
__global__ void k_hot_keys(const uint8_t* __restrict__ I, const uint8_t* __restrict__ J, const uint32_t* __restrict__ rdeg, const uint32_t* __restrict__ cdeg, uint32_t th, int cf,
                           uint64_t* __restrict__ keys, uint8_t* __restrict__ cls, unsigned int* __restrict__ counts) {
    unsigned int nreg = 0, nsrc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        const bool r = I[i], c = J[i];
        const uint32_t d = (uint32_t) min((uint64_t) rdeg[i] + cdeg[i], (uint64_t) 0x3fffffffu);
        const uint64_t k = (cf && !(r && c)) ? (r ? 1ull : 2ull) : 0ull;
        keys[i] = (r || c) ? ((k << 62) | ((uint64_t) (0x3fffffffu - d) << 32) | i) : ~0ull;
        if (cls) cls[i] = (uint8_t) (r && c ? 1 : r ? 2 : c ? 3 : 0);       // matrix.hpp:1135-1144
        nreg += r && c; nsrc += r && !c;
    }
    if (nreg) atomicAdd(counts, nreg);
    if (nsrc) atomicAdd(counts + 1, nsrc);
}
__global__ void k_hot_count(const uint64_t* __restrict__ keys, uint32_t th, uint32_t* __restrict__ n) {
    uint32_t lo = 0, hi = th;                      // first all-ones key (degrees are >= 1, so real keys are smaller)
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < 0xffffffff00000000ull) lo = mid + 1; else hi = mid;
    }
    *n = lo;
}
__global__ void k_hot_finish(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ ids, uint32_t* __restrict__ pos) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t id = (uint32_t) keys[k];
        ids[k] = id;
        pos[id] = k;
    }
}

__global__ void k_tile_bounds(const uint64_t* keys, uint64_t n, int ntiles, int shift, uint64_t* bounds) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const uint64_t target = (uint64_t) t << shift;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    bounds[t] = lo;
}

// IA[e] = IV[row]                                      compressed_column.hpp:392
__global__ void k_tile_IA(const uint64_t* keys, uint64_t n, uint64_t row_mask, const uint32_t* IV, uint32_t* IA) {
    for (uint64_t e = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; e < n; e += (uint64_t) gridDim.x * blockDim.x)
        IA[e] = IV[(uint32_t) (keys[e] & row_mask)];
}

// JA[j] = first entry whose column is >= JC[j] (empty columns get JA[j] == JA[j+1])   :381-398
__global__ void k_tile_JA(const uint64_t* keys, uint64_t n, uint64_t tile_prefix, int bw, const uint32_t* JC, uint32_t ncols, uint32_t* JA) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j <= ncols; j += gridDim.x * blockDim.x) {
        if (j == ncols) { JA[j] = (uint32_t) n; continue; }
        const uint64_t target = tile_prefix | ((uint64_t) JC[j] << bw);
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        JA[j] = (uint32_t) lo;
    }
}

__global__ void k_max_col_entries(const uint32_t* __restrict__ JA, uint32_t ncols, unsigned int* __restrict__ out) {
    unsigned int m = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x) m = max(m, JA[j + 1] - JA[j]);
    if (m) atomicMax(out, m);
}

// chunk_col[k] = column that holds edge k*GT_PUSH_CHUNK (last column with JA[c] <= e)
__global__ void k_chunk_cols(const uint32_t* JA, uint32_t ncols, uint64_t nnz, uint32_t nchunks, uint32_t* chunk_col) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k <= nchunks; k += gridDim.x * blockDim.x) {
        if (k == nchunks) { chunk_col[k] = ncols; continue; }
        const uint64_t e = (uint64_t) k * GT_PUSH_CHUNK;
        uint32_t lo = 0, hi = ncols;       // upper_bound over JA[0..ncols)
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t) JA[mid] <= e) lo = mid + 1; else hi = mid;
        }
        chunk_col[k] = lo - 1;
    }
}

static inline int grid_for(uint64_t n, int block, int sm_count, int per_sm = 8) {
    uint64_t g = (n + block - 1) / block;
    uint64_t cap = (uint64_t) sm_count * per_sm;
    return (int) std::max<uint64_t>(1, std::min(g, cap));
}

static int bits_for(uint32_t v) {   // bits needed to represent values in [0, v)
    int b = 1;
    while (b < 32 && (1ull << b) < v) b++;
    return b;
}

// ---------------------------------------------------------------------------------------------
// _TCSC_CF_: TCSC_CF_BASE::populate on the device (src/ds/compressed_column.hpp:671-1114)
// ---------------------------------------------------------------------------------------------
// F[e] = 1 iff entry e sits in a source row (row non-empty, column empty) of the tile's row group
__global__ void k_cf_flags(const uint32_t* __restrict__ IA, uint64_t nnz, const uint32_t* __restrict__ IR, const uint8_t* __restrict__ rcls, uint8_t* __restrict__ F) {
    for (uint64_t e = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; e <= nnz; e += (uint64_t) gridDim.x * blockDim.x)
        F[e] = e < nnz ? (rcls[IR[IA[e]]] == 2) : 0;
}
struct U8ToU32 { __host__ __device__ uint32_t operator()(const uint8_t& v) const { return v; } };
struct U32ToU64 { __host__ __device__ unsigned long long operator()(const uint32_t& v) const { return v; } };
__device__ __forceinline__ uint32_t cf_col_of(const uint32_t* __restrict__ JA, uint32_t ncols, uint64_t e) {
    uint32_t lo = 0, hi = ncols;                      // upper_bound(JA, e) - 1 = the column that holds entry e
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t) JA[mid] <= e) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}
// Stable partition permutation of every column: P[JA[j] + rho] = position of the column's rho-th regular entry,
// P[JA[j] + B + sigma] = position of its sigma-th source entry (B = number of regular entries); S = exclusive scan of F.
__global__ void k_cf_partition(const uint32_t* __restrict__ JA, uint32_t ncols, uint64_t nnz, const uint8_t* __restrict__ F, const uint32_t* __restrict__ S,
                               uint32_t* __restrict__ P) {
    for (uint64_t e = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; e < nnz; e += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t j = cf_col_of(JA, ncols, e);
        const uint32_t b = JA[j], en = JA[j + 1];
        const uint32_t nsrc = S[en] - S[b], B = (en - b) - nsrc, sigma = S[e] - S[b];
        if (F[e]) P[b + B + sigma] = (uint32_t) e; else P[b + ((uint32_t) e - b - sigma)] = (uint32_t) e;
    }
}
// "Moving source rows to the end" (:671-708) in closed form.  The reference walks the column's source entries in
// ascending position and swaps each with the LAST regular entry still to its right; with B regular entries that pairs
// the i-th source among the first B positions with the regular entry of rank B-1-i, and leaves everything else alone.
__global__ void k_cf_swap(const uint32_t* __restrict__ JA, uint32_t ncols, uint64_t nnz, const uint8_t* __restrict__ F, const uint32_t* __restrict__ S,
                          const uint32_t* __restrict__ P, const uint32_t* __restrict__ IA, const uint32_t* __restrict__ A,
                          uint32_t* __restrict__ IA_out, uint32_t* __restrict__ A_out) {
    for (uint64_t e = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; e < nnz; e += (uint64_t) gridDim.x * blockDim.x) {
        if (!F[e]) continue;
        const uint32_t j = cf_col_of(JA, ncols, e);
        const uint32_t b = JA[j], en = JA[j + 1];
        const uint32_t nsrc = S[en] - S[b], B = (en - b) - nsrc, sigma = S[e] - S[b];
        if ((uint32_t) e - b >= B) continue;            // already in the tail
        const uint32_t t = P[b + B - 1 - sigma];
        IA_out[e] = IA[t]; IA_out[t] = IA[e];
        if (A) { A_out[e] = A[t]; A_out[t] = A[e]; }
    }
}
// per compressed column: which of the four lists it joins (bit k = kind k), and its source-entry count
__global__ void k_cf_col_kinds(const uint32_t* __restrict__ JA, uint32_t ncols, const uint32_t* __restrict__ S, const uint32_t* __restrict__ JC,
                               const uint8_t* __restrict__ ccls, uint8_t* __restrict__ kind0, uint8_t* __restrict__ kind1, uint8_t* __restrict__ kind2,
                               uint8_t* __restrict__ kind3, unsigned long long* __restrict__ snk_src_edges) {
    unsigned long long local = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x) {
        const uint32_t b = JA[j], en = JA[j + 1], len = en - b, nsrc = S[en] - S[b];
        const uint8_t c = ccls[JC[j]];
        const bool local_col = len > 0, reg = c == 1, snk = c == 3;
        kind0[j] = local_col && reg && nsrc < len;      // :749-833
        kind1[j] = local_col && snk && nsrc < len;      // :862-946
        kind2[j] = local_col && reg && nsrc > 0;        // :949-1022
        kind3[j] = local_col && snk && nsrc > 0;        // :1025-1108
        if (local_col && snk) local += nsrc;            // NC_SRC_R_SNK_C counts EDGES (:1046-1049)
    }
    if (local) atomicAdd(snk_src_edges, local);
}
__global__ void k_cf_pairs(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ S, const uint32_t* __restrict__ JCk, uint32_t n, int kind,
                           uint32_t* __restrict__ JAk) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t j = JCk[i];
        const uint32_t b = JA[j], en = JA[j + 1], nsrc = S[en] - S[b];
        uint32_t lo, hi;
        if (kind <= 1) { lo = b; hi = en - nsrc; }      // regular rows of the column
        else if (kind == 2) { lo = en - nsrc; hi = en; }
        else { lo = b + nsrc; hi = en; }                // the reference's own start (:1094)
        JAk[2 * i] = lo; JAk[2 * i + 1] = hi;
    }
}
__global__ void k_cf_owned(const uint8_t* __restrict__ cls, uint32_t th, uint8_t* __restrict__ reg, uint8_t* __restrict__ src, uint8_t* __restrict__ snk) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        const uint8_t c = cls[i];
        reg[i] = c == 1; src[i] = c == 2; snk[i] = c == 3;
    }
}

static uint32_t select_flagged(gt_ctx* ctx, const uint8_t* flags, uint32_t n, DevBuf<uint32_t>& out, uint32_t min_alloc = 0) {
    cudaStream_t st = ctx->stream;
    DevBuf<uint32_t> tmp_out; tmp_out.alloc(std::max<uint32_t>(n, 1));
    DevBuf<unsigned int> d_n; d_n.alloc(1);
    size_t tb = 0;
    cub::CountingInputIterator<uint32_t> it(0);
    GT_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, flags, tmp_out.p, d_n.p, (int) n, st));
    DevBuf<uint8_t> tmp; tmp.alloc(tb);
    GT_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, it, flags, tmp_out.p, d_n.p, (int) n, st));
    unsigned int h = 0;
    GT_CUDA(cudaMemcpyAsync(&h, d_n.p, 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));
    out.alloc(std::max<uint32_t>(std::max(h, min_alloc), 1));
    GT_CUDA(cudaMemsetAsync(out.p, 0, out.bytes(), st));
    if (h) GT_CUDA(cudaMemcpyAsync(out.p, tmp_out.p, (size_t) h * 4, cudaMemcpyDeviceToDevice, st));
    GT_CUDA(cudaStreamSynchronize(st));
    ctx->kernel_launches += 2;
    return h;
}

static void build_cf(gt_graph* g) {
    gt_ctx* ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t th = g->lay.info.tile_height;
    // classify_vertices of the owned segment
    {
        const uint8_t* cls = g->cls[g->hot_of_row_slot[g->lay.info.accu_segment_row]].p;
        DevBuf<uint8_t> f; f.alloc(3 * (size_t) th);
        k_cf_owned<<<grid_for(th, 256, ctx->sm_count), 256, 0, st>>>(cls, th, f.p, f.p + th, f.p + 2 * (size_t) th);
        ctx->kernel_launches++;
        g->cf_owned.nreg = select_flagged(ctx, f.p, th, g->cf_owned.regular_rows);
        g->cf_owned.nsrc = select_flagged(ctx, f.p + th, th, g->cf_owned.source_rows);
        g->cf_owned.nsnk = select_flagged(ctx, f.p + 2 * (size_t) th, th, g->cf_owned.sink_columns);
    }
    g->cf_tiles.resize(g->tiles.size());
    for (size_t k = 0; k < g->tiles.size(); k++) {
        const Tile& T = g->tiles[k];
        CfTile& C = g->cf_tiles[k];
        if (!T.nnz) continue;
        const SegMaps& R = g->rows[T.row_slot];
        const SegMaps& Cs = g->cols[T.col_slot];
        const uint8_t* rcls = g->cls[g->hot_of_row_slot[T.row_slot]].p;
        const uint8_t* ccls = g->cls[g->hot_of_col_slot[T.col_slot]].p;
        uint32_t* IA = g->IA_pool.p + T.offset;
        uint32_t* A = g->weighted ? g->A_pool.p + T.offset : nullptr;
        const uint32_t ncols = Cs.nnz;
        DevBuf<uint8_t> F; F.alloc(T.nnz + 1);
        DevBuf<uint32_t> S; S.alloc(T.nnz + 1);
        k_cf_flags<<<grid_for(T.nnz + 1, 256, ctx->sm_count, 16), 256, 0, st>>>(IA, T.nnz, R.ids.p, rcls, F.p);
        {
            size_t tb = 0;
            cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> in(F.p, U8ToU32());
            GT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, S.p, (int64_t) T.nnz + 1, st));
            DevBuf<uint8_t> tmp; tmp.alloc(tb);
            GT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, S.p, (int64_t) T.nnz + 1, st));
            GT_CUDA(cudaStreamSynchronize(st));
        }
        {   // IA (and A) in the reference's order
            DevBuf<uint32_t> P, IA2, A2;
            P.alloc(T.nnz); IA2.alloc(T.nnz);
            if (A) A2.alloc(T.nnz);
            k_cf_partition<<<grid_for(T.nnz, 256, ctx->sm_count, 16), 256, 0, st>>>(T.JA.p, ncols, T.nnz, F.p, S.p, P.p);
            GT_CUDA(cudaMemcpyAsync(IA2.p, IA, T.nnz * 4, cudaMemcpyDeviceToDevice, st));
            if (A) GT_CUDA(cudaMemcpyAsync(A2.p, A, T.nnz * 4, cudaMemcpyDeviceToDevice, st));
            k_cf_swap<<<grid_for(T.nnz, 256, ctx->sm_count, 16), 256, 0, st>>>(T.JA.p, ncols, T.nnz, F.p, S.p, P.p, IA, A, IA2.p, A ? A2.p : nullptr);
            GT_CUDA(cudaMemcpyAsync(IA, IA2.p, T.nnz * 4, cudaMemcpyDeviceToDevice, st));
            if (A) GT_CUDA(cudaMemcpyAsync(A, A2.p, T.nnz * 4, cudaMemcpyDeviceToDevice, st));
            GT_CUDA(cudaGetLastError());
            GT_CUDA(cudaStreamSynchronize(st));
        }
        // the four lists
        DevBuf<uint8_t> kinds; kinds.alloc(4 * (size_t) std::max<uint32_t>(ncols, 1));
        DevBuf<unsigned long long> d_e; d_e.alloc(1);
        GT_CUDA(cudaMemsetAsync(d_e.p, 0, 8, st));
        k_cf_col_kinds<<<grid_for(ncols, 256, ctx->sm_count), 256, 0, st>>>(T.JA.p, ncols, S.p, Cs.ids.p, ccls, kinds.p, kinds.p + ncols, kinds.p + 2 * (size_t) ncols,
                                                                          kinds.p + 3 * (size_t) ncols, d_e.p);
        unsigned long long snk_src_edges = 0;
        GT_CUDA(cudaMemcpyAsync(&snk_src_edges, d_e.p, 8, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        for (int kind = 0; kind < 4; kind++) {
            const uint32_t want = kind == 3 ? (uint32_t) snk_src_edges : 0;      // NC_SRC_R_SNK_C is an edge count; the tail stays zero
            C.filled[kind] = select_flagged(ctx, kinds.p + (size_t) kind * ncols, ncols, C.JC[kind], want);
            C.NC[kind] = kind == 3 ? want : C.filled[kind];
            C.JA[kind].alloc(2 * (size_t) std::max<uint32_t>(C.NC[kind], 1));
            GT_CUDA(cudaMemsetAsync(C.JA[kind].p, 0, C.JA[kind].bytes(), st));
            if (C.filled[kind])
                k_cf_pairs<<<grid_for(C.filled[kind], 256, ctx->sm_count), 256, 0, st>>>(T.JA.p, S.p, C.JC[kind].p, C.filled[kind], kind, C.JA[kind].p);
        }
        ctx->kernel_launches += 8;
        GT_CUDA(cudaGetLastError());
        GT_CUDA(cudaStreamSynchronize(st));
    }
}

// ---------------------------------------------------------------------------------------------
// `partitioned`: `triples` / the generator range [gen_first, gen_first + ntriples) is THIS RANK'S SHARE of the records.
static gt_graph* build(gt_ctx* ctx, const void* triples, uint64_t ntriples, int weighted, int on_device,
                       const RmatParams* gen, uint32_t nvertices, const gt_graph_flags* flags, int compression,
                       int partitioned = 0, uint64_t gen_first = 0) {
    GT_REQUIRE(ctx, "gt_graph_build: ctx is NULL");
    GT_REQUIRE(flags, "gt_graph_build: flags is NULL");
    GT_REQUIRE(compression == GT_TCSC || compression == GT_TCSC_CF, "gt_graph_build: only _TCSC_ / _TCSC_CF_ are GPU formats");
    GT_REQUIRE(gen || triples || ntriples == 0, "gt_graph_build: triples is NULL");
    GT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::unique_ptr<gt_graph> g(new gt_graph());
    g->ctx = ctx;
    g->flags = *flags;
    g->weighted = weighted;
    g->compression = compression;
    g->nvertices = nvertices;
    g->nedges_input = ntriples;
    g->lay = make_layout(nvertices, ctx->nranks, ctx->rank);
    const Layout& L = g->lay;
    const uint32_t p = L.info.nranks, th = L.info.tile_height;
    const uint64_t nall = (uint64_t) p * th;
    const int ntiles = (int) L.local_tiles_row_order.size();
    const int bw = bits_for(th);
    int bt = 0;
    while ((1 << bt) < ntiles) bt++;
    GT_REQUIRE(bt + 2 * bw <= 64, "gt_graph_build: tile coordinates do not fit a 64-bit sort key");

    // ownership table
    std::vector<int32_t> tile_local((size_t) p * p, -1);
    for (int k = 0; k < ntiles; k++) tile_local[L.local_tiles_row_order[k]] = k;
    DevBuf<int32_t> d_tile_local; d_tile_local.alloc(tile_local.size());
    GT_CUDA(cudaMemcpyAsync(d_tile_local.p, tile_local.data(), tile_local.size() * 4, cudaMemcpyHostToDevice, st));

    DevBuf<uint8_t> I_all, J_all;
    I_all.alloc(nall); J_all.alloc(nall);
    GT_CUDA(cudaMemsetAsync(I_all.p, 0, nall, st));
    GT_CUDA(cudaMemsetAsync(J_all.p, 0, nall, st));
    DevBuf<uint32_t> rdeg_all, cdeg_all;
    rdeg_all.alloc(nall); cdeg_all.alloc(nall);
    GT_CUDA(cudaMemsetAsync(rdeg_all.p, 0, nall * 4, st));
    GT_CUDA(cudaMemsetAsync(cdeg_all.p, 0, nall * 4, st));
    DevBuf<unsigned long long> counters; counters.alloc(4);
    GT_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned long long), st));

    IngestParams Q{};
    Q.th = th; Q.p = p;
    Q.self_loops = flags->self_loops; Q.acyclic = flags->acyclic; Q.transpose = flags->transpose; Q.directed = flags->directed;
    Q.weighted = weighted; Q.bw = bw;
    Q.tile_local = d_tile_local.p; Q.I_all = I_all.p; Q.J_all = J_all.p; Q.rdeg_all = rdeg_all.p; Q.cdeg_all = cdeg_all.p; Q.counters = counters.p;
    RmatParams G{};
    if (gen) G = *gen;

    // host input is staged through a bounded device buffer
    const uint64_t rec_words = weighted ? 3 : 2;
    const uint64_t CH = 1ull << 26;
    DevBuf<uint32_t> stage;
    if (!gen && !on_device && ntriples) stage.alloc(std::min(CH, ntriples) * rec_words);

    // one pass over the records: `route` = the partitioned pre-stage (k_route), else the ingest proper (k_ingest)
    auto run_pass = [&](bool fill, bool route = false) {
        for (uint64_t first = 0; first < ntriples; first += CH) {
            const uint64_t n = std::min(CH, ntriples - first);
            const uint32_t* src = nullptr;
            if (!G.enabled) {
                if (on_device) src = (const uint32_t*) triples + first * rec_words;
                else {
                    GT_CUDA(cudaMemcpyAsync(stage.p, (const uint32_t*) triples + first * rec_words, n * rec_words * 4, cudaMemcpyHostToDevice, st));
                    src = stage.p;
                }
            }
            const int grid = grid_for(n, 256, ctx->sm_count, 16);
            if (route) {
                if (fill) k_route<true><<<grid, 256, 0, st>>>(Q, G, src, gen_first + first, n);
                else k_route<false><<<grid, 256, 0, st>>>(Q, G, src, gen_first + first, n);
            } else {
                if (fill) k_ingest<true><<<grid, 256, 0, st>>>(Q, G, src, gen_first + first, n);
                else k_ingest<false><<<grid, 256, 0, st>>>(Q, G, src, gen_first + first, n);
            }
            ctx->kernel_launches++;
            GT_CUDA(cudaGetLastError());
            if (!G.enabled && !on_device) GT_CUDA(cudaStreamSynchronize(st));   // stage is reused
        }
    };

    // ---- partitioned ingest: flags on the share, entries to their owners, marks and degrees all-reduced ---------------
    DevBuf<uint32_t> routed;                               // the entries this rank owns, {row, col[, w]} records (NCCL exchange) ...
    PeerWindow* route_win = nullptr;                       // ... or the peer window they were copied into
    struct WinGuard { gt_ctx* c; PeerWindow*& w; ~WinGuard() { if (w) { peer_window_destroy(c, w); w = nullptr; } } } route_guard{ctx, route_win};
    if (partitioned && ctx->nranks > 1) {
        GT_REQUIRE(ctx->comm, "gt_graph_build_partitioned: the context has no communicator");
        const int nr = ctx->nranks;
        DevBuf<int32_t> d_tile_rank; d_tile_rank.alloc(L.tile_rank.size());
        GT_CUDA(cudaMemcpyAsync(d_tile_rank.p, L.tile_rank.data(), L.tile_rank.size() * 4, cudaMemcpyHostToDevice, st));
        DevBuf<unsigned long long> dest_count, dest_off, matrix;
        dest_count.alloc(nr); dest_off.alloc(nr); matrix.alloc((size_t) nr * (nr + 1));
        GT_CUDA(cudaMemsetAsync(dest_count.p, 0, (size_t) nr * 8, st));
        Q.tile_rank = d_tile_rank.p; Q.dest_count = dest_count.p; Q.dest_off = dest_off.p;
        run_pass(false, true);
        // what every rank sends to every rank (+ its share of the record count), one all-gather
        std::vector<unsigned long long> h_matrix((size_t) nr * (nr + 1));
        {
            unsigned long long* mine = matrix.p + (size_t) ctx->rank * (nr + 1);
            GT_CUDA(cudaMemcpyAsync(mine, dest_count.p, (size_t) nr * 8, cudaMemcpyDeviceToDevice, st));
            const unsigned long long share = ntriples;
            GT_CUDA(cudaMemcpyAsync(mine + nr, &share, 8, cudaMemcpyHostToDevice, st));
            comm_allgather_inplace(ctx->comm, COMM_WORLD, matrix.p, (size_t) nr + 1, CT_U64, st);
            GT_CUDA(cudaMemcpyAsync(h_matrix.data(), matrix.p, h_matrix.size() * 8, cudaMemcpyDeviceToHost, st));
        }
        // marks, degrees and the bad-record count of the WHOLE list (every rank needs the rows/columns of all its segments)
        comm_allreduce(ctx->comm, COMM_WORLD, I_all.p, I_all.p, nall, CT_U8, CO_MAX, st);
        comm_allreduce(ctx->comm, COMM_WORLD, J_all.p, J_all.p, nall, CT_U8, CO_MAX, st);
        comm_allreduce(ctx->comm, COMM_WORLD, rdeg_all.p, rdeg_all.p, nall, CT_U32, CO_SUM, st);
        comm_allreduce(ctx->comm, COMM_WORLD, cdeg_all.p, cdeg_all.p, nall, CT_U32, CO_SUM, st);
        comm_allreduce(ctx->comm, COMM_WORLD, counters.p + 2, counters.p + 2, 1, CT_U64, CO_SUM, st);
        GT_CUDA(cudaStreamSynchronize(st));
        // the plan (host arithmetic, gt_layout.cpp; entries -> bytes here)
        std::vector<uint64_t> cmat((size_t) nr * nr);
        uint64_t nrecords = 0;
        for (int r = 0; r < nr; r++) {
            for (int q = 0; q < nr; q++) cmat[(size_t) r * nr + q] = h_matrix[(size_t) r * (nr + 1) + q];
            nrecords += h_matrix[(size_t) r * (nr + 1) + nr];
        }
        const RoutePlan plan = make_route_plan(nr, ctx->rank, cmat.data());
        const uint64_t rec_bytes = rec_words * 4, nrecv = plan.nrecv;
        std::vector<unsigned long long> h_off(plan.send_offset.begin(), plan.send_offset.end());
        g->nedges_input = nrecords;
        DevBuf<uint32_t> sendbuf; sendbuf.alloc(std::max<uint64_t>(plan.nsend, 1) * rec_words);
        GT_CUDA(cudaMemcpyAsync(dest_off.p, h_off.data(), (size_t) nr * 8, cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemsetAsync(dest_count.p, 0, (size_t) nr * 8, st));
        Q.sendbuf = sendbuf.p;
        run_pass(true, true);
        // The exchange.  Where the ranks can map each other's memory (one node, NVLink), every rank copies its blocks straight
        // into the owners' receive buffers — a world peer window sized for the largest receiver — and one world fence tells
        // everybody that all copies have landed; a block from rank r sits behind the blocks of the ranks before r, the
        // order the receive offsets assume.  No NCCL point-to-point connections are built (the first grouped
        // send/recv of a job costs ~6 s at 4-8 ranks).  Otherwise: one grouped ncclSend/ncclRecv exchange.
        const char* pe = getenv("GT_PEER");
        route_win = (pe && atoi(pe) == 0) ? nullptr : peer_window_create(ctx, COMM_WORLD, std::max<uint64_t>(plan.max_recv, 1) * rec_bytes);
        const uint32_t* received = nullptr;
        auto mine = [&](int q) { return cmat[(size_t) ctx->rank * nr + q] * rec_bytes; };      // bytes this rank holds for q
        if (route_win) {
            for (int j = 0; j < nr; j++) {
                const int q = (ctx->rank + j) % nr;                    // start with the own block, then walk the ring
                if (!mine(q)) continue;
                GT_CUDA(cudaMemcpyAsync(route_win->remote[q] + plan.remote_offset[q] * rec_bytes, (const uint8_t*) sendbuf.p + plan.send_offset[q] * rec_bytes,
                                        mine(q), cudaMemcpyDefault, st));
            }
            peer_fence_world(ctx, st);
            received = (const uint32_t*) route_win->local;
        } else {
            routed.alloc(std::max<uint64_t>(nrecv, 1) * rec_words);
            std::vector<uint64_t> scount(nr), sdispl(nr), rcount(nr), rdispl(nr);
            for (int q = 0; q < nr; q++) {
                scount[q] = mine(q); sdispl[q] = plan.send_offset[q] * rec_bytes;
                rcount[q] = cmat[(size_t) q * nr + ctx->rank] * rec_bytes; rdispl[q] = plan.recv_offset[q] * rec_bytes;
            }
            if (scount[ctx->rank])
                GT_CUDA(cudaMemcpyAsync((uint8_t*) routed.p + rdispl[ctx->rank], (const uint8_t*) sendbuf.p + sdispl[ctx->rank], scount[ctx->rank], cudaMemcpyDeviceToDevice, st));
            scount[ctx->rank] = rcount[ctx->rank] = 0;
            comm_alltoallv_bytes(ctx->comm, (const uint8_t*) sendbuf.p, scount.data(), sdispl.data(), (uint8_t*) routed.p, rcount.data(), rdispl.data(), st);
            received = routed.p;
        }
        GT_CUDA(cudaStreamSynchronize(st));
        // from here on: the ordinary build over the routed entries (flags already applied, marks already complete)
        triples = received; ntriples = nrecv; on_device = 1;
        G.enabled = 0;
        Q.self_loops = 1; Q.acyclic = 0; Q.transpose = 0; Q.directed = 1; Q.no_marks = 1;
        gen_first = 0;
        stage.release();
    }

    run_pass(false);
    unsigned long long h_counters[4];
    GT_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaStreamSynchronize(st));
    if (h_counters[2])
        throw Error(GT_ERR_INVALID, "gt_graph_build: " + std::to_string(h_counters[2]) + " records name a vertex id beyond nvertices");
    uint64_t nloc = h_counters[0];

    DevBuf<uint64_t> keys, keys_alt;
    DevBuf<uint32_t> wts, wts_alt;
    keys.alloc(nloc); keys_alt.alloc(nloc);
    if (weighted) { wts.alloc(nloc); wts_alt.alloc(nloc); }
    Q.keys = keys.p; Q.wts = wts.p;
    run_pass(true);
    GT_CUDA(cudaStreamSynchronize(st));
    stage.release();
    routed.release();
    if (route_win) {                                       // every rank has consumed its window; none is written any more
        peer_fence_world(ctx, st);
        GT_CUDA(cudaStreamSynchronize(st));
        peer_window_destroy(ctx, route_win);
        route_win = nullptr;
    }

    // ---- sort (+ dedup) ---------------------------------------------------------------------
    uint64_t* sorted_keys = keys.p;
    uint32_t* sorted_wts = wts.p;
    if (nloc) {
        DevBuf<uint8_t> tmp;
        size_t tmp_bytes = 0;
        const int key_bits = bt + 2 * bw;
        if (!weighted) {
            cub::DoubleBuffer<uint64_t> db(keys.p, keys_alt.p);
            GT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, (int64_t) nloc, 0, key_bits, st));
            tmp.alloc(tmp_bytes);
            GT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, db, (int64_t) nloc, 0, key_bits, st));
            sorted_keys = db.Current();
        } else {
            // stable: by weight first, then by (tile, col, row) -> ascending weight inside every (row,col)
            cub::DoubleBuffer<uint32_t> dw(wts.p, wts_alt.p);
            cub::DoubleBuffer<uint64_t> dk(keys.p, keys_alt.p);
            GT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dw, dk, (int64_t) nloc, 0, 32, st));
            size_t tmp2 = 0;
            GT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, dk, dw, (int64_t) nloc, 0, key_bits, st));
            tmp.alloc(std::max(tmp_bytes, tmp2));
            GT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dw, dk, (int64_t) nloc, 0, 32, st));
            GT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp2, dk, dw, (int64_t) nloc, 0, key_bits, st));
            sorted_keys = dk.Current();
            sorted_wts = dw.Current();
        }
        ctx->kernel_launches += 8;
        if (!flags->parallel_edges) {          // matrix.hpp:552-555 (std::unique on adjacent (row,col))
            uint64_t* other_k = (sorted_keys == keys.p) ? keys_alt.p : keys.p;
            uint32_t* other_w = weighted ? ((sorted_wts == wts.p) ? wts_alt.p : wts.p) : nullptr;
            size_t ub = 0;
            unsigned long long* d_nsel = counters.p + 3;
            if (!weighted) {
                GT_CUDA(cub::DeviceSelect::Unique(nullptr, ub, sorted_keys, other_k, d_nsel, (int64_t) nloc, st));
                if (ub > tmp.n) tmp.alloc(ub);
                GT_CUDA(cub::DeviceSelect::Unique(tmp.p, ub, sorted_keys, other_k, d_nsel, (int64_t) nloc, st));
            } else {
                GT_CUDA(cub::DeviceSelect::UniqueByKey(nullptr, ub, sorted_keys, sorted_wts, other_k, other_w, d_nsel, (int64_t) nloc, st));
                if (ub > tmp.n) tmp.alloc(ub);
                GT_CUDA(cub::DeviceSelect::UniqueByKey(tmp.p, ub, sorted_keys, sorted_wts, other_k, other_w, d_nsel, (int64_t) nloc, st));
            }
            ctx->kernel_launches += 2;
            unsigned long long nsel = 0;
            GT_CUDA(cudaMemcpyAsync(&nsel, d_nsel, sizeof(nsel), cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            nloc = nsel;
            sorted_keys = other_k;
            sorted_wts = other_w;
        }
        GT_CUDA(cudaStreamSynchronize(st));
    }
    g->nnz_local = nloc;

    // ---- index maps ---------------------------------------------------------------------------
    {
        DevBuf<uint32_t> scan; scan.alloc(nall + 1);
        DevBuf<uint8_t> tmp;
        size_t tb = 0;
        GT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, scan.p, scan.p, (int64_t) (nall + 1), st));
        tmp.alloc(tb);
        auto make = [&](const DevBuf<uint8_t>& bits_all, const std::vector<int32_t>& segs, std::vector<SegMaps>& out) {
            GT_CUDA(cudaMemsetAsync(scan.p + nall, 0, 4, st));
            k_u8_to_u32<<<grid_for(nall, 256, ctx->sm_count), 256, 0, st>>>(bits_all.p, scan.p, nall);
            GT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, scan.p, scan.p, (int64_t) (nall + 1), st));
            ctx->kernel_launches += 2;
            out.resize(segs.size());
            std::vector<uint32_t> ends(segs.size() * 2);
            for (size_t k = 0; k < segs.size(); k++) {
                const uint64_t b = (uint64_t) segs[k] * th;
                GT_CUDA(cudaMemcpyAsync(&ends[2 * k], scan.p + b, 4, cudaMemcpyDeviceToHost, st));
                GT_CUDA(cudaMemcpyAsync(&ends[2 * k + 1], scan.p + b + th, 4, cudaMemcpyDeviceToHost, st));
            }
            GT_CUDA(cudaStreamSynchronize(st));
            for (size_t k = 0; k < segs.size(); k++) {
                SegMaps& m = out[k];
                m.segment = segs[k];
                m.nnz = ends[2 * k + 1] - ends[2 * k];
                m.bits.alloc(th); m.prefix.alloc(th); m.ids.alloc(m.nnz);
                k_seg_maps<<<grid_for(th, 256, ctx->sm_count), 256, 0, st>>>(bits_all.p, scan.p, (uint32_t) ((uint64_t) segs[k] * th), th,
                                                                          m.bits.p, m.prefix.p, m.ids.p);
                ctx->kernel_launches++;
            }
            GT_CUDA(cudaGetLastError());
            GT_CUDA(cudaStreamSynchronize(st));
        };
        make(I_all, L.local_row_segments, g->rows);
        make(J_all, L.local_col_segments, g->cols);
    }
    // ---- hot order of every local segment (rows and columns of a segment share it) ------------------
    {
        std::vector<int32_t> segs = L.local_row_segments;
        for (int32_t c : L.local_col_segments) if (std::find(segs.begin(), segs.end(), c) == segs.end()) segs.push_back(c);
        g->hot.resize(segs.size());
        DevBuf<uint64_t> keys, alt; keys.alloc(th); alt.alloc(th);
        DevBuf<uint8_t> tmp;
        size_t tb = 0;
        {
            cub::DoubleBuffer<uint64_t> db(keys.p, alt.p);
            GT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, (int64_t) th, 0, 64, st));
            tmp.alloc(tb);
        }
        DevBuf<uint32_t> d_n; d_n.alloc(3);
        const int cf = compression == GT_TCSC_CF;
        if (cf) g->cls.resize(segs.size());
        for (size_t h = 0; h < segs.size(); h++) {
            HotOrder& H = g->hot[h];
            H.segment = segs[h];
            const uint64_t b = (uint64_t) segs[h] * th;
            if (cf) g->cls[h].alloc(th);
            GT_CUDA(cudaMemsetAsync(d_n.p, 0, 12, st));
            k_hot_keys<<<grid_for(th, 256, ctx->sm_count), 256, 0, st>>>(I_all.p + b, J_all.p + b, rdeg_all.p + b, cdeg_all.p + b, th, cf, keys.p, cf ? g->cls[h].p : nullptr, d_n.p + 1);
            cub::DoubleBuffer<uint64_t> db(keys.p, alt.p);
            GT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, (int64_t) th, 0, 64, st));
            k_hot_count<<<1, 1, 0, st>>>(db.Current(), th, d_n.p);
            uint32_t h_n[3];
            GT_CUDA(cudaMemcpyAsync(h_n, d_n.p, 12, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            H.n = h_n[0];
            H.nreg = cf ? h_n[1] : H.n;
            H.nsrc = cf ? h_n[2] : 0;
            H.ids.alloc(H.n); H.pos.alloc(th);
            GT_CUDA(cudaMemsetAsync(H.pos.p, 0xff, (size_t) th * 4, st));
            if (H.n) k_hot_finish<<<grid_for(H.n, 256, ctx->sm_count), 256, 0, st>>>(db.Current(), H.n, H.ids.p, H.pos.p);
            ctx->kernel_launches += 6;
            GT_CUDA(cudaGetLastError());
            GT_CUDA(cudaStreamSynchronize(st));
        }
        auto find_hot = [&](int32_t seg) { for (size_t h = 0; h < segs.size(); h++) if (segs[h] == seg) return (int) h; return -1; };
        for (int32_t r : L.local_row_segments) g->hot_of_row_slot.push_back(find_hot(r));
        for (int32_t c : L.local_col_segments) g->hot_of_col_slot.push_back(find_hot(c));
    }
    // column degrees of the local column segments (raw record counts over the WHOLE matrix): the non-stationary engine
    // sizes a frontier by its edges, not only by its columns
    g->col_deg.resize(L.local_col_segments.size());
    g->col_edges.assign(L.local_col_segments.size(), 0);
    {
        DevBuf<unsigned long long> d_sum; d_sum.alloc(1);
        DevBuf<uint8_t> tmp;
        size_t tb = 0;
        GT_CUDA(cub::DeviceReduce::Sum(nullptr, tb, cdeg_all.p, d_sum.p, (int) th, st));
        tmp.alloc(tb);
        for (size_t k = 0; k < L.local_col_segments.size(); k++) {
            const uint64_t b = (uint64_t) L.local_col_segments[k] * th;
            g->col_deg[k].alloc(th);
            GT_CUDA(cudaMemcpyAsync(g->col_deg[k].p, cdeg_all.p + b, (size_t) th * 4, cudaMemcpyDeviceToDevice, st));
            cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> in(cdeg_all.p + b, U32ToU64());
            GT_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, in, d_sum.p, (int) th, st));
            unsigned long long h = 0;
            GT_CUDA(cudaMemcpyAsync(&h, d_sum.p, 8, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            g->col_edges[k] = h;
        }
        ctx->kernel_launches += L.local_col_segments.size();
    }
    I_all.release(); J_all.release(); rdeg_all.release(); cdeg_all.release();

    // ---- tiles ------------------------------------------------------------------------------------
    std::vector<uint64_t> bounds(ntiles + 1, 0);
    if (nloc) {
        DevBuf<uint64_t> d_bounds; d_bounds.alloc(ntiles + 1);
        k_tile_bounds<<<1, 128, 0, st>>>(sorted_keys, nloc, ntiles, 2 * bw, d_bounds.p);
        ctx->kernel_launches++;
        GT_REQUIRE(ntiles + 1 <= 128, "gt_graph_build: more than 127 local tiles");
        GT_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds.p, (ntiles + 1) * 8, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
    }
    // pool offsets are padded to 4 entries so that every tile's IA / A start 16-byte aligned
    uint64_t pool = 0;
    std::vector<uint64_t> pool_off(ntiles);
    for (int k = 0; k < ntiles; k++) { pool_off[k] = pool; pool += (bounds[k + 1] - bounds[k] + 3) / 4 * 4; }
    g->IA_pool.alloc(pool);
    if (weighted) g->A_pool.alloc(pool);
    g->tiles.resize(ntiles);
    const uint64_t row_mask = (1ull << bw) - 1;
    for (int k = 0; k < ntiles; k++) {
        Tile& T = g->tiles[k];
        const int kth = L.local_tiles_row_order[k];
        T.rg = kth / p; T.cg = kth % p;
        T.row_slot = (uint32_t) L.row_slot_of((int) T.rg);
        T.col_slot = (uint32_t) L.col_slot_of((int) T.cg);
        T.offset = pool_off[k];
        T.nnz = bounds[k + 1] - bounds[k];
        GT_REQUIRE(T.nnz < (1ull << 32), "gt_graph_build: a tile holds 2^32 or more entries (JA is uint32, compressed_column.hpp:293-296)");
        const SegMaps& R = g->rows[T.row_slot];
        const SegMaps& C = g->cols[T.col_slot];
        T.JA.alloc((size_t) C.nnz + 1);
        const uint64_t* tk = sorted_keys + bounds[k];
        if (T.nnz) {
            k_tile_IA<<<grid_for(T.nnz, 256, ctx->sm_count), 256, 0, st>>>(tk, T.nnz, row_mask, R.prefix.p, g->IA_pool.p + T.offset);
            ctx->kernel_launches++;
            if (weighted) GT_CUDA(cudaMemcpyAsync(g->A_pool.p + T.offset, sorted_wts + bounds[k], T.nnz * 4, cudaMemcpyDeviceToDevice, st));
        }
        k_tile_JA<<<grid_for((uint64_t) C.nnz + 1, 256, ctx->sm_count), 256, 0, st>>>(tk, T.nnz, (uint64_t) k << (2 * bw), bw, C.ids.p, C.nnz, T.JA.p);
        const uint32_t nchunks = (uint32_t) ((T.nnz + GT_PUSH_CHUNK - 1) / GT_PUSH_CHUNK);
        T.chunk_col.alloc((size_t) nchunks + 1);
        k_chunk_cols<<<grid_for((uint64_t) nchunks + 1, 256, ctx->sm_count), 256, 0, st>>>(T.JA.p, C.nnz, T.nnz, nchunks, T.chunk_col.p);
        ctx->kernel_launches += 2;
    }
    GT_CUDA(cudaGetLastError());
    GT_CUDA(cudaStreamSynchronize(st));
    {   // longest column per tile + scratch of the heavy-column path of the frontier SpMSpV
        uint64_t maxnnz = 0;                      // every listed chunk holds > kHeavyChunk/2 entries of its tile
        for (const Tile& T : g->tiles) maxnnz = std::max(maxnnz, T.nnz);
        g->heavy_list.alloc(maxnnz / 4096 + 64);
        g->heavy_count.alloc(1);
        DevBuf<unsigned int> d_m; d_m.alloc(ntiles);
        GT_CUDA(cudaMemsetAsync(d_m.p, 0, (size_t) ntiles * 4, st));
        for (int k = 0; k < ntiles; k++) {
            const Tile& T = g->tiles[k];
            const uint32_t nc = g->cols[T.col_slot].nnz;
            if (T.nnz && nc) { k_max_col_entries<<<grid_for(nc, 256, ctx->sm_count), 256, 0, st>>>(T.JA.p, nc, d_m.p + k); ctx->kernel_launches++; }
        }
        std::vector<unsigned int> h_m(ntiles, 0);
        GT_CUDA(cudaMemcpyAsync(h_m.data(), d_m.p, (size_t) ntiles * 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < ntiles; k++) g->tiles[k].max_col_entries = h_m[k];
    }

    if (compression == GT_TCSC_CF) build_cf(g.get());

    // nnz over all ranks
    g->nnz_global = g->nnz_local;
    if (ctx->comm) {
        DevBuf<unsigned long long> v; v.alloc(1);
        unsigned long long h = g->nnz_local;
        GT_CUDA(cudaMemcpyAsync(v.p, &h, 8, cudaMemcpyHostToDevice, st));
        comm_allreduce(ctx->comm, COMM_WORLD, v.p, v.p, 1, CT_U64, CO_SUM, st);
        GT_CUDA(cudaMemcpyAsync(&h, v.p, 8, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        g->nnz_global = h;
    }
    return g.release();
}

}  // namespace gt

// ---------------------------------------------------------------------------------------------------
extern "C" int gt_graph_build(gt_ctx* ctx, const void* triples, uint64_t ntriples, int weighted, int on_device,
                              uint32_t nvertices, const gt_graph_flags* flags, int compression, gt_graph** out) {
    return gt::guarded([&] {
        GT_REQUIRE(out, "gt_graph_build: out is NULL");
        *out = gt::build(ctx, triples, ntriples, weighted, on_device, nullptr, nvertices, flags, compression);
    });
}

extern "C" int gt_graph_build_rmat(gt_ctx* ctx, uint32_t scale, uint64_t nedges, uint64_t seed, int weighted,
                                   const gt_graph_flags* flags, int compression, gt_graph** out) {
    return gt::guarded([&] {
        GT_REQUIRE(out, "gt_graph_build_rmat: out is NULL");
        GT_REQUIRE(scale >= 1 && scale <= 31, "gt_graph_build_rmat: scale must be in [1, 31]");
        gt::RmatParams P = gt::make_rmat_params(scale, seed, weighted);
        *out = gt::build(ctx, nullptr, nedges, weighted, 1, &P, 1u << scale, flags, compression);
    });
}

// Partitioned ingest: every rank passes ITS SHARE of the records (Graph::parread_binary reads 1/p of the file per rank,
// src/mat/graph.hpp:307-335) and the entries travel to the owners of their tiles (Matrix::distribute, matrix.hpp:692-810).
extern "C" int gt_graph_build_partitioned(gt_ctx* ctx, const void* share, uint64_t nshare, int weighted, int on_device,
                                          uint32_t nvertices, const gt_graph_flags* flags, int compression, gt_graph** out) {
    return gt::guarded([&] {
        GT_REQUIRE(out, "gt_graph_build_partitioned: out is NULL");
        *out = gt::build(ctx, share, nshare, weighted, on_device, nullptr, nvertices, flags, compression, 1);
    });
}

extern "C" int gt_graph_build_rmat_partitioned(gt_ctx* ctx, uint32_t scale, uint64_t nedges, uint64_t seed, int weighted,
                                               const gt_graph_flags* flags, int compression, gt_graph** out) {
    return gt::guarded([&] {
        GT_REQUIRE(out && ctx, "gt_graph_build_rmat_partitioned: NULL argument");
        GT_REQUIRE(scale >= 1 && scale <= 31, "gt_graph_build_rmat_partitioned: scale must be in [1, 31]");
        gt::RmatParams P = gt::make_rmat_params(scale, seed, weighted);
        // rank r generates records [r * (nedges / p), ...) and the last rank the remainder: the reference's split of the file
        const uint64_t share = nedges / (uint64_t) ctx->nranks, first = share * (uint64_t) ctx->rank;
        const uint64_t n = ctx->rank == ctx->nranks - 1 ? nedges - first : share;
        *out = gt::build(ctx, nullptr, n, weighted, 1, &P, 1u << scale, flags, compression, 1, first);
    });
}

extern "C" int gt_rmat_generate(gt_ctx* ctx, uint32_t scale, uint64_t first_edge, uint64_t nedges, uint64_t seed,
                                int weighted, void* triples_dev) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx && triples_dev, "gt_rmat_generate: NULL argument");
        GT_REQUIRE(scale >= 1 && scale <= 31, "gt_rmat_generate: scale must be in [1, 31]");
        GT_CUDA(cudaSetDevice(ctx->device));
        gt::RmatParams P = gt::make_rmat_params(scale, seed, weighted);
        if (!nedges) return;
        gt::k_rmat_generate<<<gt::grid_for(nedges, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(P, first_edge, nedges, (uint32_t*) triples_dev);
        ctx->kernel_launches++;
        GT_CUDA(cudaGetLastError());
    });
}

extern "C" int gt_graph_free(gt_graph* g) {
    return gt::guarded([&] {
        if (!g) return;
        cudaSetDevice(g->ctx->device);
        delete g;
    });
}

extern "C" int gt_graph_info_get(gt_graph* g, gt_graph_info* out) {
    return gt::guarded([&] {
        GT_REQUIRE(g && out, "gt_graph_info_get: NULL argument");
        out->nedges_input = g->nedges_input;
        out->nnz_local = g->nnz_local;
        out->nnz_global = g->nnz_global;
        out->ntiles_local = (uint32_t) g->tiles.size();
        out->weighted = (uint32_t) g->weighted;
        out->nvertices = g->nvertices;
        out->layout = g->lay.info;
    });
}

extern "C" int gt_graph_tile_view(gt_graph* g, uint32_t local_tile, gt_tile_view* out) {
    return gt::guarded([&] {
        GT_REQUIRE(g && out, "gt_graph_tile_view: NULL argument");
        GT_REQUIRE(local_tile < g->tiles.size(), "gt_graph_tile_view: tile index out of range");
        const gt::Tile& T = g->tiles[local_tile];
        out->rg = T.rg; out->cg = T.cg; out->row_slot = T.row_slot; out->col_slot = T.col_slot;
        out->nnz = T.nnz;
        out->nnzcols = g->cols[T.col_slot].nnz;
        out->nnzrows = g->rows[T.row_slot].nnz;
        out->JA = T.JA.p;
        out->IA = g->IA_pool.p ? g->IA_pool.p + T.offset : nullptr;
        out->A = g->A_pool.p ? g->A_pool.p + T.offset : nullptr;
        out->JC = g->cols[T.col_slot].ids.p;
        out->IR = g->rows[T.row_slot].ids.p;
    });
}

namespace gt {
__global__ void k_classify(const uint8_t* __restrict__ I, const uint8_t* __restrict__ J, uint32_t th, unsigned int* __restrict__ out) {
    unsigned int reg = 0, src = 0, snk = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        const bool r = I[i], c = J[i];
        reg += r && c; src += r && !c; snk += !r && c;
    }
    if (reg) atomicAdd(out, reg);
    if (src) atomicAdd(out + 1, src);
    if (snk) atomicAdd(out + 2, snk);
}
}  // namespace gt

extern "C" int gt_graph_classify(gt_graph* g, uint32_t* regular, uint32_t* source_rows, uint32_t* sink_columns) {
    return gt::guarded([&] {
        GT_REQUIRE(g, "gt_graph_classify: NULL graph");
        GT_CUDA(cudaSetDevice(g->ctx->device));
        const uint32_t th = g->lay.info.tile_height;
        gt::DevBuf<unsigned int> d; d.alloc(3);
        GT_CUDA(cudaMemsetAsync(d.p, 0, 12, g->ctx->stream));
        gt::k_classify<<<gt::grid_for(th, 256, g->ctx->sm_count), 256, 0, g->ctx->stream>>>(
            g->rows[g->lay.info.accu_segment_row].bits.p, g->cols[g->lay.info.accu_segment_col].bits.p, th, d.p);
        g->ctx->kernel_launches++;
        unsigned int h[3];
        GT_CUDA(cudaMemcpyAsync(h, d.p, 12, cudaMemcpyDeviceToHost, g->ctx->stream));
        GT_CUDA(cudaStreamSynchronize(g->ctx->stream));
        if (regular) *regular = h[0];
        if (source_rows) *source_rows = h[1];
        if (sink_columns) *sink_columns = h[2];
    });
}

extern "C" int gt_graph_classify_lists(gt_graph* g, const uint32_t** regular_rows, uint32_t* nregular, const uint32_t** source_rows, uint32_t* nsource,
                                       const uint32_t** sink_columns, uint32_t* nsink) {
    return gt::guarded([&] {
        GT_REQUIRE(g, "gt_graph_classify_lists: NULL graph");
        if (g->compression != GT_TCSC_CF)
            throw gt::Error(GT_ERR_INVALID, "gt_graph_classify_lists: the graph was not built with GT_TCSC_CF (the reference keeps these lists for _TCSC_CF_ only)");
        if (regular_rows) *regular_rows = g->cf_owned.regular_rows.p;
        if (source_rows) *source_rows = g->cf_owned.source_rows.p;
        if (sink_columns) *sink_columns = g->cf_owned.sink_columns.p;
        if (nregular) *nregular = g->cf_owned.nreg;
        if (nsource) *nsource = g->cf_owned.nsrc;
        if (nsink) *nsink = g->cf_owned.nsnk;
    });
}

extern "C" int gt_graph_tile_cf_view(gt_graph* g, uint32_t local_tile, gt_tile_cf_view* out) {
    return gt::guarded([&] {
        GT_REQUIRE(g && out, "gt_graph_tile_cf_view: NULL argument");
        GT_REQUIRE(local_tile < g->tiles.size(), "gt_graph_tile_cf_view: tile index out of range");
        if (g->compression != GT_TCSC_CF)
            throw gt::Error(GT_ERR_INVALID, "gt_graph_tile_cf_view: the graph was not built with GT_TCSC_CF");
        const gt::CfTile& C = g->cf_tiles[local_tile];
        for (int k = 0; k < 4; k++) {
            out->NC[k] = C.NC[k]; out->filled[k] = C.filled[k];
            out->JA[k] = C.JA[k].p; out->JC[k] = C.JC[k].p;
        }
    });
}

extern "C" int gt_graph_rowgrp_maps(gt_graph* g, uint32_t row_slot, const uint8_t** I, const uint32_t** IV, uint32_t* nnzrows) {
    return gt::guarded([&] {
        GT_REQUIRE(g && row_slot < g->rows.size(), "gt_graph_rowgrp_maps: slot out of range");
        if (I) *I = g->rows[row_slot].bits.p;
        if (IV) *IV = g->rows[row_slot].prefix.p;
        if (nnzrows) *nnzrows = g->rows[row_slot].nnz;
    });
}

extern "C" int gt_graph_colgrp_maps(gt_graph* g, uint32_t col_slot, const uint8_t** J, const uint32_t** JV, uint32_t* nnzcols) {
    return gt::guarded([&] {
        GT_REQUIRE(g && col_slot < g->cols.size(), "gt_graph_colgrp_maps: slot out of range");
        if (J) *J = g->cols[col_slot].bits.p;
        if (JV) *JV = g->cols[col_slot].prefix.p;
        if (nnzcols) *nnzcols = g->cols[col_slot].nnz;
    });
}
