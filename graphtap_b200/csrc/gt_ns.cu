// gt_ns.cu — BFS / CC / SSSP: the non-stationary iteration of Vertex_Program::execute on the device.
//
// The reference loop (src/vp/vertex_program.hpp:407-441) per iteration, for a non-stationary program:
//   scatter_gather   x[j] = C[v] ? messenger(V[v]) : infinity() over the owned column segment, plus the compacted
//                    (xi, xv) pairs and the 0.6 activity rule (:710-784); Ibcast / Send of dense-or-sparse x along
//                    the column group with the item count ahead of the payload (:864-1013)
//   combine          per local tile the frontier SpMSpV when the column segment's owner was sparse, else the dense
//                    SpMV skipping infinity() (:1437-1506); follower -> leader send of the partial y, dense or as the
//                    touched rows (yi, yv) (:1330-1434), leader-side min (:1543-1573)
//   apply            applicator on the non-empty rows (:1695-1802), activity flags C
//   has_converged    all C == 0 on all ranks (:1884-1923)
// and what it becomes here:
//   * ONE batched launch per kernel for all local tiles (grid.y = tile, descriptors in device memory), each tile
//     reading the {mode, count} header its column segment's owner published — the SpMSpV and the dense kernel are both
//     enqueued and the one whose mode it is not returns at once.  The host never needs a frontier size, so nothing in
//     the iteration synchronises with it; the convergence count of iteration k is read while iteration k+1 (an empty
//     frontier, a no-op, if k was the last) is already running.
//   * the applicator also writes the NEXT iteration's dense x and frontier list (changed vertices are exactly the
//     active ones), so scatter_gather is only the exchange;
//   * SpMSpV: a CTA takes 256 frontier columns, block-scans their lengths and spreads the EDGES over its threads
//     (binary search in the scanned offsets), so short columns do not waste a warp each; columns longer than 8192
//     entries are cut into 4096-entry chunks for a second launch (block path for RMAT hubs);
//   * BFS on an undirected graph held by one GPU: above `bfs_bottom_up_ratio` the pass runs bottom-up — every unvisited
//     vertex walks its (ascending) neighbour list and stops at the first active one, which IS the minimum parent id
//     the reference's min-combiner (src/apps/bfs.h:61-63) would leave in y;
//   * multi-GPU: the owner's SMs store x — list or dense, by the 0.6 rule — into the column group's NVLink peer windows;
//     partial y travels to the row segment's leader the same way, as the rows a RED went out for in this iteration (y
//     is a running minimum that is never reset, :1785-1793, so a row whose current value already beats the candidate
//     has nothing new to say) or dense when more than 60 % did; the leader merges with a scatter-min.  A world all-reduce per iteration (the convergence
//     count, or a 1-element fence in fixed-iteration mode) orders every rank's next put behind every reader of the
//     current one, including across execute() / run_phase() calls, so the windows need a single buffer.
#include "gt_program.h"
#include <cub/cub.cuh>
#include <chrono>

namespace gt {

constexpr uint32_t NS_DENSE = 0, NS_SPARSE = 1, NS_BOTTOM_UP = 2;
constexpr uint32_t kNsHeavyColumn = 8192;        // frontier columns longer than this are cut into chunks ...
constexpr uint32_t kNsHeavyChunk = 4096;         // ... of this many entries, one CTA each
constexpr int kNsBatch = 256;                    // frontier columns per CTA batch = threads per CTA

struct NsTile {                                  // one local tile (device array, local_tiles_row_order)
    const uint32_t* JA; const uint32_t* IA; const uint32_t* A; const uint32_t* chunk_col;
    uint64_t nnz; uint32_t nchunks, ncols;
    const uint32_t* x; const uint32_t* xi; const uint32_t* xv; const uint32_t* hdr;   // its column segment: dense x, frontier list, {mode, count}
    uint32_t* y; uint8_t* t;                     // running-min y of its row segment; improved-row flags (segments led by another rank) or nullptr
    uint64_t dense_bytes;                        // algorithmic bytes of one dense pass
};
// The tile passes take their descriptors BY VALUE, eight at a time: kernel parameters live in the constant bank and cost
// no registers (the first version loaded them from a device array and lost two CTAs per SM of occupancy to the
// pointers it kept live).
constexpr int kNsGroup = 8;
struct NsTiles { NsTile t[kNsGroup]; };
struct NsYSend {                                 // a row segment led by another rank: its partial y goes to the leader
    uint32_t* y; uint8_t* t; uint32_t n;
    uint32_t* yi; uint32_t* yv; unsigned int* count;     // local staging list of the improved rows
    uint32_t* dst_dense; uint32_t* dst_yi; uint32_t* dst_yv; uint32_t* dst_hdr; uint32_t* dst_flag;   // this rank's slot in the leader's window
    unsigned int* done;
};
struct NsYRecv { const uint32_t* dense; const uint32_t* yi; const uint32_t* yv; const uint32_t* hdr; uint32_t n; };   // a follower's slot in MY window

struct NsState {
    size_t S = 0, R = 0, xchunk = 0, ychunk = 0;
    int G = 1;                                   // members of the reduce (row) group
    std::vector<int> xq, yq;                     // chunk of every x / y slot = group rank of the segment's leader
    PeerWindow* wx = nullptr;                    // column group: [dense S x chunk][xi S x chunk][xv S x chunk][hdr S x 4]
    PeerWindow* wy = nullptr;                    // row group, read by the leader: [dense G x chunk][yi ..][yv ..][hdr G x 4], slot = sender
    DevBuf<uint32_t> xlocal;                     // the same x layout without a window
    uint32_t* xbase = nullptr; uint32_t* hdr = nullptr;
    std::vector<uint32_t*> x, xi, xv;            // per x slot
    DevBuf<uint32_t> Y;                          // R x ychunk
    std::vector<uint32_t*> y;
    DevBuf<uint8_t> T;                           // improved-row flags, R x ychunk (used for the slots this rank does not lead)
    DevBuf<uint32_t> ystage;                     // (yi, yv) staging, 2 x ychunk per send slot
    DevBuf<NsTile> tiles; int ntiles = 0;        // device copy (the heavy-column kernel indexes it by list entry)
    std::vector<NsTile> htiles;                  // host copy, handed to the tile passes by value
    DevBuf<NsYSend> ysend; DevBuf<NsYRecv> yrecv; int nsend = 0, nrecv = 0;
    DevBuf<uint2> heavy_list;
    DevBuf<unsigned int> slot_counts;            // NCCL exchange only: frontier size of every x slot
    DevBuf<unsigned int> counters;               // [0] heavy count, [1] own frontier count, [2] x put done, [3 + 2 i] y count i, [4 + 2 i] y put done i
    DevBuf<unsigned long long> stats;            // [0] algorithmic bytes of the tile passes, [1] iterations with a sparse tile, [2] flag of the running one,
                                                 // [3] size in edges of the frontier being built
    uint32_t x_epoch = 0, y_epoch = 0;
    int active_slot = 0;                         // d_active slot of the iteration being enqueued
    bool any_heavy = false, bottom_up_ok = false, nccl_exchange = false;
    bool publish_via_put = false;                // this execute(): the convergence count of iteration n is written to the host by iteration n + 1's put kernel
    bool publish_prev = false;                   // ... and the iteration being enqueued has a predecessor in this execute()
    cudaEvent_t ev_it[2] = {nullptr, nullptr};
    unsigned long long* h_stats = nullptr;       // pinned
};

// ---- scatter_gather from the vertex state (first iteration of an execute(), run_phase(0)) ----------------------------------
// non-stationary messenger: x[j] = C[v] ? messenger(V[v]) : infinity() (:737-751)
// also sums the column degrees of the active vertices (the frontier's size in edges, see ns_mode)
__global__ void k_ns_messenger(VState V, int app, uint32_t vid0, const uint32_t* __restrict__ JC, uint32_t nc, uint32_t* __restrict__ x,
                               const uint32_t* __restrict__ cdeg, unsigned long long* __restrict__ frontier_edges) {
    unsigned long long fe = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nc; j += gridDim.x * blockDim.x) {
        const uint32_t v = JC[j];
        uint32_t m = GT_INF_U32;
        if (V.C[v]) { m = (app == GT_APP_BFS) ? vid0 + v : V.a[v]; fe += cdeg[v]; }         // bfs.h:52-54, cc.h:37-39, sssp.h:45-47
        x[j] = m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fe += __shfl_xor_sync(0xffffffffu, fe, o);
    if ((threadIdx.x & 31) == 0 && fe) atomicAdd(frontier_edges, fe);
}
// frontier list of one x segment: xi = compressed ids with x != INF, xv = their values (:744-748); order inside
// the list is irrelevant to a min reduction.  One atomic per CTA per 4096 elements (block scan of the per-thread counts).
__global__ void __launch_bounds__(1024) k_ns_frontier(const uint32_t* __restrict__ x, uint32_t nc, uint32_t* __restrict__ xi, uint32_t* __restrict__ xv,
                                                       unsigned int* __restrict__ count) {
    typedef cub::BlockScan<unsigned, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ unsigned base_s;
    const uint32_t per_iter = 1024 * 4;
    for (uint32_t start = blockIdx.x * per_iter; start < nc; start += gridDim.x * per_iter) {
        const uint32_t j0 = start + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = (j0 + u < nc) ? x[j0 + u] : GT_INF_U32;
        unsigned mine = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) mine += v[u] != GT_INF_U32;
        unsigned off, total;
        BS(tmp).ExclusiveSum(mine, off, total);
        if (threadIdx.x == 0 && total) base_s = atomicAdd(count, total);
        __syncthreads();
        if (total) {
            unsigned pos = base_s + off;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (v[u] != GT_INF_U32) { xi[pos] = j0 + u; xv[pos] = v[u]; pos++; }
        }
        __syncthreads();
    }
}

// ---- x exchange: header + stores into the column group's windows ----------------------------------------------------------
// The reference ships a column segment's x either dense or as the compacted (xi, xv) pair, by the owner's 0.6 rule,
// with the item count ahead of the payload (:760-784,864-1013).  Here the owner's SMs store straight into the other
// column-group members' windows: the frontier list if the rule says sparse, the dense segment otherwise, then the
// header {mode, count} and the arrival counter (gt_peer.cu).  The size never visits the host.
struct PutTargets { uint32_t* dense[8]; uint32_t* xi[8]; uint32_t* xv[8]; uint32_t* hdr[8]; uint32_t* flag[8]; int n; };
__device__ __forceinline__ void copy_u32(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n, uint32_t tid, uint32_t nth) {
    const uint32_t n4 = n >> 2;                    // both sides are 16-byte aligned (chunks are multiples of 4 elements)
    const uint4* s4 = (const uint4*) src;
    uint4* d4 = (uint4*) dst;
    for (uint32_t i = tid; i < n4; i += nth) d4[i] = s4[i];
    for (uint32_t i = (n4 << 2) + tid; i < n; i += nth) dst[i] = src[i];
}
// The reference's rule is by columns: sparse when at most `ratio` (0.6) of the segment's columns are active (:768-772).
// On RMAT that calls "sparse" a frontier of a third of the columns that holds three quarters of the edges, for which the
// frontier kernel (one RED per edge into random rows) takes twice as long as the dense pass — so the owner also sizes
// the frontier in EDGES (sum of the active columns' degrees) and goes dense above `edge_ratio` of the segment's edges.
// The choice changes which kernel runs and what travels, never the result.
struct NsRule { double ratio, bu_ratio, edge_ratio; unsigned long long seg_edges; };
__device__ __forceinline__ uint32_t ns_mode(unsigned k, uint32_t n, unsigned long long fe, const NsRule& R) {
    uint32_t mode = (n && ((double) k / (double) n <= R.ratio)) ? NS_SPARSE : NS_DENSE;
    if (mode == NS_SPARSE && R.edge_ratio > 0.0 && (double) fe > R.edge_ratio * (double) R.seg_edges) mode = NS_DENSE;
    if (R.bu_ratio > 0.0 && (double) k > R.bu_ratio * (double) n) mode = NS_BOTTOM_UP;
    return mode;
}
__global__ void __launch_bounds__(256) k_ns_put_x(const uint32_t* __restrict__ dense, const uint32_t* __restrict__ xi, const uint32_t* __restrict__ xv,
                                                   unsigned int* __restrict__ count, uint32_t n, unsigned long long* __restrict__ frontier_edges,
                                                   NsRule rule, uint32_t* own_hdr, PutTargets T, uint32_t epoch, unsigned int* done,
                                                   unsigned long long* __restrict__ next_active, const unsigned long long* __restrict__ prev_active,
                                                   volatile unsigned long long* host_prev) {
    const unsigned k = *count;
    const uint32_t mode = ns_mode(k, n, *frontier_edges, rule);
    // only as many CTAs as the payload can keep busy take part (32 KB each): with a frontier of a few hundred columns the
    // system-scope fences and the hand-shake of 2 x 148 CTAs cost more than the transfer
    const uint32_t words = T.n ? (mode == NS_SPARSE ? 2 * k : n) : 0;
    const uint32_t nblk = max(1u, min(gridDim.x, (words + 8191u) / 8192u));
    if (blockIdx.x >= nblk) return;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nblk * blockDim.x;
    for (int j = 0; j < T.n; j++) {
        if (mode == NS_SPARSE) { copy_u32(T.xi[j], xi, k, tid, nth); copy_u32(T.xv[j], xv, k, tid, nth); }
        else copy_u32(T.dense[j], dense, n, tid, nth);
    }
    __shared__ bool last;
    if (T.n) {                                     // (one GPU: a single CTA and nothing to order)
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(done, 1u) == nblk - 1;
        __syncthreads();
        if (!last) return;                         // the last CTA to finish publishes: every payload store is ordered before
        __threadfence_system();
    }
    if (threadIdx.x == 0) {
        own_hdr[0] = mode; own_hdr[1] = k; *done = 0;
        *count = 0; *frontier_edges = 0; *next_active = 0;      // what this iteration's applicator accumulates into
        // the (all-reduced) active count of the previous iteration goes straight to the pinned host word the convergence
        // check reads (visible to the host once this kernel has completed, which is what its event waits for): no
        // device-to-host copy sits between an iteration's all-reduce and the next iteration's first kernel
        if (host_prev) *host_prev = *prev_active;
    }
    if ((int) threadIdx.x < T.n) {
        volatile uint32_t* h = T.hdr[threadIdx.x];
        h[0] = mode; h[1] = k;
        __threadfence_system();
        volatile uint32_t* f = T.flag[threadIdx.x];
        *f = epoch & (kPeerSeqLen - 1);
    }
}
// NCCL fallback (GT_PEER=0): dense x arrived by all-gather, every rank rebuilds the lists and applies the rule itself
__global__ void k_ns_header(unsigned int* __restrict__ count, uint32_t n, NsRule rule, uint32_t* hdr, unsigned long long* __restrict__ frontier_edges,
                            unsigned long long* __restrict__ next_active) {
    const unsigned k = *count;
    hdr[0] = ns_mode(k, n, 0ull, rule); hdr[1] = k;
    *count = 0;
    if (frontier_edges) { *frontier_edges = 0; *next_active = 0; }      // own segment: see k_ns_put_x
}

// ---- combine: the tile passes ----------------------------------------------------------------------------------------------
// `filter`: read y first and skip the RED when it cannot win (y only ever decreases, so a stale read is safe).  Pays
// when the frontier is large — RMAT's hub rows settle after a few updates and stop serialising in L2 — and costs an
// extra L2 round trip when most updates are first visits, so the caller switches it on by frontier size.
template <int S>
__device__ __forceinline__ void ns_reduce(typename Semiring<S>::T* y, uint8_t* t, uint32_t r, typename Semiring<S>::T v, bool filter) {
    if (filter && v >= __ldcg(y + r)) return;
    if (!t) { Semiring<S>::reduce(y + r, v); return; }
    // the row goes to the leader: with the filter, every row a RED went out for (fire and forget; a superset of the rows
    // that improved); without it — small frontiers — exactly the rows whose atomic won
    if (filter) { atomicMin(y + r, v); t[r] = 1; }
    else if (v < atomicMin(y + r, v)) t[r] = 1;
}

// frontier SpMSpV over every local tile whose column segment travelled as a list (:1476-1488).
// A CTA takes a batch of B frontier columns, block-scans their lengths and spreads the batch's EDGES over its 256
// threads.  B follows the frontier: with few columns (the first iterations out of a hub) every CTA gets one or a few of
// them, so no CTA is left walking hundreds of long columns alone; with millions, B = 256.
template <int S, bool WEIGHTED>
__global__ void __launch_bounds__(kNsBatch, 6)
k_ns_spmspv(const __grid_constant__ NsTiles tiles, int tile0, uint2* __restrict__ heavy_list, unsigned int* __restrict__ heavy_count,
            unsigned long long* __restrict__ stats) {
    typedef Semiring<S> SR;
    typedef typename SR::T T;
    const NsTile& Q = tiles.t[blockIdx.y];
    if (!Q.nnz || Q.hdr[0] != NS_SPARSE) return;
    const uint32_t k = Q.hdr[1];
    if (blockIdx.x == 0 && threadIdx.x == 0 && k) stats[2] = 1;
    typedef cub::BlockScan<uint32_t, kNsBatch> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ uint32_t s_start[kNsBatch], s_off[kNsBatch];
    __shared__ T s_val[kNsBatch];
    const uint32_t tid = threadIdx.x;
    const bool filter = k > (Q.ncols >> 5);                           // more than 3 % of the columns active
    uint32_t B = (k + gridDim.x - 1) / gridDim.x;                     // columns per batch: 1 .. 256
    B = B < 1 ? 1 : B > (uint32_t) kNsBatch ? (uint32_t) kNsBatch : B;
    unsigned long long bytes = 0;
    for (uint32_t base = blockIdx.x * B; base < k; base += gridDim.x * B) {
        const uint32_t f = base + tid;
        uint32_t len = 0, b = 0;
        T v = SR::identity();
        if (tid < B && f < k) {
            const uint32_t j = Q.xi[f];
            v = Q.xv[f];
            b = Q.JA[j];
            len = Q.JA[j + 1] - b;
            bytes += 16 + (WEIGHTED ? 8ull : 4ull) * len;          // xi, xv, JA pair, IA (+A) of the column
            if (len > kNsHeavyColumn) {
                const uint32_t nch = (len + kNsHeavyChunk - 1) / kNsHeavyChunk;
                const unsigned hb = atomicAdd(heavy_count, nch);
                for (uint32_t c = 0; c < nch; c++) heavy_list[hb + c] = make_uint2(f, ((uint32_t) (tile0 + blockIdx.y) << 24) | c);
                len = 0;
            }
        }
        uint32_t off, total;
        BS(tmp).ExclusiveSum(len, off, total);
        s_start[tid] = b; s_val[tid] = v; s_off[tid] = off;
        __syncthreads();
        for (uint32_t idx = tid; idx < total; idx += kNsBatch) {
            uint32_t lo = 0, hi = B;                                 // first c with s_off[c] > idx; the column is the one before
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_off[mid] <= idx) lo = mid + 1; else hi = mid;
            }
            const uint32_t c = lo - 1;
            const uint32_t i = s_start[c] + (idx - s_off[c]);
            const uint32_t r = ld_stream_u32(Q.IA + i);
            ns_reduce<S>(Q.y, Q.t, r, WEIGHTED ? SR::mul(s_val[c], ld_stream_u32(Q.A + i)) : s_val[c], filter);
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    if ((tid & 31) == 0 && bytes) atomicAdd(stats, bytes);
}

template <int S, bool WEIGHTED>
__global__ void __launch_bounds__(256)
k_ns_heavy(const NsTile* __restrict__ tiles, const uint2* __restrict__ heavy_list, const unsigned int* __restrict__ heavy_count) {
    typedef Semiring<S> SR;
    typedef typename SR::T T;
    const unsigned int n = *heavy_count;
    for (unsigned int h = blockIdx.x; h < n; h += gridDim.x) {          // one CTA per (heavy column, chunk)
        const uint2 fc = heavy_list[h];
        const NsTile& Q = tiles[fc.y >> 24];
        const uint32_t j = Q.xi[fc.x];
        const T v = Q.xv[fc.x];
        const uint32_t b = Q.JA[j] + (fc.y & 0xffffffu) * kNsHeavyChunk, e = min(Q.JA[j + 1], b + kNsHeavyChunk);
        const bool filter = Q.hdr[1] > (Q.ncols >> 5);
        for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
            const uint32_t r = ld_stream_u32(Q.IA + i);
            ns_reduce<S>(Q.y, Q.t, r, WEIGHTED ? SR::mul(v, ld_stream_u32(Q.A + i)) : v, filter);
        }
    }
}

// dense SpMV skipping infinity() over every local tile whose column segment travelled dense (:1491-1502)
template <int S, bool WEIGHTED>
__global__ void __launch_bounds__(kPushThreads, 8)
k_ns_dense(const __grid_constant__ NsTiles tiles, unsigned long long* __restrict__ stats, unsigned int* __restrict__ heavy_count) {
    if (heavy_count && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *heavy_count = 0;   // consumed by the heavy kernel before this launch
    const NsTile& Q = tiles.t[blockIdx.y];
    if (!Q.nnz || Q.hdr[0] != NS_DENSE) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats, (unsigned long long) Q.dense_bytes);
    spmv_push_chunks<S, WEIGHTED, true, true>(Q.JA, Q.IA, Q.A, Q.chunk_col, Q.nchunks, Q.nnz, (const typename Semiring<S>::T*) Q.x,
                                              (typename Semiring<S>::T*) Q.y, Q.t, blockIdx.x, gridDim.x);
}

// BFS bottom-up pass (one GPU, undirected graph: the single tile is symmetric, compressed row ids == compressed column
// ids, and a column's entries are its vertex's neighbours in ascending id order).  y[j] of an unvisited vertex is
// still infinity() (any earlier active neighbour would have visited it), so the first active neighbour found is the
// minimum the reference's push would have left there (src/apps/bfs.h:61-63).
__global__ void __launch_bounds__(256)
k_ns_bfs_bottom_up(const __grid_constant__ NsTiles tiles, const uint32_t* __restrict__ JC, const uint32_t* __restrict__ hops, unsigned long long* __restrict__ stats) {
    const NsTile& Q = tiles.t[0];
    if (Q.hdr[0] != NS_BOTTOM_UP) return;
    unsigned long long bytes = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < Q.ncols; j += gridDim.x * blockDim.x) {
        bytes += 8;                                                  // JC + hops of the vertex
        if (hops[JC[j]] != GT_INF_U32) continue;
        const uint32_t b = Q.JA[j], e = Q.JA[j + 1];
        bytes += 8;
        for (uint32_t i = b; i < e; i++) {
            const uint32_t xv = Q.x[Q.IA[i]];
            bytes += 8;
            if (xv != GT_INF_U32) { Q.y[j] = xv; break; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    if ((threadIdx.x & 31) == 0 && bytes) atomicAdd(stats, bytes);
}

// ---- y exchange: improved rows (or the dense segment) to the leader, scatter-min there (:1405-1424,1543-1573) ---------------
__global__ void __launch_bounds__(1024) k_ns_pack_y(const NsYSend* __restrict__ sends) {
    const NsYSend Q = sends[blockIdx.y];
    typedef cub::BlockScan<unsigned, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ unsigned base_s;
    // 16 flags per thread and round (four independent 4-byte loads in flight); a round without any flag — most of them in
    // the long tail of small iterations — costs one barrier and no scan
    const uint32_t per_iter = 1024 * 16;
    for (uint32_t start = blockIdx.x * per_iter; start < Q.n; start += gridDim.x * per_iter) {
        uint32_t f4[4];
        unsigned mine = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t i0 = start + w * 4096 + threadIdx.x * 4;      // n is padded to a multiple of 4 flags
            f4[w] = (i0 < Q.n) ? *reinterpret_cast<const uint32_t*>(Q.t + i0) : 0u;
            mine += __popc(f4[w] & 0x01010101u);
        }
        if (!__syncthreads_or(mine != 0)) continue;
        unsigned off, total;
        BS(tmp).ExclusiveSum(mine, off, total);
        if (threadIdx.x == 0) base_s = atomicAdd(Q.count, total);
        __syncthreads();
        if (mine) {
            unsigned pos = base_s + off;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                if (!f4[w]) continue;
                const uint32_t i0 = start + w * 4096 + threadIdx.x * 4;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if ((f4[w] >> (8 * u)) & 1u) { Q.yi[pos] = i0 + u; Q.yv[pos] = Q.y[i0 + u]; pos++; }
                *reinterpret_cast<uint32_t*>(Q.t + i0) = 0u;      // cleared for the next iteration (:1795-1801)
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_ns_put_y(const NsYSend* __restrict__ sends, double ratio, uint32_t epoch) {
    const NsYSend Q = sends[blockIdx.y];
    const unsigned k = *Q.count;
    const bool sparse = Q.n && ((double) k / (double) Q.n <= ratio);
    const uint32_t words = sparse ? 2 * k : Q.n;
    const uint32_t nblk = max(1u, min(gridDim.x, (words + 8191u) / 8192u));      // see k_ns_put_x
    if (blockIdx.x >= nblk) return;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nblk * blockDim.x;
    if (sparse) { copy_u32(Q.dst_yi, Q.yi, k, tid, nth); copy_u32(Q.dst_yv, Q.yv, k, tid, nth); }
    else copy_u32(Q.dst_dense, Q.y, Q.n, tid, nth);
    __threadfence_system();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(Q.done, 1u) == nblk - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x == 0) {
        *Q.done = 0; *Q.count = 0;
        volatile uint32_t* h = Q.dst_hdr;
        h[0] = sparse ? NS_SPARSE : NS_DENSE; h[1] = k;
        __threadfence_system();
        *(volatile uint32_t*) Q.dst_flag = epoch & (kPeerSeqLen - 1);
    }
}
__global__ void __launch_bounds__(256) k_ns_merge_y(const NsYRecv* __restrict__ recvs, uint32_t* __restrict__ y) {
    const NsYRecv Q = recvs[blockIdx.y];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (Q.hdr[0] == NS_SPARSE) {
        const uint32_t k = Q.hdr[1];
        for (uint32_t i = tid; i < k; i += nth) atomicMin(y + Q.yi[i], Q.yv[i]);                     // :1559-1562
    } else {
        for (uint32_t i = tid; i < Q.n; i += nth) { const uint32_t v = Q.dense[i]; if (v < y[i]) atomicMin(y + i, v); }   // :1568-1569
    }
}

// ---- apply ------------------------------------------------------------------------------------------------------------------
// applicator on the rows of rowgrp_nnz_rows (:1739-1751,1768-1780), fused with the next iteration's messenger and frontier
// compaction: a vertex is active next iteration iff its applicator returned true now.
// (A pass over only the rows whose y improved was built and measured: with the flag traffic it adds to the tile passes
// and its reset pass it LOSES on one GPU — SSSP RMAT-25 12.3 vs 11.4 ms, BFS RMAT-22 0.67 vs 0.51 ms,
// profiles/r02_ncu_ns_sssp25_sparse_apply_experiment.csv — so the applicator walks every non-empty row, as the reference does.)
__global__ void __launch_bounds__(256)
k_ns_apply(VState V, int app, int weighted, uint32_t vid0, const uint32_t* __restrict__ IR, uint32_t nr, const uint32_t* __restrict__ y, uint32_t iteration,
           const uint8_t* __restrict__ J, const uint32_t* __restrict__ JV, const uint32_t* __restrict__ cdeg, uint32_t* __restrict__ x,
           uint32_t* __restrict__ xi, uint32_t* __restrict__ xv, unsigned int* __restrict__ count, unsigned long long* __restrict__ active,
           unsigned long long* __restrict__ stats) {
    typedef cub::BlockScan<unsigned, 256> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ unsigned base_s;
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats[2]) { stats[1]++; stats[2] = 0; }
    unsigned changed = 0;
    unsigned long long fe = 0;                                         // the next frontier's size in edges -> stats[3]
    const uint32_t per_iter = 256 * 4;
    for (uint32_t start = blockIdx.x * per_iter; start < nr; start += gridDim.x * per_iter) {
        uint32_t cj[4], cm[4];
        unsigned mine = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t r = start + u * 256 + threadIdx.x;
            cj[u] = 0xffffffffu;
            if (r >= nr) continue;
            const uint32_t v = IR[r];
            const uint32_t yy = y[r];
            bool ch = false;
            uint32_t msg = 0;
            if (app == GT_APP_BFS) {                                   // bfs.h:65-77
                if (V.b[v] == GT_INF_U32 && yy != GT_INF_U32) { V.b[v] = iteration + 1; V.a[v] = yy; ch = true; }
                msg = vid0 + v;                                        // bfs.h:52-54
            } else if (app == GT_APP_CC) {                             // cc.h:51-55
                const uint32_t old = V.a[v];
                if (yy < old) { V.a[v] = yy; ch = true; }
                msg = ch ? yy : old;                                   // cc.h:37-39
            } else {                                                   // sssp.h:58-66
                const uint32_t old = V.a[v];
                const uint32_t nw = (yy < old) ? (weighted ? yy : yy + 1) : old;
                if (nw != old) { V.a[v] = nw; ch = true; }
                msg = nw;                                              // sssp.h:45-47
            }
            changed += ch;
            // x[j] != infinity() exactly where C[v] is set (the messenger and every earlier pass keep it so), so a vertex
            // that was inactive and stays inactive — nearly all of them in the long tail — has nothing to write
            if (!ch && !V.C[v]) continue;
            V.C[v] = ch;
            if (J[v]) {                                                // the vertex has a column: next iteration's x (:737-751)
                const uint32_t j = JV[v];
                x[j] = ch ? msg : GT_INF_U32;
                if (ch) { cj[u] = j; cm[u] = msg; mine++; fe += cdeg[v]; }
            }
        }
        if (!__syncthreads_or(mine != 0)) continue;                    // no new frontier entry in these 1024 rows: no scan
        unsigned off, total;
        BS(tmp).ExclusiveSum(mine, off, total);
        if (threadIdx.x == 0) base_s = atomicAdd(count, total);
        __syncthreads();
        if (mine) {
            unsigned pos = base_s + off;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (cj[u] != 0xffffffffu) { xi[pos] = cj[u]; xv[pos] = cm[u]; pos++; }
        }
        __syncthreads();
    }
    typedef cub::BlockReduce<unsigned, 256> BR;
    __shared__ typename BR::TempStorage tmp2;
    const unsigned tot = BR(tmp2).Sum(changed);
    if (threadIdx.x == 0 && tot) atomicAdd(active, (unsigned long long) tot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fe += __shfl_xor_sync(0xffffffffu, fe, o);
    if ((threadIdx.x & 31) == 0 && fe) atomicAdd(stats + 3, fe);
}
// After the first applicator pass: vertices whose row is empty everywhere take applicator(state) -> false (:1726-1738,
// :38), so they are never active again and their x stays infinity() from now on.
__global__ void k_ns_clear_empty(uint8_t* __restrict__ C, const uint8_t* __restrict__ I, const uint8_t* __restrict__ J, const uint32_t* __restrict__ JV,
                                 uint32_t th, uint32_t* __restrict__ x) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x)
        if (!I[i]) { C[i] = 0; if (J[i]) x[JV[i]] = GT_INF_U32; }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
static inline size_t round4(size_t n) { return (n + 3) / 4 * 4; }

void ns_alloc(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    gt_graph* g = P->g;
    cudaStream_t st = ctx->stream;
    std::unique_ptr<NsState> Np(new NsState());
    NsState& N = *Np;
    N.S = P->pcol->size(); N.R = P->prow->size();
    for (const SegMaps& s : *P->pcol) N.xchunk = std::max<size_t>(N.xchunk, s.nnz);
    for (const SegMaps& s : *P->prow) N.ychunk = std::max<size_t>(N.ychunk, s.nnz);
    N.xchunk = std::max<size_t>(4, round4(N.xchunk)); N.ychunk = std::max<size_t>(4, round4(N.ychunk));
    auto chunk_of = [&](CommGroup grp, int segment, size_t k) {
        return ctx->comm ? (size_t) comm_index_of_world_rank(ctx->comm, grp, g->lay.leader_ranks[segment]) : k;
    };
    N.xq.resize(N.S); N.yq.resize(N.R);
    for (size_t k = 0; k < N.S; k++) N.xq[k] = (int) chunk_of(P->bcast_group, (*P->pcol)[k].segment, k);
    for (size_t k = 0; k < N.R; k++) N.yq[k] = (int) chunk_of(P->reduce_group, (*P->prow)[k].segment, k);
    const int Sg = ctx->comm ? comm_size_in(ctx->comm, P->bcast_group) : 1;
    N.G = ctx->comm ? comm_size_in(ctx->comm, P->reduce_group) : 1;
    const char* e = getenv("GT_PEER");
    const bool want_peer = ctx->comm && !(e && atoi(e) == 0) && N.S <= 8 && N.G <= 8;
    const size_t xwords = 3 * N.S * N.xchunk + 4 * N.S;
    // both windows are created (or not) on every rank alike: peer_window_create is collective over the world
    if (want_peer && Sg > 1) N.wx = peer_window_create(ctx, P->bcast_group, xwords * sizeof(uint32_t));
    if (want_peer && N.G > 1 && (Sg == 1 || N.wx)) {
        N.wy = peer_window_create(ctx, P->reduce_group, (3 * (size_t) N.G * N.ychunk + 4 * (size_t) N.G) * sizeof(uint32_t));
        if (!N.wy && N.wx) { peer_window_destroy(ctx, N.wx); N.wx = nullptr; }
    }
    N.nccl_exchange = ctx->comm && ((Sg > 1 && !N.wx) || (N.G > 1 && !N.wy));
    if (N.wx) N.xbase = (uint32_t*) N.wx->local;
    else {
        N.xlocal.alloc(xwords);
        GT_CUDA(cudaMemsetAsync(N.xlocal.p, 0, N.xlocal.bytes(), st));
        N.xbase = N.xlocal.p;
    }
    N.hdr = N.xbase + 3 * N.S * N.xchunk;
    N.x.resize(N.S); N.xi.resize(N.S); N.xv.resize(N.S);
    for (size_t k = 0; k < N.S; k++) {
        N.x[k] = N.xbase + (size_t) N.xq[k] * N.xchunk;
        N.xi[k] = N.xbase + (N.S + N.xq[k]) * N.xchunk;
        N.xv[k] = N.xbase + (2 * N.S + N.xq[k]) * N.xchunk;
    }
    N.Y.alloc(N.R * N.ychunk);
    N.y.resize(N.R);
    for (size_t k = 0; k < N.R; k++) N.y[k] = N.Y.p + (size_t) N.yq[k] * N.ychunk;
    // tiles
    N.ntiles = (int) g->tiles.size();
    const bool y_sparse = N.wy != nullptr;
    if (y_sparse) {
        N.T.alloc(N.R * N.ychunk);
        GT_CUDA(cudaMemsetAsync(N.T.p, 0, N.T.bytes(), st));
    }
    std::vector<NsTile> ht(N.ntiles);
    uint64_t nnz_all = 0;
    for (int k = 0; k < N.ntiles; k++) {
        const Tile& T = g->tiles[k];
        NsTile& Q = ht[k];
        Q.JA = T.JA.p; Q.IA = g->IA_pool.p + T.offset; Q.A = g->weighted ? g->A_pool.p + T.offset : nullptr; Q.chunk_col = T.chunk_col.p;
        Q.nnz = T.nnz; Q.nchunks = (uint32_t) ((T.nnz + GT_PUSH_CHUNK - 1) / GT_PUSH_CHUNK); Q.ncols = g->cols[T.col_slot].nnz;
        Q.x = N.x[T.col_slot]; Q.xi = N.xi[T.col_slot]; Q.xv = N.xv[T.col_slot]; Q.hdr = N.hdr + 4 * (size_t) N.xq[T.col_slot];
        Q.y = N.y[T.row_slot];
        Q.t = (y_sparse && (int) T.row_slot != P->own_row_slot) ? N.T.p + (size_t) N.yq[T.row_slot] * N.ychunk : nullptr;
        Q.dense_bytes = T.nnz ? (g->weighted ? 8ull : 4ull) * T.nnz + 4ull * ((uint64_t) Q.ncols + 1) + 4ull * Q.ncols : 0;
        N.any_heavy |= T.max_col_entries > kNsHeavyColumn;
        nnz_all += T.nnz;
    }
    N.tiles.alloc(N.ntiles);
    GT_CUDA(cudaMemcpyAsync(N.tiles.p, ht.data(), ht.size() * sizeof(NsTile), cudaMemcpyHostToDevice, st));
    N.htiles = ht;
    N.heavy_list.alloc(nnz_all / kNsHeavyChunk + nnz_all / kNsHeavyColumn + 64);
    N.nsend = N.nrecv = N.wy ? N.G - 1 : 0;
    N.counters.alloc(3 + 2 * (size_t) std::max(1, N.nsend));
    GT_CUDA(cudaMemsetAsync(N.counters.p, 0, N.counters.bytes(), st));
    N.slot_counts.alloc(std::max<size_t>(1, N.S));
    N.stats.alloc(4);
    GT_CUDA(cudaMemsetAsync(N.stats.p, 0, N.stats.bytes(), st));
    GT_CUDA(cudaMallocHost((void**) &N.h_stats, 4 * sizeof(unsigned long long)));
    if (N.wy) {
        N.ystage.alloc(2 * N.ychunk * (size_t) N.nsend);
        std::vector<NsYSend> hs;
        std::vector<NsYRecv> hr;
        const int me = N.wy->me;
        const size_t G = (size_t) N.G;
        for (size_t r = 0; r < N.R; r++) {
            if ((int) r == P->own_row_slot) continue;
            const int q = N.yq[r];                                    // the leader's group rank
            NsYSend s{};
            s.y = N.y[r]; s.t = N.T.p + (size_t) q * N.ychunk; s.n = (*P->prow)[r].nnz;
            s.yi = N.ystage.p + 2 * N.ychunk * hs.size(); s.yv = s.yi + N.ychunk;
            s.count = N.counters.p + 3 + 2 * hs.size(); s.done = s.count + 1;
            uint32_t* base = (uint32_t*) N.wy->remote[q];
            s.dst_dense = base + (size_t) me * N.ychunk; s.dst_yi = base + (G + me) * N.ychunk; s.dst_yv = base + (2 * G + me) * N.ychunk;
            s.dst_hdr = base + 3 * G * N.ychunk + 4 * (size_t) me; s.dst_flag = N.wy->flag(q, me);
            hs.push_back(s);
        }
        const uint32_t* mine = (const uint32_t*) N.wy->local;
        for (int m = 0; m < N.G; m++) {
            if (m == me) continue;
            NsYRecv r{};
            r.dense = mine + (size_t) m * N.ychunk; r.yi = mine + (G + m) * N.ychunk; r.yv = mine + (2 * G + m) * N.ychunk;
            r.hdr = mine + 3 * G * N.ychunk + 4 * (size_t) m; r.n = (*P->prow)[P->own_row_slot].nnz;
            hr.push_back(r);
        }
        GT_REQUIRE((int) hs.size() == N.nsend && (int) hr.size() == N.nrecv, "non-stationary engine: every member of a row group leads exactly one of its segments");
        N.ysend.alloc(hs.size()); N.yrecv.alloc(hr.size());
        GT_CUDA(cudaMemcpyAsync(N.ysend.p, hs.data(), hs.size() * sizeof(NsYSend), cudaMemcpyHostToDevice, st));
        GT_CUDA(cudaMemcpyAsync(N.yrecv.p, hr.data(), hr.size() * sizeof(NsYRecv), cudaMemcpyHostToDevice, st));
    }
    // bottom-up BFS needs the symmetric single tile of an undirected graph on one GPU
    N.bottom_up_ok = P->app == GT_APP_BFS && ctx->nranks == 1 && !g->flags.directed && !g->weighted && N.ntiles == 1 &&
                     (*P->prow)[0].nnz == (*P->pcol)[0].nnz;
    for (int i = 0; i < 2; i++) GT_CUDA(cudaEventCreateWithFlags(&N.ev_it[i], cudaEventDisableTiming));
    GT_CUDA(cudaStreamSynchronize(st));                               // the host vectors above go out of scope
    P->ns = Np.release();
}

void ns_free(gt_program* P) {
    NsState* N = P->ns;
    if (!N) return;
    if (N->wx || N->wy) {                      // every put into these windows was consumed before execute() returned
        peer_window_destroy(P->ctx, N->wx);
        peer_window_destroy(P->ctx, N->wy);
    }
    for (int i = 0; i < 2; i++) if (N->ev_it[i]) cudaEventDestroy(N->ev_it[i]);
    if (N->h_stats) cudaFreeHost(N->h_stats);
    delete N;
    P->ns = nullptr;
}

void ns_initialize(gt_program* P) {            // Y starts at infinity() (:625-635)
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    k_fill<uint32_t><<<grid_for(N.Y.n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(N.Y.p, GT_INF_U32, N.Y.n);
    ctx->kernel_launches++;
    if (N.T.p) GT_CUDA(cudaMemsetAsync(N.T.p, 0, N.T.bytes(), ctx->stream));
    GT_CUDA(cudaGetLastError());
}

// x of the owned column segment + its frontier list from (V, C): scatter_gather_nonstationary (:710-758)
static void ns_x_from_state(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    cudaStream_t st = ctx->stream;
    const SegMaps& own = (*P->pcol)[P->own_col_slot];
    const int k = P->own_col_slot;
    GT_CUDA(cudaMemsetAsync(N.counters.p + 1, 0, sizeof(unsigned int), st));
    GT_CUDA(cudaMemsetAsync(N.stats.p + 3, 0, sizeof(unsigned long long), st));
    if (own.nnz) {
        k_ns_messenger<<<grid_for(own.nnz, 256, ctx->sm_count), 256, 0, st>>>(P->vs(), P->app, P->vid0, own.ids.p, own.nnz, N.x[k], P->g->col_deg[k].p, N.stats.p + 3);
        k_ns_frontier<<<grid_for((own.nnz + 3) / 4, 1024, ctx->sm_count, 2), 1024, 0, st>>>(N.x[k], own.nnz, N.xi[k], N.xv[k], N.counters.p + 1);
        ctx->kernel_launches += 2;
    }
    GT_CUDA(cudaGetLastError());
}

// the exchange half of scatter_gather: header + puts (:760-784,864-1013)
static void ns_exchange_x(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    cudaStream_t st = ctx->stream;
    const SegMaps& own = (*P->pcol)[P->own_col_slot];
    const int k = P->own_col_slot;
    const double bu = N.bottom_up_ok ? P->bfs_bottom_up_ratio : 0.0;
    const NsRule rule{P->activity_filtering_ratio, bu, P->activity_filtering_ratio >= 1.0 ? 0.0 : P->dense_edge_ratio, P->g->col_edges[k]};
    uint32_t* own_hdr = N.hdr + 4 * (size_t) N.xq[k];
    if (N.nccl_exchange) {
        // GT_PEER=0 or no peer mapping: ONE in-place all-gather of the dense segments; every rank rebuilds the lists of
        // the other segments and applies the owner's rule itself
        if (comm_size_in(ctx->comm, P->bcast_group) > 1) comm_allgather_inplace(ctx->comm, P->bcast_group, N.xbase, N.xchunk, CT_U32, st);
        GT_CUDA(cudaMemsetAsync(N.slot_counts.p, 0, N.slot_counts.bytes(), st));
        for (size_t s = 0; s < N.S; s++) {
            const SegMaps& seg = (*P->pcol)[s];
            unsigned int* cnt = (int) s == k ? N.counters.p + 1 : N.slot_counts.p + s;
            if ((int) s != k && seg.nnz) {
                k_ns_frontier<<<grid_for((seg.nnz + 3) / 4, 1024, ctx->sm_count, 2), 1024, 0, st>>>(N.x[s], seg.nnz, N.xi[s], N.xv[s], cnt);
                ctx->kernel_launches++;
            }
            k_ns_header<<<1, 1, 0, st>>>(cnt, seg.nnz, NsRule{P->activity_filtering_ratio, (int) s == k ? bu : 0.0, 0.0, 0ull}, N.hdr + 4 * (size_t) N.xq[s],
                                         (int) s == k ? N.stats.p + 3 : nullptr, P->d_active.p + N.active_slot);
            ctx->kernel_launches++;
        }
        GT_CUDA(cudaGetLastError());
        return;
    }
    PutTargets T{};
    if (N.wx) {
        const int me = N.wx->me;
        for (int j = 1; j < N.wx->size; j++) {
            const int q = (me + j) % N.wx->size;
            uint32_t* base = (uint32_t*) N.wx->remote[q];
            T.dense[T.n] = base + (size_t) me * N.xchunk;
            T.xi[T.n] = base + (N.S + me) * N.xchunk;
            T.xv[T.n] = base + (2 * N.S + me) * N.xchunk;
            T.hdr[T.n] = base + 3 * N.S * N.xchunk + 4 * (size_t) me;
            T.flag[T.n] = N.wx->flag(q, me);
            T.n++;
        }
    }
    N.x_epoch++;
    const int prev = N.active_slot ^ 1;
    k_ns_put_x<<<T.n ? 2 * ctx->sm_count : 1, 256, 0, st>>>(N.x[k], N.xi[k], N.xv[k], N.counters.p + 1, own.nnz, N.stats.p + 3, rule, own_hdr, T,
                                                            N.x_epoch, N.counters.p + 2, P->d_active.p + N.active_slot,
                                                            N.publish_prev ? P->d_active.p + prev : nullptr, N.publish_prev ? P->h_active + prev : nullptr);
    ctx->kernel_launches++;
    if (N.publish_prev) GT_CUDA(cudaEventRecord(N.ev_it[prev], st));
    tl_mark(P, "x_put", st);
    if (N.wx) { peer_wait_all(ctx, N.wx, N.x_epoch, st); tl_mark(P, "x_arrived", st); }
    GT_CUDA(cudaGetLastError());
}

static void ns_combine(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    cudaStream_t st = ctx->stream;
    const gt_graph* g = P->g;
    // (the counters this pass and the applicator accumulate into were zeroed by the kernels that consumed them: k_ns_put_x /
    // k_ns_header the frontier count, its edge count and the active count; k_ns_dense the heavy-column count)
    const int hgrid = ctx->sm_count * 4;
    const int nper = std::min(N.ntiles, kNsGroup);
    const int gx = std::max(1, ctx->sm_count * 8 / nper);                          // 8 CTAs of 256 threads per SM over the tiles of a launch
    auto group = [&](int t0) {
        NsTiles Tg{};
        for (int i = 0; i < kNsGroup && t0 + i < N.ntiles; i++) Tg.t[i] = N.htiles[t0 + i];
        return Tg;
    };
#define GT_NS_PASS(S, W) do { \
        for (int t0 = 0; t0 < N.ntiles; t0 += kNsGroup) { \
            k_ns_spmspv<S, W><<<dim3(gx, std::min(kNsGroup, N.ntiles - t0)), kNsBatch, 0, st>>>(group(t0), t0, N.heavy_list.p, N.counters.p, N.stats.p); \
            ctx->kernel_launches++; \
        } \
        if (N.any_heavy) { k_ns_heavy<S, W><<<hgrid, 256, 0, st>>>(N.tiles.p, N.heavy_list.p, N.counters.p); ctx->kernel_launches++; } \
        for (int t0 = 0; t0 < N.ntiles; t0 += kNsGroup) { \
            k_ns_dense<S, W><<<dim3(gx, std::min(kNsGroup, N.ntiles - t0)), kPushThreads, 0, st>>>(group(t0), N.stats.p, t0 == 0 ? N.counters.p : nullptr); \
            ctx->kernel_launches++; \
        } \
    } while (0)
    if (P->semiring == GT_MIN_PLUS_U32) GT_NS_PASS(GT_MIN_PLUS_U32, true);
    else if (g->weighted) GT_NS_PASS(GT_MIN_SELECT_U32, true);
    else GT_NS_PASS(GT_MIN_SELECT_U32, false);
#undef GT_NS_PASS
    if (N.bottom_up_ok && P->bfs_bottom_up_ratio > 0.0) {
        k_ns_bfs_bottom_up<<<grid_for((*P->pcol)[0].nnz, 256, ctx->sm_count, 8), 256, 0, st>>>(group(0), (*P->pcol)[0].ids.p, P->b.p, N.stats.p);
        ctx->kernel_launches++;
    }
    tl_mark(P, "tiles_done", st);
    if (N.wy) {
        N.y_epoch++;
        const dim3 gp(std::max(1, std::min<int>(ctx->sm_count, (int) ((N.ychunk + 16383) / 16384))), N.nsend);
        k_ns_pack_y<<<gp, 1024, 0, st>>>(N.ysend.p);
        k_ns_put_y<<<dim3(ctx->sm_count, N.nsend), 256, 0, st>>>(N.ysend.p, P->activity_filtering_ratio, N.y_epoch);
        tl_mark(P, "y_put", st);
        peer_wait_all(ctx, N.wy, N.y_epoch, st);
        tl_mark(P, "y_arrived", st);
        k_ns_merge_y<<<dim3(ctx->sm_count, N.nrecv), 256, 0, st>>>(N.yrecv.p, N.y[P->own_row_slot]);
        ctx->kernel_launches += 3;
    } else if (ctx->comm && N.G > 1) {
        comm_reduce_scatter_inplace(ctx->comm, P->reduce_group, N.Y.p, N.ychunk, CT_U32, CO_MIN, st);
    }
    GT_CUDA(cudaGetLastError());
}

// applicator (+ next x); the active count of this iteration goes to d_active[slot] and, all-reduced, to h_active[slot]
static void ns_apply(gt_program* P, uint32_t iteration, int slot) {
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    cudaStream_t st = ctx->stream;
    const SegMaps& row = (*P->prow)[P->own_row_slot];
    const SegMaps& col = (*P->pcol)[P->own_col_slot];
    const int k = P->own_col_slot;
    if (row.nnz) {
        k_ns_apply<<<grid_for((row.nnz + 3) / 4, 256, ctx->sm_count), 256, 0, st>>>(P->vs(), P->app, P->g->weighted, P->vid0, row.ids.p, row.nnz, N.y[P->own_row_slot],
                                                                                  iteration, col.bits.p, col.prefix.p, P->g->col_deg[k].p, N.x[k], N.xi[k], N.xv[k], N.counters.p + 1,
                                                                                  P->d_active.p + slot, N.stats.p);
        ctx->kernel_launches++;
    }
    if (!P->empty_cleared) {
        k_ns_clear_empty<<<grid_for(P->th, 256, ctx->sm_count), 256, 0, st>>>(P->C.p, row.bits.p, col.bits.p, col.prefix.p, P->th, N.x[k]);
        ctx->kernel_launches++;
        P->empty_cleared = true;
    }
    // has_converged (:1884-1923): the world all-reduce is also the ordering point that lets the windows have one buffer
    tl_mark(P, "applied", st);
    if (ctx->comm) { comm_allreduce(ctx->comm, COMM_WORLD, P->d_active.p + slot, P->d_active.p + slot, 1, CT_U64, CO_SUM, st); tl_mark(P, "allreduced", st); }
    if (!N.publish_via_put) {         // otherwise the next iteration's put kernel hands the count to the host (k_ns_put_x)
        GT_CUDA(cudaMemcpyAsync(P->h_active + slot, P->d_active.p + slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaEventRecord(N.ev_it[slot], st));
    }
    GT_CUDA(cudaGetLastError());
}

static void ns_check_peer_error(gt_program* P) {
    if (P->ns->wx || P->ns->wy) {
        if ((uint32_t) P->h_active[2]) {
            const uint32_t who = (uint32_t) P->h_active[2] - 1;
            P->h_active[2] = 0;
            P->poisoned = true;
            cudaMemsetAsync(peer_error_word(P->ctx), 0, 4, P->ctx->stream);
            throw Error(GT_ERR_NCCL, "gt_program_execute: NVLink peer exchange timed out waiting for group member " + std::to_string(who) +
                                         " (results are invalid; free this program and create a new one)");
        }
    }
}

void ns_execute(gt_program* P, uint32_t num_iterations) {
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    cudaStream_t st = ctx->stream;
    const bool check = num_iterations == 0;
    const bool peer = N.wx || N.wy;
    uint32_t sample_it = 0;                              // iteration (from 0) the -DTIMING sample belongs to
    auto timed = [&](int which, double& acc, auto&& fn) {
        if (!P->timing) { fn(); return; }
        GT_CUDA(cudaStreamSynchronize(st));
        const auto t0 = std::chrono::steady_clock::now();
        fn();
        GT_CUDA(cudaStreamSynchronize(st));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        acc += ms;
        P->add_sample(which, sample_it, ms);
    };
    GT_CUDA(cudaMemsetAsync(N.stats.p, 0, 3 * sizeof(unsigned long long), st));
    if (peer) peer_fence_world(ctx, st);                 // every rank has left whatever used the windows before (run_phase, an earlier execute)
    timed(0, P->tm.scatter_gather_ms, [&] { ns_x_from_state(P); });
    const uint32_t it0 = P->iteration;
    // run-ahead convergence mode: iteration n + 1 is always enqueued before the count of iteration n is read, so its put
    // kernel can deliver that count (fixed-iteration runs never read it; the timing knob serialises and keeps the copy)
    N.publish_via_put = check && !P->timing && !N.nccl_exchange;
    auto enqueue = [&](uint32_t n) {                     // iteration it0 + n
        P->iteration = it0 + n;
        sample_it = n;
        N.active_slot = (int) (n & 1);
        N.publish_prev = N.publish_via_put && n > 0;
        timed(0, P->tm.scatter_gather_ms, [&] { ns_exchange_x(P); });
        timed(1, P->tm.combine_ms, [&] { ns_combine(P); });
        timed(2, P->tm.apply_ms, [&] { ns_apply(P, it0 + n, (int) (n & 1)); });
    };
    // vertex-phase bytes per iteration: JC + state for the messenger, IR + y + state for the applicator
    const uint64_t vertex_bytes = 8ull * (*P->pcol)[P->own_col_slot].nnz + 16ull * (*P->prow)[P->own_row_slot].nnz;
    uint32_t done = 0;
    enqueue(0);
    while (true) {
        if (check) {
            // iteration done+1 is enqueued before the count of iteration `done` is read: if `done` turns out to be the last
            // one, the extra iteration finds an empty frontier everywhere and changes nothing
            if (!P->timing) enqueue(done + 1);
            GT_CUDA(cudaEventSynchronize(N.ev_it[done & 1]));
            const bool conv = P->h_active[done & 1] == 0;
            done++;
            if (conv) { P->converged = true; break; }
            if (P->timing) enqueue(done);
        } else {
            done++;
            if (it0 + done >= num_iterations) break;
            enqueue(done);
        }
    }
    P->iteration = it0 + done;
    N.publish_via_put = N.publish_prev = false;
    GT_CUDA(cudaMemcpyAsync(N.h_stats, N.stats.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (peer) GT_CUDA(cudaMemcpyAsync(&P->h_active[2], peer_error_word(ctx), 4, cudaMemcpyDeviceToHost, st));
    GT_CUDA(cudaEventRecord(P->ev1, st));
    GT_CUDA(cudaStreamSynchronize(st));
    ns_check_peer_error(P);
    tl_dump(P);
    P->tm.bytes_algorithmic = N.h_stats[0] + (uint64_t) done * vertex_bytes;
    P->tm.sparse_iterations = (uint32_t) N.h_stats[1];
}

void ns_run_phase(gt_program* P, int phase) {
    gt_ctx* ctx = P->ctx;
    NsState& N = *P->ns;
    N.active_slot = 0;
    N.publish_via_put = N.publish_prev = false;
    if (phase == 0) {
        if (N.wx || N.wy) peer_fence_world(ctx, ctx->stream);
        ns_x_from_state(P);
        ns_exchange_x(P);
        GT_CUDA(cudaStreamSynchronize(ctx->stream));
    } else if (phase == 1) ns_combine(P);
    else ns_apply(P, P->iteration, 0);
}

}  // namespace gt
