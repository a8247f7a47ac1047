// gt_kernels.cuh — the hot path: per-semiring SpMV / SpMSpV over TCSC tiles, hand-written for sm_100a.
//
// Reference loops being replaced (src/vp/vertex_program.hpp):
//   _ROW_ push   for j<ncols: for i in [JA[j],JA[j+1]): combiner(y[IA[i]], x[j][, A[i]])        :1164-1172
//   _COL_ pull   for j<ncols: for i in [JA[j],JA[j+1]): combiner(y[j], x[IA[i]][, A[i]])        :1175-1183
//   dense  ns    same as push, skipping columns with x[j] == infinity()                         :1491-1502
//   sparse ns    for k<frontier: j = xi[k]; same inner loop with xv[k]; t[IA[i]] = 1             :1476-1488
// The per-edge virtual combiner() becomes a compile-time semiring.
#pragma once
#include "gt_graph.h"

namespace gt {

// ---- semirings -------------------------------------------------------------------------------------
template <int S> struct Semiring;
template <> struct Semiring<GT_PLUS_TIMES_F64> {          // src/apps/pr.h:35-41, deg.h:41-47
    typedef double T;
    static constexpr bool kSkipInf = false;
    __device__ static __forceinline__ T mul(T x, uint32_t w) { return x * (double) w; }
    __device__ static __forceinline__ bool skip(T) { return false; }
    __device__ static __forceinline__ void reduce(T* y, T v) { atomicAdd(y, v); }     // RED.E.ADD.F64
    __device__ static __forceinline__ void reduce_dense(T* y, T v) { atomicAdd(y, v); }
    __device__ static __forceinline__ T identity() { return 0.0; }
    __device__ static __forceinline__ T plus(T a, T b) { return a + b; }
};
template <> struct Semiring<GT_MIN_PLUS_U32> {            // src/apps/sssp.h:49-52
    typedef uint32_t T;
    __device__ static __forceinline__ T mul(T x, uint32_t w) { return x + w; }
    __device__ static __forceinline__ bool skip(T x) { return x == GT_INF_U32; }
    __device__ static __forceinline__ void reduce(T* y, T v) { atomicMin(y, v); }     // RED.E.MIN
    // Dense passes only: y only ever decreases, so a (possibly stale) L2 read that is already <= v proves the
    // RED useless; hub rows settle after a few updates and stop serialising in L2 (CC RMAT-24: 12.0 -> 9.1 ms).
    // Not used by the frontier kernel, where most updates are first visits and the extra read costs 2-3x.
    __device__ static __forceinline__ void reduce_dense(T* y, T v) { if (v < __ldcg(y)) atomicMin(y, v); }
    __device__ static __forceinline__ T identity() { return GT_INF_U32; }
    __device__ static __forceinline__ T plus(T a, T b) { return a < b ? a : b; }
};
template <> struct Semiring<GT_MIN_SELECT_U32> {          // src/apps/bfs.h:61-63, cc.h:47-49
    typedef uint32_t T;
    __device__ static __forceinline__ T mul(T x, uint32_t w) { return x + w; }      // bfs.h:56-59 (weighted build)
    __device__ static __forceinline__ bool skip(T x) { return x == GT_INF_U32; }
    __device__ static __forceinline__ void reduce(T* y, T v) { atomicMin(y, v); }
    __device__ static __forceinline__ void reduce_dense(T* y, T v) { if (v < __ldcg(y)) atomicMin(y, v); }
    __device__ static __forceinline__ T identity() { return GT_INF_U32; }
    __device__ static __forceinline__ T plus(T a, T b) { return a < b ? a : b; }
};

// streaming 128-bit load that does not pollute L1 (index arrays are read exactly once per pass)
__device__ __forceinline__ uint4 ld_stream_u4(const uint32_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// ---- push SpMV, load-balanced over edges ---------------------------------------------------------------
// One CTA work item = GT_PUSH_CHUNK consecutive edges of the tile, whatever columns they belong to, so
// RMAT's 1..10^5 column-degree skew never unbalances the grid.  Per chunk:
//   1. the chunk's first column comes from a build-time table (no search on the hot path);
//   2. the columns overlapping the chunk are read from JA once, coalesced, and every non-empty one
//      drops its id at its first edge's slot in shared memory;
//   3. an inclusive max-scan over the slots turns those markers into a column id per edge;
//   4. IA (and A) are read with 128-bit streaming loads, x[col] with cached loads (consecutive edges
//      share columns), and the result is combined into y with one RED per edge.
constexpr int kPushThreads = 256;
constexpr int kPushPerThread = GT_PUSH_CHUNK / kPushThreads;   // 8 edges = two 128-bit loads

// IMPROVED (engine, min semirings): t marks the rows whose y really decreased — what a follower owes its leader —
// instead of every visited row (the kernel-level entry point keeps the reference's meaning, :1486).
template <int S, bool WEIGHTED, bool SKIP_INF, bool IMPROVED>
__device__ __forceinline__ void
spmv_push_chunks(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ IA, const uint32_t* __restrict__ A,
                 const uint32_t* __restrict__ chunk_col, uint32_t nchunks, uint64_t nnz,
                 const typename Semiring<S>::T* __restrict__ x, typename Semiring<S>::T* __restrict__ y, uint8_t* __restrict__ t,
                 uint32_t first_chunk, uint32_t chunk_stride) {
    typedef Semiring<S> SR;
    typedef typename SR::T T;
    __shared__ uint32_t colof[GT_PUSH_CHUNK];
    __shared__ uint32_t warp_max[kPushThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (uint32_t chunk = first_chunk; chunk < nchunks; chunk += chunk_stride) {
        const uint64_t e0 = (uint64_t) chunk * GT_PUSH_CHUNK;
        const uint32_t cnt = (uint32_t) min((uint64_t) GT_PUSH_CHUNK, nnz - e0);
        const uint32_t c0 = chunk_col[chunk];
        const uint32_t c1 = chunk_col[chunk + 1];      // column of the next chunk's first edge (or ncols)
#pragma unroll
        for (int k = 0; k < kPushPerThread; k++) colof[tid + k * kPushThreads] = 0;
        __syncthreads();
        // 2. markers.  Column c covers [JA[c], JA[c+1]); columns c0..c1 can overlap this chunk.
        for (uint32_t c = c0 + tid; c <= c1; c += kPushThreads) {
            const uint64_t s = JA[c];
            if (s >= e0 + cnt) break;                   // JA is monotone: later columns start later
            const uint64_t en = JA[c + 1];              // c1 <= ncols-1 when this executes (s < nnz)
            if (en > s && en > e0) colof[s > e0 ? (uint32_t) (s - e0) : 0u] = c;
        }
        __syncthreads();
        // 3. inclusive max-scan, 8 consecutive slots per thread
        uint32_t v[kPushPerThread];
        {
            const uint4 a = *reinterpret_cast<const uint4*>(&colof[tid * kPushPerThread]);
            const uint4 b = *reinterpret_cast<const uint4*>(&colof[tid * kPushPerThread + 4]);
            v[0] = a.x; v[1] = max(v[0], a.y); v[2] = max(v[1], a.z); v[3] = max(v[2], a.w);
            v[4] = max(v[3], b.x); v[5] = max(v[4], b.y); v[6] = max(v[5], b.z); v[7] = max(v[6], b.w);
        }
        uint32_t run = v[kPushPerThread - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, run, o);
            if (lane >= o) run = max(run, n);
        }
        if (lane == 31) warp_max[wid] = run;
        __syncthreads();
        uint32_t carry = __shfl_up_sync(0xffffffffu, run, 1);
        if (lane == 0) carry = 0;
        for (int w = 0; w < wid; w++) carry = max(carry, warp_max[w]);
#pragma unroll
        for (int k = 0; k < kPushPerThread; k++) v[k] = max(v[k], carry);
        __syncthreads();                                // every thread has read its slots and warp_max
        *reinterpret_cast<uint4*>(&colof[tid * kPushPerThread]) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(&colof[tid * kPushPerThread + 4]) = make_uint4(v[4], v[5], v[6], v[7]);
        __syncthreads();
        // 4. edges: thread handles slots [4*tid, 4*tid+4) and [1024+4*tid, ...)
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t s0 = half * (GT_PUSH_CHUNK / 2) + tid * 4;
            if (s0 >= cnt) continue;
            uint32_t rows[4], wts[4] = {1, 1, 1, 1};
            const uint32_t m = min(4u, cnt - s0);
            if (m == 4) {                               // e0 and s0 are multiples of 4: 16-byte aligned
                const uint4 r4 = ld_stream_u4(IA + e0 + s0);
                rows[0] = r4.x; rows[1] = r4.y; rows[2] = r4.z; rows[3] = r4.w;
                if (WEIGHTED) { const uint4 w4 = ld_stream_u4(A + e0 + s0); wts[0] = w4.x; wts[1] = w4.y; wts[2] = w4.z; wts[3] = w4.w; }
            } else {
                for (uint32_t k = 0; k < m; k++) { rows[k] = IA[e0 + s0 + k]; if (WEIGHTED) wts[k] = A[e0 + s0 + k]; }
            }
            const uint4 c4 = *reinterpret_cast<const uint4*>(&colof[s0]);
            const uint32_t cols[4] = {c4.x, c4.y, c4.z, c4.w};
            T xv = SR::identity();
            uint32_t last = 0xffffffffu;
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                if (k >= m) break;
                if (cols[k] != last) { xv = __ldg(x + cols[k]); last = cols[k]; }
                if (SKIP_INF && SR::skip(xv)) continue;
                const T val = WEIGHTED ? SR::mul(xv, wts[k]) : xv;
                if constexpr (IMPROVED) {
                    // the filter read goes to L2, one per edge, in line: through L1 (stale lines are safe, y never increases)
                    // and / or with a thread's four reads issued ahead of its REDs it measured 2-7 % slower
                    // (profiles/r02_ns_filter_read_variants.md)
                    // t marks every row a RED went out for: a superset of the rows that improved (all a follower owes its
                    // leader) and a subset of the reference's touched rows (:1486).  Waiting for the atomic's old value to
                    // mark only the winners turns every RED into a round trip (ATOM), which is what bounded these passes
                    if (t) { if (val < __ldcg(y + rows[k])) { atomicMin(y + rows[k], val); t[rows[k]] = 1; } }
                    else SR::reduce_dense(y + rows[k], val);
                } else {
                    SR::reduce_dense(y + rows[k], val);
                    if (t) t[rows[k]] = 1;
                }
            }
        }
        __syncthreads();                                // colof is rewritten by the next chunk
    }
}

template <int S, bool WEIGHTED, bool SKIP_INF>
__global__ void __launch_bounds__(kPushThreads)
k_spmv_push(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ IA, const uint32_t* __restrict__ A,
            const uint32_t* __restrict__ chunk_col, uint32_t nchunks, uint64_t nnz,
            const typename Semiring<S>::T* __restrict__ x, typename Semiring<S>::T* __restrict__ y, uint8_t* __restrict__ t) {
    spmv_push_chunks<S, WEIGHTED, SKIP_INF, false>(JA, IA, A, chunk_col, nchunks, nnz, x, y, t, blockIdx.x, gridDim.x);
}

// ---- pull SpMV (_COL_ ordering): y[j] (+)= sum over column j of x[IA[i]] ---------------------------------
// One warp per column, lanes stride the column's edges, shuffle reduction, one plain RMW per column
// (a column belongs to exactly one warp and tiles run back to back on one stream).
template <int S, bool WEIGHTED>
__global__ void __launch_bounds__(256)
k_spmv_pull(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ IA, const uint32_t* __restrict__ A, uint32_t ncols,
            const typename Semiring<S>::T* __restrict__ x, typename Semiring<S>::T* __restrict__ y) {
    typedef Semiring<S> SR;
    typedef typename SR::T T;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j = warp; j < ncols; j += nwarps) {
        const uint32_t b = JA[j], e = JA[j + 1];
        if (b == e) continue;
        T acc = SR::identity();
        for (uint32_t i = b + lane; i < e; i += 32) {
            const T xv = __ldg(x + IA[i]);
            if (SR::skip(xv)) continue;
            acc = SR::plus(acc, WEIGHTED ? SR::mul(xv, A[i]) : xv);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = SR::plus(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) y[j] = SR::plus(y[j], acc);
    }
}

// ---- frontier SpMSpV: only the k active columns ---------------------------------------------------------
// Body: one warp per frontier column, lanes stride the column's entries (columns of a frontier are independent,
// so the grid is as wide as the frontier).  RMAT hubs: a column with more than kHeavyColumn entries is not walked
// by its warp; the warp appends one (frontier position, chunk) pair per kHeavyChunk entries to a list, and a second
// launch gives every pair a CTA (block path for heavy columns), so a 10^5..10^6-entry hub is spread over the whole
// grid instead of serialising an iteration behind a single warp.
constexpr uint32_t kHeavyColumn = 16384;
constexpr uint32_t kHeavyChunk = 8192;

template <int S, bool WEIGHTED>
__global__ void __launch_bounds__(256)
k_spmspv_push(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ IA, const uint32_t* __restrict__ A,
              const uint32_t* __restrict__ xi, const typename Semiring<S>::T* __restrict__ xv, uint32_t k,
              typename Semiring<S>::T* __restrict__ y, uint8_t* __restrict__ t,
              uint2* __restrict__ heavy_list, unsigned int* __restrict__ heavy_count) {
    typedef Semiring<S> SR;
    typedef typename SR::T T;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t f = warp; f < k; f += nwarps) {
        const uint32_t j = xi[f];
        const T v = xv[f];
        const uint32_t b = JA[j], e = JA[j + 1];
        if (heavy_list && e - b > kHeavyColumn) {
            const uint32_t nchunks = (e - b + kHeavyChunk - 1) / kHeavyChunk;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(heavy_count, nchunks);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t c = lane; c < nchunks; c += 32) heavy_list[base + c] = make_uint2(f, c);
            continue;
        }
        for (uint32_t i = b + lane; i < e; i += 32) {
            const uint32_t r = IA[i];
            SR::reduce(y + r, WEIGHTED ? SR::mul(v, A[i]) : v);
            if (t) t[r] = 1;
        }
    }
}

template <int S, bool WEIGHTED>
__global__ void __launch_bounds__(256)
k_spmspv_heavy(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ IA, const uint32_t* __restrict__ A,
               const uint32_t* __restrict__ xi, const typename Semiring<S>::T* __restrict__ xv,
               typename Semiring<S>::T* __restrict__ y, uint8_t* __restrict__ t,
               const uint2* __restrict__ heavy_list, const unsigned int* __restrict__ heavy_count) {
    typedef Semiring<S> SR;
    typedef typename SR::T T;
    const unsigned int n = *heavy_count;
    for (unsigned int h = blockIdx.x; h < n; h += gridDim.x) {          // one CTA per (heavy column, chunk)
        const uint2 fc = heavy_list[h];
        const uint32_t j = xi[fc.x];
        const T v = xv[fc.x];
        const uint32_t b = JA[j] + fc.y * kHeavyChunk, e = min(JA[j + 1], b + kHeavyChunk);
        for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
            const uint32_t r = IA[i];
            SR::reduce(y + r, WEIGHTED ? SR::mul(v, A[i]) : v);
            if (t) t[r] = 1;
        }
    }
}

}  // namespace gt
