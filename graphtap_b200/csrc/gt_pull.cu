// gt_pull.cu — the plus-times SpMV of the stationary programs (PageRank) as a PULL over a layout derived
// from the TCSC tiles, designed for what a B200 can and cannot do (profiles/r01_microbench_*.log):
//
//   * f64 RED.ADD into y collapses to 6-20 Gop/s under RMAT's row skew (hub rows serialise in L2), while
//     random 8-byte gathers run at 275-565 Gop/s from L2 and ~1000 Gop/s from shared memory.  So the
//     reference's push loop  y[IA[i]] += x[j]  (src/vp/vertex_program.hpp:1164-1172) is turned around:
//     every row sums its own x values, no atomics on the common path.
//   * Every vertex segment gets ONE "hot order" (gt_graph.h HotOrder): its vertices with any entry, by
//     decreasing global degree, computed at ingest from the whole edge list so all ranks agree without
//     talking.  Both the x and the y vector of the segment are indexed in that order.  Hot x values are
//     packed densely at the front, so they stay L1/L2 resident and only the cold tail goes to HBM; and
//     because x and y of the owned segment share the order, applicator + messenger fuse into one
//     perfectly sequential pass over the vertex state (gt_engine.cu).  TCSC already renumbers rows and
//     columns to dense ranges (src/ds/compressed_column.hpp:381-416); this is one more renumbering of
//     the same kind.
//   * Rows are stored as SELL-32: virtual rows (below) sorted by decreasing length, 32 of them per slice,
//     column-major inside the slice, so lane l of a warp walks row l with perfectly coalesced 128-byte
//     index loads and neighbouring lanes have (nearly) equal trip counts.  Rows longer than `vrow`
//     entries are cut into virtual rows whose partial sums meet in y through one RED.ADD each
//     (<= degree/vrow per row), which bounds the skew a warp can see.
//   * No shared-memory staging of x: a large carve-out shrinks L1, and L1's capacity bounds the number of
//     gather misses in flight (ncu: MIO throttle 100 %, L1TEX 23 % with a 200 KB cache vs 84 % without;
//     profiles/r01_ncu_pull_s26_hot*.txt).  The kernel is bound by the L1-miss -> L2 gather path: one 32-byte
//     sector per edge, ~280 G gathers/s per GPU (~12.9 TB/s of L2 sector traffic), the structural limit of a
//     gather design.  L2 eviction hints (index stream evict-first, x evict-last) are worth 8 %.
//   * Multi-GPU: x is one equal-sized chunk per member of the column group; entries whose column lies in this
//     rank's own chunk form a separate SELL array that runs while the other chunks are still arriving over NVLink
//     (copy-engine puts into peer windows, gt_peer.cu; or one in-place all-gather with GT_PEER=0).
//
// Results: each row's sum is formed in a fixed order by one lane (split rows excepted), so it differs from
// the reference's column-order sum only by f64 rounding, ~1e-16 relative — inside the 1e-6 contract.
#include "gt_pull.h"
#include <cub/cub.cuh>
#include <algorithm>

namespace gt {

static inline int grid_for(uint64_t n, int block, int sm_count, int per_sm = 8) {
    uint64_t g = (n + block - 1) / block;
    uint64_t cap = (uint64_t) sm_count * per_sm;
    return (int) std::max<uint64_t>(1, std::min(g, cap));
}

// ---- build kernels -------------------------------------------------------------------------------------
// compressed id -> index in the segment's hot order (composition of JC / IR with HotOrder::pos)
__global__ void k_compose(const uint32_t* __restrict__ ids, uint32_t n, const uint32_t* __restrict__ pos, uint32_t add, uint32_t* __restrict__ out) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) out[j] = add + pos[ids[j]];
}
// column code of a compressed column: index into the concatenated x, | kPullHotBit for the segment's hottest
__global__ void k_col_codes(const uint32_t* __restrict__ ids, uint32_t n, const uint32_t* __restrict__ pos, uint32_t l1hot, uint32_t xoff,
                            uint32_t* __restrict__ code) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t r = pos[ids[j]];
        code[j] = (xoff + r) | (r < l1hot ? kPullHotBit : 0u);
    }
}
// one (row', colcode) key per stored entry of a tile
__global__ void k_expand(const uint32_t* __restrict__ JA, const uint32_t* __restrict__ IA, uint64_t nnz, uint32_t ncols,
                         const uint32_t* __restrict__ col_code, const uint32_t* __restrict__ row_rank, uint64_t* __restrict__ keys) {
    for (uint64_t e = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; e < nnz; e += (uint64_t) gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = ncols;                      // upper_bound(JA, e) - 1 = the column that holds entry e
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t) JA[mid] <= e) lo = mid + 1; else hi = mid;
        }
        keys[e] = ((uint64_t) row_rank[IA[e]] << 32) | col_code[lo - 1];
    }
}
__global__ void k_row_ptr(const uint64_t* __restrict__ keys, uint64_t n, uint32_t nrows, uint64_t* __restrict__ rowptr) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += gridDim.x * blockDim.x) {
        const uint64_t target = (uint64_t) r << 32;
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        rowptr[r] = lo;
    }
}
__global__ void k_vrow_counts(const uint64_t* __restrict__ rowptr, uint32_t nrows, uint32_t vlen, uint32_t* __restrict__ nv) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        const uint64_t len = rowptr[r + 1] - rowptr[r];
        nv[r] = (uint32_t) std::max<uint64_t>(1, (len + vlen - 1) / vlen);
    }
}
__global__ void k_vrow_make(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ vbase, uint32_t nrows, uint32_t vlen,
                            uint32_t* __restrict__ vkey, uint32_t* __restrict__ vid) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        const uint64_t len = rowptr[r + 1] - rowptr[r];
        const uint32_t b = vbase[r], n = vbase[r + 1] - b;
        for (uint32_t c = 0; c < n; c++) {
            const uint64_t l = std::min<uint64_t>(vlen, len - (uint64_t) c * vlen);
            vkey[b + c] = vlen - (uint32_t) l;           // ascending sort = decreasing length
            vid[b + c] = b + c;
        }
    }
}
// per sorted virtual row: start in the CSR key array, length, target row (+ split flag)
__global__ void k_vrow_meta(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ vbase, uint32_t nrows, uint32_t vlen,
                            const uint32_t* __restrict__ rank_of_v, uint64_t* __restrict__ vstart, uint32_t* __restrict__ vl, uint32_t* __restrict__ vtgt) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        const uint64_t len = rowptr[r + 1] - rowptr[r];
        const uint32_t b = vbase[r], n = vbase[r + 1] - b;
        for (uint32_t c = 0; c < n; c++) {
            const uint32_t pos = rank_of_v[b + c];
            vstart[pos] = rowptr[r] + (uint64_t) c * vlen;
            vl[pos] = (uint32_t) std::min<uint64_t>(vlen, len - (uint64_t) c * vlen);
            vtgt[pos] = r | (n > 1 ? kPullSplit : 0u);
        }
    }
}
__global__ void k_invert(const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ inv) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) inv[perm[i]] = i;
}
__global__ void k_slice_sizes(const uint32_t* __restrict__ vl, uint32_t nv, uint32_t nslices, uint64_t* __restrict__ sizes) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s <= nslices; s += gridDim.x * blockDim.x)
        sizes[s] = (s < nslices) ? 32ull * vl[(uint64_t) s * 32] : 0ull;     // rows are sorted by decreasing length
}
// SELL-32 fill: one warp per slice, lane l copies virtual row 32 s + l column-major
__global__ void __launch_bounds__(256) k_sell_fill(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vstart, const uint32_t* __restrict__ vl,
                                                    uint32_t nv, const uint64_t* __restrict__ slice_ptr, uint32_t nslices, uint32_t pad_code,
                                                    uint32_t* __restrict__ sell) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t s = warp; s < nslices; s += nwarps) {
        const uint64_t base = slice_ptr[s];
        const uint32_t L = (uint32_t) ((slice_ptr[s + 1] - base) >> 5);
        const uint32_t v = s * 32 + lane;
        const uint32_t len = v < nv ? vl[v] : 0;
        const uint64_t st = v < nv ? vstart[v] : 0;
        for (uint32_t k = 0; k < L; k++) sell[base + (uint64_t) k * 32 + lane] = k < len ? (uint32_t) keys[st + k] : pad_code;
    }
}
__global__ void k_count_nonzero_slices(const uint64_t* __restrict__ slice_ptr, uint32_t nslices, unsigned int* __restrict__ out) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nslices; s += gridDim.x * blockDim.x)
        if (slice_ptr[s + 1] > slice_ptr[s]) atomicMax(out, s + 1);
}

// ---- the hot kernel ----------------------------------------------------------------------------------------
// Persistent CTAs, one per SM.  Prologue: the hot x values of every local column segment -> shared memory.
// Then each warp takes slices round-robin (slices are sorted by decreasing length, so every warp gets the
// same mix of long and short ones).  Inner loop, per lane: 8 independent 4-byte index loads (streaming,
// coalesced across the warp), 8 independent gathers (shared memory for hot codes, read-only global for the
// rest), 8 adds.  One coalesced store of y per slice; virtual rows of split rows use RED.ADD.
// Column code = index into the concatenated x | kPullHotBit when the column is among the `l1hot` hottest of its
// segment.  Hot gathers use allocating loads (they are re-read by every warp of the SM, so L1 keeps them), cold
// gathers use L1::no_allocate so that they do not evict the hot lines.  With L2HINT the index stream is read
// evict-first and x evict-last in L2, so the 4 B/edge stream does not push x out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() { uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
template <bool L2HINT>
__device__ __forceinline__ uint32_t ld_index(const uint32_t* p, uint64_t pol) {
    uint32_t r;
    if (L2HINT) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
template <bool L1SPLIT, bool L2HINT>
__device__ __forceinline__ double ld_x(const double* __restrict__ x, uint32_t code, uint64_t pol) {
    double v;
    const double* p = x + (code & ~kPullHotBit);
    if (!L1SPLIT) {
        if (L2HINT) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        else v = __ldg(p);
    } else if (code & kPullHotBit) {       // hot: keep in L1 as long as possible
        if (L2HINT) asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        else asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    } else {                               // cold: still allocates (L1 lines track the misses in flight; no_allocate
                                           // gathers were 1.5x slower, profiles/r01_sweep_pull_cachepolicy_s26.log) but leaves first
        if (L2HINT) asm volatile("ld.global.nc.L1::evict_first.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        else asm volatile("ld.global.nc.L1::evict_first.f64 %0, [%1];" : "=d"(v) : "l"(p));
    }
    return v;
}

// Persistent CTAs.  Each warp takes slices round-robin; slices are sorted by decreasing length and dealt to
// consecutive CTAs (not to consecutive warps of one CTA), so every SM receives the same mix of long and short
// slices: with the naive order the longest slices all landed on the first few SMs and the rest of the chip idled
// (profiles/r01_ncu_pull_v0_s22.txt).  Inner loop, per lane: UNROLL independent 4-byte index loads (coalesced
// across the warp), UNROLL independent 8-byte gathers, UNROLL adds.  One store of y per virtual row; virtual rows
// of split rows use RED.ADD.
// ACCUM: 0 = y[t] = acc (first pass over the row), 1 = y[t] += acc (a row's owner lane is unique within a pass)
template <int UNROLL, bool L1SPLIT, bool L2HINT, int ACCUM>
__global__ void __launch_bounds__(kPullThreads)
k_spmv_pull_sell(const uint32_t* __restrict__ sell, const uint64_t* __restrict__ slice_ptr, uint32_t nslices,
                 const uint32_t* __restrict__ vtgt, uint32_t nv, const double* __restrict__ x, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nwarps = gridDim.x * (blockDim.x >> 5);
    const uint64_t pol_idx = L2HINT ? l2_policy_evict_first() : 0, pol_x = L2HINT ? l2_policy_evict_last() : 0;
    for (uint32_t s = warp; s < nslices; s += nwarps) {
        const uint64_t base = slice_ptr[s];
        const uint32_t L = (uint32_t) ((slice_ptr[s + 1] - base) >> 5);
        const uint32_t* p = sell + base + lane;
        double acc = 0.0;
        uint32_t k = 0;
        for (; k + UNROLL <= L; k += UNROLL) {
            uint32_t c[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) c[u] = ld_index<L2HINT>(p + (uint64_t) (k + u) * 32, pol_idx);
            double v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) v[u] = ld_x<L1SPLIT, L2HINT>(x, c[u], pol_x);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) acc += v[u];
        }
        for (; k < L; k++) acc += ld_x<L1SPLIT, L2HINT>(x, ld_index<L2HINT>(p + (uint64_t) k * 32, pol_idx), pol_x);
        const uint32_t v = s * 32 + lane;
        if (v < nv) {
            const uint32_t t = vtgt[v];
            if (t & kPullSplit) atomicAdd(y + (t & ~kPullSplit), acc);
            else if (ACCUM) y[t] += acc;              // later passes: a row's owner lane is unique within a pass
            else y[t] = acc;
        }
    }
}

// Hot-band variant (single GPU, GT_PULL_BAND + GT_PULL_BAND_SMEM=1): the band's x values [0, band) are staged in
// shared memory once per CTA and every gather of the pass is an LDS; the cold pass keeps the whole L1 for its misses
// (the kernel that mixed both starved L1, profiles/r01_ncu_pull_s26_hot*.txt).  Codes >= band are padding (0.0).
template <int UNROLL>
__global__ void __launch_bounds__(kPullThreads)
k_spmv_pull_sell_smem(const uint32_t* __restrict__ sell, const uint64_t* __restrict__ slice_ptr, uint32_t nslices,
                      const uint32_t* __restrict__ vtgt, uint32_t nv, const double* __restrict__ x, uint32_t band, double* __restrict__ y) {
    extern __shared__ double xs[];
    for (uint32_t i = threadIdx.x; i < band; i += blockDim.x) xs[i] = x[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nwarps = gridDim.x * (blockDim.x >> 5);
    const uint64_t pol_idx = l2_policy_evict_first();
    for (uint32_t s = warp; s < nslices; s += nwarps) {
        const uint64_t base = slice_ptr[s];
        const uint32_t L = (uint32_t) ((slice_ptr[s + 1] - base) >> 5);
        const uint32_t* p = sell + base + lane;
        double acc = 0.0;
        uint32_t k = 0;
        for (; k + UNROLL <= L; k += UNROLL) {
            uint32_t c[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) c[u] = ld_index<true>(p + (uint64_t) (k + u) * 32, pol_idx);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) acc += c[u] < band ? xs[c[u]] : 0.0;
        }
        for (; k < L; k++) { const uint32_t c = ld_index<true>(p + (uint64_t) k * 32, pol_idx); acc += c < band ? xs[c] : 0.0; }
        const uint32_t v = s * 32 + lane;
        if (v < nv) {
            const uint32_t t = vtgt[v];
            if (t & kPullSplit) atomicAdd(y + (t & ~kPullSplit), acc);
            else y[t] = acc;
        }
    }
}

struct CodeInRange {                            // column code (hot flag masked) inside [lo, hi)
    uint32_t lo, hi; bool want;
    __device__ bool operator()(const uint64_t& k) const {
        const uint32_t c = (uint32_t) k & ~kPullHotBit;
        return (c >= lo && c < hi) == want;
    }
};

// _TCSC_CF_ split of the (row', code) keys: 3 = source row, 2 = regular row x sink column, 0 = regular x regular
struct CfPart {
    uint32_t yreg, xchunk, xsnk0[kPullMaxSegs];      // first source-row position of the row segment; first sink position of every x chunk
    int want;
    __device__ int kind(const uint64_t& k) const {
        if ((uint32_t) (k >> 32) >= yreg) return 3;
        const uint32_t c = (uint32_t) k & ~kPullHotBit, q = c / xchunk;
        return (c - q * xchunk) >= xsnk0[q] ? 2 : 0;
    }
    __device__ bool operator()(const uint64_t& k) const { return kind(k) == want; }
};

// ---- build ---------------------------------------------------------------------------------------------------
// (row', code)-sorted keys of one row segment -> virtual rows (<= vrow entries) sorted by decreasing length -> SELL-32
static void build_sell(gt_ctx* ctx, const uint64_t* sorted, uint64_t total, uint32_t nr, uint32_t kVRow, uint32_t pad_code, PullSell& Q) {
    cudaStream_t st = ctx->stream;
    Q.nnz = total;
    if (!total || !nr) return;
    {
        DevBuf<uint64_t> rowptr; rowptr.alloc((size_t) nr + 1);
        k_row_ptr<<<grid_for((uint64_t) nr + 1, 256, ctx->sm_count), 256, 0, st>>>(sorted, total, nr, rowptr.p);
        // virtual rows
        DevBuf<uint32_t> nvr; nvr.alloc((size_t) nr + 1);
        GT_CUDA(cudaMemsetAsync(nvr.p + nr, 0, 4, st));
        k_vrow_counts<<<grid_for(nr, 256, ctx->sm_count), 256, 0, st>>>(rowptr.p, nr, kVRow, nvr.p);
        DevBuf<uint32_t> vbase; vbase.alloc((size_t) nr + 1);
        {
            size_t tb = 0;
            GT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nvr.p, vbase.p, (int64_t) nr + 1, st));
            DevBuf<uint8_t> tmp; tmp.alloc(tb);
            GT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, nvr.p, vbase.p, (int64_t) nr + 1, st));
            GT_CUDA(cudaStreamSynchronize(st));
        }
        uint32_t nv = 0;
        GT_CUDA(cudaMemcpyAsync(&nv, vbase.p + nr, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        Q.nv = nv;
        DevBuf<uint32_t> vkey, vkey_alt, vid, vid_alt;
        vkey.alloc(nv); vkey_alt.alloc(nv); vid.alloc(nv); vid_alt.alloc(nv);
        k_vrow_make<<<grid_for(nr, 256, ctx->sm_count), 256, 0, st>>>(rowptr.p, vbase.p, nr, kVRow, vkey.p, vid.p);
        uint32_t* vid_sorted = nullptr;
        {
            int lb = 1;
            while ((1u << lb) <= kVRow) lb++;
            cub::DoubleBuffer<uint32_t> dk(vkey.p, vkey_alt.p), dv(vid.p, vid_alt.p);
            size_t tb = 0;
            GT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int64_t) nv, 0, lb, st));
            DevBuf<uint8_t> tmp; tmp.alloc(tb);
            GT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int64_t) nv, 0, lb, st));
            GT_CUDA(cudaStreamSynchronize(st));
            vid_sorted = dv.Current();
        }
        DevBuf<uint32_t> vpos; vpos.alloc(nv);          // unsorted virtual row -> sorted position
        k_invert<<<grid_for(nv, 256, ctx->sm_count), 256, 0, st>>>(vid_sorted, nv, vpos.p);
        DevBuf<uint64_t> vstart; vstart.alloc(nv);
        DevBuf<uint32_t> vl; vl.alloc(nv);
        Q.vtgt.alloc(nv);
        k_vrow_meta<<<grid_for(nr, 256, ctx->sm_count), 256, 0, st>>>(rowptr.p, vbase.p, nr, kVRow, vpos.p, vstart.p, vl.p, Q.vtgt.p);
        const uint32_t nslices = (nv + 31) / 32;
        DevBuf<uint64_t> sizes; sizes.alloc((size_t) nslices + 1);
        Q.slice_ptr.alloc((size_t) nslices + 1);
        k_slice_sizes<<<grid_for((uint64_t) nslices + 1, 256, ctx->sm_count), 256, 0, st>>>(vl.p, nv, nslices, sizes.p);
        {
            size_t tb = 0;
            GT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, sizes.p, Q.slice_ptr.p, (int64_t) nslices + 1, st));
            DevBuf<uint8_t> tmp; tmp.alloc(tb);
            GT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, sizes.p, Q.slice_ptr.p, (int64_t) nslices + 1, st));
            GT_CUDA(cudaStreamSynchronize(st));
        }
        uint64_t sell_len = 0;
        GT_CUDA(cudaMemcpyAsync(&sell_len, Q.slice_ptr.p + nslices, 8, cudaMemcpyDeviceToHost, st));
        DevBuf<unsigned int> cnt; cnt.alloc(1);
        GT_CUDA(cudaMemsetAsync(cnt.p, 0, 4, st));
        k_count_nonzero_slices<<<grid_for(nslices, 256, ctx->sm_count), 256, 0, st>>>(Q.slice_ptr.p, nslices, cnt.p);
        unsigned int active = 0;
        GT_CUDA(cudaMemcpyAsync(&active, cnt.p, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        Q.nslices = active;
        Q.sell_len = sell_len;
        Q.sell.alloc(sell_len);
        k_sell_fill<<<grid_for((uint64_t) nslices * 32, 256, ctx->sm_count, 8), 256, 0, st>>>(sorted, vstart.p, vl.p, nv, Q.slice_ptr.p, active, pad_code, Q.sell.p);
        ctx->kernel_launches += 12;
        GT_CUDA(cudaGetLastError());
        GT_CUDA(cudaStreamSynchronize(st));
    }
}

PullLayout* pull_build(gt_graph* g) {
    gt_ctx* ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    GT_REQUIRE(!g->weighted, "pull layout: unweighted graphs only");
    std::unique_ptr<PullLayout> P(new PullLayout());
    const size_t S = g->cols.size(), R = g->rows.size();
    GT_REQUIRE(S <= kPullMaxSegs, "pull layout: too many local column segments");
    bool vrow_fixed = false;
    if (const char* e = getenv("GT_PULL_VROW")) { P->vrow = std::max(8, atoi(e)); vrow_fixed = true; }
    if (const char* e = getenv("GT_PULL_BAND")) P->band = (uint32_t) std::max(0, atoi(e));
    if (const char* e = getenv("GT_PULL_BAND_SMEM")) P->band_smem = atoi(e) != 0;
    GT_REQUIRE(!P->band_smem || (P->band && P->band <= 28000), "pull layout: GT_PULL_BAND_SMEM needs 0 < GT_PULL_BAND <= 28000 (224 KB of f64)");
    const bool verbose = getenv("GT_PULL_VERBOSE") && atoi(getenv("GT_PULL_VERBOSE"));
    if (const char* e = getenv("GT_PULL_L1HOT")) P->l1hot = (uint32_t) std::max(0, atoi(e));
    if (const char* e = getenv("GT_PULL_L2HINT")) P->l2hint = atoi(e) != 0;
    if (const char* e = getenv("GT_PULL_UNROLL")) P->unroll = atoi(e) == 4 ? 4 : 8;
    if (const char* e = getenv("GT_PULL_THREADS")) P->threads = std::min(1024, std::max(32, atoi(e) / 32 * 32));
    if (const char* e = getenv("GT_PULL_CTAS")) P->ctas_per_sm = std::min(8, std::max(1, atoi(e)));
    // A warp's longest slice bounds the tail of a launch, so the virtual-row length follows the launch size: one GPU's
    // share at p = 8 is 2^25 entries per launch, where 128 beats 512 by 14 % (profiles/r01_sweep_vrow.log).
    auto vrow_for = [&](uint64_t entries) -> uint32_t {
        if (vrow_fixed) return P->vrow;
        return entries < (48ull << 20) ? 128u : entries < (96ull << 20) ? 256u : P->vrow;
    };

    // concatenated x / y spaces (see PullLayout): chunk index = group rank of the segment's leader
    P->xoff.resize(S); P->xn.resize(S); P->yoff.resize(R); P->yn.resize(R);
    P->xreg.resize(S); P->xsnk0.resize(S); P->yreg.resize(R); P->ysrc.resize(R);
    P->cf = g->compression == GT_TCSC_CF;
    for (size_t k = 0; k < S; k++) {
        const HotOrder& H = g->hot[g->hot_of_col_slot[k]];
        P->xn[k] = H.n; P->xchunk = std::max(P->xchunk, P->xn[k]);
        P->xreg[k] = H.nreg; P->xsnk0[k] = P->cf ? H.nreg + H.nsrc : H.n;
    }
    for (size_t k = 0; k < R; k++) {
        const HotOrder& H = g->hot[g->hot_of_row_slot[k]];
        P->yn[k] = H.n; P->ychunk = std::max(P->ychunk, P->yn[k]);
        P->yreg[k] = H.nreg; P->ysrc[k] = H.nsrc;
    }
    P->xchunk = (P->xchunk + 1) / 2 * 2;       // keep every chunk 16-byte aligned
    P->ychunk = (P->ychunk + 1) / 2 * 2;
    for (size_t k = 0; k < S; k++) {
        const int q = ctx->comm ? comm_index_of_world_rank(ctx->comm, COMM_COLGRP, g->lay.leader_ranks[g->cols[k].segment]) : (int) k;
        P->xoff[k] = (uint32_t) q * P->xchunk;
    }
    for (size_t k = 0; k < R; k++) {
        const int q = ctx->comm ? comm_index_of_world_rank(ctx->comm, COMM_ROWGRP, g->lay.leader_ranks[g->rows[k].segment]) : (int) k;
        P->yoff[k] = (uint32_t) q * P->ychunk;
    }
    GT_REQUIRE((uint64_t) S * P->xchunk + 2 < (1ull << 31), "pull layout: x space exceeds 31-bit codes");
    P->xlen = (uint32_t) S * P->xchunk;
    P->ylen = (uint32_t) R * P->ychunk;
    const uint32_t pad_code = P->xlen;                         // x[xlen] is a permanent 0.0

    // 1. compressed column id -> code, compressed row id -> y index
    std::vector<DevBuf<uint32_t>> col_code(S), row_rank(R);
    for (size_t k = 0; k < S; k++) {
        const uint32_t n = g->cols[k].nnz;
        col_code[k].alloc(n);
        if (n) {
            k_col_codes<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(g->cols[k].ids.p, n, g->hot[g->hot_of_col_slot[k]].pos.p,
                                                                       P->l1hot / (uint32_t) S, P->xoff[k], col_code[k].p);
            ctx->kernel_launches++;
        }
    }
    P->rows.resize(R);
    for (size_t k = 0; k < R; k++) {
        const uint32_t n = g->rows[k].nnz;
        row_rank[k].alloc(n);
        P->rows[k].ny = P->yn[k];
        if (n) {
            k_compose<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(g->rows[k].ids.p, n, g->hot[g->hot_of_row_slot[k]].pos.p, 0, row_rank[k].p);
            ctx->kernel_launches++;
        }
    }
    GT_CUDA(cudaGetLastError());

    // 2. per row slot: expand -> sort by (row', code) -> virtual rows -> SELL-32
    const int code_bits = 32;                                  // bit 31 carries the hot flag
    for (size_t k = 0; k < R; k++) {
        PullRows& Q = P->rows[k];
        const uint32_t nr = Q.ny;          // rows are addressed by y index (position in the segment's hot order)
        uint64_t total = 0;
        for (const Tile& T : g->tiles) if (T.row_slot == k) total += T.nnz;
        Q.nnz = total;
        if (!nr || !total) continue;
        DevBuf<uint64_t> keys, alt; keys.alloc(total); alt.alloc(total);
        uint64_t off = 0;
        for (const Tile& T : g->tiles) {
            if (T.row_slot != k || !T.nnz) continue;
            k_expand<<<grid_for(T.nnz, 256, ctx->sm_count, 16), 256, 0, st>>>(T.JA.p, g->IA_pool.p + T.offset, T.nnz, g->cols[T.col_slot].nnz,
                                                                          col_code[T.col_slot].p, row_rank[k].p, keys.p + off);
            ctx->kernel_launches++;
            off += T.nnz;
        }
        GT_CUDA(cudaGetLastError());
        int row_bits = 1;
        while (row_bits < 32 && (1ull << row_bits) < nr) row_bits++;
        uint64_t* sorted = nullptr;
        // rows occupy bits [32, 32+row_bits); codes bits [0, code_bits): sort the low field, then the high one
        {
            cub::DoubleBuffer<uint64_t> db(keys.p, alt.p);
            size_t tb = 0, tb2 = 0;
            GT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, (int64_t) total, 0, code_bits, st));
            GT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb2, db, (int64_t) total, 32, 32 + row_bits, st));
            DevBuf<uint8_t> tmp; tmp.alloc(std::max(tb, tb2));
            GT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, (int64_t) total, 0, code_bits, st));
            GT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb2, db, (int64_t) total, 32, 32 + row_bits, st));
            GT_CUDA(cudaStreamSynchronize(st));
            ctx->kernel_launches += 8;
            sorted = db.Current();
        }
        uint64_t* other = (sorted == keys.p) ? alt.p : keys.p;
        // _TCSC_CF_: [regular x regular | regular rows x sink columns | source rows], each still sorted by (row', code)
        uint64_t n_rr = total, n_rs = 0, n_sx = 0;
        if (P->cf) {
            CfPart pred{};
            pred.yreg = P->yreg[k]; pred.xchunk = P->xchunk;
            for (uint32_t q = 0; q < kPullMaxSegs; q++) pred.xsnk0[q] = 0xffffffffu;
            for (size_t c = 0; c < S; c++) pred.xsnk0[P->xoff[c] / P->xchunk] = P->xsnk0[c];
            DevBuf<unsigned long long> d_n; d_n.alloc(1);
            size_t tb = 0;
            pred.want = 0;
            GT_CUDA(cub::DeviceSelect::If(nullptr, tb, sorted, other, d_n.p, (int64_t) total, pred, st));
            DevBuf<uint8_t> tmp; tmp.alloc(tb);
            uint64_t done = 0, cnt[3] = {0, 0, 0};
            const int wants[3] = {0, 2, 3};
            for (int w = 0; w < 3; w++) {
                pred.want = wants[w];
                unsigned long long h_n = 0;
                GT_CUDA(cub::DeviceSelect::If(tmp.p, tb, sorted, other + done, d_n.p, (int64_t) total, pred, st));
                GT_CUDA(cudaMemcpyAsync(&h_n, d_n.p, 8, cudaMemcpyDeviceToHost, st));
                GT_CUDA(cudaStreamSynchronize(st));
                cnt[w] = h_n; done += h_n;
            }
            GT_REQUIRE(done == total, "pull layout: computation-filtering split lost entries");
            n_rr = cnt[0]; n_rs = cnt[1]; n_sx = cnt[2];
            ctx->kernel_launches += 3;
            std::swap(sorted, other);
        }
        Q.nnz_rr = n_rr;
        const uint64_t* cf_tail = sorted + n_rr;          // [RS | SX] stay here; the RR range may move to `other` below
        // own x chunk apart from the rest (multi-GPU), or the hottest `band` columns apart from the tail (GT_PULL_BAND)
        uint64_t* rr = sorted;
        uint64_t n_own = 0;
        const bool multi = ctx->comm && comm_size_in(ctx->comm, COMM_COLGRP) > 1;
        if ((multi || P->band) && n_rr) {
            const uint32_t lo = multi ? P->xoff[g->lay.info.accu_segment_col] : 0, hi = multi ? lo + P->xchunk : P->band;
            DevBuf<unsigned long long> d_n; d_n.alloc(2);
            unsigned long long h_n = 0;
            size_t tb = 0;
            GT_CUDA(cub::DeviceSelect::If(nullptr, tb, sorted, other, d_n.p, (int64_t) n_rr, CodeInRange{lo, hi, true}, st));
            DevBuf<uint8_t> tmp; tmp.alloc(tb);
            GT_CUDA(cub::DeviceSelect::If(tmp.p, tb, sorted, other, d_n.p, (int64_t) n_rr, CodeInRange{lo, hi, true}, st));
            GT_CUDA(cudaMemcpyAsync(&h_n, d_n.p, 8, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            n_own = h_n;
            GT_CUDA(cub::DeviceSelect::If(tmp.p, tb, sorted, other + n_own, d_n.p + 1, (int64_t) n_rr, CodeInRange{lo, hi, false}, st));
            GT_CUDA(cudaStreamSynchronize(st));
            ctx->kernel_launches += 2;
            rr = other;
        }
        if (n_own) build_sell(ctx, rr, n_own, nr, vrow_for(n_own), pad_code, Q.own);
        build_sell(ctx, rr + n_own, n_rr - n_own, nr, vrow_for(n_rr - n_own), pad_code, Q.rest);
        if (n_rs) build_sell(ctx, cf_tail, n_rs, nr, vrow_for(n_rs), pad_code, Q.snk);
        if (n_sx) build_sell(ctx, cf_tail + n_rs, n_sx, nr, vrow_for(n_sx), pad_code, Q.src);
        if (verbose)
            fprintf(stderr, "[gt pull] rank %d row slot %zu: rows %u entries %llu | part0 entries %llu (%.1f %%) vrows %u slices %u sell_len %llu (pad %.1f %%) | "
                            "part1 entries %llu vrows %u slices %u sell_len %llu (pad %.1f %%)\n", ctx->rank, k, nr, (unsigned long long) total,
                    (unsigned long long) Q.own.nnz, 100.0 * Q.own.nnz / total, Q.own.nv, Q.own.nslices, (unsigned long long) Q.own.sell_len,
                    Q.own.sell_len ? 100.0 * (Q.own.sell_len - Q.own.nnz) / Q.own.sell_len : 0.0, (unsigned long long) Q.rest.nnz, Q.rest.nv, Q.rest.nslices,
                    (unsigned long long) Q.rest.sell_len, Q.rest.sell_len ? 100.0 * (Q.rest.sell_len - Q.rest.nnz) / Q.rest.sell_len : 0.0);
        if (verbose && P->cf)
            fprintf(stderr, "[gt pull] rank %d row slot %zu: computation filtering: regular rows %u source rows %u | REG x REG %llu entries, REG x SNK %llu, SRC rows %llu\n",
                    ctx->rank, k, P->yreg[k], P->ysrc[k], (unsigned long long) Q.nnz_rr, (unsigned long long) Q.snk.nnz, (unsigned long long) Q.src.nnz);
    }
    return P.release();
}

void pull_free(PullLayout* P) { delete P; }

// y[row slot] = sum over the slot's rows; x = the concatenated, hot-ordered x buffer (x[xlen] == 0.0)
// x = the concatenated, hot-ordered x buffer (x[xlen] == 0.0); y zero-filled by the caller
void pull_spmv(gt_ctx* ctx, const PullLayout* P, uint32_t row_slot, int part, const double* x, double* y) {
    const PullRows& R = P->rows[row_slot];
    const PullSell& Q = part == 0 ? R.own : part == 1 ? R.rest : part == 2 ? R.snk : R.src;
    // y is zero-filled before the pass.  Only the first part launched may use the plain store: every SELL array carries a
    // (possibly empty) virtual row for EVERY row of the segment, so a later part that stored instead of adding would wipe
    // the rows that merely share its last slice — the hottest ones, which sort first among the empty rows.
    const int accum = (part == 1 && R.own.nslices > 0) || part >= 2;
    if (!Q.nslices) return;
    if (part == 0 && P->band_smem) {                  // hot band from shared memory: one CTA per SM, the whole carve-out
        const size_t smem = (size_t) P->band * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            GT_CUDA(cudaFuncSetAttribute(k_spmv_pull_sell_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = true;
        }
        k_spmv_pull_sell_smem<8><<<ctx->sm_count, kPullThreads, smem, ctx->stream>>>(Q.sell.p, Q.slice_ptr.p, Q.nslices, Q.vtgt.p, Q.nv, x, P->band, y);
        ctx->kernel_launches++;
        GT_CUDA(cudaGetLastError());
        return;
    }
    const int grid = ctx->sm_count * P->ctas_per_sm;
#define GT_PULL_LAUNCH(U, A, B, C) k_spmv_pull_sell<U, A, B, C><<<grid, P->threads, 0, ctx->stream>>>(Q.sell.p, Q.slice_ptr.p, Q.nslices, Q.vtgt.p, Q.nv, x, y)
#define GT_PULL_AB(U, A, B) do { if (accum) GT_PULL_LAUNCH(U, A, B, 1); else GT_PULL_LAUNCH(U, A, B, 0); } while (0)
    const bool split = P->l1hot > 0;
    if (P->unroll == 4) {
        if (split) { if (P->l2hint) GT_PULL_AB(4, true, true); else GT_PULL_AB(4, true, false); }
        else { if (P->l2hint) GT_PULL_AB(4, false, true); else GT_PULL_AB(4, false, false); }
    } else {
        if (split) { if (P->l2hint) GT_PULL_AB(8, true, true); else GT_PULL_AB(8, true, false); }
        else { if (P->l2hint) GT_PULL_AB(8, false, true); else GT_PULL_AB(8, false, false); }
    }
#undef GT_PULL_AB
#undef GT_PULL_LAUNCH
    ctx->kernel_launches++;
    GT_CUDA(cudaGetLastError());
}

}  // namespace gt
