// gt_engine.cu — Vertex_Program::execute on the device.
//
// The reference loop (src/vp/vertex_program.hpp:407-441):
//     while (true) { scatter_gather(); combine(); apply(); iteration++; converged? }
// and what each phase becomes here:
//   scatter_gather   messenger kernel over the owned column segment's non-empty columns (:687-758), then ONE in-place
//                    ncclAllGather along the column group: x is laid out as one chunk per group member, chunk q =
//                    the segment led by group rank q (the reference: one Ibcast per segment, :843-862,970-1013)
//   combine          per local tile, in local_tiles_row_order: push SpMV / frontier SpMSpV (:1057-1113,
//                    :1330-1434) — PageRank: the pull SpMV of gt_pull.cu — then ONE in-place ncclReduceScatter
//                    along the row group (the follower->leader Isend + leader-side combine of :1083-1108,1522-1573)
//   apply            applicator kernel on the owned segment (:1640-1802), activity flags C
//   has_converged    device count of C, ncclAllReduce over the world, one 8-byte D2H (:1884-1923); for the
//                    non-stationary programs it shares one stream synchronisation with the frontier sizes
// The five shipped programs are recognised by enum; their messenger/combiner/applicator bodies
// (src/apps/{deg,pr,bfs,cc,sssp}.h) are the __device__ functions below.  With GT_COL every row/column
// notion swaps, including the two communicators (:279-325).
//
// Non-stationary exchange: the reference ships (index,value) pairs when at most 60 % of a segment is
// active (:760-784,970-1013).  On NVLink the dense u32 segment is cheap, so x always travels dense and
// every rank rebuilds the frontier list locally; the 0.6 rule still picks SpMSpV vs dense SpMV per
// column segment (:1475), so `sparse_iterations` matches the reference's schedule.
#include "gt_program.h"
#include <cub/cub.cuh>
#include <memory>
#include <algorithm>
#include <cmath>
#include <chrono>

namespace gt {

__global__ void k_init_state(VState V, int app, uint32_t th, uint32_t vid0, uint32_t root, double alpha, int stationary) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        const uint32_t vid = vid0 + i;
        switch (app) {
            case GT_APP_DEG: V.a[i] = 0; V.C[i] = 1; break;                                  // deg.h:32-35
            case GT_APP_PR: V.a[i] = 0; V.rank[i] = alpha; V.C[i] = stationary ? 1 : 0; break;   // pr.h:15-19, base initializer :32
            case GT_APP_BFS:                                                                  // bfs.h:37-49
                if (vid == root) { V.a[i] = vid; V.b[i] = 0; V.C[i] = 1; }
                else { V.a[i] = 0; V.b[i] = GT_INF_U32; V.C[i] = 0; }
                break;
            case GT_APP_CC: V.a[i] = vid; V.C[i] = 1; break;                                   // cc.h:32-35
            case GT_APP_SSSP:                                                                 // sssp.h:34-43
                if (vid == root) { V.a[i] = 0; V.C[i] = 1; } else { V.a[i] = GT_INF_U32; V.C[i] = 0; }
                break;
        }
    }
}

// ---- messenger ---------------------------------------------------------------------------------------
// stationary: x[j] = messenger(V[JC[j]]) over the owned segment's non-empty columns (:699-705)
__global__ void k_messenger_f64(VState V, int app, const uint32_t* __restrict__ JC, uint32_t nc, double* __restrict__ x) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nc; j += gridDim.x * blockDim.x) {
        if (app == GT_APP_DEG) { x[j] = 1.0; continue; }                                   // deg.h:37-39
        const uint32_t v = JC[j];
        const uint32_t d = V.a[v];
        x[j] = d ? V.rank[v] / (double) d : 0.0;                                           // pr.h:31-33
    }
}
// ---- applicator ------------------------------------------------------------------------------------------
// stationary, TCSC: rows with I[i] take y[j++] (here y[r] with v = IR[r]) (:1655-1670)
// `cls` (PageRank on a _TCSC_CF_ graph): regular rows every iteration, source rows only on the last (:1671-1692)
__global__ void k_apply_f64(VState V, int app, const uint32_t* __restrict__ IR, uint32_t nr, const double* __restrict__ y,
                            double alpha, double tol, const uint8_t* __restrict__ cls, int last) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nr; r += gridDim.x * blockDim.x) {
        const uint32_t v = IR[r];
        if (cls) { const uint8_t c = cls[v]; if (!(c == 1 || (c == 2 && last))) continue; }
        const double yy = y[r];
        if (app == GT_APP_DEG) { V.a[v] = (uint32_t) yy; V.C[v] = 0; continue; }           // deg.h:49-52
        const double old = V.rank[v];
        const double nw = __dadd_rn(alpha, __dmul_rn(1.0 - alpha, yy));                      // pr.h:45 (no fma contraction)
        V.rank[v] = nw;
        V.C[v] = fabs(nw - old) > tol;                                                       // pr.h:46
    }
}
// vertices whose row is empty everywhere: applicator(state) -> false (:1666-1667, :38)
__global__ void k_clear_C_empty(uint8_t* C, const uint8_t* __restrict__ I, uint32_t th) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x)
        if (!I[i]) C[i] = 0;
}
// has_converged: vertices still moving; on a _TCSC_CF_ graph only the regular rows count (:1902-1916)
__global__ void __launch_bounds__(256) k_count_u8(const uint8_t* __restrict__ C, uint32_t n, const uint8_t* __restrict__ cls, unsigned long long* __restrict__ out) {
    unsigned local = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) local += (C[i] && (!cls || cls[i] == 1)) ? 1 : 0;
    typedef cub::BlockReduce<unsigned, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const unsigned tot = BR(tmp).Sum(local);
    if (threadIdx.x == 0 && tot) atomicAdd(out, (unsigned long long) tot);
}

// ---- PageRank in the pull layout: the owned segment's state lives in hot order while execute() runs, so
// applicator (pr.h:43-47) and next iteration's messenger (pr.h:31-33) are one sequential pass ------------
__global__ void k_pr_to_hot(const uint32_t* __restrict__ ids, uint32_t n, const double* __restrict__ rank, const uint32_t* __restrict__ deg,
                            const uint8_t* __restrict__ I, double* __restrict__ rank_h, uint32_t* __restrict__ deg_h, uint8_t* __restrict__ flag_h) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t v = ids[k];
        rank_h[k] = rank[v]; deg_h[k] = deg[v]; flag_h[k] = I[v];
    }
}
__global__ void k_pr_from_hot(const uint32_t* __restrict__ ids, uint32_t n, const double* __restrict__ rank_h, const uint8_t* __restrict__ C_h,
                              double* __restrict__ rank, uint8_t* __restrict__ C) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t v = ids[k];
        rank[v] = rank_h[k]; C[v] = C_h[k];
    }
}
__global__ void k_pr_messenger_h(const double* __restrict__ rank_h, const uint32_t* __restrict__ deg_h, uint32_t n, double* __restrict__ x) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t d = deg_h[k];
        x[k] = d ? rank_h[k] / (double) d : 0.0;
    }
}
// partial y vectors of the owned row segment that the other members of the row group have put into this rank's
// window (gt_peer.cu); the leader's combine (:1522-1541) is folded into the applicator's read
struct YParts { const double* p[8]; int n; };
__global__ void __launch_bounds__(256) k_pr_apply_h(const double* __restrict__ y, YParts yr, double* __restrict__ rank_h, const uint32_t* __restrict__ deg_h,
                                                     const uint8_t* __restrict__ flag_h, uint8_t* __restrict__ C_h, double* __restrict__ x, uint32_t n,
                                                     double alpha, double tol, unsigned long long* __restrict__ active) {
    unsigned local = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double r = rank_h[k];
        uint8_t c = 0;
        if (flag_h[k]) {                                   // row non-empty: applicator(state, y)
            double yk = y[k];
            for (int j = 0; j < yr.n; j++) yk += yr.p[j][k];
            const double nw = __dadd_rn(alpha, __dmul_rn(1.0 - alpha, yk));
            c = fabs(nw - r) > tol;
            rank_h[k] = nw;
            r = nw;
        }                                                  // else applicator(state) -> false (:1666-1667)
        C_h[k] = c;
        local += c;
        const uint32_t d = deg_h[k];
        x[k] = d ? r / (double) d : 0.0;                   // next iteration's messenger
    }
    if (active) {
        typedef cub::BlockReduce<unsigned, 256> BR;
        __shared__ typename BR::TempStorage tmp;
        const unsigned tot = BR(tmp).Sum(local);
        if (threadIdx.x == 0 && tot) atomicAdd(active, (unsigned long long) tot);
    }
}

// _TCSC_CF_ in convergence mode: after has_converged() the reference's combine() does nothing for this compression
// (:1036-1043) and apply() then hands every SOURCE row the y left by the last iteration — zero, because only the
// REG x REG list ran (:1282) — so the row ends at alpha + (1 - alpha) * 0 (:1683-1690).  Reproduced as is.
__global__ void k_pr_cf_sources_h(double* __restrict__ rank_h, uint8_t* __restrict__ C_h, uint32_t lo, uint32_t hi, double alpha) {
    for (uint32_t k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
        rank_h[k] = __dadd_rn(alpha, __dmul_rn(1.0 - alpha, 0.0));
        C_h[k] = 0;
    }
}
__global__ void k_pr_cf_sources(VState V, const uint8_t* __restrict__ cls, uint32_t th, double alpha) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x)
        if (cls[i] == 2) { V.rank[i] = __dadd_rn(alpha, __dmul_rn(1.0 - alpha, 0.0)); V.C[i] = 0; }
}

// initialize(other): degree hand-over where the row is non-empty (:476-483, pr.h:24-28)
__global__ void k_init_from_deg(VState V, const uint32_t* __restrict__ other_deg, const uint8_t* __restrict__ I, uint32_t th, double alpha) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x)
        if (I[i]) { V.a[i] = other_deg[i]; V.rank[i] = alpha; V.C[i] = 1; }
}

// AoS <-> SoA at the boundary
__global__ void k_pack_state(VState V, int app, uint32_t th, uint32_t vid0, uint32_t* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        switch (app) {
            case GT_APP_PR: {
                out[4 * i] = V.a[i]; out[4 * i + 1] = 0;
                const unsigned long long bits = (unsigned long long) __double_as_longlong(V.rank[i]);
                out[4 * i + 2] = (uint32_t) bits; out[4 * i + 3] = (uint32_t) (bits >> 32);
                break;
            }
            case GT_APP_BFS: out[3 * i] = V.a[i]; out[3 * i + 1] = V.b[i]; out[3 * i + 2] = vid0 + i; break;
            default: out[i] = V.a[i]; break;
        }
    }
}
__global__ void k_unpack_state(VState V, int app, uint32_t th, const uint32_t* in) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < th; i += gridDim.x * blockDim.x) {
        switch (app) {
            case GT_APP_PR: {
                V.a[i] = in[4 * i];
                const unsigned long long bits = (unsigned long long) in[4 * i + 2] | ((unsigned long long) in[4 * i + 3] << 32);
                V.rank[i] = __longlong_as_double((long long) bits);
                break;
            }
            case GT_APP_BFS: V.a[i] = in[3 * i]; V.b[i] = in[3 * i + 1]; break;
            default: V.a[i] = in[i]; break;
        }
    }
}

// ---- tile kernel dispatch ---------------------------------------------------------------------------------
static void launch_spmv(gt_ctx* ctx, const gt_graph* g, const Tile& T, int semiring, int ordering, bool skip_inf,
                        const void* x, void* y, uint8_t* t) {
    if (!T.nnz) return;
    cudaStream_t st = ctx->stream;
    const uint32_t* IA = g->IA_pool.p + T.offset;
    const uint32_t* A = g->weighted ? g->A_pool.p + T.offset : nullptr;
    const uint32_t ncols = g->cols[T.col_slot].nnz;
    if (ordering == GT_ROW) {
        const uint32_t nchunks = (uint32_t) ((T.nnz + GT_PUSH_CHUNK - 1) / GT_PUSH_CHUNK);
        const int grid = (int) std::min<uint64_t>(nchunks, (uint64_t) ctx->sm_count * 8);
#define GT_PUSH(S, W, K) k_spmv_push<S, W, K><<<grid, kPushThreads, 0, st>>>(T.JA.p, IA, A, T.chunk_col.p, nchunks, T.nnz, \
            (const Semiring<S>::T*) x, (Semiring<S>::T*) y, t)
        if (semiring == GT_PLUS_TIMES_F64) { if (A) GT_PUSH(GT_PLUS_TIMES_F64, true, false); else GT_PUSH(GT_PLUS_TIMES_F64, false, false); }
        else if (semiring == GT_MIN_PLUS_U32) {
            GT_REQUIRE(A, "min-plus needs a weighted graph");
            if (skip_inf) GT_PUSH(GT_MIN_PLUS_U32, true, true); else GT_PUSH(GT_MIN_PLUS_U32, true, false);
        } else {
            if (A) { if (skip_inf) GT_PUSH(GT_MIN_SELECT_U32, true, true); else GT_PUSH(GT_MIN_SELECT_U32, true, false); }
            else { if (skip_inf) GT_PUSH(GT_MIN_SELECT_U32, false, true); else GT_PUSH(GT_MIN_SELECT_U32, false, false); }
        }
#undef GT_PUSH
    } else {
        const int grid = grid_for((uint64_t) ncols * 32, 256, ctx->sm_count, 8);
#define GT_PULL(S, W) k_spmv_pull<S, W><<<grid, 256, 0, st>>>(T.JA.p, IA, A, ncols, (const Semiring<S>::T*) x, (Semiring<S>::T*) y)
        if (semiring == GT_PLUS_TIMES_F64) { if (A) GT_PULL(GT_PLUS_TIMES_F64, true); else GT_PULL(GT_PLUS_TIMES_F64, false); }
        else if (semiring == GT_MIN_PLUS_U32) { GT_REQUIRE(A, "min-plus needs a weighted graph"); GT_PULL(GT_MIN_PLUS_U32, true); }
        else { if (A) GT_PULL(GT_MIN_SELECT_U32, true); else GT_PULL(GT_MIN_SELECT_U32, false); }
#undef GT_PULL
    }
    ctx->kernel_launches++;
    GT_CUDA(cudaGetLastError());
}

static void launch_spmspv(gt_ctx* ctx, const gt_graph* g, const Tile& T, int semiring, const uint32_t* xi, const void* xv, uint32_t k,
                          void* y, uint8_t* t) {
    if (!T.nnz || !k) return;
    cudaStream_t st = ctx->stream;
    const uint32_t* IA = g->IA_pool.p + T.offset;
    const uint32_t* A = g->weighted ? g->A_pool.p + T.offset : nullptr;
    const int grid = grid_for((uint64_t) k * 32, 256, ctx->sm_count, 8);
    // heavy-column path only where a column of this tile can exceed the threshold (one extra launch otherwise)
    const bool heavy = T.max_col_entries > kHeavyColumn;
    uint2* hl = heavy ? g->heavy_list.p : nullptr;
    unsigned int* hc = heavy ? g->heavy_count.p : nullptr;
    if (heavy) GT_CUDA(cudaMemsetAsync(hc, 0, sizeof(unsigned int), st));
    const int hgrid = ctx->sm_count * 4;
#define GT_SP(S, W) do { \
        k_spmspv_push<S, W><<<grid, 256, 0, st>>>(T.JA.p, IA, A, xi, (const Semiring<S>::T*) xv, k, (Semiring<S>::T*) y, t, hl, hc); \
        if (heavy) { k_spmspv_heavy<S, W><<<hgrid, 256, 0, st>>>(T.JA.p, IA, A, xi, (const Semiring<S>::T*) xv, (Semiring<S>::T*) y, t, hl, hc); ctx->kernel_launches++; } \
    } while (0)
    if (semiring == GT_PLUS_TIMES_F64) { if (A) GT_SP(GT_PLUS_TIMES_F64, true); else GT_SP(GT_PLUS_TIMES_F64, false); }
    else if (semiring == GT_MIN_PLUS_U32) { GT_REQUIRE(A, "min-plus needs a weighted graph"); GT_SP(GT_MIN_PLUS_U32, true); }
    else { if (A) GT_SP(GT_MIN_SELECT_U32, true); else GT_SP(GT_MIN_SELECT_U32, false); }
#undef GT_SP
    ctx->kernel_launches++;
    GT_CUDA(cudaGetLastError());
}

}  // namespace gt

namespace gt {

static void prog_alloc(gt_program* P) {
    gt_graph* g = P->g;
    gt_ctx* ctx = P->ctx;
    P->th = g->lay.info.tile_height;
    P->vid0 = (uint32_t) g->lay.info.owned_segment * P->th;
    if (P->ordering == GT_ROW) {
        P->prow = &g->rows; P->pcol = &g->cols;
        P->own_row_slot = g->lay.info.accu_segment_row; P->own_col_slot = g->lay.info.accu_segment_col;
        P->bcast_group = COMM_COLGRP; P->reduce_group = COMM_ROWGRP;
    } else {
        P->prow = &g->cols; P->pcol = &g->rows;
        P->own_row_slot = g->lay.info.accu_segment_col; P->own_col_slot = g->lay.info.accu_segment_row;
        P->bcast_group = COMM_ROWGRP; P->reduce_group = COMM_COLGRP;
    }
    if (P->app == GT_APP_PR) P->rank.alloc(P->th);
    P->a.alloc(P->th);
    if (P->app == GT_APP_BFS) P->b.alloc(P->th);
    P->C.alloc(P->th);
    P->d_active.alloc(4);
    GT_CUDA(cudaMemsetAsync(P->d_active.p, 0, P->d_active.bytes(), ctx->stream));
    GT_CUDA(cudaMallocHost((void**) &P->h_active, 4 * sizeof(unsigned long long)));
    memset(P->h_active, 0, 4 * sizeof(unsigned long long));
    GT_CUDA(cudaEventCreate(&P->ev0));
    GT_CUDA(cudaEventCreate(&P->ev1));
    if (!P->stationary) { ns_alloc(P); return; }          // BFS / CC / SSSP: gt_ns.cu
    P->X.resize(P->pcol->size());
    P->Y.resize(P->prow->size());
    // x: one equal-sized chunk per member of the broadcast group, chunk q = the segment led by group rank q, so the
    // whole exchange is ONE in-place ncclAllGather (the reference: one Ibcast per segment, :843-862,970-1013); y: the
    // same along the reduce group with ONE in-place ncclReduceScatter (the reference: Isend to the leader + host-side
    // combine, :1083-1108,1522-1573).  Every member leads exactly one of its group's segments (tests/test_layout.py).
    auto chunk_of = [&](CommGroup grp, int segment, size_t k) {
        return ctx->comm ? (size_t) comm_index_of_world_rank(ctx->comm, grp, g->lay.leader_ranks[segment]) : k;
    };
    for (const SegMaps& s : *P->pcol) P->xchunk = std::max<size_t>(P->xchunk, s.nnz);
    for (const SegMaps& s : *P->prow) P->ychunk = std::max<size_t>(P->ychunk, s.nnz);
    P->xchunk = (P->xchunk + 3) / 4 * 4;           // chunks stay 16-byte aligned
    P->ychunk = (P->ychunk + 3) / 4 * 4;
    const size_t S = P->X.size();
    P->Xcat.alloc((S * P->xchunk + 1) * P->esize());
    GT_CUDA(cudaMemsetAsync(P->Xcat.p, 0, P->Xcat.n, ctx->stream));
    P->Ycat.alloc(std::max<size_t>(1, P->Y.size() * P->ychunk) * P->esize());
    GT_CUDA(cudaMemsetAsync(P->Ycat.p, 0, P->Ycat.n, ctx->stream));
    for (size_t k = 0; k < S; k++) {
        P->X[k].p = P->Xcat.p + chunk_of(P->bcast_group, (*P->pcol)[k].segment, k) * P->xchunk * P->esize();
        P->X[k].n = (size_t) (*P->pcol)[k].nnz * P->esize();
    }
    for (size_t k = 0; k < P->Y.size(); k++) {
        P->Y[k].p = P->Ycat.p + chunk_of(P->reduce_group, (*P->prow)[k].segment, k) * P->ychunk * P->esize();
        P->Y[k].n = (size_t) (*P->prow)[k].nnz * P->esize();
    }
}

// x / y of the pull path.  Multi-GPU: windows the other group members write into over NVLink (GT_PEER=0 or a failed
// cudaIpc mapping -> plain buffers + NCCL collectives; every rank takes the same branch, peer_window_create agrees).
static void pull_alloc(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const PullLayout* L = P->pull;
    const int S = ctx->comm ? comm_size_in(ctx->comm, P->bcast_group) : 1;
    const int G = ctx->comm ? comm_size_in(ctx->comm, P->reduce_group) : 1;
    const char* e = getenv("GT_PEER");
    const bool want_peer = ctx->comm && !(e && atoi(e) == 0) && S <= 8 && G <= 8;
    P->x_stride = ((size_t) L->xlen + 2) / 2 * 2;                 // x[xlen] is the permanent 0.0 the padding codes point at
    if (want_peer && S > 1) P->wx = peer_window_create(ctx, P->bcast_group, 2 * P->x_stride * sizeof(double));
    if (want_peer && G > 1) P->wy = peer_window_create(ctx, P->reduce_group, 2 * (size_t) G * L->ychunk * sizeof(double));
    if (P->wx) {
        P->xbuf[0] = (double*) P->wx->local;
        P->xbuf[1] = P->xbuf[0] + P->x_stride;
    } else {
        P->Xh.alloc((size_t) L->xlen + 1);
        GT_CUDA(cudaMemsetAsync(P->Xh.p, 0, P->Xh.bytes(), st));
        P->xbuf[0] = P->xbuf[1] = P->Xh.p;
    }
    P->Yh.alloc(L->ylen);
    P->own_hot = &P->g->hot[P->g->hot_of_row_slot[P->own_row_slot]];
    P->rank_h.alloc(P->own_hot->n); P->deg_h.alloc(P->own_hot->n); P->flag_h.alloc(P->own_hot->n); P->C_h.alloc(P->own_hot->n);
    GT_CUDA(cudaEventCreateWithFlags(&P->ev_b, cudaEventDisableTiming));
    for (int i = 0; i < GT_PEER_MAX_LANES; i++) GT_CUDA(cudaEventCreateWithFlags(&P->ev_yput[i], cudaEventDisableTiming));
    P->pull_ready = true;
}

// init_stationary + init_nonstationary (:504-636)
static void prog_initialize(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const auto t_init = std::chrono::steady_clock::now();     // init_time of -DTIMING (:445-470)
    struct InitClock { gt_program* P; cudaStream_t st; std::chrono::steady_clock::time_point t0;
        ~InitClock() { cudaStreamSynchronize(st); P->init_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } } init_clock{P, st, t_init};
    k_init_state<<<grid_for(P->th, 256, ctx->sm_count), 256, 0, st>>>(P->vs(), P->app, P->th, P->vid0, P->prm.root, P->prm.alpha, P->stationary);
    ctx->kernel_launches++;
    if (!P->stationary) ns_initialize(P);       // Y starts at infinity() (:625-635)
    GT_CUDA(cudaGetLastError());
    // PageRank's plus-times SpMV runs as a pull over the derived layout (gt_pull.cu) unless pr_layout = 0
    P->pull = nullptr;
    if (P->app == GT_APP_PR && P->ordering == GT_ROW && !P->g->weighted && P->pr_layout == 1) {
        if (!P->g->pull) P->g->pull = pull_build(P->g);
        P->pull = P->g->pull;
        if (!P->pull_ready) pull_alloc(P);
    }
    P->cf = P->app == GT_APP_PR && P->g->compression == GT_TCSC_CF && P->ordering == GT_ROW;
    P->hot_valid = false;
    P->x_ready = false;
    P->initialized = true;
    P->iteration = 0;
    P->converged = false;
    P->empty_cleared = false;
}

// the running iteration's place in the computation-filtering schedule (:1246,1282): iteration 0 adds REG x SNK, the last
// iteration of a fixed-count run adds the source rows; phases driven through run_phase() are a middle iteration
static inline bool cf_first(const gt_program* P) { return P->cf && P->in_execute && P->iteration == 0; }
static inline bool cf_last(const gt_program* P) { return P->cf && P->in_execute && !P->check_mode && P->iteration + 1 == P->num_iterations; }

// SURVEY.md §8(d), one iteration: per tile IA (+A) + JA + one read of its x segment; per row group one write of y;
// vertex phase on the owned segment: JC + state read, IR + y + state read/write.  On a _TCSC_CF_ graph the pull path
// moves only what the schedule touches: the REG x REG entries (+ REG x SNK first, + source rows last), the regular part
// of x and of y.  `combine_only`: without the vertex phase.
static uint64_t algorithmic_bytes_iteration(const gt_program* P, bool first, bool last, bool combine_only) {
    const gt_graph* g = P->g;
    const uint64_t es = P->esize();
    uint64_t b = 0;
    if (P->pull && P->pull->cf) {
        const PullLayout* L = P->pull;
        for (size_t k = 0; k < L->rows.size(); k++) {
            const PullRows& Q = L->rows[k];
            b += 4 * (Q.nnz_rr + (first ? Q.snk.nnz : 0) + (last ? Q.src.nnz : 0));
            b += 8ull * (L->yreg[k] + (last ? L->ysrc[k] : 0));
        }
        for (size_t c = 0; c < L->xn.size(); c++)
            b += (4 + 8) * (uint64_t) (L->xreg[c] + ((first || last) ? L->xn[c] - L->xsnk0[c] : 0)) * L->rows.size();   // column pointer + x, once per tile of the column
        if (!combine_only) {
            const HotOrder& H = *P->own_hot;
            b += (4 + 12) * (uint64_t) H.nreg + (4 + 8 + 16) * (uint64_t) (H.nreg + (last ? H.nsrc : 0));
        }
        return b;
    }
    for (const Tile& T : g->tiles) {
        const uint64_t nc = (P->ordering == GT_ROW ? g->cols[T.col_slot].nnz : g->rows[T.row_slot].nnz);
        if (!T.nnz) continue;
        b += (g->weighted ? 8 : 4) * T.nnz + 4 * ((uint64_t) g->cols[T.col_slot].nnz + 1) + es * nc;
    }
    for (const SegMaps& r : *P->prow) b += es * r.nnz;
    if (combine_only) return b;
    const uint64_t state = (P->app == GT_APP_PR) ? 12 : 4;
    b += (4 + state) * (uint64_t) (*P->pcol)[P->own_col_slot].nnz;
    b += (4 + es + (P->app == GT_APP_PR ? 16 : 8)) * (uint64_t) (*P->prow)[P->own_row_slot].nnz;
    return b;
}

// natural-order state -> hot-order working state of the owned segment (once per execute)
static void pull_state_in(gt_program* P) {
    if (P->hot_valid) return;
    gt_ctx* ctx = P->ctx;
    const HotOrder& H = *P->own_hot;
    if (H.n) {
        const SegMaps& own = (*P->prow)[P->own_row_slot];
        k_pr_to_hot<<<grid_for(H.n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(H.ids.p, H.n, P->rank.p, P->a.p, own.bits.p, P->rank_h.p, P->deg_h.p, P->flag_h.p);
        ctx->kernel_launches++;
        GT_CUDA(cudaMemsetAsync(P->C_h.p, 0, H.n, ctx->stream));
    }
    P->hot_valid = true;
    P->x_ready = false;
}
// hot-order state -> natural-order V / C (at the end of execute)
static void pull_state_out(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    const HotOrder& H = *P->own_hot;
    if (!P->empty_cleared) {                          // vertices outside the hot order never see an applicator with y
        const SegMaps& own = (*P->prow)[P->own_row_slot];
        k_clear_C_empty<<<grid_for(P->th, 256, ctx->sm_count), 256, 0, ctx->stream>>>(P->C.p, own.bits.p, P->th);
        ctx->kernel_launches++;
        P->empty_cleared = true;
    }
    if (H.n) {
        k_pr_from_hot<<<grid_for(H.n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(H.ids.p, H.n, P->rank_h.p, P->C_h.p, P->rank.p, P->C.p);
        ctx->kernel_launches++;
    }
    GT_CUDA(cudaGetLastError());
}

// ---- PageRank iteration on the pull layout, with the exchange over NVLink peer windows ------------------------------
// Why two buffers per window are enough (no barrier anywhere in the loop):
//  * x: a column-group peer Q puts x(k+2) into the parity that held x(k) only after its applicator of iteration k+1,
//    which needs this rank's x(k+1) for Q's SpMV — and this rank puts x(k+1) only after its own applicator of
//    iteration k, i.e. after its last read of x(k).  Likewise this rank's applicator overwrites its local copy of the
//    parity of x(k-1) only after every peer's x(k) has arrived, which they sent after consuming x(k-1).
//  * y: a row-group peer F puts y(k+2) into the parity of y(k) only after its applicator of iteration k+1, which
//    needs this rank's partial y(k+1), sent after this rank's applicator of iteration k (stream order) — the reader
//    of y(k).  One buffer would not do: F's y(k+1) depends on x(k+1) from F's column group, not on this rank.
//  * every put is consumed (its counter waited for) before execute() / run_phase() returns, so a window is never
//    written after its owner could have freed it.
// x buffer the current iteration reads (parity of the last put) / the one the next messenger writes
void tl_mark(gt_program* P, const char* tag, cudaStream_t s) {
    if (!P->timeline_on) return;
    cudaEvent_t ev;
    if (!P->timeline_pool.empty()) { ev = P->timeline_pool.back(); P->timeline_pool.pop_back(); }
    else GT_CUDA(cudaEventCreate(&ev));
    GT_CUDA(cudaEventRecord(ev, s));
    P->timeline.push_back({ev, tag, P->iteration});
}
void tl_dump(gt_program* P) {
    if (!P->timeline_on) return;
    const char* prefix = getenv("GT_TIMELINE");
    std::string path = std::string(prefix ? prefix : "gt_timeline") + ".r" + std::to_string(P->ctx->rank) + ".jsonl";
    FILE* f = fopen(path.c_str(), "a");
    if (f) fprintf(f, "{\"rank\": %d, \"marks\": [", P->ctx->rank);
    bool firstm = true;
    for (const gt_program::Mark& m : P->timeline) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, P->ev0, m.ev) == cudaSuccess && f) {
            fprintf(f, "%s[%u, \"%s\", %.4f]", firstm ? "" : ", ", m.iteration, m.tag, ms);
            firstm = false;
        }
        P->timeline_pool.push_back(m.ev);
    }
    cudaGetLastError();
    if (f) { fprintf(f, "]}\n"); fclose(f); }
    P->timeline.clear();
}
static inline double* pull_x_cur(gt_program* P) { return P->xbuf[P->wx ? (P->x_epoch & 1) : 0]; }
static inline double* pull_x_next(gt_program* P) { return P->xbuf[P->wx ? ((P->x_epoch + 1) & 1) : 0]; }

static void pull_scatter_gather(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    pull_state_in(P);
    const PullLayout* L = P->pull;
    const uint32_t n = P->own_hot->n;
    const uint32_t xo_off = L->xoff[P->own_col_slot];
    double* xo = pull_x_next(P) + xo_off;
    // Computation filtering: after the first exchange only the REGULAR part of x changes (the applicator writes nothing
    // else), and the sink columns' x — constant, their vertices have no row and are never applied — must sit in both
    // parity buffers.  So the first exchange of an execute() ships the whole chunk and the sink range twice.
    const bool first = !P->x_ready;
    const uint32_t snk0 = L->xsnk0[P->own_col_slot], nsnk = n - std::min(n, snk0);
    double* xo_other = P->wx ? P->xbuf[P->x_epoch & 1] + xo_off : nullptr;       // the parity the next exchange does NOT use
    if (first && n) {                                 // later iterations: x was written by the fused applicator
        k_pr_messenger_h<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(P->rank_h.p, P->deg_h.p, n, xo);
        ctx->kernel_launches++;
        if (L->cf && xo_other && nsnk) GT_CUDA(cudaMemcpyAsync(xo_other + snk0, xo + snk0, (size_t) nsnk * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    P->x_ready = true;
    P->ag_pending = false;
    const uint32_t nput = (L->cf && !first) ? P->own_hot->nreg : n;
    if (P->wx) {
        // bcast_stationary (:843-862) as puts: the own chunk goes into every other member's window by copy engine,
        // on the side streams, while this rank's SpMV over its own chunk is already running
        P->x_epoch++;
        GT_CUDA(cudaEventRecord(ctx->ev_x, st));
        peer_put_begin(ctx, ctx->ev_x);
        const size_t off = ((size_t) (P->x_epoch & 1) * P->x_stride + xo_off) * sizeof(double);
        const size_t off_other = ((size_t) ((P->x_epoch + 1) & 1) * P->x_stride + xo_off + snk0) * sizeof(double);
        for (int j = 1; j < P->wx->size; j++) {
            const int q = (P->wx->me + j) % P->wx->size;
            if (first && L->cf && nsnk) peer_put(ctx, P->wx, q, off_other, xo + snk0, (size_t) nsnk * sizeof(double), P->x_epoch, false);
            peer_put(ctx, P->wx, q, off, xo, (size_t) nput * sizeof(double), P->x_epoch);
        }
        peer_put_end(ctx, nullptr);
        for (int j = 0; j < std::min(ctx->peer_lanes, P->wx->size - 1); j++) tl_mark(P, "x_put_done", ctx->put_stream[j]);
        P->x_wait_pending = true;
    } else if (ctx->comm && comm_size_in(ctx->comm, P->bcast_group) > 1) {      // every member leads exactly one of the group's segments
        // the all-gather runs on its own stream; the SpMV over the own chunk does not wait for it
        GT_CUDA(cudaEventRecord(ctx->ev_x, st));
        GT_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_x, 0));
        comm_allgather_inplace(ctx->comm, P->bcast_group, P->Xh.p, P->pull->xchunk, CT_F64, ctx->comm_stream);
        GT_CUDA(cudaEventRecord(ctx->ev_ag, ctx->comm_stream));
        P->ag_pending = true;
    }
    GT_CUDA(cudaGetLastError());
}

// the other members' x chunks have landed (no-op when nothing is in flight)
static void pull_x_arrived(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    if (P->x_wait_pending) { peer_wait_all(ctx, P->wx, P->x_epoch, ctx->stream); P->x_wait_pending = false; }
    if (P->ag_pending) { GT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_ag, 0)); P->ag_pending = false; }
}

static void pull_combine(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const PullLayout* L = P->pull;
    const double* x = pull_x_cur(P);
    const size_t R = L->rows.size();
    const size_t own = (size_t) P->own_row_slot;
    const bool first = cf_first(P), last = cf_last(P);
    P->combine_bytes = algorithmic_bytes_iteration(P, first, last, true);
    if (P->ypush_pending) { peer_puts_done(ctx, P->ev_yput, st); P->ypush_pending = false; }   // the last puts have read Yh
    if (P->Yh.n) GT_CUDA(cudaMemsetAsync(P->Yh.p, 0, P->Yh.bytes(), st));     // std::fill(y, 0) (:1026-1032)
    // everything of a row segment that needs the other members' x: the rest of REG x REG every iteration, REG x SNK in
    // iteration 0 (:1246-1262), the source rows in the last one (:1282-1317)
    auto remote_parts = [&](size_t k) {
        double* y = P->Yh.p + L->yoff[k];
        pull_spmv(ctx, L, (uint32_t) k, 1, x, y);
        if (first) pull_spmv(ctx, L, (uint32_t) k, 2, x, y);
        if (last) pull_spmv(ctx, L, (uint32_t) k, 3, x, y);
    };
    // Row segments led by other ranks first: their partial y leaves for the leader while the owned segment is computed.
    // Part 0 needs only this rank's own x chunk, so it runs while the other chunks are still arriving.
    for (size_t k = 0; k < R; k++) if (k != own && L->yn[k]) pull_spmv(ctx, L, (uint32_t) k, 0, x, P->Yh.p + L->yoff[k]);
    if (L->yn[own]) pull_spmv(ctx, L, (uint32_t) own, 0, x, P->Yh.p + L->yoff[own]);
    tl_mark(P, "own_parts_done", st);
    pull_x_arrived(P);
    tl_mark(P, "x_arrived", st);
    for (size_t k = 0; k < R; k++) if (k != own && L->yn[k]) remote_parts(k);
    tl_mark(P, "follower_rows_done", st);
    if (P->wy) {
        // combine_2d_stationary's follower -> leader sends (:1083-1108) as puts into slot `me` of the leader's window;
        // with computation filtering only the rows the leader will apply: the regular ones, plus the source rows at the end
        P->y_epoch++;
        GT_CUDA(cudaEventRecord(P->ev_b, st));
        peer_put_begin(ctx, P->ev_b);
        const int G = P->wy->size;
        for (size_t k = 0; k < R; k++) {
            if (k == own) continue;
            const int q = (int) (L->yoff[k] / L->ychunk);                       // group rank of the segment's leader
            const size_t off = ((size_t) (P->y_epoch & 1) * G + P->wy->me) * L->ychunk * sizeof(double);
            const uint32_t ny = L->cf ? L->yreg[k] + (last ? L->ysrc[k] : 0) : L->yn[k];
            peer_put(ctx, P->wy, q, off, P->Yh.p + L->yoff[k], (size_t) ny * sizeof(double), P->y_epoch);
        }
        peer_put_end(ctx, P->ev_yput);
        tl_mark(P, "y_put_done", ctx->put_stream[0]);
        P->ypush_pending = true;
    }
    if (L->yn[own]) remote_parts(own);
    tl_mark(P, "own_rows_done", st);
    if (P->wy) { peer_wait_all(ctx, P->wy, P->y_epoch, st); tl_mark(P, "y_arrived", st); }     // the followers' partials of the owned segment
    else if (ctx->comm && comm_size_in(ctx->comm, P->reduce_group) > 1)
        comm_reduce_scatter_inplace(ctx->comm, P->reduce_group, P->Yh.p, L->ychunk, CT_F64, CO_SUM, st);
}

static void pull_apply(gt_program* P, bool count_active) {
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const PullLayout* L = P->pull;
    // apply_stationary's _TCSC_CF_ branch (:1671-1692): the regular rows every iteration, the source rows on the last one;
    // the positions beyond (sink columns: no row) are never applied and their x never changes
    const uint32_t n = L->cf ? P->own_hot->nreg + (cf_last(P) ? P->own_hot->nsrc : 0) : P->own_hot->n;
    if (count_active) GT_CUDA(cudaMemsetAsync(P->d_active.p, 0, sizeof(unsigned long long), st));
    if (n) {
        YParts yr{};
        if (P->wy)                                    // combine_postprocess_stationary_for_all (:1522-1541): y += y_follower
            for (int m = 0; m < P->wy->size; m++)
                if (m != P->wy->me) yr.p[yr.n++] = (const double*) P->wy->local + ((size_t) (P->y_epoch & 1) * P->wy->size + m) * L->ychunk;
        k_pr_apply_h<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(P->Yh.p + L->yoff[P->own_row_slot], yr, P->rank_h.p, P->deg_h.p, P->flag_h.p, P->C_h.p,
                                                                   pull_x_next(P) + L->xoff[P->own_col_slot], n, P->prm.alpha, P->prm.tol,
                                                                   count_active ? P->d_active.p : nullptr);
        ctx->kernel_launches++;
    }
    GT_CUDA(cudaGetLastError());
}

static void scatter_gather(gt_program* P) {
    if (P->pull) { pull_scatter_gather(P); return; }
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const SegMaps& own = (*P->pcol)[P->own_col_slot];
    if (own.nnz) {
        k_messenger_f64<<<grid_for(own.nnz, 256, ctx->sm_count), 256, 0, st>>>(P->vs(), P->app, own.ids.p, own.nnz, (double*) P->X[P->own_col_slot].p);
        ctx->kernel_launches++;
    }
    if (ctx->comm && comm_size_in(ctx->comm, P->bcast_group) > 1)      // bcast_stationary (:843-862)
        comm_allgather_inplace(ctx->comm, P->bcast_group, P->Xcat.p, P->xchunk, CT_F64, st);
    GT_CUDA(cudaGetLastError());
}

static void combine(gt_program* P) {
    if (P->pull) { pull_combine(P); return; }
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    gt_graph* g = P->g;
    P->combine_bytes = algorithmic_bytes_iteration(P, false, false, true);
    GT_CUDA(cudaMemsetAsync(P->Ycat.p, 0, P->Ycat.n, st));     // std::fill(y, 0) (:1026-1032)
    // The reference walks local_tiles_row_order (_ROW_) or local_tiles_col_order (_COL_); the order only
    // fixes when a segment's partial is shipped, which the grouped reduce below does for all at once.
    for (const Tile& T : g->tiles) {
        const uint32_t xs = (P->ordering == GT_ROW) ? T.col_slot : T.row_slot;
        const uint32_t ys = (P->ordering == GT_ROW) ? T.row_slot : T.col_slot;
        if (!T.nnz) continue;
        launch_spmv(ctx, g, T, P->semiring, P->ordering, false, P->X[xs].p, P->Y[ys].p, nullptr);
    }
    if (ctx->comm && comm_size_in(ctx->comm, P->reduce_group) > 1)
        comm_reduce_scatter_inplace(ctx->comm, P->reduce_group, P->Ycat.p, P->ychunk, CT_F64, CO_SUM, st);
}

static void apply(gt_program* P, bool count_active = false) {
    if (P->pull) { pull_apply(P, count_active); return; }
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const SegMaps& own = (*P->prow)[P->own_row_slot];
    if (!P->empty_cleared) {
        k_clear_C_empty<<<grid_for(P->th, 256, ctx->sm_count), 256, 0, st>>>(P->C.p, own.bits.p, P->th);
        ctx->kernel_launches++;
        P->empty_cleared = true;
    }
    if (own.nnz) {
        // push path on a _TCSC_CF_ graph: the SpMV covers every entry each iteration (sink columns carry x = 0 under the
        // reference's drivers, so y is the same), the applicator and the convergence test follow the CF row classes
        const uint8_t* cls = P->cf ? P->g->cls[P->g->hot_of_row_slot[P->own_row_slot]].p : nullptr;
        k_apply_f64<<<grid_for(own.nnz, 256, ctx->sm_count), 256, 0, st>>>(P->vs(), P->app, own.ids.p, own.nnz, (const double*) P->Y[P->own_row_slot].p, P->prm.alpha,
                                                                         P->prm.tol, cls, cf_last(P));
        ctx->kernel_launches++;
    }
    GT_CUDA(cudaGetLastError());
}

// all C == 0 on all ranks (:1884-1923): device count (+ ncclAllReduce), one 8-byte D2H
static void has_converged_begin(gt_program* P) {
    gt_ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    if (!P->pull) {
        GT_CUDA(cudaMemsetAsync(P->d_active.p, 0, sizeof(unsigned long long), st));
        const uint8_t* cls = P->cf ? P->g->cls[P->g->hot_of_row_slot[P->own_row_slot]].p : nullptr;
        k_count_u8<<<grid_for(P->th, 256, ctx->sm_count), 256, 0, st>>>(P->C.p, P->th, cls, P->d_active.p);
        ctx->kernel_launches++;
    }
    if (ctx->comm) comm_allreduce(ctx->comm, COMM_WORLD, P->d_active.p, P->d_active.p, 1, CT_U64, CO_SUM, st);
    GT_CUDA(cudaMemcpyAsync(P->h_active, P->d_active.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
}
static bool has_converged_end(gt_program* P) {
    GT_CUDA(cudaStreamSynchronize(P->ctx->stream));
    return P->h_active[0] == 0;
}

}  // namespace gt

// ---------------------------------------------------------------------------------------------------------
extern "C" int gt_program_create(gt_graph* g, int app, int stationary, int gather_depends_on_apply,
                                 int apply_depends_on_iter, int ordering, const gt_params* params, gt_program** out) {
    return gt::guarded([&] {
        GT_REQUIRE(g && out, "gt_program_create: NULL argument");
        GT_REQUIRE(app >= GT_APP_DEG && app <= GT_APP_SSSP, "gt_program_create: unknown app (only the five shipped programs run on the device)");
        GT_REQUIRE(ordering == GT_ROW || ordering == GT_COL, "gt_program_create: bad ordering");
        const bool want_stationary = (app == GT_APP_DEG || app == GT_APP_PR);
        GT_REQUIRE((stationary != 0) == want_stationary, "gt_program_create: this app's stationary flag differs from the reference driver's");
        if (ordering == GT_COL && !stationary)
            throw gt::Error(GT_ERR_UNSUPPORTED, "gt_program_create: _COL_ ordering exists for stationary programs only — the reference's spmv_nonstationary has no "
                                                "_COL_ branch either (src/vp/vertex_program.hpp:1437-1506), it would index y by row ids with the vectors swapped");
        GT_CUDA(cudaSetDevice(g->ctx->device));
        std::unique_ptr<gt_program> P(new gt_program());
        P->g = g; P->ctx = g->ctx; P->app = app; P->stationary = stationary;
        P->gather_depends_on_apply = gather_depends_on_apply; P->apply_depends_on_iter = apply_depends_on_iter;
        P->ordering = ordering;
        P->prm.alpha = 0.15; P->prm.tol = 1e-5; P->prm.root = 0;
        if (params) P->prm = *params;
        P->f64 = want_stationary;
        P->semiring = want_stationary ? GT_PLUS_TIMES_F64 : (g->weighted ? GT_MIN_PLUS_U32 : GT_MIN_SELECT_U32);
        gt::prog_alloc(P.get());
        *out = P.release();
    });
}

extern "C" int gt_program_free(gt_program* p) {
    return gt::guarded([&] {
        if (!p) return;
        cudaSetDevice(p->ctx->device);
        gt::ns_free(p);
        for (cudaEvent_t e : p->timeline_pool) cudaEventDestroy(e);
        if (p->h_active) cudaFreeHost(p->h_active);
        if (p->ev0) cudaEventDestroy(p->ev0);
        if (p->ev1) cudaEventDestroy(p->ev1);
        if (p->wx || p->wy) {                     // every put into these windows was consumed before execute() returned
            gt::peer_window_destroy(p->ctx, p->wx);
            gt::peer_window_destroy(p->ctx, p->wy);
        }
        if (p->ev_b) cudaEventDestroy(p->ev_b);
        for (int i = 0; i < GT_PEER_MAX_LANES; i++) if (p->ev_yput[i]) cudaEventDestroy(p->ev_yput[i]);
        delete p;
    });
}

extern "C" int gt_program_set(gt_program* p, const char* name, double value) {
    return gt::guarded([&] {
        GT_REQUIRE(p && name, "gt_program_set: NULL argument");
        std::string n(name);
        if (n == "activity_filtering_ratio") p->activity_filtering_ratio = value;
        else if (n == "timing") p->timing = value != 0;
        else if (n == "iteration") { p->iteration = (uint32_t) value; p->converged = false; }   // public member, vertex_program.hpp:60
        else if (n == "pr_layout") { p->pr_layout = (int) value; p->initialized = false; }
        else if (n == "bfs_bottom_up_ratio") p->bfs_bottom_up_ratio = value;
        else if (n == "dense_edge_ratio") p->dense_edge_ratio = value;
        else throw gt::Error(GT_ERR_INVALID, "gt_program_set: unknown knob " + n);
    });
}

extern "C" int gt_program_init_from(gt_program* p, gt_program* other) {
    return gt::guarded([&] {
        GT_REQUIRE(p && other, "gt_program_init_from: NULL argument");
        GT_REQUIRE(p->app == GT_APP_PR && other->app == GT_APP_DEG, "gt_program_init_from: only PR <- Deg is defined (src/apps/pr.h:24-28)");
        GT_REQUIRE(p->g->lay.info.tile_height == other->g->lay.info.tile_height, "gt_program_init_from: programs live on different layouts");
        GT_CUDA(cudaSetDevice(p->ctx->device));
        gt::prog_initialize(p);
        const gt::SegMaps& own = (*p->prow)[p->own_row_slot];
        gt::k_init_from_deg<<<gt::grid_for(p->th, 256, p->ctx->sm_count), 256, 0, p->ctx->stream>>>(p->vs(), other->a.p, own.bits.p, p->th, p->prm.alpha);
        p->ctx->kernel_launches++;
        GT_CUDA(cudaGetLastError());
    });
}

extern "C" int gt_program_execute(gt_program* p, uint32_t num_iterations, uint32_t* iters_done) {
    return gt::guarded([&] {
        GT_REQUIRE(p, "gt_program_execute: NULL program");
        gt_ctx* ctx = p->ctx;
        if (p->poisoned)
            throw gt::Error(GT_ERR_NCCL, "gt_program_execute: an earlier peer exchange of this program timed out; free it and create a new one");
        GT_CUDA(cudaSetDevice(ctx->device));
        if (!p->initialized) gt::prog_initialize(p);
        const bool check = num_iterations == 0;
        const uint64_t launches0 = ctx->kernel_launches;
        p->tm.bytes_algorithmic = 0;
        p->tm.sparse_iterations = 0;
        const uint32_t it0 = p->iteration;
        GT_CUDA(cudaEventRecord(p->ev0, ctx->stream));
        p->tm.scatter_gather_ms = p->tm.combine_ms = p->tm.apply_ms = 0;
        for (auto& v : p->phase_samples) v.clear();
        if (!p->stationary) {
            p->timeline_on = getenv("GT_TIMELINE") != nullptr;
            gt::ns_execute(p, num_iterations);            // records ev1, drains the stream
        } else {
            const bool peer_used = p->wx || p->wy;
            p->in_execute = true; p->check_mode = check; p->num_iterations = num_iterations;
            // the members of a group may enter execute() far apart (one still building its pull layout): meet once, on the
            // device, before the first arrival counter is polled, so the poll timeout only ever measures a real stall
            if (peer_used) gt::peer_fence_world(ctx, ctx->stream);
            // -DTIMING counters of the reference (:640-684,1018-1054,1611-1637): wall clock around each phase with
            // the stream drained, only when the "timing" knob is on (it serialises host and device)
            auto phase = [&](int which, double& acc, auto&& fn) {
                if (!p->timing) { fn(); return; }
                GT_CUDA(cudaStreamSynchronize(ctx->stream));
                const auto t0 = std::chrono::steady_clock::now();
                fn();
                GT_CUDA(cudaStreamSynchronize(ctx->stream));
                const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                acc += ms;
                p->add_sample(which, p->iteration - it0, ms);
            };
            p->timeline_on = p->pull && getenv("GT_TIMELINE") != nullptr;
            while (true) {
                phase(0, p->tm.scatter_gather_ms, [&] { gt::scatter_gather(p); });
                phase(1, p->tm.combine_ms, [&] { gt::combine(p); });
                p->tm.bytes_algorithmic += gt::algorithmic_bytes_iteration(p, gt::cf_first(p), gt::cf_last(p), false);
                phase(2, p->tm.apply_ms, [&] { gt::apply(p, check); });
                gt::tl_mark(p, "applied", ctx->stream);
                p->iteration++;
                if (check) {
                    gt::has_converged_begin(p);
                    p->converged = gt::has_converged_end(p);
                    if (p->converged) break;              // the post-convergence combine()+apply() (:425-429) changes no state for _TCSC_ ...
                } else if (p->iteration >= num_iterations) break;
            }
            p->in_execute = false;
            if (p->cf && check && p->converged) {         // ... and leaves the source rows of a _TCSC_CF_ graph at alpha (see k_pr_cf_sources)
                if (p->pull) {
                    const uint32_t lo = p->own_hot->nreg, hi = lo + p->own_hot->nsrc;
                    if (hi > lo) { gt::k_pr_cf_sources_h<<<gt::grid_for(hi - lo, 256, ctx->sm_count), 256, 0, ctx->stream>>>(p->rank_h.p, p->C_h.p, lo, hi, p->prm.alpha); ctx->kernel_launches++; }
                } else {
                    gt::k_pr_cf_sources<<<gt::grid_for(p->th, 256, ctx->sm_count), 256, 0, ctx->stream>>>(p->vs(), p->g->cls[p->g->hot_of_row_slot[p->own_row_slot]].p, p->th, p->prm.alpha);
                    ctx->kernel_launches++;
                }
            }
            if (p->pull) { gt::pull_x_arrived(p); gt::pull_state_out(p); }   // hot-order working state -> V (one pass per execute, inside the timed window)
            GT_CUDA(cudaEventRecord(p->ev1, ctx->stream));
            if (peer_used) GT_CUDA(cudaMemcpyAsync(&p->h_active[2], gt::peer_error_word(ctx), 4, cudaMemcpyDeviceToHost, ctx->stream));
            GT_CUDA(cudaStreamSynchronize(ctx->stream));
            if (peer_used && (uint32_t) p->h_active[2]) {
                const uint32_t who = (uint32_t) p->h_active[2] - 1;
                p->h_active[2] = 0;
                p->poisoned = true;
                cudaMemsetAsync(gt::peer_error_word(ctx), 0, 4, ctx->stream);
                throw gt::Error(GT_ERR_NCCL, "gt_program_execute: NVLink peer exchange timed out waiting for group member " + std::to_string(who) +
                                                 " (results are invalid; free this program and create a new one)");
            }
        }
        float ms = 0;
        GT_CUDA(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
        p->tm.execute_ms = ms;
        if (p->timeline_on) { for (int i = 0; i < GT_PEER_MAX_LANES; i++) if (ctx->put_stream[i]) cudaStreamSynchronize(ctx->put_stream[i]); gt::tl_dump(p); }
        p->tm.kernel_launches = ctx->kernel_launches - launches0;
        p->tm.iterations = p->iteration - it0;
        if (iters_done) *iters_done = p->iteration;
    });
}

extern "C" int gt_program_run_phase(gt_program* p, int phase) {
    return gt::guarded([&] {
        GT_REQUIRE(p, "gt_program_run_phase: NULL program");
        GT_REQUIRE(phase >= 0 && phase <= 2, "gt_program_run_phase: phase must be 0, 1 or 2");
        GT_CUDA(cudaSetDevice(p->ctx->device));
        if (!p->initialized) gt::prog_initialize(p);
        if (!p->stationary) { gt::ns_run_phase(p, phase); return; }
        if (p->pull) gt::pull_state_in(p);
        if (phase == 0) {
            p->x_ready = false;
            gt::scatter_gather(p);
            if (p->pull) gt::pull_x_arrived(p);
            GT_CUDA(cudaStreamSynchronize(p->ctx->stream));
        }
        else if (phase == 1) gt::combine(p);
        else gt::apply(p);
    });
}

extern "C" uint32_t gt_program_state_bytes(gt_program* p) {
    if (!p) return 0;
    return p->app == GT_APP_PR ? 16 : p->app == GT_APP_BFS ? 12 : 4;
}

extern "C" int gt_program_state_to_host(gt_program* p, void* V_out, uint64_t cap_bytes) {
    return gt::guarded([&] {
        GT_REQUIRE(p && V_out, "gt_program_state_to_host: NULL argument");
        GT_REQUIRE(p->initialized, "gt_program_state_to_host: program not initialized");
        const uint64_t bytes = (uint64_t) p->th * gt_program_state_bytes(p);
        GT_REQUIRE(cap_bytes >= bytes, "gt_program_state_to_host: buffer too small");
        GT_CUDA(cudaSetDevice(p->ctx->device));
        if (p->stage.n < bytes / 4) p->stage.alloc(bytes / 4);
        gt::DevBuf<uint32_t>& stage = p->stage;
        gt::k_pack_state<<<gt::grid_for(p->th, 256, p->ctx->sm_count), 256, 0, p->ctx->stream>>>(p->vs(), p->app, p->th, p->vid0, stage.p);
        p->ctx->kernel_launches++;
        GT_CUDA(cudaGetLastError());
        GT_CUDA(cudaMemcpyAsync(V_out, stage.p, bytes, cudaMemcpyDeviceToHost, p->ctx->stream));
        GT_CUDA(cudaStreamSynchronize(p->ctx->stream));
    });
}

extern "C" int gt_program_state_from_host(gt_program* p, const void* V_in, uint64_t bytes) {
    return gt::guarded([&] {
        GT_REQUIRE(p && V_in, "gt_program_state_from_host: NULL argument");
        const uint64_t need = (uint64_t) p->th * gt_program_state_bytes(p);
        GT_REQUIRE(bytes == need, "gt_program_state_from_host: size is not tile_height states");
        GT_CUDA(cudaSetDevice(p->ctx->device));
        if (!p->initialized) gt::prog_initialize(p);
        if (p->stage.n < need / 4) p->stage.alloc(need / 4);
        gt::DevBuf<uint32_t>& stage = p->stage;
        GT_CUDA(cudaMemcpyAsync(stage.p, V_in, need, cudaMemcpyHostToDevice, p->ctx->stream));
        p->hot_valid = false; p->x_ready = false;
        gt::k_unpack_state<<<gt::grid_for(p->th, 256, p->ctx->sm_count), 256, 0, p->ctx->stream>>>(p->vs(), p->app, p->th, stage.p);
        p->ctx->kernel_launches++;
        GT_CUDA(cudaGetLastError());
        GT_CUDA(cudaStreamSynchronize(p->ctx->stream));
    });
}

// checksum() is end-of-run reporting: the reference's u64 accumulator truncates the
// running sum at every addition, so the value depends on the order; reproduce it with the same
// sequential loop over the downloaded states.
extern "C" int gt_program_checksum(gt_program* p, uint64_t* value_sum, uint64_t* reachable) {
    return gt::guarded([&] {
        GT_REQUIRE(p && p->initialized, "gt_program_checksum: program not initialized");
        GT_CUDA(cudaSetDevice(p->ctx->device));
        cudaStream_t st = p->ctx->stream;
        const uint32_t th = p->th, nrows = p->g->lay.info.nrows;
        uint64_t sum = 0, cnt = 0;
        if (p->app == GT_APP_PR) {
            std::vector<double> r(th);
            GT_CUDA(cudaMemcpyAsync(r.data(), p->rank.p, (size_t) th * 8, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < th; i++)
                if (r[i] != 0.0 && (uint64_t) p->vid0 + i < nrows) { sum = (uint64_t) ((double) sum + r[i]); cnt++; }
        } else {
            std::vector<uint32_t> v(th);
            const uint32_t* src = (p->app == GT_APP_BFS) ? p->b.p : p->a.p;
            const uint32_t inf = (p->app == GT_APP_DEG) ? 0u : GT_INF_U32;
            GT_CUDA(cudaMemcpyAsync(v.data(), src, (size_t) th * 4, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < th; i++)
                if (v[i] != inf && (uint64_t) p->vid0 + i < nrows) { sum += v[i]; cnt++; }
        }
        if (p->ctx->comm) {
            unsigned long long h[2] = {sum, cnt};
            GT_CUDA(cudaMemcpyAsync(p->d_active.p, h, 16, cudaMemcpyHostToDevice, st));
            gt::comm_allreduce(p->ctx->comm, gt::COMM_WORLD, p->d_active.p, p->d_active.p, 2, gt::CT_U64, gt::CO_SUM, st);
            GT_CUDA(cudaMemcpyAsync(h, p->d_active.p, 16, cudaMemcpyDeviceToHost, st));
            GT_CUDA(cudaStreamSynchronize(st));
            sum = h[0]; cnt = h[1];
        }
        if (value_sum) *value_sum = sum;
        if (reachable) *reachable = cnt;
    });
}

extern "C" int gt_program_timing_samples(gt_program* p, int phase, double* out_ms, uint32_t cap, uint32_t* n) {
    return gt::guarded([&] {
        GT_REQUIRE(p && n, "gt_program_timing_samples: NULL argument");
        GT_REQUIRE(phase >= 0 && phase <= 3, "gt_program_timing_samples: phase must be 0 (scatter_gather), 1 (combine), 2 (apply) or 3 (init)");
        if (phase == 3) { *n = 1; if (out_ms && cap) out_ms[0] = p->init_ms; return; }
        const std::vector<double>& v = p->phase_samples[phase];
        *n = (uint32_t) v.size();
        for (uint32_t i = 0; out_ms && i < cap && i < v.size(); i++) out_ms[i] = v[i];
    });
}

extern "C" int gt_program_timing(gt_program* p, gt_timing* out) {
    return gt::guarded([&] {
        GT_REQUIRE(p && out, "gt_program_timing: NULL argument");
        p->tm.combine_bytes = p->combine_bytes;
        *out = p->tm;
    });
}

// ---- kernel-level entry points -------------------------------------------------------------------------------
extern "C" int gt_tile_spmv(gt_graph* g, uint32_t local_tile, int semiring, int ordering, const void* x, void* y) {
    return gt::guarded([&] {
        GT_REQUIRE(g && x && y, "gt_tile_spmv: NULL argument");
        GT_REQUIRE(local_tile < g->tiles.size(), "gt_tile_spmv: tile index out of range");
        GT_REQUIRE(semiring >= GT_PLUS_TIMES_F64 && semiring <= GT_MIN_SELECT_U32, "gt_tile_spmv: unknown semiring");
        GT_CUDA(cudaSetDevice(g->ctx->device));
        gt::launch_spmv(g->ctx, g, g->tiles[local_tile], semiring, ordering, semiring != GT_PLUS_TIMES_F64, x, y, nullptr);
    });
}

extern "C" int gt_tile_spmspv(gt_graph* g, uint32_t local_tile, int semiring, const uint32_t* xi, const void* xv,
                              uint32_t k, void* y, uint8_t* t) {
    return gt::guarded([&] {
        GT_REQUIRE(g && y && (k == 0 || (xi && xv)), "gt_tile_spmspv: NULL argument");
        GT_REQUIRE(local_tile < g->tiles.size(), "gt_tile_spmspv: tile index out of range");
        GT_REQUIRE(semiring >= GT_PLUS_TIMES_F64 && semiring <= GT_MIN_SELECT_U32, "gt_tile_spmspv: unknown semiring");
        GT_CUDA(cudaSetDevice(g->ctx->device));
        gt::launch_spmspv(g->ctx, g, g->tiles[local_tile], semiring, xi, xv, k, y, t);
    });
}
