// gt_internal.h — shared declarations of libgraphtap_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <memory>
#include "../../include/graphtap_b200.h"

namespace gt {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);

#define GT_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw gt::Error(GT_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e__) + \
                                             " (" __FILE__ ":" + std::to_string(__LINE__) + ")");   \
    } while (0)

#define GT_REQUIRE(cond, msg)                                                       \
    do {                                                                            \
        if (!(cond)) throw gt::Error(GT_ERR_INVALID, std::string(msg));             \
    } while (0)

// Wraps every extern "C" body: exceptions -> status + gt_last_error().
template <typename F>
static inline int guarded(F&& f) {
    try {
        f();
        return GT_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return GT_ERR_OOM;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return GT_ERR_INVALID;
    }
}

// ---- device buffer ------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) {
            cudaError_t e = cudaMalloc((void**) &p, count * sizeof(T));
            if (e != cudaSuccess) {
                p = nullptr; n = 0;
                throw Error(e == cudaErrorMemoryAllocation ? GT_ERR_OOM : GT_ERR_CUDA,
                            std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e));
            }
        }
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
};

// ---- layout (gt_layout.cpp): host restatement of Matrix::init_matrix -----------------------------
struct Layout {
    gt_layout info{};
    std::vector<int32_t> tile_rank;              // [nrowgrps * ncolgrps] after the leader swap
    std::vector<int32_t> leader_ranks;           // [nrowgrps]
    std::vector<int32_t> local_tiles_row_order;  // kth = rg * ncolgrps + cg
    std::vector<int32_t> local_tiles_col_order;
    std::vector<int32_t> local_row_segments, local_col_segments;
    std::vector<int32_t> all_rowgrp_ranks, all_colgrp_ranks;         // sorted
    std::vector<int32_t> follower_rowgrp_ranks, follower_colgrp_ranks;
    int row_slot_of(int seg) const {
        for (size_t i = 0; i < local_row_segments.size(); i++) if (local_row_segments[i] == seg) return (int) i;
        return -1;
    }
    int col_slot_of(int seg) const {
        for (size_t i = 0; i < local_col_segments.size(); i++) if (local_col_segments[i] == seg) return (int) i;
        return -1;
    }
};
Layout make_layout(uint32_t nvertices, int nranks, int rank);
struct RoutePlan {                               // partitioned ingest: where this rank's blocks go (entries, not bytes)
    std::vector<uint64_t> send_offset;           // [nranks] first entry of the block for q in my send buffer
    std::vector<uint64_t> recv_offset;           // [nranks] first entry of the block from q in my receive buffer
    std::vector<uint64_t> remote_offset;         // [nranks] first entry of MY block in q's receive buffer
    uint64_t nsend = 0, nrecv = 0, max_recv = 0; // my totals; the largest receive buffer of any rank
};
RoutePlan make_route_plan(int nranks, int rank, const uint64_t* counts);

// ---- NCCL through dlopen (gt_comm.cpp) -------------------------------------------------------------
struct Comm;   // opaque: world + row-group + col-group communicators
Comm* comm_create(int rank, int nranks, const void* unique_id, const Layout& lay, cudaStream_t stream);
void comm_destroy(Comm* c);
void comm_unique_id(void* out128);
enum CommGroup { COMM_WORLD = 0, COMM_ROWGRP = 1, COMM_COLGRP = 2 };
enum CommType { CT_U32 = 0, CT_F64 = 1, CT_U64 = 2, CT_U8 = 3 };
enum CommOp { CO_SUM = 0, CO_MIN = 1, CO_MAX = 2 };
int comm_rank_in(Comm* c, CommGroup g);
int comm_size_in(Comm* c, CommGroup g);
int comm_index_of_world_rank(Comm* c, CommGroup g, int world_rank);
void comm_group_start(Comm* c);
void comm_group_end(Comm* c);
void comm_bcast(Comm* c, CommGroup g, void* buf, size_t count, CommType t, int root, cudaStream_t s);
void comm_reduce(Comm* c, CommGroup g, const void* send, void* recv, size_t count, CommType t, CommOp op, int root, cudaStream_t s);
void comm_allgather_inplace(Comm* c, CommGroup g, void* buf, size_t count, CommType t, cudaStream_t s);
void comm_reduce_scatter_inplace(Comm* c, CommGroup g, void* buf, size_t count, CommType t, CommOp op, cudaStream_t s);
void comm_allreduce(Comm* c, CommGroup g, const void* send, void* recv, size_t count, CommType t, CommOp op, cudaStream_t s);
void comm_alltoallv_bytes(Comm* c, const uint8_t* send, const uint64_t* scount, const uint64_t* sdispl, uint8_t* recv, const uint64_t* rcount,
                          const uint64_t* rdispl, cudaStream_t s);

}  // namespace gt

#define GT_PEER_MAX_LANES 4
// ---- the three opaque handle types ------------------------------------------------------------------
struct gt_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    gt::Comm* comm = nullptr;
    int sm_count = 0;
    uint64_t kernel_launches = 0;     // every launch this library makes increments this
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // gt_ctx_timer_*
    cudaStream_t comm_stream = nullptr;         // collectives that overlap compute (x all-gather of the pull path)
    cudaEvent_t ev_x = nullptr, ev_ag = nullptr; // own x chunk written / all-gather landed
    // NVLink peer exchange (gt_peer.cu): source table of the arrival counters' values, error word, poll timeout, and
    // the streams ("lanes") a put is spread over so that several copy engines drive the links at once
    gt::DevBuf<uint32_t> peer_seq, peer_err, peer_fence, barrier_word;
    double peer_timeout_ms = 30000.0;
    int peer_lanes = 0;
    cudaStream_t put_stream[GT_PEER_MAX_LANES] = {};
};
