// gt_comm.cpp — NCCL over NVLink 5 / NVSwitch in place of the reference's MPI back end.
//
// The reference builds a world communicator plus one row-group and one column-group communicator
// from explicit rank lists (src/mpi/env.hpp:104-124, lists from src/mat/matrix.hpp:382-465) and then
// issues, per iteration: Ibcast of every x segment along the column group
// (src/vp/vertex_program.hpp:843-862,970-1013), a follower->leader send of every partial y along the
// row group with the reduction done by the leader (:1083-1108,1522-1573), and a world Allreduce for
// convergence (:1918).  Here those are ONE in-place ncclAllGather (x, column group), ONE in-place
// ncclReduceScatter (y, row group) and an ncclAllReduce per iteration, on communicators split with the
// same rank lists (every member of a group leads exactly one of the group's segments).  PageRank's iteration no
// longer uses them: its x / y exchange runs over NVLink peer windows (gt_peer.cu), NCCL only carries the window
// handles at start-up; BFS / CC / SSSP and the GT_PEER=0 mode still do.
//
// NCCL is bound with dlopen("libnccl.so.2") so that (a) inside a Python process the copy torch already
// loaded is reused (one NCCL per process), (b) a single-GPU run needs no NCCL at all, and (c) the
// library still loads on a machine without NCCL.
#include "gt_internal.h"
#include <dlfcn.h>
#include <algorithm>

namespace gt {

// Minimal NCCL declarations (ABI-stable since 2.x); avoids a build-time dependency on nccl.h.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // 0 = success
// ncclDataType_t / ncclRedOp_t values from nccl.h
enum { NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_PROD = 1, NCCL_MAX = 2, NCCL_MIN = 3 };

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static Nccl& nccl() {
    static Nccl n;
    if (n.h) return n;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.h) break;
    }
    if (!n.h) throw Error(GT_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(n.h, s);
        if (!p) throw Error(GT_ERR_NCCL, std::string("NCCL symbol missing: ") + s);
        return p;
    };
    n.GetUniqueId = (decltype(n.GetUniqueId)) sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank)) sym("ncclCommInitRank");
    n.CommSplit = (decltype(n.CommSplit)) sym("ncclCommSplit");
    n.CommDestroy = (decltype(n.CommDestroy)) sym("ncclCommDestroy");
    n.Broadcast = (decltype(n.Broadcast)) sym("ncclBroadcast");
    n.Reduce = (decltype(n.Reduce)) sym("ncclReduce");
    n.AllReduce = (decltype(n.AllReduce)) sym("ncclAllReduce");
    n.AllGather = (decltype(n.AllGather)) sym("ncclAllGather");
    n.ReduceScatter = (decltype(n.ReduceScatter)) sym("ncclReduceScatter");
    n.Send = (decltype(n.Send)) sym("ncclSend");
    n.Recv = (decltype(n.Recv)) sym("ncclRecv");
    n.GroupStart = (decltype(n.GroupStart)) sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd)) sym("ncclGroupEnd");
    n.GetErrorString = (decltype(n.GetErrorString)) sym("ncclGetErrorString");
    return n;
}

#define GT_NCCL(call)                                                                                 \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != 0) throw Error(GT_ERR_NCCL, std::string(#call) + " failed: " + nccl().GetErrorString(r__)); \
    } while (0)

struct Comm {
    int rank = 0, nranks = 1;
    ncclComm_t world = nullptr, rowgrp = nullptr, colgrp = nullptr;
    std::vector<int32_t> row_ranks, col_ranks;   // group rank -> world rank (sorted lists)
};

void comm_unique_id(void* out128) {
    ncclUniqueId id;
    GT_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
}

Comm* comm_create(int rank, int nranks, const void* unique_id, const Layout& lay, cudaStream_t) {
    GT_REQUIRE(unique_id, "gt_ctx_create: nranks > 1 needs the 128-byte NCCL unique id");
    Comm* c = new Comm();
    c->rank = rank;
    c->nranks = nranks;
    c->row_ranks = lay.all_rowgrp_ranks;
    c->col_ranks = lay.all_colgrp_ranks;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    GT_NCCL(nccl().CommInitRank(&c->world, nranks, id, rank));
    // Every member of a group derives the same colour (the smallest world rank in its list) and its key
    // is its position in the sorted list, so group ranks equal the reference's rank_rg / rank_cg.
    int row_key = (int) (std::find(c->row_ranks.begin(), c->row_ranks.end(), rank) - c->row_ranks.begin());
    int col_key = (int) (std::find(c->col_ranks.begin(), c->col_ranks.end(), rank) - c->col_ranks.begin());
    GT_NCCL(nccl().CommSplit(c->world, c->row_ranks.front(), row_key, &c->rowgrp, nullptr));
    GT_NCCL(nccl().CommSplit(c->world, c->col_ranks.front(), col_key, &c->colgrp, nullptr));
    return c;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->rowgrp) nccl().CommDestroy(c->rowgrp);
    if (c->colgrp) nccl().CommDestroy(c->colgrp);
    if (c->world) nccl().CommDestroy(c->world);
    delete c;
}

static ncclComm_t pick(Comm* c, CommGroup g) { return g == COMM_WORLD ? c->world : g == COMM_ROWGRP ? c->rowgrp : c->colgrp; }
static int nccl_type(CommType t) { return t == CT_U32 ? NCCL_UINT32 : t == CT_F64 ? NCCL_FLOAT64 : t == CT_U64 ? NCCL_UINT64 : NCCL_UINT8; }
static int nccl_op(CommOp o) { return o == CO_SUM ? NCCL_SUM : o == CO_MIN ? NCCL_MIN : NCCL_MAX; }

int comm_size_in(Comm* c, CommGroup g) {
    return g == COMM_WORLD ? c->nranks : g == COMM_ROWGRP ? (int) c->row_ranks.size() : (int) c->col_ranks.size();
}
int comm_index_of_world_rank(Comm* c, CommGroup g, int world_rank) {
    if (g == COMM_WORLD) return world_rank;
    const auto& v = g == COMM_ROWGRP ? c->row_ranks : c->col_ranks;
    auto it = std::find(v.begin(), v.end(), world_rank);
    GT_REQUIRE(it != v.end(), "comm: world rank is not a member of the group");
    return (int) (it - v.begin());
}
int comm_rank_in(Comm* c, CommGroup g) { return comm_index_of_world_rank(c, g, c->rank); }

void comm_group_start(Comm*) { GT_NCCL(nccl().GroupStart()); }
void comm_group_end(Comm*) { GT_NCCL(nccl().GroupEnd()); }

void comm_bcast(Comm* c, CommGroup g, void* buf, size_t count, CommType t, int root, cudaStream_t s) {
    GT_NCCL(nccl().Broadcast(buf, buf, count, nccl_type(t), root, pick(c, g), s));
}
void comm_reduce(Comm* c, CommGroup g, const void* send, void* recv, size_t count, CommType t, CommOp op, int root, cudaStream_t s) {
    GT_NCCL(nccl().Reduce(send, recv, count, nccl_type(t), nccl_op(op), root, pick(c, g), s));
}
// in place: rank r contributes buf[r*count .. (r+1)*count) and receives everything
void comm_allgather_inplace(Comm* c, CommGroup g, void* buf, size_t count, CommType t, cudaStream_t s) {
    const size_t es = t == CT_F64 || t == CT_U64 ? 8 : t == CT_U32 ? 4 : 1;
    GT_NCCL(nccl().AllGather((const char*) buf + (size_t) comm_rank_in(c, g) * count * es, buf, count, nccl_type(t), pick(c, g), s));
}
// in place: every rank passes size*count elements, rank r ends up with the reduction of chunk r at buf[r*count ..)
void comm_reduce_scatter_inplace(Comm* c, CommGroup g, void* buf, size_t count, CommType t, CommOp op, cudaStream_t s) {
    const size_t es = t == CT_F64 || t == CT_U64 ? 8 : t == CT_U32 ? 4 : 1;
    GT_NCCL(nccl().ReduceScatter(buf, (char*) buf + (size_t) comm_rank_in(c, g) * count * es, count, nccl_type(t), nccl_op(op), pick(c, g), s));
}
// world all-to-all-v of raw bytes: rank r sends send[sdispl[q] .. + scount[q]) to rank q and receives rcount[q] bytes from q
// at recv[rdispl[q]] (the reference's pairwise Sendrecv redistribution, src/mat/matrix.hpp:692-810, as one grouped exchange)
void comm_alltoallv_bytes(Comm* c, const uint8_t* send, const uint64_t* scount, const uint64_t* sdispl, uint8_t* recv, const uint64_t* rcount,
                          const uint64_t* rdispl, cudaStream_t s) {
    GT_NCCL(nccl().GroupStart());
    for (int q = 0; q < c->nranks; q++) {
        if (scount[q]) GT_NCCL(nccl().Send(send + sdispl[q], scount[q], NCCL_UINT8, q, c->world, s));
        if (rcount[q]) GT_NCCL(nccl().Recv(recv + rdispl[q], rcount[q], NCCL_UINT8, q, c->world, s));
    }
    GT_NCCL(nccl().GroupEnd());
}
void comm_allreduce(Comm* c, CommGroup g, const void* send, void* recv, size_t count, CommType t, CommOp op, cudaStream_t s) {
    GT_NCCL(nccl().AllReduce(send, recv, count, nccl_type(t), nccl_op(op), pick(c, g), s));
}

}  // namespace gt
