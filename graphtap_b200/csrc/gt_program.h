// gt_program.h — the vertex-program handle shared by the stationary engine (gt_engine.cu: Deg / PageRank) and the
// non-stationary engine (gt_ns.cu: BFS / CC / SSSP).
#pragma once
#include "gt_kernels.cuh"
#include "gt_pull.h"
#include "gt_peer.h"
#include <algorithm>

namespace gt {

static inline int grid_for(uint64_t n, int block, int sm_count, int per_sm = 8) {
    uint64_t g = (n + block - 1) / block;
    uint64_t cap = (uint64_t) sm_count * per_sm;
    return (int) std::max<uint64_t>(1, std::min(g, cap));
}

// ---- vertex state, SoA on the device (the reference's AoS std::vector<Vertex_State> V is produced on
// demand by gt_program_state_to_host) --------------------------------------------------------------
struct VState {
    double* rank;        // PR
    uint32_t* a;         // Deg/PR degree | BFS parent | CC label | SSSP distance
    uint32_t* b;         // BFS hops
    uint8_t* C;          // activity / convergence flags (:161)
};

struct NsState;          // device descriptors, windows and counters of a non-stationary program (gt_ns.cu)

}  // namespace gt

struct gt_program {
    gt_graph* g = nullptr;
    gt_ctx* ctx = nullptr;
    int app = 0, stationary = 0, gather_depends_on_apply = 0, apply_depends_on_iter = 0, ordering = GT_ROW;
    gt_params prm{};
    int semiring = 0;
    bool f64 = false;
    uint32_t th = 0, vid0 = 0;
    // program-level views: under GT_COL "rows" are the matrix's column groups (:279-325)
    std::vector<gt::SegMaps>* prow = nullptr;
    std::vector<gt::SegMaps>* pcol = nullptr;
    int own_row_slot = 0, own_col_slot = 0;
    gt::CommGroup bcast_group = gt::COMM_COLGRP, reduce_group = gt::COMM_ROWGRP;
    // state
    gt::DevBuf<double> rank;
    gt::DevBuf<uint32_t> a, b;
    gt::DevBuf<uint8_t> C;
    struct Span { uint8_t* p = nullptr; size_t n = 0; };
    // ---- stationary programs (gt_engine.cu) -------------------------------------------------------------------
    gt::DevBuf<uint8_t> Xcat;                      // all local x segments back to back + one trailing zero element
    std::vector<Span> X;                           // views into Xcat, |x|*esize bytes each
    gt::DevBuf<uint8_t> Ycat;                      // all local y segments, same chunking along the row group
    std::vector<Span> Y;                           // views into Ycat, |y|*esize bytes each
    size_t xchunk = 0, ychunk = 0;                 // chunk sizes in elements
    int pr_layout = 1;                             // 1: derived pull layout for the plus-times SpMV (gt_pull.cu), 0: push over TCSC
    const gt::PullLayout* pull = nullptr;          // owned by the graph
    // pull mode: x / y in hot order, and the owned segment's state in hot order while execute() runs
    gt::DevBuf<double> Xh;                         // concatenated hot-ordered x of the local column segments (+ one 0.0)
    gt::DevBuf<double> Yh;                         // concatenated y chunks of the local row segments
    // NVLink peer exchange (gt_peer.cu).  wx: the members of the column group put their x chunk into each other's
    // window, two buffers alternating by epoch parity (a rank one iteration ahead writes x(k+1) while a slower one
    // still reads x(k)).  wy: followers put the partial y of a row segment into its leader's window, one slot per
    // sender, again two parities.  Without a window the same exchange is one ncclAllGather / ncclReduceScatter.
    gt::PeerWindow* wx = nullptr;
    gt::PeerWindow* wy = nullptr;
    double* xbuf[2] = {nullptr, nullptr};          // the x buffer of each parity (both = Xh.p without wx)
    size_t x_stride = 0;                           // doubles between the two x buffers inside wx
    uint32_t x_epoch = 0, y_epoch = 0;             // puts issued so far (= the value the arrival counters must reach)
    bool x_wait_pending = false, ypush_pending = false, pull_ready = false;
    cudaEvent_t ev_b = nullptr, ev_yput[GT_PEER_MAX_LANES] = {};   // y complete for the follower segments / their puts have read Yh
    gt::DevBuf<double> rank_h;
    gt::DevBuf<uint32_t> deg_h;
    gt::DevBuf<uint8_t> flag_h, C_h;
    const gt::HotOrder* own_hot = nullptr;
    gt::DevBuf<uint32_t> stage;                    // AoS staging of V for gt_program_state_{to,from}_host (kept: no malloc per call)
    bool hot_valid = false, x_ready = false, ag_pending = false;
    // _TCSC_CF_ graphs: the computation-filtering schedule of the running execute() (vertex_program.hpp:1218-1325,1671-1692)
    bool cf = false;                               // PageRank on a GT_TCSC_CF graph
    bool in_execute = false, check_mode = false;
    uint32_t num_iterations = 0;
    uint64_t combine_bytes = 0;                    // algorithmic bytes of the most recent combine phase
    // GT_TIMELINE=<path prefix>: CUDA events at the phase boundaries of every iteration of the pull path, on the main
    // stream and on the put streams, written as one JSON line per execute() and rank (there is no nsys in the image)
    struct Mark { cudaEvent_t ev; const char* tag; uint32_t iteration; };
    std::vector<Mark> timeline;
    std::vector<cudaEvent_t> timeline_pool;
    bool timeline_on = false;
    // ---- non-stationary programs (gt_ns.cu) ------------------------------------------------------------------
    gt::NsState* ns = nullptr;
    // ---- both -----------------------------------------------------------------------------------------------------
    gt::DevBuf<unsigned long long> d_active;      // [0], [1]: active count of the iteration of that parity, [2..3]: scratch
    unsigned long long* h_active = nullptr;       // pinned mirror + [2]: peer error word
    bool initialized = false, converged = false, empty_cleared = false, poisoned = false;
    uint32_t iteration = 0;
    double activity_filtering_ratio = 0.6;        // :194
    double dense_edge_ratio = 0.5;                // non-stationary: a frontier holding more than this share of the segment's edges runs the dense pass (0 = columns only)
    double bfs_bottom_up_ratio = 0.05;            // BFS on an undirected single-GPU graph: bottom-up pass above this frontier share (0 = never)
    bool timing = false;
    gt_timing tm{};
    // -DTIMING vectors of the reference (:202-208): one wall-clock sample per iteration and phase of the latest execute()
    // (0 scatter_gather, 1 combine, 2 apply), and init_time; filled only while the "timing" knob is on
    std::vector<double> phase_samples[3];
    double init_ms = 0;
    void add_sample(int phase, uint32_t it, double ms) {
        std::vector<double>& v = phase_samples[phase];
        if (v.size() <= it) v.resize((size_t) it + 1, 0.0);
        v[it] += ms;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    gt::VState vs() { return gt::VState{rank.p, a.p, b.p, C.p}; }
    size_t esize() const { return f64 ? 8 : 4; }
};

namespace gt {

template <typename T>
__global__ void k_fill(T* p, T v, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) p[i] = v;
}

// GT_TIMELINE marks (gt_engine.cu)
void tl_mark(gt_program* P, const char* tag, cudaStream_t s);
void tl_dump(gt_program* P);

// non-stationary engine (gt_ns.cu)
void ns_alloc(gt_program* P);                      // buffers, windows, tile descriptors
void ns_free(gt_program* P);
void ns_initialize(gt_program* P);                 // Y <- infinity (:625-635), counters
void ns_execute(gt_program* P, uint32_t num_iterations);     // the iteration loop of execute() (:416-433)
void ns_run_phase(gt_program* P, int phase);

}  // namespace gt
