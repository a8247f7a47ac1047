/*
 * graphtap_b200.h — C ABI of libgraphtap_b200.so: the B200-native replacement for the one hot path of
 * hmofrad/GraphTap, the vertex-program SpMV/SpMSpV engine over 2D-partitioned TCSC tiles.
 *
 * The reference has no FFI: it is header-only C++ templates whose hot loops call virtuals per edge
 * (SURVEY.md §8b).  This header is the seam a maintainer binds instead: plain pointers and sizes,
 * opaque handles, `int` status (0 = ok) + gt_last_error().  Every entry point cites the reference
 * interface it replaces (paths relative to the reference repo).  The source-compatible C++ shim that
 * keeps `Graph<>::load`, `Vertex_Program<>` and the five `*_Program` classes on top of these calls is
 * include/graphtap/graphtap.hpp; INTEGRATION.md shows the binding.
 *
 * One host thread per GPU, one process per GPU (the reference: one MPI rank per process,
 * src/mpi/env.hpp:77-93).  There is NO CPU fallback: every compute entry point fails with
 * GT_ERR_NO_DEVICE when no sm_100 device is present.
 */
#ifndef GRAPHTAP_B200_H
#define GRAPHTAP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT_ABI_VERSION 1
#if defined(__GNUC__)
#define GT_API __attribute__((visibility("default")))
#else
#define GT_API
#endif

/* ---- status ------------------------------------------------------------------------------- */
enum {
    GT_OK = 0,
    GT_ERR_INVALID = 1,      /* bad argument / unsupported combination                          */
    GT_ERR_NO_DEVICE = 2,    /* no CUDA device, or not sm_100                                   */
    GT_ERR_CUDA = 3,         /* a CUDA runtime call failed (message in gt_last_error)            */
    GT_ERR_NCCL = 4,         /* NCCL missing or a collective failed                              */
    GT_ERR_OOM = 5,
    GT_ERR_UNSUPPORTED = 6   /* valid in the reference, deliberately not provided (see message)  */
};
/* The reference reports errors with fprintf(stderr)+Env::exit(1) (src/mpi/env.hpp:159-162); the
 * shim does the same with this string. Thread-local. */
GT_API const char* gt_last_error(void);
GT_API int gt_abi_version(void);

/* ---- enums mirroring the reference's ------------------------------------------------------- */
/* src/ds/compressed_column.hpp:17-23 (Compression_type). _CSC_/_DCSC_ are not GPU formats. */
enum { GT_TCSC = 2, GT_TCSC_CF = 3 };
/* src/vp/vertex_program.hpp:17-21 (Ordering_type) */
enum { GT_ROW = 0, GT_COL = 1 };
/* the five shipped vertex programs, src/apps/{deg,pr,bfs,cc,sssp}.h */
enum { GT_APP_DEG = 0, GT_APP_PR = 1, GT_APP_BFS = 2, GT_APP_CC = 3, GT_APP_SSSP = 4 };
/* the (⊕,⊗) pairs those programs' combiner() overloads implement */
enum {
    GT_PLUS_TIMES_F64 = 0,   /* y += x [* w]            src/apps/pr.h:35-41, deg.h:41-47 */
    GT_MIN_PLUS_U32 = 1,     /* y = min(y, x + w)       src/apps/sssp.h:49-52             */
    GT_MIN_SELECT_U32 = 2    /* y = min(y, x)           src/apps/bfs.h:61-63, cc.h:47-49   */
};
#define GT_INF_U32 2147483647u   /* src/apps/bfs.h:12 */

typedef struct gt_ctx gt_ctx;
typedef struct gt_graph gt_graph;
typedef struct gt_program gt_program;

/* ---- context: replaces Env::init / rowgrps_init / colgrps_init / finalize --------------------
 * (src/mpi/env.hpp:77-124,140-157).  nranks > 1 builds a world NCCL communicator from `nccl_id`
 * (128 bytes from gt_nccl_unique_id on rank 0, distributed by the host program) and splits the
 * row-group and column-group communicators from the reference's rank lists
 * (src/mat/matrix.hpp:447-465).  The per-iteration exchanges of the reference (Ibcast of x along the column
 * group, Isend/Irecv of partial y along the row group, src/vp/vertex_program.hpp:843-1013,1083-1108) run as
 * stores into cudaIpc-mapped NVLink peer windows of the other group members where the ranks are processes of one
 * node (PageRank: x and y; BFS/CC/SSSP: the sparse-or-dense frontier), and as NCCL collectives otherwise or with
 * the environment variable GT_PEER=0; the communicators also carry the window handles and the convergence
 * all-reduce (:1918). */
GT_API int gt_nccl_unique_id(void* out128);
GT_API int gt_ctx_create(int device, int rank, int nranks, const void* nccl_id, gt_ctx** out);
GT_API int gt_ctx_destroy(gt_ctx* ctx);
/* CUDA-event stopwatch on the engine stream (the stream every kernel of this library is launched on):
 * begin records an event, end records another, synchronises and returns the elapsed device time. */
GT_API int gt_ctx_timer_begin(gt_ctx* ctx);
GT_API int gt_ctx_timer_end(gt_ctx* ctx, double* elapsed_ms);
GT_API int gt_ctx_sync(gt_ctx* ctx);                      /* cudaStreamSynchronize on the engine stream */
/* Env::barrier (src/mpi/env.hpp:164-166, MPI_Barrier on the world): drains this rank's streams and returns only
 * when every rank of the job has done the same (a 1-element all-reduce over the world communicator). */
GT_API int gt_ctx_barrier(gt_ctx* ctx);
GT_API void* gt_ctx_stream(gt_ctx* ctx);                  /* the cudaStream_t every kernel is launched on */
/* device scratch the caller may use for staging (cudaMalloc / cudaFree / copies on the ctx stream) */
GT_API int gt_dev_alloc(gt_ctx* ctx, size_t bytes, void** out);
GT_API int gt_dev_free(gt_ctx* ctx, void* p);
GT_API int gt_dev_upload(gt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
GT_API int gt_dev_download(gt_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
GT_API int gt_dev_memset(gt_ctx* ctx, void* dst_dev, int byte, size_t bytes);
GT_API int gt_host_alloc_pinned(size_t bytes, void** out);
GT_API int gt_host_free_pinned(void* p);

/* ---- 2D-transposed tile layout: replaces Matrix::Matrix + init_matrix + Tiling -----------------
 * (src/mat/matrix.hpp:184-202,272-495; src/mat/tiling.hpp:39-73).  Pure host arithmetic; usable
 * without a GPU. */
typedef struct {
    uint32_t nranks, rank;
    uint32_t nrows;                 /* nvertices + 1                 src/mat/graph.hpp:89-90      */
    uint32_t nrowgrps, ncolgrps;    /* = nranks each                 src/mat/graph.hpp:98         */
    uint32_t tile_height;           /* nrows / nrowgrps + 1          src/mat/matrix.hpp:193-194   */
    uint32_t rowgrp_nranks, colgrp_nranks;      /* src/mat/tiling.hpp:65-73 */
    uint32_t rank_nrowgrps, rank_ncolgrps;      /* src/mat/tiling.hpp:55-56 */
    int32_t owned_segment;          /* src/mat/matrix.hpp:367-370 */
    int32_t accu_segment_rg, accu_segment_cg, accu_segment_row, accu_segment_col; /* :466-485 */
} gt_layout;
enum {
    GT_LT_TILE_RANK = 0,            /* nrowgrps*ncolgrps owners after the leader swap  :299-341 */
    GT_LT_LEADER_RANKS = 1,         /* :338 */
    GT_LT_LOCAL_TILES_ROW_ORDER = 2,/* :355 */
    GT_LT_LOCAL_TILES_COL_ORDER = 3,/* :374-380 */
    GT_LT_LOCAL_ROW_SEGMENTS = 4,   /* :359-360 */
    GT_LT_LOCAL_COL_SEGMENTS = 5,   /* :356-357 */
    GT_LT_ALL_ROWGRP_RANKS = 6,     /* sorted, :447 */
    GT_LT_ALL_COLGRP_RANKS = 7,     /* sorted, :454 */
    GT_LT_FOLLOWER_ROWGRP_RANKS = 8,/* :405,451 */
    GT_LT_FOLLOWER_COLGRP_RANKS = 9 /* :433,458 */
};
GT_API int gt_layout_query(uint32_t nvertices, int nranks, int rank, gt_layout* out);
GT_API int gt_layout_table(uint32_t nvertices, int nranks, int rank, int which, int32_t* out, uint32_t cap, uint32_t* n);

/* ---- graph: replaces Graph::load / load_binary / free -------------------------------------------
 * (src/mat/graph.hpp:41-46,75-81,104-191,307-372) and everything they drive: edge pre-processing
 * flags, distribute, sort + dedup (src/mat/matrix.hpp:537-560), filter_vertices (:860-1122),
 * classify_vertices (:1124-1282), TCSC populate (src/ds/compressed_column.hpp:370-417). */
typedef struct {
    int directed, transpose, self_loops, acyclic, parallel_edges;   /* src/mat/graph.hpp:41-43 */
} gt_graph_flags;

/* `triples` is the reference's on-disk record array (src/ds/triple.hpp:9-18,40-49):
 * {u32 row, u32 col} or, weighted, {u32 row, u32 col, u32 w}.  It is the GLOBAL edge list: every rank
 * passes the same records (or generates them) and keeps the tiles it owns — no exchange, the cheap way for
 * generated graphs; gt_graph_build_partitioned below is the reference's read-a-share-then-redistribute
 * (src/mat/matrix.hpp:692-810).
 * `on_device` != 0: `triples` is a device pointer (already resident in HBM). */
GT_API int gt_graph_build(gt_ctx* ctx, const void* triples, uint64_t ntriples, int weighted, int on_device,
                   uint32_t nvertices, const gt_graph_flags* flags, int compression, gt_graph** out);
/* Same, from the counter-based synthetic RMAT generator (graphtap_b200/rmat.py documents the stream;
 * Graph500 a,b,c,d, label permutation, weights in [1,128]).  Bench tooling: the reference ships no
 * generator. */
GT_API int gt_graph_build_rmat(gt_ctx* ctx, uint32_t scale, uint64_t nedges, uint64_t seed, int weighted,
                        const gt_graph_flags* flags, int compression, gt_graph** out);
/* Partitioned ingest = the reference's own scheme: every rank passes ITS SHARE of the records (Graph::parread_binary reads
 * bytes [rank*share, ...) of the file, src/mat/graph.hpp:307-335), the per-edge flags run on the share and every entry
 * travels to the rank that owns its tile (Matrix::distribute's pairwise Sendrecv rounds, src/mat/matrix.hpp:692-810, as
 * copies into the owners' NVLink peer windows, or one grouped NCCL send/recv exchange where the ranks cannot map each
 * other); the non-empty row/column marks and degrees are all-reduced (the reference OR-reduces them along the
 * row/column groups, :973-1083).  Any split of the global list gives the same graph as gt_graph_build on
 * the whole list: tiles, maps and orders are bit-identical.  Collective over the context's ranks; with one rank it is
 * gt_graph_build.  nedges_input of the result = records over all shares. */
GT_API int gt_graph_build_partitioned(gt_ctx* ctx, const void* share, uint64_t nshare, int weighted, int on_device,
                   uint32_t nvertices, const gt_graph_flags* flags, int compression, gt_graph** out);
/* gt_graph_build_rmat with every rank generating records [rank * (nedges / nranks), ...) (the last rank also the
 * remainder) and routing them as above. */
GT_API int gt_graph_build_rmat_partitioned(gt_ctx* ctx, uint32_t scale, uint64_t nedges, uint64_t seed, int weighted,
                        const gt_graph_flags* flags, int compression, gt_graph** out);
/* The exchange plan gt_graph_build_partitioned follows, as host arithmetic (callable without a GPU; tests/test_ingest_share.py
 * drives it over gloo): counts[r * nranks + q] = entries rank r holds for rank q (what Matrix::distribute's Sendrecv
 * rounds exchange pairwise, src/mat/matrix.hpp:692-810).  Per destination / source q: where the block for q starts in this
 * rank's send buffer, where the block from q starts in its receive buffer, and where this rank's block starts in q's
 * receive buffer (blocks lie in sender order); totals, and the largest receive buffer of any rank.  Units: entries. */
GT_API int gt_ingest_route_plan(int nranks, int rank, const uint64_t* counts, uint64_t* send_offset, uint64_t* recv_offset,
                         uint64_t* remote_offset, uint64_t* nsend, uint64_t* nrecv, uint64_t* max_recv);
GT_API int gt_rmat_generate(gt_ctx* ctx, uint32_t scale, uint64_t first_edge, uint64_t nedges, uint64_t seed,
                     int weighted, void* triples_dev);
GT_API int gt_graph_free(gt_graph* g);

typedef struct {
    uint64_t nedges_input;          /* raw records                      src/mat/graph.hpp:335      */
    uint64_t nnz_local;             /* stored entries in this rank's tiles                         */
    uint64_t nnz_global;            /* sum over ranks (== nnz_local at nranks 1)                   */
    uint32_t ntiles_local;          /* = nranks                                                    */
    uint32_t weighted;
    uint32_t nvertices;
    gt_layout layout;
} gt_graph_info;
GT_API int gt_graph_info_get(gt_graph* g, gt_graph_info* out);

/* One local tile in `local_tiles_row_order`; device pointers into the graph (valid until
 * gt_graph_free).  Field-for-field TCSC_BASE (src/ds/compressed_column.hpp:287-296). */
typedef struct {
    uint32_t rg, cg;                /* tile coordinates in the grid                                */
    uint32_t row_slot, col_slot;    /* index into this rank's local row / col segment lists        */
    uint64_t nnz;
    uint32_t nnzcols, nnzrows;      /* group-wide non-empty counts = |x|, |y|                      */
    const uint32_t* JA;             /* [nnzcols+1] column pointers over compressed columns         */
    const uint32_t* IA;             /* [nnz] compressed row ids                                    */
    const uint32_t* A;              /* [nnz] weights or NULL                                       */
    const uint32_t* JC;             /* [nnzcols] compressed -> local column id                     */
    const uint32_t* IR;             /* [nnzrows] compressed -> local row id                        */
} gt_tile_view;
GT_API int gt_graph_tile_view(gt_graph* g, uint32_t local_tile, gt_tile_view* out);
/* Index maps of a local segment slot (src/mat/matrix.hpp:82-85): bitvector I/J (u8[tile_height]) and
 * prefix map IV/JV (u32[tile_height], 0 where empty).  Device pointers. */
GT_API int gt_graph_rowgrp_maps(gt_graph* g, uint32_t row_slot, const uint8_t** I, const uint32_t** IV, uint32_t* nnzrows);
GT_API int gt_graph_colgrp_maps(gt_graph* g, uint32_t col_slot, const uint8_t** J, const uint32_t** JV, uint32_t* nnzcols);

/* classify_vertices on the owned segment (src/mat/matrix.hpp:1124-1144): regular = row and column non-empty, source
 * rows = row non-empty / column empty, sink columns = row empty / column non-empty.  These are the sets the reference's
 * _TCSC_CF_ "computation filtering" schedules by (src/vp/vertex_program.hpp:1218-1325,1671-1692,1902-1916); a graph built
 * with GT_TCSC_CF runs that schedule on the device (gt_program_execute).  gt_graph_classify returns the counts (any
 * compression), gt_graph_classify_lists the local vertex ids (rowgrp_regular_rows / rowgrp_source_rows /
 * colgrp_sink_columns, :853-855; device pointers, GT_TCSC_CF graphs only). */
GT_API int gt_graph_classify(gt_graph* g, uint32_t* regular, uint32_t* source_rows, uint32_t* sink_columns);
GT_API int gt_graph_classify_lists(gt_graph* g, const uint32_t** regular_rows, uint32_t* nregular, const uint32_t** source_rows, uint32_t* nsource,
                            const uint32_t** sink_columns, uint32_t* nsink);
/* TCSC_CF_BASE (src/ds/compressed_column.hpp:419-1114) of one local tile of a GT_TCSC_CF graph: gt_graph_tile_view's IA
 * (and A) are in the reference's order — every column's source-row entries moved behind its regular-row entries by
 * the reference's own swap sequence (:671-708) — and the four computation-filtering lists, kind 0 = REG_R_REG_C,
 * 1 = REG_R_SNK_C, 2 = SRC_R_REG_C, 3 = SRC_R_SNK_C: NC[k] (start, end) pairs into IA plus NC[k] compressed column ids.
 * `filled[k]` pairs are written, the rest are zero: the reference sizes SRC_R_SNK_C by an EDGE count (:1046-1049) and
 * starts its ranges at JA[j] + n (:1094); both quirks are reproduced so the arrays can be diffed. */
typedef struct {
    uint32_t NC[4], filled[4];
    const uint32_t* JA[4];          /* [2 * NC[k]] */
    const uint32_t* JC[4];          /* [NC[k]]     */
} gt_tile_cf_view;
GT_API int gt_graph_tile_cf_view(gt_graph* g, uint32_t local_tile, gt_tile_cf_view* out);

/* ---- kernel level: replaces Vertex_Program::spmv_stationary / spmv_nonstationary -----------------
 * (src/vp/vertex_program.hpp:96-105,1115-1327,1437-1506).  x and y are device vectors in the tile's
 * compressed spaces (|x| = nnzcols, |y| = nnzrows; with GT_COL the roles swap, :1175-1183).
 * gt_tile_spmv: y ⊕= A ⊗ x over every column (min semirings skip x == GT_INF_U32, :1492).
 * gt_tile_spmspv: only the k frontier columns xi[0..k) with values xv[0..k) (:1476-1488);
 *                 t (optional) receives the touched-row flags (:1486). */
GT_API int gt_tile_spmv(gt_graph* g, uint32_t local_tile, int semiring, int ordering, const void* x, void* y);
GT_API int gt_tile_spmspv(gt_graph* g, uint32_t local_tile, int semiring, const uint32_t* xi, const void* xv,
                   uint32_t k, void* y, uint8_t* t);

/* ---- engine level: replaces Vertex_Program<...> ---------------------------------------------------
 * ctor flags (src/vp/vertex_program.hpp:27-29), execute (:407-441), initialize(other) (:466-501),
 * checksum (:1926-1960), free (:335-405), public V (:61). */
typedef struct {
    double alpha;                   /* src/apps/pr.h:13 (0.15)   */
    double tol;                     /* src/apps/pr.h:12 (1e-5)   */
    uint32_t root;                  /* bfs.h:35, sssp.h:32       */
} gt_params;
GT_API int gt_program_create(gt_graph* g, int app, int stationary, int gather_depends_on_apply,
                      int apply_depends_on_iter, int ordering, const gt_params* params, gt_program** out);
GT_API int gt_program_free(gt_program* p);
/* initialize(const Vertex_Program& other): Deg -> PR degree hand-over, only where the row is
 * non-empty (:479-482). */
GT_API int gt_program_init_from(gt_program* p, gt_program* other);
/* execute(num_iterations); 0 = run until has_converged() (:412-433).  iters_done = `iteration`. */
GT_API int gt_program_execute(gt_program* p, uint32_t num_iterations, uint32_t* iters_done);
/* Bytes of one vertex state in the reference's AoS layout: Deg_State 4, PR_State 16
 * {u32 degree; u32 pad; f64 rank}, BFS_State 12 {parent,hops,vid}, CC_State 4, SSSP_State 4. */
GT_API uint32_t gt_program_state_bytes(gt_program* p);
/* V of the owned segment: tile_height states, vertex id = owned_segment*tile_height + i (:1804-1808). */
GT_API int gt_program_state_to_host(gt_program* p, void* V_out, uint64_t cap_bytes);
GT_API int gt_program_state_from_host(gt_program* p, const void* V_in, uint64_t bytes);
/* checksum(): u64 running sum with per-element truncation, and reachable count (:1929-1958);
 * allreduced over ranks. */
GT_API int gt_program_checksum(gt_program* p, uint64_t* value_sum, uint64_t* reachable);
typedef struct {
    double execute_ms;              /* the reference's "Execute time" window (:416-437), CUDA events */
    double scatter_gather_ms, combine_ms, apply_ms;   /* -DTIMING counters (:202-208), 0 unless enabled */
    uint64_t kernel_launches;       /* kernels this library launched inside execute()              */
    uint64_t bytes_algorithmic;     /* SURVEY.md §8(d) algorithmic bytes moved inside execute()    */
    uint32_t iterations;
    uint32_t sparse_iterations;     /* iterations that ran the frontier SpMSpV                     */
    uint64_t combine_bytes;         /* algorithmic bytes of the most recent combine phase (stationary programs): what the SpMV
                                       pass of that iteration had to move; on a GT_TCSC_CF graph a middle iteration skips the
                                       REG x SNK and source-row entries, as the reference does                                */
} gt_timing;
GT_API int gt_program_timing(gt_program* p, gt_timing* out);
/* The reference's -DTIMING vectors (src/vp/vertex_program.hpp:202-208) for the most recent execute(), with the "timing" knob
 * on: one wall-clock sample in ms per iteration of phase 0 scatter_gather_time, 1 combine_time, 2 apply_time; phase 3 = the
 * single init_time sample.  *n = samples available; at most `cap` are written.  display() prints its `TIMING` line from
 * these (:2134-2152). */
GT_API int gt_program_timing_samples(gt_program* p, int phase, double* out_ms, uint32_t cap, uint32_t* n);
/* One phase of one iteration in isolation, for per-phase timing (the reference's -DTIMING counters,
 * src/vp/vertex_program.hpp:640-684,1018-1054,1611-1637): 0 scatter_gather, 1 combine (the SpMV /
 * SpMSpV over every local tile + the row-group reduce), 2 apply.  Does not advance `iteration`. */
GT_API int gt_program_run_phase(gt_program* p, int phase);
/* knobs: name = "activity_filtering_ratio" (default 0.6, :194; 1.0 also switches the edge rule below off), "timing" (0/1),
 * "pr_layout" (0 = push over TCSC, 1 = derived pull layout), "iteration" (the public member, :60),
 * "dense_edge_ratio" (default 0.5: a frontier holding more than this share of its segment's edges runs the dense pass
 * although the reference's column rule calls it sparse; 0 = column rule only; results do not depend on it),
 * "bfs_bottom_up_ratio" (default 0.05: BFS on an undirected single-GPU graph runs bottom-up above this share of active
 * columns; 0 = never). */
GT_API int gt_program_set(gt_program* p, const char* name, double value);

#ifdef __cplusplus
}
#endif
#endif
