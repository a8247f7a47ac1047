/*
 * gt_oracle.c — CPU restatement of GraphTap's vertex-program SpMV path, in plain C.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this; nothing under graphtap_b200/ does, and the product has no CPU path.
 *
 * PARITY IS PINNED: tests/test_oracle.py checks this file against (a) the known answers the
 * reference's own code gives on its own fixtures, (b) per-vertex dumps and per-tile
 * TCSC arrays of the UNMODIFIED reference built into oracle/_ref (np = 1, 2, 4, 8 through the
 * fork+shm MPI stand-in), committed under tests/golden/.
 *
 * It simulates all p ranks in one process.  Every function cites the reference code it follows
 * (paths relative to the GraphTap repo).  For p > 1 the f64 summation order of the reference is kept:
 * each rank accumulates its own tiles of a tile-row in column order, then the row-group leader adds
 * the followers' partial vectors in follower order (src/vp/vertex_program.hpp:1061-1111,1522-1541).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define GTO_INF 2147483647u                 /* src/apps/bfs.h:12 */
enum { GTO_DEG = 0, GTO_PR = 1, GTO_BFS = 2, GTO_CC = 3, GTO_SSSP = 4 };

typedef struct { uint32_t row, col, w; } entry_t;

typedef struct {
    uint64_t nnz;
    uint32_t *JA, *IA, *A;                  /* TCSC_BASE, src/ds/compressed_column.hpp:287-296 */
} tile_t;

typedef struct {
    uint32_t nnz;                           /* non-empty rows (cols) of the whole row (col) group */
    uint8_t* bits;                          /* I / J  [tile_height] */
    uint32_t* prefix;                       /* IV / JV, 0 where empty (src/mat/matrix.hpp:1030-1041) */
    uint32_t* ids;                          /* IR / JC */
} seg_t;

typedef struct {
    uint32_t p, th, nrows, nvertices;
    uint32_t rowgrp_nranks, colgrp_nranks;
    int weighted;
    int32_t* tile_rank;                     /* [p*p] after the leader swap */
    int32_t* leader;                        /* [p] */
    tile_t* tiles;                          /* [p*p] */
    seg_t *rows, *cols;                     /* [p] */
    uint64_t nnz;
} gto_graph;

/* ---- layout: Matrix::init_matrix (src/mat/matrix.hpp:272-341) + Tiling (src/mat/tiling.hpp:39-73) */
static void layout(gto_graph* g) {
    uint32_t p = g->p, a = (uint32_t) sqrt((double) p), b = a;
    while (a * b != p) { b++; a = p / b; }                     /* tiling.hpp:65-73 */
    g->rowgrp_nranks = a; g->colgrp_nranks = b;
    int32_t* R = (int32_t*) malloc(sizeof(int32_t) * p * p);
    for (uint32_t i = 0; i < p; i++)
        for (uint32_t j = 0; j < p; j++) R[i * p + j] = (int32_t) ((i % b) * a + (j % a));   /* matrix.hpp:301-302 */
    g->leader = (int32_t*) malloc(sizeof(int32_t) * p);
    for (uint32_t i = 0; i < p; i++) g->leader[i] = -1;
    int32_t* tmp = (int32_t*) malloc(sizeof(int32_t) * p);
    for (uint32_t i = 0; i < p; i++) {                          /* matrix.hpp:330-341 */
        for (uint32_t j = i; j < p; j++) {
            int found = 0;
            for (uint32_t k = 0; k < p; k++) if (g->leader[k] == R[j * p + i]) found = 1;
            if (!found) {
                memcpy(tmp, R + j * p, sizeof(int32_t) * p);
                memcpy(R + j * p, R + i * p, sizeof(int32_t) * p);
                memcpy(R + i * p, tmp, sizeof(int32_t) * p);
                break;
            }
        }
        g->leader[i] = R[i * p + i];
    }
    free(tmp);
    g->tile_rank = R;
}

int gto_layout(uint32_t nvertices, int p, int32_t* tile_rank, int32_t* leader_ranks, uint32_t* tile_height,
               uint32_t* rowgrp_nranks, uint32_t* colgrp_nranks) {
    gto_graph g; memset(&g, 0, sizeof(g));
    g.p = (uint32_t) p;
    layout(&g);
    if (tile_rank) memcpy(tile_rank, g.tile_rank, sizeof(int32_t) * p * p);
    if (leader_ranks) memcpy(leader_ranks, g.leader, sizeof(int32_t) * p);
    if (tile_height) *tile_height = (nvertices + 1) / (uint32_t) p + 1;     /* matrix.hpp:193, graph.hpp:89 */
    if (rowgrp_nranks) *rowgrp_nranks = g.rowgrp_nranks;
    if (colgrp_nranks) *colgrp_nranks = g.colgrp_nranks;
    free(g.tile_rank); free(g.leader);
    return 0;
}

/* ---- sort orders: ColSort (src/ds/triple.hpp:78-98) */
static int cmp_unweighted(const void* x, const void* y) {
    const entry_t *a = (const entry_t*) x, *b = (const entry_t*) y;
    if (a->col != b->col) return a->col < b->col ? -1 : 1;
    if (a->row != b->row) return a->row < b->row ? -1 : 1;
    return 0;
}
/* HAS_WEIGHT: (col, weight); the reference's std::sort leaves ties unordered, here ties break by row so
 * the restatement is deterministic (any tie order gives the same min-plus result) */
static int cmp_weighted(const void* x, const void* y) {
    const entry_t *a = (const entry_t*) x, *b = (const entry_t*) y;
    if (a->col != b->col) return a->col < b->col ? -1 : 1;
    if (a->w != b->w) return a->w < b->w ? -1 : 1;
    if (a->row != b->row) return a->row < b->row ? -1 : 1;
    return 0;
}

void gto_free(gto_graph* g) {
    if (!g) return;
    for (uint32_t t = 0; t < g->p * g->p; t++) { free(g->tiles[t].JA); free(g->tiles[t].IA); free(g->tiles[t].A); }
    for (uint32_t s = 0; s < g->p; s++) {
        free(g->rows[s].bits); free(g->rows[s].prefix); free(g->rows[s].ids);
        free(g->cols[s].bits); free(g->cols[s].prefix); free(g->cols[s].ids);
    }
    free(g->tiles); free(g->rows); free(g->cols); free(g->tile_rank); free(g->leader); free(g);
}

static void seg_finish(seg_t* s, uint32_t th) {      /* prefix + id list, matrix.hpp:1026-1041, compressed_column.hpp:399-416 */
    uint32_t k = 0;
    for (uint32_t i = 0; i < th; i++) { if (s->bits[i]) { s->prefix[i] = k++; } else s->prefix[i] = 0; }
    s->nnz = k;
    s->ids = (uint32_t*) malloc(sizeof(uint32_t) * (k ? k : 1));
    k = 0;
    for (uint32_t i = 0; i < th; i++) if (s->bits[i]) s->ids[k++] = i;
}

/* Graph::parread_binary flags (src/mat/graph.hpp:337-356) -> Matrix::insert (:267-270) -> init_tiles sort +
 * dedup (matrix.hpp:544-557) -> filter_vertices (:860-1122) -> TCSC_BASE::populate (compressed_column.hpp:370-417) */
gto_graph* gto_build(const uint32_t* triples, uint64_t n, int weighted, uint32_t nvertices, int p_,
                     int directed, int transpose, int self_loops, int acyclic, int parallel_edges) {
    gto_graph* g = (gto_graph*) calloc(1, sizeof(gto_graph));
    const uint32_t p = (uint32_t) p_;
    g->p = p; g->nvertices = nvertices; g->nrows = nvertices + 1; g->weighted = weighted;
    g->th = g->nrows / p + 1;
    layout(g);
    const uint32_t th = g->th;
    const int rec = weighted ? 3 : 2;
    /* pass 1: count per tile */
    uint64_t* cnt = (uint64_t*) calloc((size_t) p * p + 1, sizeof(uint64_t));
    for (int pass = 0; pass < 2; pass++) {
        entry_t** fill = NULL;
        static entry_t** bufs;
        if (pass == 1) {
            bufs = (entry_t**) malloc(sizeof(entry_t*) * p * p);
            for (uint32_t t = 0; t < p * p; t++) { bufs[t] = (entry_t*) malloc(sizeof(entry_t) * (cnt[t] ? cnt[t] : 1)); cnt[t] = 0; }
            fill = bufs;
        }
        for (uint64_t e = 0; e < n; e++) {
            uint32_t r = triples[e * rec], c = triples[e * rec + 1], w = weighted ? triples[e * rec + 2] : 1;
            if (r == c && !self_loops) continue;                                    /* graph.hpp:339-342 */
            if (acyclic && c < r) { uint32_t x = r; r = c; c = x; }                 /* :344-347 */
            if (transpose) { uint32_t x = r; r = c; c = x; }                        /* :349-350 */
            for (int k = 0; k < (directed ? 1 : 2); k++) {                          /* :352-357 */
                uint32_t rr = k ? c : r, cc = k ? r : c;
                uint32_t t = (rr / th) * p + (cc / th);                             /* matrix.hpp:218-220 */
                if (fill) { entry_t en = {rr, cc, w}; fill[t][cnt[t]] = en; }
                cnt[t]++;
            }
        }
        if (pass == 1) {
            g->tiles = (tile_t*) calloc((size_t) p * p, sizeof(tile_t));
            g->rows = (seg_t*) calloc(p, sizeof(seg_t));
            g->cols = (seg_t*) calloc(p, sizeof(seg_t));
            for (uint32_t s = 0; s < p; s++) {
                g->rows[s].bits = (uint8_t*) calloc(th, 1); g->rows[s].prefix = (uint32_t*) calloc(th, 4);
                g->cols[s].bits = (uint8_t*) calloc(th, 1); g->cols[s].prefix = (uint32_t*) calloc(th, 4);
            }
            for (uint32_t t = 0; t < p * p; t++) {
                entry_t* E = bufs[t];
                uint64_t m = cnt[t];
                qsort(E, m, sizeof(entry_t), weighted ? cmp_weighted : cmp_unweighted);
                if (!parallel_edges && m) {                                        /* std::unique on (row,col), matrix.hpp:545,553 */
                    uint64_t o = 0;
                    for (uint64_t i = 1; i < m; i++)
                        if (!(E[i].row == E[o].row && E[i].col == E[o].col)) E[++o] = E[i];
                    m = o + 1;
                }
                cnt[t] = m;
                const uint32_t rg = t / p, cg = t % p;
                for (uint64_t i = 0; i < m; i++) { g->rows[rg].bits[E[i].row % th] = 1; g->cols[cg].bits[E[i].col % th] = 1; }
            }
            for (uint32_t s = 0; s < p; s++) { seg_finish(&g->rows[s], th); seg_finish(&g->cols[s], th); }
            for (uint32_t t = 0; t < p * p; t++) {                                  /* populate, compressed_column.hpp:381-398 */
                const uint32_t rg = t / p, cg = t % p;
                tile_t* T = &g->tiles[t];
                entry_t* E = bufs[t];
                const uint64_t m = cnt[t];
                const uint32_t nc = g->cols[cg].nnz;
                T->nnz = m;
                T->JA = (uint32_t*) calloc((size_t) nc + 1, 4);
                T->IA = (uint32_t*) malloc(4 * (m ? m : 1));
                T->A = (uint32_t*) malloc(4 * (m ? m : 1));
                for (uint64_t i = 0; i < m; i++) {
                    T->JA[g->cols[cg].prefix[E[i].col % th] + 1]++;
                    T->IA[i] = g->rows[rg].prefix[E[i].row % th];
                    T->A[i] = E[i].w;
                }
                for (uint32_t j = 0; j < nc; j++) T->JA[j + 1] += T->JA[j];
                g->nnz += m;
                free(E);
            }
            free(bufs);
        }
    }
    free(cnt);
    return g;
}

uint32_t gto_tile_height(const gto_graph* g) { return g->th; }
uint64_t gto_nnz(const gto_graph* g) { return g->nnz; }
uint64_t gto_tile_nnz(const gto_graph* g, uint32_t rg, uint32_t cg) { return g->tiles[rg * g->p + cg].nnz; }
const uint32_t* gto_tile_JA(const gto_graph* g, uint32_t rg, uint32_t cg) { return g->tiles[rg * g->p + cg].JA; }
const uint32_t* gto_tile_IA(const gto_graph* g, uint32_t rg, uint32_t cg) { return g->tiles[rg * g->p + cg].IA; }
const uint32_t* gto_tile_A(const gto_graph* g, uint32_t rg, uint32_t cg) { return g->tiles[rg * g->p + cg].A; }
uint32_t gto_seg_nnz(const gto_graph* g, int is_col, uint32_t s) { return (is_col ? g->cols : g->rows)[s].nnz; }
const uint8_t* gto_seg_bits(const gto_graph* g, int is_col, uint32_t s) { return (is_col ? g->cols : g->rows)[s].bits; }
const uint32_t* gto_seg_prefix(const gto_graph* g, int is_col, uint32_t s) { return (is_col ? g->cols : g->rows)[s].prefix; }
const uint32_t* gto_seg_ids(const gto_graph* g, int is_col, uint32_t s) { return (is_col ? g->cols : g->rows)[s].ids; }

/* ---- TCSC_CF: classify_vertices (src/mat/matrix.hpp:1124-1144) + TCSC_CF_BASE::populate (src/ds/compressed_column.hpp:671-1114)
 * Restated literally, quirks included: the swap runs only if the column holds a non-source entry
 * (:682), NC_SRC_R_SNK_C counts EDGES (:1046-1049) and its ranges start at JA[j] + n (:1094).  One call produces, for
 * tile (rg, cg): the CF-ordered IA (and A), and the four (start,end)-pair lists with their JC lists.
 * kind: 0 REG_R_REG_C, 1 REG_R_SNK_C, 2 SRC_R_REG_C, 3 SRC_R_SNK_C. */
typedef struct {
    uint32_t* IA; uint32_t* A;
    uint32_t NC[4];            /* allocated pairs (the reference's NC_*) */
    uint32_t filled[4];        /* pairs actually written (== NC except for SRC_R_SNK_C) */
    uint32_t* JA[4];           /* 2 * NC */
    uint32_t* JC[4];           /* NC */
} gto_cf_tile;

/* vertex classes of segment s: 1 regular (row and column non-empty), 2 source row (row only), 3 sink column (column only), 0 neither */
void gto_classify(const gto_graph* g, uint32_t s, uint8_t* cls) {
    for (uint32_t i = 0; i < g->th; i++) {
        const int r = g->rows[s].bits[i], c = g->cols[s].bits[i];
        cls[i] = (uint8_t) (r && c ? 1 : r ? 2 : c ? 3 : 0);                       /* matrix.hpp:1135-1144 */
    }
}

void gto_cf_free(gto_cf_tile* T) {
    if (!T) return;
    free(T->IA); free(T->A);
    for (int k = 0; k < 4; k++) { free(T->JA[k]); free(T->JC[k]); }
    free(T);
}

gto_cf_tile* gto_cf_build(const gto_graph* g, uint32_t rg, uint32_t cg) {
    const tile_t* B = &g->tiles[rg * g->p + cg];
    const uint32_t nnzcols = g->cols[cg].nnz;
    const uint32_t* JA = B->JA;
    const uint32_t* JC = g->cols[cg].ids;
    const uint32_t* IR = g->rows[rg].ids;
    gto_cf_tile* T = (gto_cf_tile*) calloc(1, sizeof(gto_cf_tile));
    const uint64_t m = B->nnz;
    T->IA = (uint32_t*) malloc(4 * (m ? m : 1)); T->A = (uint32_t*) malloc(4 * (m ? m : 1));
    memcpy(T->IA, B->IA, 4 * m); memcpy(T->A, B->A, 4 * m);
    uint32_t* IA = T->IA; uint32_t* A = T->A;
    uint8_t* rcls = (uint8_t*) malloc(g->th); uint8_t* ccls = (uint8_t*) malloc(g->th);
    gto_classify(g, rg, rcls); gto_classify(g, cg, ccls);
#define SRC(i) (rcls[IR[IA[i]]] == 2)                                             /* source_rows_bitvector[IR[IA[i]]] */
    /* "Moving source rows to the end" :671-708 */
    uint32_t* r = (uint32_t*) malloc(4 * (m ? m : 1));
    for (uint32_t j = 0; j < nnzcols; j++) {
        uint32_t n = 0, mm = 0;
        for (uint32_t i = JA[j]; i < JA[j + 1]; i++) if (SRC(i)) { mm = JA[j + 1] - JA[j]; r[n++] = i; }
        if (mm > 0 && mm > n) {
            for (uint32_t pp = 0; pp < n; pp++) {
                for (uint32_t q = JA[j + 1] - 1; ; q--) {                         /* the reference's q >= JA[j] cannot fail: mm > n */
                    if (!SRC(q)) { uint32_t t = IA[r[pp]]; IA[r[pp]] = IA[q]; IA[q] = t; t = A[r[pp]]; A[r[pp]] = A[q]; A[q] = t; break; }
                    else if (r[pp] == q) break;
                }
            }
        }
    }
    free(r);
    /* the four lists; a column is "local" when it has entries in this tile (:649-657) */
    for (int kind = 0; kind < 4; kind++) {
        const int want_col = (kind == 0 || kind == 2) ? 1 : 3;                     /* regular / sink columns of the column group */
        for (int pass = 0; pass < 2; pass++) {
            uint32_t l = 0, o = 0, kcnt = 0;
            for (uint32_t j = 0; j < nnzcols; j++) {
                if (JA[j] == JA[j + 1] || ccls[JC[j]] != want_col) continue;      /* the JC_LOCAL_VAL x {regular,sink}_columns merge */
                uint32_t n = 0;
                const uint32_t len = JA[j + 1] - JA[j];
                for (uint32_t i = JA[j]; i < JA[j + 1]; i++) if (SRC(i)) n++;
                if (kind <= 1) {                                                  /* :749-833, :862-946: regular rows */
                    if (n == len) continue;                                       /* all-source column: no pair */
                    if (pass) { T->JA[kind][2 * l] = JA[j]; T->JA[kind][2 * l + 1] = JA[j + 1] - n; T->JC[kind][o++] = j; }
                    l++;
                } else if (kind == 2) {                                           /* :949-1022 */
                    if (!n) continue;
                    if (pass) { T->JA[kind][2 * l] = JA[j + 1] - n; T->JA[kind][2 * l + 1] = JA[j + 1]; T->JC[kind][o++] = j; }
                    l++;
                } else {                                                          /* :1025-1108: NC counts edges, start is JA[j] + n */
                    kcnt += n;
                    if (!n) continue;
                    if (pass) { T->JA[kind][2 * l] = JA[j] + n; T->JA[kind][2 * l + 1] = JA[j + 1]; T->JC[kind][o++] = j; }
                    l++;
                }
            }
            if (!pass) {
                T->NC[kind] = kind == 3 ? kcnt : l;
                T->JA[kind] = (uint32_t*) calloc(2 * (size_t) T->NC[kind] + 1, 4);
                T->JC[kind] = (uint32_t*) calloc((size_t) T->NC[kind] + 1, 4);
            } else T->filled[kind] = l;
        }
    }
#undef SRC
    free(rcls); free(ccls);
    return T;
}
uint32_t gto_cf_nc(const gto_cf_tile* T, int kind) { return T->NC[kind]; }
uint32_t gto_cf_filled(const gto_cf_tile* T, int kind) { return T->filled[kind]; }
const uint32_t* gto_cf_ja(const gto_cf_tile* T, int kind) { return T->JA[kind]; }
const uint32_t* gto_cf_jc(const gto_cf_tile* T, int kind) { return T->JC[kind]; }
const uint32_t* gto_cf_ia(const gto_cf_tile* T) { return T->IA; }
const uint32_t* gto_cf_a(const gto_cf_tile* T) { return T->A; }

/* ---- one tile, one semiring: spmv_stationary / spmv_nonstationary ------------------------------------
 * semiring 0 plus-times f64 (pr.h:35-41), 1 min-plus u32 (sssp.h:49-52), 2 min-select u32 (bfs.h:61-63)
 * ordering 0 = _ROW_ (vertex_program.hpp:1164-1172), 1 = _COL_ (:1175-1183) */
void gto_tile_spmv_f64(const gto_graph* g, uint32_t rg, uint32_t cg, int ordering, const double* x, double* y) {
    const tile_t* T = &g->tiles[rg * g->p + cg];
    const uint32_t nc = g->cols[cg].nnz;
    for (uint32_t j = 0; j < nc; j++)
        for (uint32_t i = T->JA[j]; i < T->JA[j + 1]; i++) {
            if (ordering == 0) y[T->IA[i]] += g->weighted ? x[j] * (double) T->A[i] : x[j];
            else y[j] += g->weighted ? x[T->IA[i]] * (double) T->A[i] : x[T->IA[i]];
        }
}
/* dense non-stationary branch (:1491-1502): skip columns whose x is infinity() */
void gto_tile_spmv_u32(const gto_graph* g, uint32_t rg, uint32_t cg, int plus, const uint32_t* x, uint32_t* y, uint8_t* t) {
    const tile_t* T = &g->tiles[rg * g->p + cg];
    const uint32_t nc = g->cols[cg].nnz;
    for (uint32_t j = 0; j < nc; j++) {
        if (x[j] == GTO_INF) continue;
        for (uint32_t i = T->JA[j]; i < T->JA[j + 1]; i++) {
            const uint32_t v = plus ? x[j] + T->A[i] : x[j];
            if (v < y[T->IA[i]]) y[T->IA[i]] = v;
            if (t) t[T->IA[i]] = 1;
        }
    }
}
/* sparse branch (:1476-1488) */
void gto_tile_spmspv_u32(const gto_graph* g, uint32_t rg, uint32_t cg, int plus, const uint32_t* xi, const uint32_t* xv, uint32_t k,
                         uint32_t* y, uint8_t* t) {
    const tile_t* T = &g->tiles[rg * g->p + cg];
    for (uint32_t f = 0; f < k; f++) {
        const uint32_t j = xi[f];
        for (uint32_t i = T->JA[j]; i < T->JA[j + 1]; i++) {
            const uint32_t v = plus ? xv[f] + T->A[i] : xv[f];
            if (v < y[T->IA[i]]) y[T->IA[i]] = v;
            if (t) t[T->IA[i]] = 1;
        }
    }
}

/* ---- Deg on the stored matrix: column (ordering 1, pr.cpp:40-43) or row (ordering 0, deg.cpp) entry counts.
 * deg[v] for v in [0, p*th); stays 0 where the applicator is never called (:1666-1667). */
void gto_degree(const gto_graph* g, int ordering, uint32_t* deg) {
    const uint32_t p = g->p, th = g->th;
    memset(deg, 0, sizeof(uint32_t) * (size_t) p * th);
    for (uint32_t s = 0; s < p; s++) {
        const seg_t* S = ordering ? &g->cols[s] : &g->rows[s];
        double* y = (double*) calloc(S->nnz ? S->nnz : 1, sizeof(double));
        for (uint32_t o = 0; o < p; o++) {
            const uint32_t rg = ordering ? o : s, cg = ordering ? s : o;
            const seg_t* XS = ordering ? &g->rows[rg] : &g->cols[cg];
            double* x = (double*) malloc(sizeof(double) * (XS->nnz ? XS->nnz : 1));
            for (uint32_t j = 0; j < XS->nnz; j++) x[j] = 1.0;                      /* deg.h:37-39 */
            gto_tile_spmv_f64(g, rg, cg, ordering, x, y);
            free(x);
        }
        for (uint32_t k = 0; k < S->nnz; k++) deg[(size_t) s * th + S->ids[k]] = (uint32_t) y[k];   /* deg.h:49-52 */
        free(y);
    }
}

/* ---- PageRank, fixed iteration count or until convergence (TCSC semantics) ----------------------------------
 * pr.cpp:26-53: Deg with _COL_ on M = (dst,src), hand-over where the row is non-empty
 * (vertex_program.hpp:479-482), then `iters` iterations of scatter (pr.h:31-33) / combine / apply (pr.h:43-47).
 * rank/deg cover p*th vertices (padding ids keep rank = alpha).  iters == 0: run until no |delta| > tol. */
uint32_t gto_pagerank(const gto_graph* g, uint32_t iters, double alpha, double tol, double* rank, uint32_t* deg_out) {
    const uint32_t p = g->p, th = g->th;
    const size_t nall = (size_t) p * th;
    uint32_t* deg = (uint32_t*) malloc(sizeof(uint32_t) * nall);
    gto_degree(g, 1, deg);
    for (uint32_t s = 0; s < p; s++)
        for (uint32_t i = 0; i < th; i++)
            if (!g->rows[s].bits[i]) deg[(size_t) s * th + i] = 0;                  /* initialize(other) only where I[i] */
    for (size_t v = 0; v < nall; v++) rank[v] = alpha;
    double** X = (double**) malloc(sizeof(double*) * p);
    for (uint32_t s = 0; s < p; s++) X[s] = (double*) malloc(sizeof(double) * (g->cols[s].nnz ? g->cols[s].nnz : 1));
    double** part = (double**) malloc(sizeof(double*) * p);          /* one partial y per rank */
    uint32_t it = 0;
    for (;;) {
        for (uint32_t s = 0; s < p; s++)                              /* scatter_gather_stationary :699-705 */
            for (uint32_t j = 0; j < g->cols[s].nnz; j++) {
                const size_t v = (size_t) s * th + g->cols[s].ids[j];
                X[s][j] = deg[v] ? rank[v] / deg[v] : 0.0;
            }
        uint64_t active = 0;
        for (uint32_t rg = 0; rg < p; rg++) {
            const uint32_t nr = g->rows[rg].nnz;
            for (uint32_t r = 0; r < p; r++) part[r] = NULL;
            for (uint32_t cg = 0; cg < p; cg++) {                      /* each rank walks its tiles of this row in column order */
                const int32_t owner = g->tile_rank[rg * p + cg];
                if (!part[owner]) part[owner] = (double*) calloc(nr ? nr : 1, sizeof(double));
                if (g->tiles[rg * p + cg].nnz) gto_tile_spmv_f64(g, rg, cg, 0, X[cg], part[owner]);
            }
            double* y = part[g->leader[rg]];
            for (uint32_t r = 0; r < p; r++)                          /* followers in sorted rank order :1530-1539 */
                if (part[r] && (int32_t) r != g->leader[rg]) { for (uint32_t k = 0; k < nr; k++) y[k] += part[r][k]; }
            for (uint32_t k = 0; k < nr; k++) {                        /* apply_stationary :1655-1670, pr.h:43-47 */
                const size_t v = (size_t) rg * th + g->rows[rg].ids[k];
                const double tmp = rank[v];
                rank[v] = alpha + (1.0 - alpha) * y[k];
                if (fabs(rank[v] - tmp) > tol) active++;
            }
            for (uint32_t r = 0; r < p; r++) free(part[r]);
        }
        it++;
        if (iters ? it >= iters : active == 0) break;
    }
    if (deg_out) memcpy(deg_out, deg, sizeof(uint32_t) * nall);
    for (uint32_t s = 0; s < p; s++) free(X[s]);
    free(X); free(part); free(deg);
    return it;
}

/* ---- PageRank on _TCSC_CF_ tiles: the computation-filtering schedule ----------------------------------------------------
 * spmv_stationary's TCSC_CF branch (vertex_program.hpp:1218-1325): iteration 0 adds REG_R x SNK_C, every (non-converged)
 * iteration REG_R x REG_C, the last iteration (fixed count) SRC_R x REG_C and SRC_R x SNK_C — the latter guarded by
 * NC_SRC_R_REG_C (:1302) and walking NC_SRC_R_SNK_C pairs.  apply_stationary's CF branch (:1671-1692): regular rows every
 * iteration, source rows only on the last.  has_converged over the regular rows only (:1902-1916).  In convergence mode
 * combine() does NOTHING after convergence for _TCSC_CF_ (:1036-1043), so the final apply() gives every source row
 * alpha + (1 - alpha) * 0: the "quirky path" of SURVEY.md §8c (fixture: 12 iterations, checksum 51). */
uint32_t gto_pagerank_cf(const gto_graph* g, uint32_t iters, double alpha, double tol, double* rank, uint32_t* deg_out) {
    const uint32_t p = g->p, th = g->th;
    const size_t nall = (size_t) p * th;
    uint32_t* deg = (uint32_t*) malloc(sizeof(uint32_t) * nall);
    gto_degree(g, 1, deg);
    for (uint32_t s = 0; s < p; s++)
        for (uint32_t i = 0; i < th; i++)
            if (!g->rows[s].bits[i]) deg[(size_t) s * th + i] = 0;
    for (size_t v = 0; v < nall; v++) rank[v] = alpha;
    gto_cf_tile** CF = (gto_cf_tile**) calloc((size_t) p * p, sizeof(gto_cf_tile*));
    for (uint32_t t = 0; t < p * p; t++) if (g->tiles[t].nnz) CF[t] = gto_cf_build(g, t / p, t % p);
    uint8_t** cls = (uint8_t**) malloc(sizeof(uint8_t*) * p);
    for (uint32_t s = 0; s < p; s++) { cls[s] = (uint8_t*) malloc(th); gto_classify(g, s, cls[s]); }
    double** X = (double**) malloc(sizeof(double*) * p);
    for (uint32_t s = 0; s < p; s++) X[s] = (double*) malloc(sizeof(double) * (g->cols[s].nnz ? g->cols[s].nnz : 1));
    double** part = (double**) malloc(sizeof(double*) * p);
    double** Ylead = (double**) calloc(p, sizeof(double*));          /* the leader's y of every row group, kept for the final apply() */
    uint32_t it = 0;
    int converged = 0;
    const int check = iters == 0;
    for (;;) {
        for (uint32_t s = 0; s < p; s++)
            for (uint32_t j = 0; j < g->cols[s].nnz; j++) {
                const size_t v = (size_t) s * th + g->cols[s].ids[j];
                X[s][j] = deg[v] ? rank[v] / deg[v] : 0.0;
            }
        const int last = !check && it + 1 == iters;
        uint64_t moving = 0, nreg = 0;
        for (uint32_t rg = 0; rg < p; rg++) {
            const uint32_t nr = g->rows[rg].nnz;
            for (uint32_t r = 0; r < p; r++) part[r] = NULL;
            for (uint32_t cg = 0; cg < p; cg++) {
                const int32_t owner = g->tile_rank[rg * p + cg];
                if (!part[owner]) part[owner] = (double*) calloc(nr ? nr : 1, sizeof(double));
                const gto_cf_tile* T = CF[rg * p + cg];
                if (!T) continue;
                double* y = part[owner];
                const double* x = X[cg];
#define GTO_CF_RUN(kind, count) for (uint32_t j = 0; j < (count); j++) { const uint32_t l = T->JC[kind][j]; \
                    for (uint32_t i = T->JA[kind][2 * j]; i < T->JA[kind][2 * j + 1]; i++) y[T->IA[i]] += x[l]; }
                if (it == 0) GTO_CF_RUN(1, T->NC[1])                              /* :1246-1262 */
                GTO_CF_RUN(0, T->NC[0])                                           /* :1264-1281 */
                if (last) {                                                       /* :1282-1317 */
                    GTO_CF_RUN(2, T->NC[2])
                    if (T->NC[2]) GTO_CF_RUN(3, T->NC[3])
                }
#undef GTO_CF_RUN
            }
            double* y = part[g->leader[rg]];
            for (uint32_t r = 0; r < p; r++)
                if (part[r] && (int32_t) r != g->leader[rg]) { for (uint32_t k = 0; k < nr; k++) y[k] += part[r][k]; }
            for (uint32_t k = 0; k < nr; k++) {                        /* apply_stationary, CF branch :1671-1692 */
                const uint32_t i = g->rows[rg].ids[k];
                const size_t v = (size_t) rg * th + i;
                const int c = cls[rg][i];
                if (c == 1 || (c == 2 && last)) {
                    const double tmp = rank[v];
                    rank[v] = alpha + (1.0 - alpha) * y[k];
                    if (c == 1) { nreg++; if (fabs(rank[v] - tmp) > tol) moving++; }
                }
            }
            free(Ylead[rg]); Ylead[rg] = y;
            for (uint32_t r = 0; r < p; r++) if ((int32_t) r != g->leader[rg]) free(part[r]);
        }
        it++;
        if (check) {
            if (moving == 0) { converged = 1; break; }                 /* :1902-1921 */
        } else if (it >= iters) break;
    }
    if (converged)                                                     /* execute(): combine() (a no-op here) + apply() :425-429 */
        for (uint32_t rg = 0; rg < p; rg++)
            for (uint32_t k = 0; k < g->rows[rg].nnz; k++) {
                const uint32_t i = g->rows[rg].ids[k];
                if (cls[rg][i] == 2) rank[(size_t) rg * th + i] = alpha + (1.0 - alpha) * Ylead[rg][k];   /* y of a source row is still 0 */
            }
    if (deg_out) memcpy(deg_out, deg, sizeof(uint32_t) * nall);
    for (uint32_t s = 0; s < p; s++) { free(X[s]); free(cls[s]); free(Ylead[s]); }
    for (uint32_t t = 0; t < p * p; t++) gto_cf_free(CF[t]);
    free(X); free(part); free(deg); free(CF); free(cls); free(Ylead);
    return it;
}

/* ---- BFS / CC / SSSP: the non-stationary loop (vertex_program.hpp:710-784,1330-1506,1695-1802) ----------------
 * out_a: BFS parent | CC label | SSSP distance; out_b: BFS hops (may be NULL otherwise).  All p*th vertices.
 * Returns `iteration` (includes the final no-change iteration).  sparse_iters counts iterations in which at
 * least one column segment took the (xi,xv) branch under the 0.6 rule (:768-772). */
uint32_t gto_nonstationary(const gto_graph* g, int app, uint32_t root, double ratio, uint32_t* out_a, uint32_t* out_b, uint32_t* sparse_iters) {
    const uint32_t p = g->p, th = g->th;
    const size_t nall = (size_t) p * th;
    const int plus = g->weighted;
    uint8_t* C = (uint8_t*) calloc(nall, 1);
    for (size_t v = 0; v < nall; v++) {
        if (app == GTO_BFS) { out_a[v] = (v == root) ? (uint32_t) v : 0; out_b[v] = (v == root) ? 0 : GTO_INF; C[v] = v == root; }   /* bfs.h:37-49 */
        else if (app == GTO_CC) { out_a[v] = (uint32_t) v; C[v] = 1; }                                                               /* cc.h:32-35 */
        else { out_a[v] = (v == root) ? 0 : GTO_INF; C[v] = v == root; }                                                             /* sssp.h:34-43 */
    }
    uint32_t **X = (uint32_t**) malloc(sizeof(uint32_t*) * p), **XI = (uint32_t**) malloc(sizeof(uint32_t*) * p),
             **XV = (uint32_t**) malloc(sizeof(uint32_t*) * p), **Y = (uint32_t**) malloc(sizeof(uint32_t*) * p);
    uint32_t* K = (uint32_t*) calloc(p, 4);
    for (uint32_t s = 0; s < p; s++) {
        const uint32_t nc = g->cols[s].nnz, nr = g->rows[s].nnz;
        X[s] = (uint32_t*) malloc(4 * (nc ? nc : 1)); XI[s] = (uint32_t*) malloc(4 * (nc ? nc : 1)); XV[s] = (uint32_t*) malloc(4 * (nc ? nc : 1));
        Y[s] = (uint32_t*) malloc(4 * (nr ? nr : 1));
        for (uint32_t k = 0; k < nr; k++) Y[s][k] = GTO_INF;                        /* :625-635; never reset afterwards (:1785) */
    }
    uint32_t it = 0, nsparse = 0;
    for (;;) {
        int any_sparse = 0;
        for (uint32_t s = 0; s < p; s++) {                                           /* scatter_gather_nonstationary :737-751 */
            uint32_t k = 0;
            for (uint32_t j = 0; j < g->cols[s].nnz; j++) {
                const size_t v = (size_t) s * th + g->cols[s].ids[j];
                if (C[v]) {
                    X[s][j] = (app == GTO_BFS) ? (uint32_t) v : out_a[v];            /* bfs.h:52-54, cc.h:37-39, sssp.h:45-47 */
                    XV[s][k] = X[s][j]; XI[s][k] = j; k++;
                } else X[s][j] = GTO_INF;
            }
            K[s] = k;
        }
        for (uint32_t rg = 0; rg < p; rg++)
            for (uint32_t cg = 0; cg < p; cg++) {
                if (!g->tiles[rg * p + cg].nnz) continue;
                const uint32_t nc = g->cols[cg].nnz;
                const int sparse = nc && ((double) K[cg] / nc <= ratio);             /* :768-772, :1475 */
                if (sparse) { any_sparse = 1; gto_tile_spmspv_u32(g, rg, cg, plus, XI[cg], XV[cg], K[cg], Y[rg], NULL); }
                else gto_tile_spmv_u32(g, rg, cg, plus, X[cg], Y[rg], NULL);
            }
        nsparse += any_sparse;
        uint64_t active = 0;
        for (uint32_t s = 0; s < p; s++) {                                           /* apply_nonstationary :1695-1783 */
            if (it == 0)
                for (uint32_t i = 0; i < th; i++) if (!g->rows[s].bits[i]) C[(size_t) s * th + i] = 0;   /* applicator(state) -> false */
            for (uint32_t k = 0; k < g->rows[s].nnz; k++) {
                const size_t v = (size_t) s * th + g->rows[s].ids[k];
                const uint32_t y = Y[s][k];
                int ch = 0;
                if (app == GTO_BFS) { if (out_b[v] == GTO_INF && y != GTO_INF) { out_b[v] = it + 1; out_a[v] = y; ch = 1; } }   /* bfs.h:65-77 */
                else if (app == GTO_CC) { if (y < out_a[v]) { out_a[v] = y; ch = 1; } }                                       /* cc.h:51-55 */
                else { const uint32_t old = out_a[v]; const uint32_t nw = (y < old) ? (plus ? y : y + 1) : old; if (nw != old) { out_a[v] = nw; ch = 1; } }   /* sssp.h:58-66 */
                C[v] = (uint8_t) ch;
                active += ch;
            }
        }
        it++;
        if (!active) break;                                                          /* has_converged :1884-1923 */
    }
    if (sparse_iters) *sparse_iters = nsparse;
    for (uint32_t s = 0; s < p; s++) { free(X[s]); free(XI[s]); free(XV[s]); free(Y[s]); }
    free(X); free(XI); free(XV); free(Y); free(K); free(C);
    return it;
}

/* checksum(): u64 accumulator that truncates after every addition (:1929-1958) */
void gto_checksum_f64(const double* v, uint64_t n_valid, uint64_t* sum, uint64_t* count) {
    uint64_t s = 0, c = 0;
    for (uint64_t i = 0; i < n_valid; i++) if (v[i] != 0.0) { s = (uint64_t) ((double) s + v[i]); c++; }
    *sum = s; *count = c;
}
void gto_checksum_u32(const uint32_t* v, uint64_t n_valid, uint32_t infinity, uint64_t* sum, uint64_t* count) {
    uint64_t s = 0, c = 0;
    for (uint64_t i = 0; i < n_valid; i++) if (v[i] != infinity) { s += v[i]; c++; }
    *sum = s; *count = c;
}

/* ---- counter-based RMAT stream (host copy of graphtap_b200/rmat.py, used to write CPU-baseline samples) */
static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void gto_rmat(uint32_t scale, uint64_t first, uint64_t n, uint64_t seed, int weighted, uint32_t* out) {
    uint64_t mul[3], add[3];
    const uint64_t mask = (1ull << scale) - 1;
    const uint32_t sh = scale / 2 > 1 ? scale / 2 : 1;
    uint64_t k = splitmix64(seed * 0x632BE59BD9B4E019ull + 0x1234567ull);
    for (int r = 0; r < 3; r++) { k = splitmix64(k + (uint64_t) r); mul[r] = k | 1ull; add[r] = k >> 17; }
    const uint64_t base = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull);
    uint64_t zero = (1ull << (scale / 4 > 1 ? scale / 4 : 1)) - 1;
    for (int r = 0; r < 3; r++) { zero = (zero * mul[r] + add[r]) & mask; zero ^= zero >> sh; }
    const int stride = weighted ? 3 : 2;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t ctr = base + (first + i) * 32ull;
        uint64_t s = 0, d = 0, r = 0;
        for (uint32_t lvl = 0; lvl < scale; lvl++) {
            uint32_t u;
            if ((lvl & 1) == 0) { r = splitmix64(ctr + (lvl >> 1)); u = (uint32_t) r; } else u = (uint32_t) (r >> 32);
            s = (s << 1) | (u >= 3264175144u);
            d = (d << 1) | ((u >= 2448131358u && u < 3264175144u) || u >= 4080218931u);
        }
        for (int q = 0; q < 3; q++) { s = (s * mul[q] + add[q]) & mask; s ^= s >> sh; d = (d * mul[q] + add[q]) & mask; d ^= d >> sh; }
        out[i * stride] = (uint32_t) (s ^ zero);
        out[i * stride + 1] = (uint32_t) (d ^ zero);
        if (weighted) out[i * stride + 2] = (uint32_t) ((splitmix64(ctr + 31) >> 33) % 128ull) + 1u;
    }
}
