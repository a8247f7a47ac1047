/*
 * ref_dump.cpp — driver that runs the UNMODIFIED GraphTap reference (headers included in place from
 * /root/reference/src, never copied) and dumps what its stdout does not show: every vertex state,
 * and on request every TCSC tile array and index map.   TEST INFRASTRUCTURE (oracle/_ref).
 *
 * Built once per app by oracle/ref_build/Makefile with -DAPP_PR / -DAPP_PR1 / -DAPP_DEG / -DAPP_BFS /
 * -DAPP_CC / -DAPP_SSSP (the last with -DHAS_WEIGHT, as the reference Makefile:27-28 does).
 * The per-app graph flags and engine flags below are the literals of the reference mains
 * (src/apps/pr.cpp:26-53, pr1.cpp:26-50, deg.cpp:24-39, bfs.cpp:26-45, cc.cpp:25-44, sssp.cpp:26-44).
 *
 * usage: ref_<app> <file> <nvertices> [iters|root] [--dump PREFIX] [--tiles] [--ct tcsc|tcsc_cf]
 *   PREFIX.r<rank>.V.bin      raw std::vector<Vertex_State> of the owned segment (reference AoS layout)
 *   PREFIX.r<rank>.meta       text: nranks rank owned_segment tile_height nrows iterations state_bytes
 *   PREFIX.r<rank>.t<k>.{JA,IA,A,JC,IR}.bin and PREFIX.r<rank>.{I,IV,J,JV}<k>.bin   with --tiles
 * Graph::load() shells out to file(1), which the image lacks, so the public Graph::load_binary
 * (src/mat/graph.hpp:44) is called directly — same code path after the type sniffing.
 */
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <deque>
#include <algorithm>
#include <numeric>
#include <unordered_set>
#include <set>
#include <cmath>
#include <cstring>
#include <cassert>
#include <type_traits>
#include <unistd.h>
#include <sys/mman.h>
#include <mpi.h>

#define private public
#define protected public
#include "mpi/env.hpp"
#include "mat/graph.hpp"
#if defined(APP_PR) || defined(APP_PR1)
#include "apps/pr.h"
#elif defined(APP_DEG)
#include "apps/deg.h"
#elif defined(APP_BFS)
#include "apps/bfs.h"
#elif defined(APP_CC)
#include "apps/cc.h"
#elif defined(APP_SSSP)
#include "apps/sssp.h"
#elif defined(APP_LAYOUT)
#include "apps/deg.h"
#else
#error "define one of APP_PR APP_PR1 APP_DEG APP_BFS APP_CC APP_SSSP APP_LAYOUT"
#endif
#undef private
#undef protected

static std::string g_prefix;
static bool g_tiles = false;
static int g_repeat = 1;      /* --repeat R: run the timed execute() R times on the loaded graph (bench.py --impl reference) */

template <typename T>
static void write_raw(const std::string& path, const T* p, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    if (n) fwrite(p, sizeof(T), n, f);
    fclose(f);
}

template <typename G>
static void dump_tiles(G& graph) {
    if (g_prefix.empty() || !g_tiles) return;
    auto* A = graph.A;
    std::string base = g_prefix + ".r" + std::to_string(Env::rank);
    uint32_t k = 0;
    for (uint32_t t : A->local_tiles_row_order) {
        auto pair = A->tile_of_local_tile(t);
        auto& tile = A->tiles[pair.row][pair.col];
        std::string tb = base + ".t" + std::to_string(k);
        std::ofstream m(tb + ".meta");
        m << pair.row << " " << pair.col << " " << tile.nedges << " ";
        if (tile.nedges && A->compression_type == _TCSC_) {
            auto* c = static_cast<TCSC_BASE<wp, ip>*>(tile.compressor);
            m << c->nnzcols << " " << c->nnzrows << "\n";
            write_raw(tb + ".JA.bin", c->JA, (size_t) c->nnzcols + 1);
            write_raw(tb + ".IA.bin", c->IA, (size_t) c->nnz);
            write_raw(tb + ".JC.bin", c->JC, (size_t) c->nnzcols);
            write_raw(tb + ".IR.bin", c->IR, (size_t) c->nnzrows);
            #ifdef HAS_WEIGHT
            write_raw(tb + ".A.bin", c->A, (size_t) c->nnz);
            #endif
        } else if (tile.nedges && A->compression_type == _TCSC_CF_) {
            auto* c = static_cast<TCSC_CF_BASE<wp, ip>*>(tile.compressor);
            m << c->nnzcols << " " << c->nnzrows << "\n";
            write_raw(tb + ".JA.bin", c->JA, (size_t) c->nnzcols + 1);
            write_raw(tb + ".IA.bin", c->IA, (size_t) c->nnz);
            write_raw(tb + ".JC.bin", c->JC, (size_t) c->nnzcols);
            write_raw(tb + ".IR.bin", c->IR, (size_t) c->nnzrows);
            /* the four computation-filtering lists (src/ds/compressed_column.hpp:749-1114): NC pairs + NC column ids each */
            std::ofstream cf(tb + ".cf.meta");
            cf << c->NC_REG_R_REG_C << " " << c->NC_REG_R_SNK_C << " " << c->NC_SRC_R_REG_C << " " << c->NC_SRC_R_SNK_C << "\n";
            write_raw(tb + ".cf0.JA.bin", c->JA_REG_R_REG_C, c->NC_REG_R_REG_C ? 2 * (size_t) c->NC_REG_R_REG_C : 0);
            write_raw(tb + ".cf0.JC.bin", c->JC_REG_R_REG_C, c->NC_REG_R_REG_C ? (size_t) c->NC_REG_R_REG_C : 0);
            write_raw(tb + ".cf1.JA.bin", c->JA_REG_R_SNK_C, c->NC_REG_R_SNK_C ? 2 * (size_t) c->NC_REG_R_SNK_C : 0);
            write_raw(tb + ".cf1.JC.bin", c->JC_REG_R_SNK_C, c->NC_REG_R_SNK_C ? (size_t) c->NC_REG_R_SNK_C : 0);
            write_raw(tb + ".cf2.JA.bin", c->JA_SRC_R_REG_C, c->NC_SRC_R_REG_C ? 2 * (size_t) c->NC_SRC_R_REG_C : 0);
            write_raw(tb + ".cf2.JC.bin", c->JC_SRC_R_REG_C, c->NC_SRC_R_REG_C ? (size_t) c->NC_SRC_R_REG_C : 0);
            write_raw(tb + ".cf3.JA.bin", c->JA_SRC_R_SNK_C, c->NC_SRC_R_SNK_C ? 2 * (size_t) c->NC_SRC_R_SNK_C : 0);
            write_raw(tb + ".cf3.JC.bin", c->JC_SRC_R_SNK_C, c->NC_SRC_R_SNK_C ? (size_t) c->NC_SRC_R_SNK_C : 0);
        } else m << "0 0\n";
        k++;
    }
    for (uint32_t i = 0; i < A->I.size(); i++) {
        write_raw(base + ".I" + std::to_string(i) + ".bin", A->I[i].data(), A->I[i].size());
        write_raw(base + ".IV" + std::to_string(i) + ".bin", A->IV[i].data(), A->IV[i].size());
    }
    for (uint32_t i = 0; i < A->J.size(); i++) {
        write_raw(base + ".J" + std::to_string(i) + ".bin", A->J[i].data(), A->J[i].size());
        write_raw(base + ".JV" + std::to_string(i) + ".bin", A->JV[i].data(), A->JV[i].size());
    }
    if (A->compression_type == _TCSC_CF_) {          /* classify_vertices (src/mat/matrix.hpp:1124-1144,853-855): the owned segment's lists */
        write_raw(base + ".regrows.bin", A->rowgrp_regular_rows.data(), A->rowgrp_regular_rows.size());
        write_raw(base + ".srcrows.bin", A->rowgrp_source_rows.data(), A->rowgrp_source_rows.size());
        write_raw(base + ".snkcols.bin", A->colgrp_sink_columns.data(), A->colgrp_sink_columns.size());
    }
    std::ofstream lay(base + ".layout");
    lay << "local_row_segments";
    for (auto v : A->local_row_segments) lay << " " << v;
    lay << "\nlocal_col_segments";
    for (auto v : A->local_col_segments) lay << " " << v;
    lay << "\nleader_ranks";
    for (auto v : A->leader_ranks) lay << " " << v;
    lay << "\nnnz_row_sizes_loc";
    for (auto v : A->nnz_row_sizes_loc) lay << " " << v;
    lay << "\nnnz_col_sizes_loc";
    for (auto v : A->nnz_col_sizes_loc) lay << " " << v;
    lay << "\n";
}

template <typename P>
static void dump_states(P& prog, const char* tag = "") {
    if (g_prefix.empty()) return;
    std::string base = g_prefix + ".r" + std::to_string(Env::rank) + tag;
    write_raw(base + ".V.bin", prog.V.data(), prog.V.size());
    std::ofstream m(base + ".meta");
    m << Env::nranks << " " << Env::rank << " " << prog.owned_segment << " " << prog.tile_height << " "
      << prog.nrows << " " << prog.iteration << " " << sizeof(prog.V[0]) << "\n";
}

int main(int argc, char** argv) {
    Env::init();
    std::vector<std::string> pos;
    Compression_type CT_override = _CSC_;
    bool have_ct = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--dump" && i + 1 < argc) g_prefix = argv[++i];
        else if (a == "--tiles") g_tiles = true;
        else if (a == "--repeat" && i + 1 < argc) g_repeat = atoi(argv[++i]);
        else if (a == "--ct" && i + 1 < argc) {
            std::string c = argv[++i];
            have_ct = true;
            CT_override = (c == "tcsc_cf") ? _TCSC_CF_ : (c == "csc") ? _CSC_ : (c == "dcsc") ? _DCSC_ : _TCSC_;
        }
        else pos.push_back(a);
    }
    if (pos.size() < 2) {
        if (Env::is_master) fprintf(stderr, "usage: %s <file> <nvertices> [iters|root] [--dump PREFIX] [--tiles] [--ct tcsc|tcsc_cf]\n", argv[0]);
        Env::exit(1);
    }
    std::string file_path = pos[0];
    ip num_vertices = std::atoi(pos[1].c_str());
    ip arg3 = (pos.size() > 2) ? (uint32_t) atoi(pos[2].c_str()) : 0;
    Tiling_type TT = _2DT_;
    double t0 = Env::clock();

#if defined(APP_LAYOUT)
    /* Layout only: build the Matrix (src/mat/matrix.hpp:184-202 -> init_matrix :272-495) for the
       (rank, nranks) given by GT_MPI_FAKE_RANK/GT_MPI_FAKE_NRANKS and print every table. */
    {
        auto* A = new Matrix<wp, ip, fp>(num_vertices + 1, num_vertices + 1, Env::nranks * Env::nranks, true, false, true, TT, _TCSC_);
        auto pv = [](const char* name, const std::vector<int32_t>& v) { printf("%s", name); for (auto x : v) printf(" %d", x); printf("\n"); };
        auto pu = [](const char* name, const std::vector<uint32_t>& v) { printf("%s", name); for (auto x : v) printf(" %u", x); printf("\n"); };
        printf("LAYOUT nranks %d rank %d\n", Env::nranks, Env::rank);
        printf("tile_height %u\n", A->tile_height);
        printf("grid %u %u %u %u %u %u\n", A->nrowgrps, A->ncolgrps, A->tiling->rowgrp_nranks, A->tiling->colgrp_nranks, A->tiling->rank_nrowgrps, A->tiling->rank_ncolgrps);
        printf("tile_rank");
        for (uint32_t i = 0; i < A->nrowgrps; i++) for (uint32_t j = 0; j < A->ncolgrps; j++) printf(" %d", A->tiles[i][j].rank);
        printf("\n");
        pv("leader_ranks", A->leader_ranks);
        pv("leader_ranks_rg", A->leader_ranks_rg);
        pv("leader_ranks_cg", A->leader_ranks_cg);
        pu("local_tiles_row_order", A->local_tiles_row_order);
        pu("local_tiles_col_order", A->local_tiles_col_order);
        pv("local_row_segments", A->local_row_segments);
        pv("local_col_segments", A->local_col_segments);
        pv("all_rowgrp_ranks", A->all_rowgrp_ranks);
        pv("all_colgrp_ranks", A->all_colgrp_ranks);
        pv("follower_rowgrp_ranks", A->follower_rowgrp_ranks);
        pv("follower_colgrp_ranks", A->follower_colgrp_ranks);
        pv("follower_rowgrp_ranks_accu_seg", A->follower_rowgrp_ranks_accu_seg);
        pv("follower_rowgrp_ranks_accu_seg_rg", A->follower_rowgrp_ranks_accu_seg_rg);
        pv("follower_rowgrp_ranks_rg", A->follower_rowgrp_ranks_rg);
        pv("follower_colgrp_ranks_cg", A->follower_colgrp_ranks_cg);
        printf("owned_segment %d\n", A->owned_segment);
        printf("accu %d %d %d %d\n", A->accu_segment_rg, A->accu_segment_cg, A->accu_segment_row, A->accu_segment_col);
        printf("rank_rg_cg %d %d\n", Env::rank_rg, Env::rank_cg);
    }
#elif defined(APP_PR)
    /* src/apps/pr.cpp:26-53 */
    Compression_type CT = have_ct ? CT_override : _TCSC_CF_;
    Graph<wp, ip, fp> G;
    G.load_binary(file_path, num_vertices, num_vertices, true, true, true, false, true, TT, CT);
    Env::print_time("Ingress", Env::clock() - t0);
    dump_tiles(G);
    Deg_Program<wp, ip, fp> V(G, true, false, false, _COL_);
    V.execute(1);
    V.checksum();
    dump_states(V, ".deg");
    Env::barrier();
    for (int rep = 1; rep < g_repeat; rep++) {        /* extra timed runs: same calls, fresh program each time */
        PR_Program<wp, ip, fp> W(G, true, false, false, _ROW_);
        W.initialize(V);
        W.execute(arg3);
        W.free();
    }
    PR_Program<wp, ip, fp> VR(G, true, false, false, _ROW_);
    VR.initialize(V);
    V.free();
    VR.execute(arg3);
    VR.checksum();
    VR.display();
    dump_states(VR);
    VR.free();
    G.free();
#elif defined(APP_PR1)
    /* src/apps/pr1.cpp:26-50 */
    Compression_type CT = have_ct ? CT_override : _TCSC_;
    Graph<wp, ip, fp> G;
    G.load_binary(file_path, num_vertices, num_vertices, true, false, true, false, true, TT, CT);
    Deg_Program<wp, ip, fp> V(G, true, false, false, _ROW_);
    V.execute(1);
    V.checksum();
    dump_states(V, ".deg");
    G.free();
    Env::barrier();
    Graph<wp, ip, fp> GR;
    GR.load_binary(file_path, num_vertices, num_vertices, true, true, true, false, true, TT, CT);
    Env::print_time("Ingress", Env::clock() - t0);
    dump_tiles(GR);
    PR_Program<wp, ip, fp> VR(GR, true, false, false, _ROW_);
    VR.initialize(V);
    V.free();
    VR.execute(arg3);
    VR.checksum();
    VR.display();
    dump_states(VR);
    VR.free();
    GR.free();
#elif defined(APP_DEG)
    /* src/apps/deg.cpp:24-39 (Fractional_Type = ip there) */
    Compression_type CT = have_ct ? CT_override : _TCSC_;
    Graph<wp, ip, ip> G;
    G.load_binary(file_path, num_vertices, num_vertices, true, false, true, false, true, TT, CT);
    Env::print_time("Ingress", Env::clock() - t0);
    dump_tiles(G);
    Deg_Program<wp, ip, ip> V(G, true, false, false, _ROW_);
    V.execute(1);
    V.checksum();
    V.display();
    dump_states(V);
    V.free();
    G.free();
#elif defined(APP_BFS)
    /* src/apps/bfs.cpp:26-45 */
    Compression_type CT = have_ct ? CT_override : _TCSC_;
    Graph<wp, ip, fp> G;
    G.load_binary(file_path, num_vertices, num_vertices, false, false, false, false, false, TT, CT);
    Env::print_time("Ingress", Env::clock() - t0);
    dump_tiles(G);
    BFS_Program<wp, ip, fp> V(G, false, false, true, _ROW_);
    V.root = arg3;
    V.execute();
    V.checksum();
    V.display();
    dump_states(V);
    V.free();
    G.free();
#elif defined(APP_CC)
    /* src/apps/cc.cpp:25-44 */
    Compression_type CT = have_ct ? CT_override : _TCSC_;
    Graph<wp, ip, fp> G;
    G.load_binary(file_path, num_vertices, num_vertices, false, false, true, false, false, TT, CT);
    Env::print_time("Ingress", Env::clock() - t0);
    dump_tiles(G);
    CC_Program<wp, ip, fp> V(G, false, true, false, _ROW_);
    V.execute();
    V.checksum();
    V.display();
    dump_states(V);
    V.free();
    G.free();
#elif defined(APP_SSSP)
    /* src/apps/sssp.cpp:26-44: directed, transpose flipped to true for the non-stationary engine */
    Compression_type CT = have_ct ? CT_override : _TCSC_;
    Graph<wp, ip, fp> G;
    G.load_binary(file_path, num_vertices, num_vertices, true, true, false, false, false, TT, CT);
    Env::print_time("Ingress", Env::clock() - t0);
    dump_tiles(G);
    SSSP_Program<wp, ip, fp> V(G, false, true, false, _ROW_);
    V.root = arg3;
    V.execute();
    V.checksum();
    V.display();
    dump_states(V);
    V.free();
    G.free();
#endif
    Env::print_time("end-to-end", Env::clock() - t0);
    Env::finalize();
    return 0;
}
