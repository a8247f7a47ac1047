// graphtap.hpp — source-compatible C++ face of the reference's public API, on top of the C ABI.
//
// A reference driver (src/apps/{pr,bfs,cc,sssp,deg}.cpp) compiles against this header unchanged apart
// from its include lines: the same `Env::init()`, `Graph<wp,ip,fp>::load(path, n, n,
// directed, transpose, self_loops, acyclic, parallel_edges, _2DT_, _TCSC_)`, `X_Program<wp,ip,fp> V(G,
// stationary, gather_depends_on_apply, apply_depends_on_iter, _ROW_)`, `V.root`, `V.execute([iters])`,
// `VR.initialize(V)`, `V.checksum()`, `V.display()`, `V.free()`, `G.free()`, `Env::finalize()`
// (src/mat/graph.hpp:41-43, src/vp/vertex_program.hpp:27-62, src/mpi/env.hpp:22-55).
//
// What differs by necessity: the messenger/combiner/applicator virtuals cannot run on the device, so
// the five shipped programs are recognised by type and mapped to the library's app enums; deriving a
// new program from Vertex_Program is a compile-time error (static_assert) — there is deliberately no
// CPU fallback.  One process per GPU: rank/size come from RANK / WORLD_SIZE /
// LOCAL_RANK (torchrun or any launcher that sets them) and the NCCL id travels through a file.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <ctime>
#include <unistd.h>
#include "../graphtap_b200.h"

// ---- src/mpi/env.hpp -------------------------------------------------------------------------------
class Env {
  public:
    static int rank, nranks;
    static bool is_master;
    static gt_ctx* ctx;

    static void fail(const char* what) {          // the reference: fprintf(stderr) + Env::exit(1)  (env.hpp:159-162)
        fprintf(stderr, "%s: %s\n", what, gt_last_error());
        std::exit(1);
    }
    static void init(bool = true) {
        const char* r = getenv("RANK");
        const char* n = getenv("WORLD_SIZE");
        const char* l = getenv("LOCAL_RANK");
        rank = r ? atoi(r) : 0;
        nranks = n ? atoi(n) : 1;
        is_master = rank == 0;
        unsigned char id[128] = {0};
        std::string id_path;
        if (nranks > 1) {
            // Rank 0 publishes the 128-byte NCCL id in a 0600 file inside a private 0700 directory; the name carries the
            // launcher's job identity (GT_JOB_ID, else TORCHELASTIC_RUN_ID + MASTER_PORT).  A file older than this
            // process minus the launch skew is a leftover of an earlier job and is never accepted; rank 0 removes the
            // file before publishing and again once the communicator exists (every rank has read it by then).
            id_path = nccl_id_path();
            const time_t born = ::time(nullptr);
            if (rank == 0) {
                ::unlink(id_path.c_str());
                if (gt_nccl_unique_id(id)) fail("gt_nccl_unique_id");
                const std::string tmp = id_path + ".tmp." + std::to_string((long) getpid());
                const int fd = ::open(tmp.c_str(), O_CREAT | O_EXCL | O_WRONLY, 0600);
                if (fd < 0 || ::write(fd, id, 128) != 128) { fprintf(stderr, "graphtap_b200: cannot publish the NCCL id in %s\n", tmp.c_str()); std::exit(1); }
                ::close(fd);
                if (std::rename(tmp.c_str(), id_path.c_str())) { fprintf(stderr, "graphtap_b200: cannot publish the NCCL id in %s\n", id_path.c_str()); std::exit(1); }
            } else {
                bool got = false;
                const double limit_s = getenv("GT_NCCL_ID_TIMEOUT_S") ? atof(getenv("GT_NCCL_ID_TIMEOUT_S")) : 120.0;
                for (double waited = 0; waited < limit_s && !got; waited += 0.01) {
                    struct stat sb;
                    if (::stat(id_path.c_str(), &sb) == 0 && sb.st_size == 128 && sb.st_uid == ::geteuid() && sb.st_mtime + 30 >= born) {
                        std::ifstream f(id_path, std::ios::binary);
                        got = f && f.read((char*) id, 128);
                    }
                    if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(10));
                }
                if (!got) { fprintf(stderr, "graphtap_b200: rank %d never saw the NCCL id of this job in %s\n", rank, id_path.c_str()); std::exit(1); }
            }
        }
        if (gt_ctx_create(l ? atoi(l) : 0, rank, nranks, nranks > 1 ? id : nullptr, &ctx)) fail("gt_ctx_create");
        if (nranks > 1 && rank == 0) ::unlink(id_path.c_str());
    }
    static std::string nccl_id_path() {
        if (const char* f = getenv("GT_NCCL_ID_FILE")) return f;
        std::string dir = std::string(getenv("XDG_RUNTIME_DIR") ? getenv("XDG_RUNTIME_DIR") : "/tmp") + "/graphtap_b200." + std::to_string((long) ::geteuid());
        ::mkdir(dir.c_str(), 0700);
        struct stat sb;
        if (::stat(dir.c_str(), &sb) || !S_ISDIR(sb.st_mode) || sb.st_uid != ::geteuid() || (sb.st_mode & 077)) {
            fprintf(stderr, "graphtap_b200: %s is not a private directory of this user\n", dir.c_str());
            std::exit(1);
        }
        auto env = [](const char* k, const char* d) { const char* v = getenv(k); return std::string(v ? v : d); };
        return dir + "/nccl_id." + env("GT_JOB_ID", env("TORCHELASTIC_RUN_ID", "job").c_str()) + "." + env("MASTER_PORT", "0");
    }
    // Env::barrier (src/mpi/env.hpp:164-166): MPI_Barrier on the world
    static void barrier() { if (ctx && gt_ctx_barrier(ctx)) fail("gt_ctx_barrier"); }
    static void finalize() { if (ctx) { gt_ctx_destroy(ctx); ctx = nullptr; } }
    static void exit(int code) { finalize(); std::exit(code); }
    static double clock() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    static void print_time(std::string preamble, double time) { if (is_master) printf("%s time: %f seconds\n", preamble.c_str(), time); }
    static void print_num(std::string preamble, uint32_t num) { if (is_master) printf("%s %d\n", preamble.c_str(), num); }
};
int Env::rank = 0;
int Env::nranks = 1;
bool Env::is_master = true;
gt_ctx* Env::ctx = nullptr;

// ---- enums (src/mat/tiling.hpp:12-15, src/ds/compressed_column.hpp:17-23, src/vp/vertex_program.hpp:17-21) ----
enum Tiling_type { _2D_, _2DT_ };
enum Compression_type { _CSC_, _DCSC_, _TCSC_, _TCSC_CF_ };
enum Ordering_type { _ROW_, _COL_ };
struct Empty {};                                   // src/ds/triple.hpp:38
struct State { State() {} };                       // src/vp/vertex_program.hpp:15

// ---- src/mat/graph.hpp -------------------------------------------------------------------------------------------
template <typename Weight = char, typename Integer_Type = uint32_t, typename Fractional_Type = float>
class Graph {
  public:
    gt_graph* handle = nullptr;
    static constexpr bool weighted = !std::is_same<Weight, Empty>::value;     // -DHAS_WEIGHT switches wp (src/apps/deg.h:13-17)

    // Graph::load (src/mat/graph.hpp:104-148): the reference sniffs the type with file(1) ("ASCII" -> text, "data" -> binary)
    void load(std::string filepath, Integer_Type nrows, Integer_Type, bool directed = true, bool transpose = false, bool self_loops = true,
              bool acyclic = false, bool parallel_edges = true, Tiling_type tiling = _2DT_, Compression_type compression = _CSC_) {
        double t1 = Env::clock();
        if (looks_like_text(filepath)) load_text(filepath, nrows, nrows, directed, transpose, self_loops, acyclic, parallel_edges, tiling, compression);
        else load_binary(filepath, nrows, nrows, directed, transpose, self_loops, acyclic, parallel_edges, tiling, compression);
        Env::print_time("Ingress", Env::clock() - t1);
    }
    // Graph::parread_binary (src/mat/graph.hpp:307-372): every rank reads its share of the file — equal whole-record shares,
    // the last rank also the remainder (:317-323) — and the library redistributes the entries (gt_graph_build_partitioned).
    void load_binary(std::string filepath, Integer_Type nrows, Integer_Type, bool directed, bool transpose, bool self_loops, bool acyclic,
                     bool parallel_edges, Tiling_type tiling, Compression_type compression) {
        std::ifstream fin(filepath.c_str(), std::ios_base::binary);
        if (!fin.is_open()) { fprintf(stderr, "Unable to open input file\n"); Env::exit(1); }
        fin.seekg(0, std::ios_base::end);
        const uint64_t filesize = (uint64_t) fin.tellg(), rec = weighted ? 12 : 8;
        const uint64_t share = (filesize / (uint64_t) Env::nranks) / rec * rec;
        const uint64_t offset = share * (uint64_t) Env::rank;
        const uint64_t endpos = Env::rank == Env::nranks - 1 ? filesize / rec * rec : offset + share;
        std::vector<uint32_t> triples((endpos - offset) / 4);
        fin.seekg((std::streamoff) offset, std::ios_base::beg);
        fin.read((char*) triples.data(), (std::streamsize) (endpos - offset));
        if ((uint64_t) fin.gcount() != endpos - offset) { fprintf(stderr, "read() failure\n"); Env::exit(1); }
        build(filepath, triples, (endpos - offset) / rec, nrows, directed, transpose, self_loops, acyclic, parallel_edges, tiling, compression, true);
    }
    // text edge lists (src/mat/graph.hpp:194-304): '#'/'%' header lines, then "row col[ weight]" per line
    void load_text(std::string filepath, Integer_Type nrows, Integer_Type, bool directed, bool transpose, bool self_loops, bool acyclic,
                   bool parallel_edges, Tiling_type tiling, Compression_type compression) {
        std::ifstream fin(filepath.c_str());
        if (!fin.is_open()) { fprintf(stderr, "Unable to open input file\n"); Env::exit(1); }
        std::vector<uint32_t> triples;
        std::string line;
        bool started = false;
        const long want = weighted ? 3 : 2;
        while (std::getline(fin, line)) {
            if (!started) { if (line.empty() || line[0] == '#' || line[0] == '%') continue; started = true; }
            if (line.empty()) break;
            if (std::count(line.cbegin(), line.cend(), ' ') + 1 != want) { fprintf(stderr, "read() failure \"%s\"\n", line.c_str()); Env::exit(1); }
            char* end = nullptr;
            const char* p = line.c_str();
            for (long f = 0; f < want; f++) { triples.push_back((uint32_t) std::strtoul(p, &end, 10)); p = end; }
        }
        build(filepath, triples, triples.size() / want, nrows, directed, transpose, self_loops, acyclic, parallel_edges, tiling, compression);
    }
  private:
    static bool looks_like_text(const std::string& path) {
        std::ifstream f(path.c_str(), std::ios_base::binary);
        char buf[4096];
        f.read(buf, sizeof(buf));
        const std::streamsize n = f.gcount();
        for (std::streamsize i = 0; i < n; i++) {
            const unsigned char c = (unsigned char) buf[i];
            if (!(c == '\n' || c == '\r' || c == '\t' || (c >= 32 && c < 127))) return false;
        }
        return n > 0;
    }
    // `share`: `triples` is this rank's share of the records (binary files); else every rank holds the whole list (text files)
    void build(const std::string& filepath, std::vector<uint32_t>& triples, uint64_t n, Integer_Type nrows, bool directed, bool transpose,
               bool self_loops, bool acyclic, bool parallel_edges, Tiling_type tiling, Compression_type compression, bool share = false) {
        if (tiling != _2DT_ || (compression != _TCSC_ && compression != _TCSC_CF_)) {
            fprintf(stderr, "graphtap_b200: only _2DT_ tiling with _TCSC_/_TCSC_CF_ compression runs on the device\n");
            Env::exit(1);
        }
        gt_graph_flags fl = {directed, transpose, self_loops, acyclic, parallel_edges};
        if ((share ? gt_graph_build_partitioned : gt_graph_build)(Env::ctx, triples.data(), n, weighted, 0, nrows, &fl,
                                                                  compression == _TCSC_ ? GT_TCSC : GT_TCSC_CF, &handle))
            Env::fail("gt_graph_build");
        gt_graph_info gi;
        gt_graph_info_get(handle, &gi);
        if (Env::is_master) printf("\n%s: Read %lu edges\n", filepath.c_str(), (unsigned long) gi.nedges_input);
    }
  public:
    void free() { if (handle) { gt_graph_free(handle); handle = nullptr; } }
};

// ---- vertex states (src/apps/{deg,pr,bfs,cc,sssp}.h), same fields, same layout -------------------------------------
#ifndef GT_INF
#define GT_INF 2147483647
#endif
struct Deg_State {
    uint32_t degree = 0;
    uint32_t get_state() { return degree; }
    std::string print_state() { return "Degree=" + std::to_string(degree); }
};
struct PR_State : Deg_State {
    double rank = 0.15;
    double get_state() { return rank; }
    std::string print_state() { return "Rank=" + std::to_string(rank) + ",Degree=" + std::to_string(degree); }
};
struct BFS_State {
    uint32_t parent = 0, hops = GT_INF, vid = 0;
    uint32_t get_state() { return hops; }
    std::string print_state() {
        return "Parent=" + std::to_string(parent) + ",Hops=" + (hops == GT_INF ? std::string("INF") : std::to_string(hops));
    }
};
struct CC_State {
    uint32_t label = 0;
    uint32_t get_state() { return label; }
    std::string print_state() { return "Label=" + std::to_string(label); }
};
struct SSSP_State {
    uint32_t distance = GT_INF;
    uint32_t get_state() { return distance; }
    std::string print_state() { return distance == GT_INF ? std::string("Distance=INF") : "Distance=" + std::to_string(distance); }
};

template <typename S> struct gt_app_of { static constexpr int value = -1; };
template <> struct gt_app_of<Deg_State> { static constexpr int value = GT_APP_DEG; };
template <> struct gt_app_of<PR_State> { static constexpr int value = GT_APP_PR; };
template <> struct gt_app_of<BFS_State> { static constexpr int value = GT_APP_BFS; };
template <> struct gt_app_of<CC_State> { static constexpr int value = GT_APP_CC; };
template <> struct gt_app_of<SSSP_State> { static constexpr int value = GT_APP_SSSP; };

// ---- src/vp/vertex_program.hpp ---------------------------------------------------------------------------------------------
template <typename Weight, typename Integer_Type, typename Fractional_Type, typename Vertex_State>
class Vertex_Program {
    static_assert(gt_app_of<Vertex_State>::value >= 0,
                  "graphtap_b200 runs the five shipped programs (Deg/PR/BFS/CC/SSSP) on the device; user-defined "
                  "messenger/combiner/applicator virtuals have no CPU fallback");
  public:
    Vertex_Program(Graph<Weight, Integer_Type, Fractional_Type>& G, bool stationary_ = false, bool gather_depends_on_apply_ = false,
                   bool apply_depends_on_iter_ = false, Ordering_type ordering_ = _ROW_)
        : graph(&G), stationary(stationary_), gather_depends_on_apply(gather_depends_on_apply_),
          apply_depends_on_iter(apply_depends_on_iter_), ordering(ordering_) {}
    virtual ~Vertex_Program() {}

    Integer_Type root = 0;                  // BFS_Program::root / SSSP_Program::root
    double alpha = 0.15, tol = 1e-5;        // src/apps/pr.h:12-13
    Integer_Type num_iterations = 0, iteration = 0;
    bool stationary, gather_depends_on_apply, apply_depends_on_iter;
    std::vector<Vertex_State> V;            // refreshed from the device after execute()/initialize()
    bool materialize_V = true;              // set false to skip the device->host copy of V after execute()
    double execute_ms = 0;                  // execute_time of the latest execute() (-DTIMING, :436-438)
    gt_program* handle = nullptr;

    void execute(Integer_Type num_iterations_ = 0) {
        num_iterations = num_iterations_;
        ensure();
        uint32_t done = 0;
        if (gt_program_execute(handle, num_iterations_, &done)) Env::fail("gt_program_execute");
        // the reference prints one "Iteration:" line per pass of the loop (:431); the device runs the iterations without
        // a host round trip, so the lines of this execute() come out together, before "Execute time"
        for (Integer_Type it = iteration + 1; it <= (Integer_Type) done; it++) Env::print_num("Iteration: ", it);
        iteration = done;
        gt_timing tm;
        gt_program_timing(handle, &tm);
        execute_ms = tm.execute_ms;
        Env::print_time("Execute", tm.execute_ms * 1e-3);
        if (materialize_V) pull_states();
    }
    void initialize() { ensure(); }
    template <typename W2, typename I2, typename F2, typename S2>
    void initialize(Vertex_Program<W2, I2, F2, S2>& other) {
        ensure();
        if (gt_program_init_from(handle, other.ensure())) Env::fail("gt_program_init_from");
    }
    void checksum() {
        uint64_t sum = 0, cnt = 0;
        if (gt_program_checksum(ensure(), &sum, &cnt)) Env::fail("gt_program_checksum");
        if (Env::is_master) {               // line formats of src/vp/vertex_program.hpp:1942-1958
            std::cout << "Iterations: " << iteration << std::endl;
            std::cout << std::fixed << "Value checksum: " << sum << std::endl;
            std::cout << std::fixed << "Reachable vertices: " << cnt << std::endl;
        }
    }
    void display(Integer_Type count = 31) {
        if (V.empty()) pull_states();
        Env::barrier();
        if (Env::rank) return;
#ifdef TIMING
        {   // the -DTIMING report (src/vp/vertex_program.hpp:2134-2152), all in ms
            double init = 0, st[3][3];
            uint32_t n = 0;
            gt_program_timing_samples(handle, 3, &init, 1, &n);
            for (int ph = 0; ph < 3; ph++) {
                gt_program_timing_samples(handle, ph, nullptr, 0, &n);
                std::vector<double> v(n);
                if (n) gt_program_timing_samples(handle, ph, v.data(), n, &n);
                double sum = 0, sq = 0;
                for (double x : v) { sum += x; sq += x * x; }
                const double mean = sum / v.size();
                st[ph][0] = sum; st[ph][1] = mean; st[ph][2] = std::sqrt(sq / v.size() - mean * mean);     // stats(), :2184-2190
            }
            std::cout << "Init           time: " << init << " ms" << std::endl;
            const char* names[3] = {"Scatter_gather", "Combine       ", "Apply         "};
            for (int ph = 0; ph < 3; ph++)
                std::cout << names[ph] << " time (sum: avg +/- std_dev): " << st[ph][0] << ": " << st[ph][1] << " +/- " << st[ph][2] << " ms" << std::endl;
            std::cout << "Execute        time: " << execute_ms << " ms" << std::endl;
            std::cout << "TIMING " << init;
            for (int ph = 0; ph < 3; ph++) std::cout << " " << st[ph][0] << " " << st[ph][1] << " " << st[ph][2];
            std::cout << " " << execute_ms << std::endl;
        }
#endif
        gt_graph_info gi;
        gt_graph_info_get(graph->handle, &gi);
        const uint64_t base = (uint64_t) gi.layout.owned_segment * gi.layout.tile_height;
        for (uint32_t i = 0; i < std::min<uint64_t>(count, V.size()); i++)
            std::cout << std::fixed << "vertex[" << base + i << "]:" << V[i].print_state() << std::endl;
    }
    void free() {
        V.clear(); V.shrink_to_fit();
        if (handle) { gt_program_free(handle); handle = nullptr; }
    }
    gt_program* ensure() {
        if (!handle) {
            gt_params prm = {alpha, tol, (uint32_t) root};
            if (gt_program_create(graph->handle, gt_app_of<Vertex_State>::value, stationary, gather_depends_on_apply, apply_depends_on_iter,
                                  ordering == _ROW_ ? GT_ROW : GT_COL, &prm, &handle))
                Env::fail("gt_program_create");
#ifdef TIMING
            if (gt_program_set(handle, "timing", 1.0)) Env::fail("gt_program_set");     // per-phase wall clocks, as the reference's -DTIMING build
#endif
        }
        return handle;
    }
    void pull_states() {
        static_assert(sizeof(Deg_State) == 4 && sizeof(PR_State) == 16 && sizeof(BFS_State) == 12, "state layouts must match the reference's");
        gt_graph_info gi;
        gt_graph_info_get(graph->handle, &gi);
        V.resize(gi.layout.tile_height);
        if (gt_program_state_to_host(ensure(), V.data(), V.size() * sizeof(Vertex_State))) Env::fail("gt_program_state_to_host");
    }

  protected:
    Graph<Weight, Integer_Type, Fractional_Type>* graph;
    Ordering_type ordering;
};

template <typename W, typename I, typename F> class Deg_Program : public Vertex_Program<W, I, F, Deg_State> { public: using Vertex_Program<W, I, F, Deg_State>::Vertex_Program; };
template <typename W, typename I, typename F> class PR_Program : public Vertex_Program<W, I, F, PR_State> { public: using Vertex_Program<W, I, F, PR_State>::Vertex_Program; };
template <typename W, typename I, typename F> class BFS_Program : public Vertex_Program<W, I, F, BFS_State> { public: using Vertex_Program<W, I, F, BFS_State>::Vertex_Program; };
template <typename W, typename I, typename F> class CC_Program : public Vertex_Program<W, I, F, CC_State> { public: using Vertex_Program<W, I, F, CC_State>::Vertex_Program; };
template <typename W, typename I, typename F> class SSSP_Program : public Vertex_Program<W, I, F, SSSP_State> { public: using Vertex_Program<W, I, F, SSSP_State>::Vertex_Program; };
