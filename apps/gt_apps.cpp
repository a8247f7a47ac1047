// gt_apps.cpp — the reference's four drivers behind one binary, written against include/graphtap/graphtap.hpp.
//
//   gt_apps pr   <file> <nvertices> [iterations]      (src/apps/pr.cpp)
//   gt_apps bfs  <file> <nvertices> [root]            (src/apps/bfs.cpp)
//   gt_apps cc   <file> <nvertices>                   (src/apps/cc.cpp)
//   gt_apps sssp <file> <nvertices> [root]            (src/apps/sssp.cpp; weighted records)
//
// Each branch makes the same API calls, with the same flag values, as the reference driver it names; the
// printed lines ("Execute time", "Iterations", "Value checksum", "Reachable vertices", vertex[i]:...) are
// the ones graphtap.slurm:101-104 greps.  Multi-GPU: launch one process per GPU with RANK / WORLD_SIZE /
// LOCAL_RANK set (torchrun does).  The reference selects weighted builds with -DHAS_WEIGHT at compile
// time; here both instantiations live in one binary.
#include "graphtap/graphtap.hpp"

using ip = uint32_t;

template <typename wp>
static int run(const std::string& app, const std::string& path, ip n, ip arg) {
    const double t0 = Env::clock();
    if (app == "pr") {
        using fp = double;
        Graph<wp, ip, fp> G;
        G.load(path, n, n, true, true, true, false, true, _2DT_, _TCSC_CF_);
        Deg_Program<wp, ip, fp> V(G, true, false, false, _COL_);
        V.execute(1);
        V.checksum();
        PR_Program<wp, ip, fp> VR(G, true, false, false, _ROW_);
        VR.initialize(V);
        V.free();
        VR.execute(arg);
        VR.checksum();
        VR.display();
        VR.free();
        G.free();
    } else {
        using fp = uint32_t;
        const bool bfs = app == "bfs", cc = app == "cc";
        Graph<wp, ip, fp> G;
        // bfs: undirected, no self loops, dedup; cc: undirected, self loops kept, dedup; sssp: directed,
        // transposed for the non-stationary engine, no self loops, dedup
        G.load(path, n, n, !(bfs || cc), !(bfs || cc), cc, false, false, _2DT_, _TCSC_);
        if (bfs) {
            BFS_Program<wp, ip, fp> V(G, false, false, true, _ROW_);
            V.root = arg; V.execute(); V.checksum(); V.display(); V.free();
        } else if (cc) {
            CC_Program<wp, ip, fp> V(G, false, true, false, _ROW_);
            V.execute(); V.checksum(); V.display(); V.free();
        } else {
            SSSP_Program<wp, ip, fp> V(G, false, true, false, _ROW_);
            V.root = arg; V.execute(); V.checksum(); V.display(); V.free();
        }
        G.free();
    }
    Env::print_time(app + " end-to-end", Env::clock() - t0);
    return 0;
}

int main(int argc, char** argv) {
    Env::init();
    if (argc < 4) {
        if (Env::is_master) std::cout << "\"Usage: " << argv[0] << " <pr|bfs|cc|sssp> <file_path> <num_vertices> [<iterations|root>]\"" << std::endl;
        Env::exit(1);
    }
    const std::string app = argv[1], path = argv[2];
    const ip n = (ip) std::atoi(argv[3]);
    const ip arg = argc > 4 ? (ip) std::atoi(argv[4]) : 0;
    int rc = 1;
    if (app == "sssp") rc = run<uint32_t>(app, path, n, arg);
    else if (app == "pr" || app == "bfs" || app == "cc") rc = run<Empty>(app, path, n, arg);
    else if (Env::is_master) std::cerr << "unknown app " << app << std::endl;
    Env::finalize();
    return rc;
}
