/*
 * mpi.h — minimal MPI replacement used ONLY to build the unmodified GraphTap reference
 * (/root/reference/src) as a test oracle in an image that ships no MPI.
 *
 * TEST INFRASTRUCTURE. Not part of the product; nothing under graphtap_b200/ includes it.
 *
 * Two back ends behind the same 23-function surface the reference uses
 * (Init_thread, Comm_size/rank/group/create/free, Group_incl/free, Barrier, Wtime,
 *  Finalize, Allreduce, Sendrecv, Send, Recv, Isend, Irecv, Ibcast, Wait, Waitall,
 *  Type_contiguous/commit/free):
 *
 *   GT_MPI_NP unset or 1   single rank: collectives are local copies, point-to-point aborts
 *                          (never reached at p = 1: every p2p loop in the reference is guarded by
 *                          `r != rank` or runs over p-1 = 0 followers).
 *   GT_MPI_NP = p > 1      MPI_Init_thread forks p-1 children; ranks talk through per-pair
 *                          single-producer/single-consumer byte rings in one MAP_SHARED region,
 *                          with MPI's non-overtaking matching per (source, tag, communicator),
 *                          unexpected-message buffering, and a progress engine driven from every
 *                          blocking call.  Enough for the reference's ingest + execute loops.
 *
 *   GT_MPI_FAKE_RANK / GT_MPI_FAKE_NRANKS  (single-process) make Comm_rank/size lie, so the
 *                          reference's pure layout code (Matrix::init_matrix) can be dumped for any
 *                          (p, rank) without running p processes.  No communication is possible in
 *                          this mode.
 */
#ifndef GT_ORACLE_MPI_STUB_H
#define GT_ORACLE_MPI_STUB_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <ctime>
#include <vector>
#include <deque>
#include <algorithm>
#include <unistd.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;   /* value = element size in bytes */
typedef int MPI_Op;
typedef int MPI_Request;    /* index into the per-process request table, 0 = null */
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_SUCCESS 0
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_MULTIPLE 3
#define MPI_STATUS_IGNORE ((MPI_Status*) nullptr)
#define MPI_STATUSES_IGNORE ((MPI_Status*) nullptr)
#define MPI_SUM 1
/* datatypes: the handle IS the byte size; Type_contiguous multiplies. */
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_UNSIGNED_LONG 8
#define MPI_DOUBLE 8
/* The only reductions the reference issues are SUM over MPI_UNSIGNED_LONG (u64). */

namespace gtmpi {

static const int MAX_RANKS = 64;
static const int MAX_COMMS = 8;
static const size_t RING_BYTES = 1u << 20;   /* per ordered pair */

struct Ring {                 /* SPSC byte ring, producer = src rank, consumer = dst rank */
    volatile uint64_t head;   /* bytes consumed */
    char pad0[56];
    volatile uint64_t tail;   /* bytes produced */
    char pad1[56];
    char data[RING_BYTES];
};

struct Shared {
    volatile int barrier_count;
    volatile int barrier_sense;
    char pad[56];
    Ring rings[1];            /* [src * np + dst] */
};

struct MsgHeader { int tag; int comm; uint64_t bytes; };

struct Comm {                 /* communicator = ordered list of world ranks */
    bool valid = false;
    std::vector<int> ranks;   /* comm rank -> world rank */
    int my = -1;              /* my rank inside it (-1: not a member) */
};

struct Pending {              /* a posted nonblocking operation */
    bool active = false;
    bool is_send = false;
    bool done = false;
    char* buf = nullptr;
    uint64_t bytes = 0;
    uint64_t moved = 0;       /* send: bytes (header+payload) pushed so far */
    int peer = -1;            /* world rank */
    int tag = 0;
    int comm = 0;
    bool header_sent = false;
};

struct Unexpected { int src; int tag; int comm; std::vector<char> data; };

struct State {
    int np = 1, rank = 0;
    bool fake = false;
    Shared* sh = nullptr;
    std::vector<Comm> comms;
    std::vector<std::vector<int>> groups;
    std::vector<Pending> reqs;              /* index 0 unused */
    std::vector<std::deque<int>> sendq;     /* per destination: request ids in issue order */
    std::vector<Unexpected> unexpected;
    /* partially received message per source */
    struct Rx { bool have_header = false; MsgHeader h; uint64_t got = 0; int req = 0; std::vector<char> tmp; uint64_t hgot = 0; };
    std::vector<Rx> rx;
    int local_sense = 0;
    std::vector<pid_t> children;
};

inline State& S() { static State s; return s; }

inline Ring& ring(int src, int dst) { return S().sh->rings[(size_t) src * S().np + dst]; }

inline void die(const char* what) {
    fprintf(stderr, "[gt mpi stub] rank %d: %s\n", S().rank, what);
    abort();
}

inline int new_request() {
    auto& r = S().reqs;
    if (r.empty()) r.resize(1);
    for (size_t i = 1; i < r.size(); i++)
        if (!r[i].active) { r[i] = Pending(); r[i].active = true; return (int) i; }
    r.emplace_back();
    r.back().active = true;
    return (int) r.size() - 1;
}

/* push as much of the pending sends to `dst` as the ring accepts (in issue order) */
inline bool progress_send(int dst) {
    State& s = S();
    bool moved_any = false;
    auto& q = s.sendq[dst];
    while (!q.empty()) {
        Pending& p = s.reqs[q.front()];
        Ring& rg = ring(s.rank, dst);
        uint64_t total = sizeof(MsgHeader) + p.bytes;
        MsgHeader h{p.tag, p.comm, p.bytes};
        while (p.moved < total) {
            uint64_t head = rg.head, tail = rg.tail;
            uint64_t space = RING_BYTES - (tail - head);
            if (!space) break;
            const char* src; uint64_t avail;
            if (p.moved < sizeof(MsgHeader)) { src = (const char*) &h + p.moved; avail = sizeof(MsgHeader) - p.moved; }
            else { src = p.buf + (p.moved - sizeof(MsgHeader)); avail = total - p.moved; }
            uint64_t n = std::min(avail, space);
            uint64_t off = tail % RING_BYTES;
            n = std::min<uint64_t>(n, RING_BYTES - off);
            memcpy(rg.data + off, src, n);
            __sync_synchronize();
            rg.tail = tail + n;
            p.moved += n;
            moved_any = true;
        }
        if (p.moved < total) break;
        p.done = true;
        q.pop_front();
    }
    return moved_any;
}

/* posted receives in issue order */
inline std::deque<int>& recv_order() { static std::deque<int> q; return q; }

inline bool progress_recv(int src) {
    State& s = S();
    bool moved_any = false;
    Ring& rg = ring(src, s.rank);
    State::Rx& rx = s.rx[src];
    for (;;) {
        uint64_t head = rg.head, tail = rg.tail;
        uint64_t avail = tail - head;
        if (!avail) break;
        __sync_synchronize();
        if (!rx.have_header) {
            uint64_t n = std::min<uint64_t>(avail, sizeof(MsgHeader) - rx.hgot);
            uint64_t off = head % RING_BYTES;
            n = std::min<uint64_t>(n, RING_BYTES - off);
            memcpy((char*) &rx.h + rx.hgot, rg.data + off, n);
            rx.hgot += n;
            __sync_synchronize();
            rg.head = head + n;
            moved_any = true;
            if (rx.hgot < sizeof(MsgHeader)) continue;
            rx.have_header = true;
            rx.got = 0;
            rx.req = 0;
            /* match the earliest posted receive */
            auto& order = recv_order();
            for (auto it = order.begin(); it != order.end(); ++it) {
                Pending& p = s.reqs[*it];
                if (p.peer == src && p.tag == rx.h.tag && p.comm == rx.h.comm) {
                    rx.req = *it;
                    order.erase(it);
                    break;
                }
            }
            if (rx.req) {
                if (s.reqs[rx.req].bytes < rx.h.bytes) die("message longer than posted receive");
            } else {
                rx.tmp.assign(rx.h.bytes, 0);
            }
            if (rx.h.bytes == 0) goto complete;
            continue;
        }
        {
            uint64_t n = std::min<uint64_t>(avail, rx.h.bytes - rx.got);
            uint64_t off = head % RING_BYTES;
            n = std::min<uint64_t>(n, RING_BYTES - off);
            char* dst = rx.req ? s.reqs[rx.req].buf : rx.tmp.data();
            memcpy(dst + rx.got, rg.data + off, n);
            rx.got += n;
            __sync_synchronize();
            rg.head = head + n;
            moved_any = true;
            if (rx.got < rx.h.bytes) continue;
        }
    complete:
        if (rx.req) s.reqs[rx.req].done = true;
        else {
            /* a receive may have been posted while this message was still streaming in */
            int late = 0;
            auto& order = recv_order();
            for (auto it = order.begin(); it != order.end(); ++it) {
                Pending& p = s.reqs[*it];
                if (p.peer == src && p.tag == rx.h.tag && p.comm == rx.h.comm) { late = *it; order.erase(it); break; }
            }
            bool earlier_unexpected = false;   /* keep arrival order among same-envelope messages */
            for (auto& u : s.unexpected)
                if (u.src == src && u.tag == rx.h.tag && u.comm == rx.h.comm) earlier_unexpected = true;
            if (late && !earlier_unexpected) {
                if (s.reqs[late].bytes < rx.h.bytes) die("message longer than posted receive");
                memcpy(s.reqs[late].buf, rx.tmp.data(), rx.h.bytes);
                s.reqs[late].done = true;
                rx.tmp.clear();
            } else {
                if (late) order.push_front(late);
                Unexpected u; u.src = src; u.tag = rx.h.tag; u.comm = rx.h.comm; u.data.swap(rx.tmp);
                s.unexpected.push_back(std::move(u));
            }
        }
        rx.have_header = false; rx.hgot = 0; rx.got = 0; rx.req = 0;
    }
    return moved_any;
}

inline void progress() {
    State& s = S();
    if (s.np <= 1 || s.fake) return;
    bool any = false;
    for (int r = 0; r < s.np; r++) {
        if (r == s.rank) continue;
        any |= progress_send(r);
        any |= progress_recv(r);
    }
    if (!any) sched_yield();
}

inline int post_send(const void* buf, uint64_t bytes, int world_dst, int tag, int comm) {
    State& s = S();
    if (s.np <= 1 || s.fake) die("point-to-point send with a single rank");
    int id = new_request();
    Pending& p = s.reqs[id];
    p.is_send = true; p.buf = (char*) buf; p.bytes = bytes; p.peer = world_dst; p.tag = tag; p.comm = comm;
    s.sendq[world_dst].push_back(id);
    progress_send(world_dst);
    return id;
}

inline int post_recv(void* buf, uint64_t bytes, int world_src, int tag, int comm) {
    State& s = S();
    if (s.np <= 1 || s.fake) die("point-to-point receive with a single rank");
    int id = new_request();
    Pending& p = s.reqs[id];
    p.is_send = false; p.buf = (char*) buf; p.bytes = bytes; p.peer = world_src; p.tag = tag; p.comm = comm;
    /* an already-arrived unexpected message? (earliest first) */
    for (auto it = s.unexpected.begin(); it != s.unexpected.end(); ++it) {
        if (it->src == world_src && it->tag == tag && it->comm == comm) {
            if (it->data.size() > bytes) die("unexpected message longer than receive");
            memcpy(buf, it->data.data(), it->data.size());
            s.unexpected.erase(it);
            p.done = true;
            return id;
        }
    }
    recv_order().push_back(id);
    return id;
}

inline void wait_req(int id) {
    if (id <= 0) return;
    State& s = S();
    while (!s.reqs[id].done) progress();
    s.reqs[id].active = false;
}

inline int world_of(int comm, int r) {
    Comm& c = S().comms[comm];
    if (!c.valid || r < 0 || r >= (int) c.ranks.size()) die("bad communicator rank");
    return c.ranks[r];
}

inline void barrier_all() {
    State& s = S();
    if (s.np <= 1 || s.fake) return;
    s.local_sense = !s.local_sense;
    if (__sync_add_and_fetch(&s.sh->barrier_count, 1) == s.np) {
        s.sh->barrier_count = 0;
        __sync_synchronize();
        s.sh->barrier_sense = s.local_sense;
    } else {
        while (s.sh->barrier_sense != s.local_sense) progress();
    }
}

} // namespace gtmpi

/* ------------------------------------------------------------------------------------------ */

inline int MPI_Init_thread(int*, char***, int required, int* provided) {
    using namespace gtmpi;
    State& s = S();
    if (provided) *provided = required;
    const char* fr = getenv("GT_MPI_FAKE_RANK");
    const char* fn = getenv("GT_MPI_FAKE_NRANKS");
    const char* np = getenv("GT_MPI_NP");
    if (fr && fn) { s.fake = true; s.rank = atoi(fr); s.np = atoi(fn); }
    else if (np && atoi(np) > 1) {
        s.np = atoi(np);
        if (s.np > MAX_RANKS) die("too many ranks");
        size_t bytes = sizeof(Shared) + sizeof(Ring) * ((size_t) s.np * s.np);
        void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) die("mmap of the shared region failed");
        s.sh = (Shared*) m;            /* anonymous shared pages are zero-filled */
        fflush(stdout); fflush(stderr);
        s.rank = 0;
        for (int r = 1; r < s.np; r++) {
            pid_t pid = fork();
            if (pid < 0) die("fork failed");
            if (pid == 0) { s.rank = r; s.children.clear(); break; }
            s.children.push_back(pid);
        }
    }
    s.comms.assign(MAX_COMMS, Comm());
    s.comms[MPI_COMM_WORLD].valid = true;
    for (int r = 0; r < s.np; r++) s.comms[MPI_COMM_WORLD].ranks.push_back(r);
    s.comms[MPI_COMM_WORLD].my = s.rank;
    s.groups.assign(1, std::vector<int>());
    s.reqs.assign(1, Pending());
    s.sendq.assign(s.np, std::deque<int>());
    s.rx.assign(s.np, State::Rx());
    return MPI_SUCCESS;
}

inline int MPI_Finalize() {
    using namespace gtmpi;
    State& s = S();
    if (s.np > 1 && !s.fake) {
        barrier_all();
        fflush(stdout); fflush(stderr);
        if (s.rank != 0) _exit(0);
        for (pid_t c : s.children) { int st; waitpid(c, &st, 0); }
    }
    return MPI_SUCCESS;
}

inline int MPI_Comm_size(MPI_Comm c, int* n) { *n = (int) gtmpi::S().comms[c].ranks.size(); return 0; }
inline int MPI_Comm_rank(MPI_Comm c, int* r) { *r = gtmpi::S().comms[c].my; return 0; }
inline double MPI_Wtime() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
inline int MPI_Barrier(MPI_Comm) { gtmpi::barrier_all(); return 0; }

inline int MPI_Comm_group(MPI_Comm c, MPI_Group* g) {
    auto& s = gtmpi::S();
    s.groups.push_back(s.comms[c].ranks);
    *g = (int) s.groups.size() - 1;
    return 0;
}
inline int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out) {
    auto& s = gtmpi::S();
    std::vector<int> v;
    for (int i = 0; i < n; i++) v.push_back(s.groups[g][ranks[i]]);
    s.groups.push_back(v);
    *out = (int) s.groups.size() - 1;
    return 0;
}
inline int MPI_Group_free(MPI_Group*) { return 0; }
/* Collective over the parent; every rank passes the group IT belongs to (the reference builds
 * disjoint row/col groups this way), so the new handle is simply the next free slot — all ranks
 * call Comm_create the same number of times, which keeps handles aligned across processes. */
inline int MPI_Comm_create(MPI_Comm, MPI_Group g, MPI_Comm* out) {
    auto& s = gtmpi::S();
    int h = 2;
    while (h < gtmpi::MAX_COMMS && s.comms[h].valid) h++;
    if (h == gtmpi::MAX_COMMS) gtmpi::die("out of communicators");
    s.comms[h].valid = true;
    s.comms[h].ranks = s.groups[g];
    s.comms[h].my = -1;
    for (size_t i = 0; i < s.comms[h].ranks.size(); i++)
        if (s.comms[h].ranks[i] == s.rank) s.comms[h].my = (int) i;
    if (s.fake && s.comms[h].my < 0) s.comms[h].my = 0;
    *out = h;
    gtmpi::barrier_all();
    return 0;
}
inline int MPI_Comm_free(MPI_Comm*) { return 0; }

inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype* t) { *t = n * old; return 0; }
inline int MPI_Type_commit(MPI_Datatype*) { return 0; }
inline int MPI_Type_free(MPI_Datatype*) { return 0; }

inline int MPI_Isend(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* rq) {
    *rq = gtmpi::post_send(buf, (uint64_t) n * t, gtmpi::world_of(c, dst), tag, c);
    return 0;
}
inline int MPI_Irecv(void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* rq) {
    *rq = gtmpi::post_recv(buf, (uint64_t) n * t, gtmpi::world_of(c, src), tag, c);
    return 0;
}
inline int MPI_Wait(MPI_Request* rq, MPI_Status*) {
    if (gtmpi::S().np <= 1 || gtmpi::S().fake) return 0;
    gtmpi::wait_req(*rq); *rq = 0; return 0;
}
inline int MPI_Waitall(int n, MPI_Request* rq, MPI_Status*) {
    if (gtmpi::S().np <= 1 || gtmpi::S().fake) return 0;
    for (int i = 0; i < n; i++) { gtmpi::wait_req(rq[i]); rq[i] = 0; }
    return 0;
}
inline int MPI_Send(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    MPI_Request r; MPI_Isend(buf, n, t, dst, tag, c, &r); return MPI_Wait(&r, nullptr);
}
inline int MPI_Recv(void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
    MPI_Request r; MPI_Irecv(buf, n, t, src, tag, c, &r); MPI_Wait(&r, nullptr);
    if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
    return 0;
}
inline int MPI_Sendrecv(const void* sb, int sn, MPI_Datatype stype, int dst, int stag,
                        void* rb, int rn, MPI_Datatype rtype, int src, int rtag, MPI_Comm c, MPI_Status*) {
    MPI_Request r[2];
    MPI_Irecv(rb, rn, rtype, src, rtag, c, &r[0]);
    MPI_Isend(sb, sn, stype, dst, stag, c, &r[1]);
    return MPI_Waitall(2, r, nullptr);
}

/* Broadcast as root-sends-to-all on a reserved tag space; several Ibcasts may be outstanding on one
 * communicator and are matched by issue order (same tag, FIFO per pair) — exactly MPI's rule. */
inline int MPI_Ibcast(void* buf, int n, MPI_Datatype t, int root, MPI_Comm c, MPI_Request* rq) {
    using namespace gtmpi;
    State& s = S();
    *rq = 0;
    if (s.np <= 1 || s.fake) return 0;
    Comm& cm = s.comms[c];
    const int BCAST_TAG = 0x7fff0000;
    if (cm.my == root) {
        /* blocking fan-out keeps the request table simple; payloads are consumed by peers that
           are themselves inside Ibcast/Wait, so this cannot deadlock with the reference's usage */
        std::vector<int> ids;
        for (size_t i = 0; i < cm.ranks.size(); i++)
            if ((int) i != root) ids.push_back(post_send(buf, (uint64_t) n * t, cm.ranks[i], BCAST_TAG, c));
        for (int id : ids) wait_req(id);
    } else {
        *rq = post_recv(buf, (uint64_t) n * t, cm.ranks[root], BCAST_TAG, c);
    }
    return 0;
}

inline int MPI_Allreduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op, MPI_Comm c) {
    using namespace gtmpi;
    State& s = S();
    if (s.np <= 1 || s.fake) { memcpy(out, in, (size_t) n * t); return 0; }
    if (t != 8) die("Allreduce: only 8-byte unsigned sums are implemented (all the reference uses)");
    Comm& cm = s.comms[c];
    const int TAG = 0x7ffe0000;
    std::vector<uint64_t> acc((const uint64_t*) in, (const uint64_t*) in + n), tmp(n);
    if (cm.my == 0) {
        for (size_t i = 1; i < cm.ranks.size(); i++) {
            wait_req(post_recv(tmp.data(), 8ull * n, cm.ranks[i], TAG, c));
            for (int k = 0; k < n; k++) acc[k] += tmp[k];
        }
        std::vector<int> ids;
        for (size_t i = 1; i < cm.ranks.size(); i++) ids.push_back(post_send(acc.data(), 8ull * n, cm.ranks[i], TAG + 1, c));
        for (int id : ids) wait_req(id);
    } else {
        wait_req(post_send(acc.data(), 8ull * n, cm.ranks[0], TAG, c));
        wait_req(post_recv(acc.data(), 8ull * n, cm.ranks[0], TAG + 1, c));
    }
    memcpy(out, acc.data(), 8ull * n);
    return 0;
}

#endif
