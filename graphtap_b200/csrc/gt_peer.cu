// gt_peer.cu — the x / y exchange of the pull path over NVLink peer memory, without a collective.
//
// The reference moves x along the column group with one Ibcast per segment and partial y along the row
// group with Isend/Irecv to the segment's leader (src/vp/vertex_program.hpp:843-862,970-1013,1083-1108);
// the first version of this library used one ncclAllGather + one ncclReduceScatter per iteration.  Both
// are bulk-synchronous: nothing downstream starts before the whole collective has finished on every
// member, and an NCCL kernel needs SM slots that the persistent SpMV kernel has taken.
//
// Here every member of a group owns a WINDOW (one cudaMalloc, exported with cudaIpcGetMemHandle and mapped
// by the other members).  A producer writes its chunk straight into the consumers' windows with the copy
// engines (cudaMemcpyAsync to the peer mapping: NVLink 5 through NVSwitch, no SM involved) and then
// advances a 32-bit arrival counter in the consumer's window with a 4-byte copy on the same stream, so
// the counter lands after the payload.  A consumer orders its stream behind the counters it needs with a
// one-warp polling kernel placed exactly where the data is first read.  There is no rendezvous: a rank
// that is ahead keeps computing, a transfer overlaps whatever kernel is running on either side, and the
// consumer waits only if the bytes really have not arrived.  Buffers that a fast peer could overwrite
// while a slow one still reads them are double-buffered by epoch parity (gt_engine.cu).
#include "gt_peer.h"

namespace gt {

__global__ void k_iota(uint32_t* p, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}

// thread q polls the counter of group member q (16 bytes apart) until it has reached `value` (modular, gt_peer.h);
// `skip` = this rank
__global__ void k_peer_wait(const uint32_t* flags, int n, int skip, uint32_t value, unsigned long long timeout_ns, uint32_t* err) {
    const int q = threadIdx.x;
    if (q < n && q != skip) {
        const volatile uint32_t* f = flags + 4 * q;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (!peer_reached(*f, value)) {
            __nanosleep(64);
            if ((++spins & 1023u) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (!t0) t0 = now;
                else if (now - t0 > timeout_ns) { atomicCAS(err, 0u, 1u + (uint32_t) q); break; }
            }
        }
    }
    __threadfence_system();
}

static void ensure_ctx_peer_state(gt_ctx* ctx) {
    if (ctx->peer_seq.p) return;
    ctx->peer_seq.alloc(kPeerSeqLen);
    k_iota<<<256, 256, 0, ctx->stream>>>(ctx->peer_seq.p, kPeerSeqLen);
    ctx->kernel_launches++;
    ctx->peer_err.alloc(1);
    GT_CUDA(cudaMemsetAsync(ctx->peer_err.p, 0, 4, ctx->stream));
    ctx->peer_fence.alloc(1);
    GT_CUDA(cudaMemsetAsync(ctx->peer_fence.p, 0, 4, ctx->stream));
    GT_CUDA(cudaGetLastError());
    GT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (const char* e = getenv("GT_PEER_TIMEOUT_MS")) ctx->peer_timeout_ms = std::max(1.0, atof(e));
    // One put stream per destination (round-robin over `peer_lanes` streams): a single stream moves ~450-500 GB/s per
    // direction (profiles/r01_peer_exchange.md), and the three puts of a column group at p = 8 queued behind each other
    // on it; on separate streams they are driven by different copy engines at the same time.
    ctx->peer_lanes = GT_PEER_MAX_LANES;
    if (const char* e = getenv("GT_PEER_LANES")) ctx->peer_lanes = std::min(GT_PEER_MAX_LANES, std::max(1, atoi(e)));
    ctx->put_stream[0] = ctx->comm_stream;
    for (int i = 1; i < GT_PEER_MAX_LANES; i++) GT_CUDA(cudaStreamCreateWithFlags(&ctx->put_stream[i], cudaStreamNonBlocking));
}

uint32_t* peer_error_word(gt_ctx* ctx) { return ctx->peer_err.p; }

static void window_release(PeerWindow* w) {
    for (int q = 0; q < w->size; q++) if (q != w->me && w->remote[q]) cudaIpcCloseMemHandle(w->remote[q]);
    if (w->local) cudaFree(w->local);
    cudaGetLastError();
    w->local = nullptr;
    w->remote.assign(w->size, nullptr);
}

PeerWindow* peer_window_create(gt_ctx* ctx, CommGroup grp, size_t data_bytes) {
    GT_REQUIRE(ctx->comm, "peer window: needs a multi-rank context");
    ensure_ctx_peer_state(ctx);
    cudaStream_t st = ctx->stream;
    std::unique_ptr<PeerWindow> w(new PeerWindow());
    w->grp = grp;
    w->size = comm_size_in(ctx->comm, grp);
    w->me = comm_rank_in(ctx->comm, grp);
    GT_REQUIRE(w->size <= 1024, "peer window: group too large");
    w->data_bytes = (data_bytes + 255) / 256 * 256;
    w->total_bytes = w->data_bytes + 16 * (size_t) w->size;
    w->remote.assign(w->size, nullptr);
    try {
        uint32_t ok = 1;
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof(mine));
        if (cudaMalloc((void**) &w->local, w->total_bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; w->local = nullptr; }
        if (ok) {
            GT_CUDA(cudaMemsetAsync(w->local, 0, w->total_bytes, st));
            if (cudaIpcGetMemHandle(&mine, w->local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
        // handles of all members, through the group's communicator
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        DevBuf<uint8_t> hb; hb.alloc(64 * (size_t) w->size);
        GT_CUDA(cudaMemcpyAsync(hb.p + 64 * (size_t) w->me, &mine, 64, cudaMemcpyHostToDevice, st));
        comm_allgather_inplace(ctx->comm, grp, hb.p, 64, CT_U8, st);
        std::vector<cudaIpcMemHandle_t> all(w->size);
        GT_CUDA(cudaMemcpyAsync(all.data(), hb.p, 64 * (size_t) w->size, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        if (ok) w->remote[w->me] = w->local;
        for (int q = 0; q < w->size && ok; q++) {
            if (q == w->me) continue;
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            w->remote[q] = (uint8_t*) p;
        }
        // all ranks of the job agree (this is also the barrier that orders every member's zero-fill before the first put)
        DevBuf<uint32_t> d_ok; d_ok.alloc(1);
        GT_CUDA(cudaMemcpyAsync(d_ok.p, &ok, 4, cudaMemcpyHostToDevice, st));
        comm_allreduce(ctx->comm, COMM_WORLD, d_ok.p, d_ok.p, 1, CT_U32, CO_MIN, st);
        uint32_t all_ok = 0;
        GT_CUDA(cudaMemcpyAsync(&all_ok, d_ok.p, 4, cudaMemcpyDeviceToHost, st));
        GT_CUDA(cudaStreamSynchronize(st));
        if (!all_ok) {
            window_release(w.get());
            return nullptr;
        }
    } catch (...) {
        window_release(w.get());
        throw;
    }
    return w.release();
}

void peer_window_destroy(gt_ctx* ctx, PeerWindow* w) {
    if (!w) return;
    // No rendezvous here: the engine consumes (waits for) every put before execute() / run_phase() returns, so no
    // transfer targets a window whose owner has reached this point; the mappings are reference-counted by the driver.
    for (int i = 0; i < GT_PEER_MAX_LANES; i++) if (ctx->put_stream[i]) cudaStreamSynchronize(ctx->put_stream[i]);
    cudaStreamSynchronize(ctx->stream);
    window_release(w);
    delete w;
}

void peer_put_begin(gt_ctx* ctx, cudaEvent_t ready) {
    for (int i = 0; i < ctx->peer_lanes; i++) GT_CUDA(cudaStreamWaitEvent(ctx->put_stream[i], ready, 0));
}

void peer_put(gt_ctx* ctx, const PeerWindow* w, int dst_member, size_t dst_offset, const void* src, size_t bytes, uint32_t value, bool advance) {
    GT_REQUIRE(dst_offset + bytes <= w->data_bytes, "peer exchange: put outside the window");
    const int d = (dst_member - w->me - 1 + w->size) % w->size;          // 0 .. size-2: which of my destinations this is
    cudaStream_t s = ctx->put_stream[d % ctx->peer_lanes];
    if (bytes) GT_CUDA(cudaMemcpyAsync(w->remote[dst_member] + dst_offset, src, bytes, cudaMemcpyDefault, s));
    if (advance) GT_CUDA(cudaMemcpyAsync(w->flag(dst_member, w->me), ctx->peer_seq.p + (value & (kPeerSeqLen - 1)), 4, cudaMemcpyDefault, s));
}

void peer_put_end(gt_ctx* ctx, cudaEvent_t* done) {
    if (!done) return;
    for (int i = 0; i < ctx->peer_lanes; i++) GT_CUDA(cudaEventRecord(done[i], ctx->put_stream[i]));
}

void peer_puts_done(gt_ctx* ctx, cudaEvent_t* done, cudaStream_t s) {
    for (int i = 0; i < ctx->peer_lanes; i++) GT_CUDA(cudaStreamWaitEvent(s, done[i], 0));
}

void peer_wait_all(gt_ctx* ctx, const PeerWindow* w, uint32_t value, cudaStream_t s) {
    if (w->size <= 1) return;
    const int n = w->size;
    k_peer_wait<<<1, (n + 31) / 32 * 32, 0, s>>>(w->flag(w->me, 0), n, w->me, value, (unsigned long long) (ctx->peer_timeout_ms * 1e6), ctx->peer_err.p);
    ctx->kernel_launches++;
    GT_CUDA(cudaGetLastError());
}

void peer_fence_world(gt_ctx* ctx, cudaStream_t s) {
    if (!ctx->comm) return;
    ensure_ctx_peer_state(ctx);
    comm_allreduce(ctx->comm, COMM_WORLD, ctx->peer_fence.p, ctx->peer_fence.p, 1, CT_U32, CO_MAX, s);
}

}  // namespace gt
