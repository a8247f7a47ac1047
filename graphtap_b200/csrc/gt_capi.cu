// gt_capi.cu — context, error reporting and raw device-memory entry points of the C ABI.
#include "gt_internal.h"

namespace gt {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
}  // namespace gt

extern "C" const char* gt_last_error(void) { return gt::g_last_error.c_str(); }
extern "C" int gt_abi_version(void) { return GT_ABI_VERSION; }

extern "C" int gt_nccl_unique_id(void* out128) {
    return gt::guarded([&] {
        GT_REQUIRE(out128, "gt_nccl_unique_id: NULL argument");
        gt::comm_unique_id(out128);
    });
}

// Env::init + Env::rowgrps_init/colgrps_init (src/mpi/env.hpp:77-124).  There is no CPU path: without
// an sm_100 device this fails loudly.
extern "C" int gt_ctx_create(int device, int rank, int nranks, const void* nccl_id, gt_ctx** out) {
    return gt::guarded([&] {
        GT_REQUIRE(out, "gt_ctx_create: out is NULL");
        GT_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "gt_ctx_create: bad rank/nranks");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw gt::Error(GT_ERR_NO_DEVICE, std::string("gt_ctx_create: no CUDA device (") + cudaGetErrorString(e) +
                                                  "); libgraphtap_b200 has no CPU fallback");
        GT_REQUIRE(device >= 0 && device < ndev, "gt_ctx_create: device index out of range");
        cudaDeviceProp prop;
        GT_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw gt::Error(GT_ERR_NO_DEVICE, std::string("gt_ctx_create: device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                                  ", this library is built for sm_100a (B200) only");
        GT_CUDA(cudaSetDevice(device));
        std::unique_ptr<gt_ctx> c(new gt_ctx());
        c->device = device; c->rank = rank; c->nranks = nranks;
        c->sm_count = prop.multiProcessorCount;
        GT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        GT_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
        GT_CUDA(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
        GT_CUDA(cudaEventCreateWithFlags(&c->ev_ag, cudaEventDisableTiming));
        if (nranks > 1) {
            // the communicators depend only on (nranks, rank): the group lists are the same for every
            // graph size (src/mat/matrix.hpp:382-465), so any nvertices gives the same lists
            gt::Layout lay = gt::make_layout(1024, nranks, rank);
            c->comm = gt::comm_create(rank, nranks, nccl_id, lay, c->stream);
        }
        *out = c.release();
    });
}

extern "C" int gt_ctx_destroy(gt_ctx* ctx) {
    return gt::guarded([&] {
        if (!ctx) return;
        cudaSetDevice(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
        for (int i = 1; i < GT_PEER_MAX_LANES; i++)
            if (ctx->put_stream[i]) { cudaStreamSynchronize(ctx->put_stream[i]); cudaStreamDestroy(ctx->put_stream[i]); }
        gt::comm_destroy(ctx->comm);
        if (ctx->ev0) { cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); }
        if (ctx->ev_x) { cudaEventDestroy(ctx->ev_x); cudaEventDestroy(ctx->ev_ag); }
        if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

extern "C" int gt_ctx_timer_begin(gt_ctx* ctx) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx, "gt_ctx_timer_begin: NULL ctx");
        GT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->ev0) { GT_CUDA(cudaEventCreate(&ctx->ev0)); GT_CUDA(cudaEventCreate(&ctx->ev1)); }
        GT_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    });
}
extern "C" int gt_ctx_timer_end(gt_ctx* ctx, double* elapsed_ms) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx && ctx->ev0 && elapsed_ms, "gt_ctx_timer_end: no timer running");
        GT_CUDA(cudaSetDevice(ctx->device));
        GT_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        GT_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        GT_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *elapsed_ms = ms;
    });
}

extern "C" int gt_ctx_sync(gt_ctx* ctx) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx, "gt_ctx_sync: NULL ctx");
        GT_CUDA(cudaSetDevice(ctx->device));
        GT_CUDA(cudaStreamSynchronize(ctx->stream));
        GT_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    });
}

extern "C" int gt_ctx_barrier(gt_ctx* ctx) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx, "gt_ctx_barrier: NULL ctx");
        GT_CUDA(cudaSetDevice(ctx->device));
        GT_CUDA(cudaStreamSynchronize(ctx->comm_stream));
        for (int i = 1; i < GT_PEER_MAX_LANES; i++) if (ctx->put_stream[i]) GT_CUDA(cudaStreamSynchronize(ctx->put_stream[i]));
        if (ctx->comm) {
            if (!ctx->barrier_word.p) { ctx->barrier_word.alloc(1); GT_CUDA(cudaMemsetAsync(ctx->barrier_word.p, 0, 4, ctx->stream)); }
            gt::comm_allreduce(ctx->comm, gt::COMM_WORLD, ctx->barrier_word.p, ctx->barrier_word.p, 1, gt::CT_U32, gt::CO_MAX, ctx->stream);
        }
        GT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

extern "C" void* gt_ctx_stream(gt_ctx* ctx) { return ctx ? (void*) ctx->stream : nullptr; }

extern "C" int gt_dev_alloc(gt_ctx* ctx, size_t bytes, void** out) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx && out, "gt_dev_alloc: NULL argument");
        GT_CUDA(cudaSetDevice(ctx->device));
        *out = nullptr;
        if (!bytes) return;
        cudaError_t e = cudaMalloc(out, bytes);
        if (e != cudaSuccess) throw gt::Error(e == cudaErrorMemoryAllocation ? GT_ERR_OOM : GT_ERR_CUDA, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    });
}
extern "C" int gt_dev_free(gt_ctx* ctx, void* p) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx, "gt_dev_free: NULL ctx");
        GT_CUDA(cudaSetDevice(ctx->device));
        if (p) GT_CUDA(cudaFree(p));
    });
}
extern "C" int gt_dev_upload(gt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "gt_dev_upload: NULL argument");
        GT_CUDA(cudaSetDevice(ctx->device));
        if (!bytes) return;
        GT_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        GT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
extern "C" int gt_dev_download(gt_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "gt_dev_download: NULL argument");
        GT_CUDA(cudaSetDevice(ctx->device));
        if (!bytes) return;
        GT_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        GT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
extern "C" int gt_dev_memset(gt_ctx* ctx, void* dst_dev, int byte, size_t bytes) {
    return gt::guarded([&] {
        GT_REQUIRE(ctx && (bytes == 0 || dst_dev), "gt_dev_memset: NULL argument");
        GT_CUDA(cudaSetDevice(ctx->device));
        if (bytes) GT_CUDA(cudaMemsetAsync(dst_dev, byte, bytes, ctx->stream));
    });
}
extern "C" int gt_host_alloc_pinned(size_t bytes, void** out) {
    return gt::guarded([&] {
        GT_REQUIRE(out, "gt_host_alloc_pinned: NULL argument");
        GT_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    });
}
extern "C" int gt_host_free_pinned(void* p) {
    return gt::guarded([&] { if (p) GT_CUDA(cudaFreeHost(p)); });
}
