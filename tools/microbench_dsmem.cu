// microbench_dsmem.cu — can a thread-block cluster's distributed shared memory serve as a big x cache?
// Each CTA of a cluster holds `slots` doubles; every thread gathers random entries from the WHOLE cluster's
// table (slots * cluster_size doubles) with ld.shared::cluster, indices streamed from HBM like the SELL index.
// Prints Gops/s.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_dsmem microbench_dsmem.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void k_fill_idx(uint32_t* idx, uint64_t n, uint32_t table) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) idx[i] = hash32((uint32_t) i * 2654435761u + 777u) % table;
}
__device__ __forceinline__ uint4 ld_stream(const uint32_t* p) {
    uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ double ld_dsmem(uint32_t base_addr, uint32_t idx, uint32_t slots) {
    // which CTA of the cluster holds it, and where
    const uint32_t cta = idx / slots, off = idx - cta * slots;
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base_addr + off * 8), "r"(cta));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote));
    return v;
}
__global__ void __launch_bounds__(1024) k_dsmem_gather(const uint32_t* __restrict__ idx, uint64_t n, const double* tab, uint32_t slots, double* out) {
    extern __shared__ double sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) sm[i] = tab[rank * slots + i];
    cluster.sync();
    const uint32_t base = (uint32_t) __cvta_generic_to_shared(sm);
    double acc = 0;
    const uint64_t n4 = n / 4;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n4; i += (uint64_t) gridDim.x * blockDim.x) {
        uint4 v = ld_stream(idx + i * 4);
        acc += ld_dsmem(base, v.x, slots) + ld_dsmem(base, v.y, slots) + ld_dsmem(base, v.z, slots) + ld_dsmem(base, v.w, slots);
    }
    if (acc == 12345.678) out[0] = acc;
    cluster.sync();
}
int main() {
    const uint64_t n = 1ull << 28;
    uint32_t* idx; double* tab; double* out;
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&tab, (1 << 22) * 8)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(tab, 0, (1 << 22) * 8));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaFuncSetAttribute(k_dsmem_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_dsmem_gather, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int csize : {1, 2, 4, 8, 16}) {
        for (uint32_t slots : {8192u, 24576u}) {
            k_fill_idx<<<148 * 8, 256>>>(idx, n, slots * csize);
            CK(cudaDeviceSynchronize());
            cudaLaunchConfig_t cfg = {};
            int grid = 148 / csize * csize;
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = slots * 8;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            float best = 1e30f; bool ok = true;
            for (int rep = 0; rep < 3 && ok; rep++) {
                cudaEventRecord(a);
                cudaError_t e = cudaLaunchKernelEx(&cfg, k_dsmem_gather, (const uint32_t*) idx, n, (const double*) tab, slots, out);
                if (e != cudaSuccess) { printf("cluster=%d slots=%u launch failed: %s\n", csize, slots, cudaGetErrorString(e)); ok = false; cudaGetLastError(); break; }
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            if (ok) printf("cluster=%2d slots/CTA=%5u table=%7.2f MB grid=%d  %8.3f ms  %7.1f Gops/s\n", csize, slots, slots * 8.0 * csize / 1048576, grid, best, n / best * 1e-6);
        }
    }
    printf("done\n");
    return 0;
}
