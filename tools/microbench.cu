// microbench.cu — B200 measurements that size the SpMV design (not part of the product):
// random 8-byte gathers, f64 RED.ADD and u32 RED.MIN against tables of growing size (L2-resident ->
// HBM-resident), with the index stream read coalesced exactly as IA is.  Prints GB/s of index stream
// and Gops/s.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void k_fill_idx(uint32_t* idx, uint64_t n, uint32_t table, int skew) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        uint32_t h = hash32((uint32_t) i * 2654435761u + 12345u);
        if (skew) {   // crude power law: square a uniform to concentrate on low ids
            double u = (h >> 8) * (1.0 / 16777216.0);
            for (int s = 0; s < skew; s++) u *= u;
            idx[i] = (uint32_t) (u * table) % table;
        } else idx[i] = h % table;
    }
}
__device__ __forceinline__ uint4 ld_stream(const uint32_t* p) {
    uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
// mode 0: gather f64 and sum (pull); 1: RED.ADD.F64 (push); 2: RED.MIN.U32; 3: index stream only
template <int MODE>
__global__ void __launch_bounds__(256) k_access(const uint32_t* __restrict__ idx, uint64_t n, double* tab, uint32_t* tab32, double* out) {
    double acc = 0;
    const uint64_t n4 = n / 4;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n4; i += (uint64_t) gridDim.x * blockDim.x) {
        uint4 v = ld_stream(idx + i * 4);
        if (MODE == 0) { acc += __ldg(tab + v.x) + __ldg(tab + v.y) + __ldg(tab + v.z) + __ldg(tab + v.w); }
        else if (MODE == 1) { atomicAdd(tab + v.x, 1.0); atomicAdd(tab + v.y, 1.0); atomicAdd(tab + v.z, 1.0); atomicAdd(tab + v.w, 1.0); }
        else if (MODE == 2) { atomicMin(tab32 + v.x, v.y); atomicMin(tab32 + v.y, v.z); atomicMin(tab32 + v.z, v.w); atomicMin(tab32 + v.w, v.x); }
        else acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 12345.678) out[0] = acc;
}
// shared-memory gather: table slice of `slots` doubles per CTA, indices streamed from HBM
__global__ void __launch_bounds__(1024) k_smem_gather(const uint32_t* __restrict__ idx, uint64_t n, const double* tab, uint32_t slots, double* out) {
    extern __shared__ double sm[];
    for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) sm[i] = tab[i];
    __syncthreads();
    double acc = 0;
    const uint64_t n4 = n / 4;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n4; i += (uint64_t) gridDim.x * blockDim.x) {
        uint4 v = ld_stream(idx + i * 4);
        acc += sm[v.x % slots] + sm[v.y % slots] + sm[v.z % slots] + sm[v.w % slots];
    }
    if (acc == 12345.678) out[0] = acc;
}
int main() {
    const uint64_t n = 1ull << 28;
    uint32_t* idx; double* tab; double* out;
    const uint64_t max_tab = 1ull << 26;           // 512 MB of f64
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&tab, max_tab * 8)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(tab, 0, max_tab * 8));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int sm = 148;
    const char* names[] = {"gather_f64", "red_add_f64", "red_min_u32", "index_only"};
    for (int skew = 0; skew <= 2; skew += 2) {
      for (uint64_t t = 1ull << 14; t <= max_tab; t <<= 2) {
        k_fill_idx<<<sm * 8, 256>>>(idx, n, (uint32_t) t, skew);
        CK(cudaDeviceSynchronize());
        for (int mode = 0; mode < 4; mode++) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(a);
                if (mode == 0) k_access<0><<<sm * 8, 256>>>(idx, n, tab, (uint32_t*) tab, out);
                if (mode == 1) k_access<1><<<sm * 8, 256>>>(idx, n, tab, (uint32_t*) tab, out);
                if (mode == 2) k_access<2><<<sm * 8, 256>>>(idx, n, tab, (uint32_t*) tab, out);
                if (mode == 3) k_access<3><<<sm * 8, 256>>>(idx, n, tab, (uint32_t*) tab, out);
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            printf("skew=%d table=%8.2f MB(f64) %-12s %8.3f ms  %7.1f Gops/s  idx %7.1f GB/s\n", skew, t * 8 / 1048576.0, names[mode], best, n / best * 1e-6, n * 4 / best * 1e-6);
        }
      }
    }
    // smem gather
    CK(cudaFuncSetAttribute(k_smem_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_fill_idx<<<sm * 8, 256>>>(idx, n, 1u << 30, 0);
    for (uint32_t slots : {4096u, 12288u, 24576u}) {
        for (int ctas = 1; ctas <= 2; ctas++) {
            if ((size_t) slots * 8 * ctas > 220 * 1024) continue;
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(a);
                k_smem_gather<<<sm * ctas, 256 * (ctas == 1 ? 4 : 2), slots * 8>>>(idx, n, tab, slots, out);
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            printf("smem_gather slots=%u ctas/sm=%d %8.3f ms %7.1f Gops/s idx %7.1f GB/s\n", slots, ctas, best, n / best * 1e-6, n * 4 / best * 1e-6);
        }
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
