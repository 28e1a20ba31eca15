// Which feature of the decode kernel's ring makes a chunk cost ~0.5 us?  Toggle them one by one.
// L2-resident source (8 MB region re-read), 32 KB x 4 stages, 104 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
// flags: 1 = 16 consumer warps, 2 = cache-hint policy, 4 = producer polls with shuffles, 8 = all lanes poll full barrier
__global__ void __launch_bounds__(544, 1) pipe_kernel(const unsigned char* src, size_t bytes, int chunk, int stages, int flags) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;
    unsigned char* ring = smem + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = (flags & 1) ? 16 : 3;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(ncons));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nchunks = (int)(bytes / chunk);
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (warp == 16) {
        for (int i = 0; i < nchunks; ++i) {
            const int s = i % stages; const uint32_t use = i / stages;
            if (use > 0) {
                if (flags & 4) {
                    for (;;) {
                        int ok = 0;
                        if (lane == 0) ok = try_wait(&empty[s], (use & 1) ^ 1);
                        ok = __shfl_sync(0xffffffffu, ok, 0);
                        if (ok) break;
                    }
                } else if (lane == 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {}
            }
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
                if (flags & 2)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                 ::"r"(s32(ring + (size_t)s * chunk)), "l"(src + (size_t)(i % 256) * chunk), "r"(chunk), "r"(s32(&full[s])), "l"(pol) : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(ring + (size_t)s * chunk)), "l"(src + (size_t)(i % 256) * chunk), "r"(chunk), "r"(s32(&full[s])) : "memory");
            }
            __syncwarp();
        }
    } else if (warp < ncons) {
        for (int i = 0; i < nchunks; ++i) {
            const int s = i % stages; const uint32_t use = i / stages;
            if (flags & 8) { while (!try_wait(&full[s], use & 1)) {} }
            else { if (lane == 0) while (!try_wait(&full[s], use & 1)) {} __syncwarp(); }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
    __syncthreads();
}
int main() {
    unsigned char* src; cudaMalloc(&src, 64u << 20); cudaMemset(src, 1, 64u << 20);
    cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    const size_t bytes = 64u << 20;   // streamed per CTA (from an 8 MB window)
    for (int cluster : {1, 8})
        for (int flags : {0, 1, 2, 4, 8, 1 | 2 | 4, 1 | 2 | 4 | 8}) {
            for (int ctas : {104}) {
                cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = 256 + 32768 * 4;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    cudaLaunchKernelEx(&cfg, pipe_kernel, (const unsigned char*)src, bytes, 32768, 4, flags);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                cudaError_t e = cudaGetLastError();
                printf("cluster=%d flags=%2d ctas=%d: %.3f ms  %.1f GB/s per SM  %.3f us/chunk  %s\n", cluster, flags, ctas, ms,
                       bytes / ms / 1e6, ms * 1e3 / (bytes / 32768), e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    return 0;
}
