// Microbenchmark: per-SM streaming bandwidth of cp.async.bulk (global -> shared ring, mbarrier completion)
// as a function of #CTAs, chunk size and ring depth.  One producer thread per CTA, consumers only wait.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__global__ void __launch_bounds__(128, 1) stream_kernel(const unsigned char* src, size_t bytes_per_cta, int chunk, int stages, int l2hit, unsigned long long* out, int nsub) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;
    unsigned char* ring = smem + 256;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 3;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t base = l2hit ? 0 : (size_t)blockIdx.x * bytes_per_cta;   // l2hit: every CTA streams the same region
    const int nchunks = (int)(bytes_per_cta / chunk);
    unsigned long long t0 = clock64();
    if (tid < 32) {                                    // producer warp: lane j issues sub-copy j
        const int sub = chunk / nsub;
        for (int i = 0; i < nchunks; ++i) {
            const int s = i % stages; const uint32_t use = i / stages;
            if (tid == 0) {
                if (use > 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
            }
            __syncwarp();
            if (tid < nsub) {
                // sub-copy j of chunk i comes from a different 'pair' region: stride bytes_per_cta / nsub
                const size_t off = (size_t)tid * (bytes_per_cta / nsub) + (size_t)i * sub;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(ring + (size_t)s * chunk + tid * sub)), "l"(src + base + off), "r"(sub), "r"(s32(&full[s])) : "memory");
            }
        }
    } else if (tid >= 32 && (tid & 31) == 0) {         // 3 consumer warps (lane 0): wait full, release
        for (int i = 0; i < nchunks; ++i) {
            const int s = i % stages; const uint32_t use = i / stages;
            while (!try_wait(&full[s], use & 1)) {}
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t per_cta = 32u << 20;                  // 32 MB per CTA
    unsigned char* src; cudaMalloc(&src, per_cta * 148); cudaMemset(src, 1, per_cta * 148);
    unsigned long long* out; cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("ctas chunkKB stages l2hit | ms  GB/s total  GB/s per SM\n");
    for (int l2hit = 0; l2hit <= 0; ++l2hit)
        for (int ctas : {64, 104, 128})
            for (int chunk : {16384, 32768})
                for (int stages : {4})
                for (int nsub : {1, 2, 8, 16}) {
                    if ((size_t)chunk * stages > 190 * 1024) continue;
                    const size_t bytes = l2hit ? (8u << 20) : per_cta;      // l2hit: 8 MB region re-read by every CTA
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        for (int k = 0; k < (l2hit ? 4 : 1); ++k)
                            stream_kernel<<<ctas, 128, 256 + chunk * stages>>>(src, bytes, chunk, stages, l2hit, out, nsub);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                    }
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    const double tot = (double)bytes * ctas * (l2hit ? 4 : 1);
                    printf("%4d %5d %6d %5d nsub=%2d | %7.3f %9.1f %8.1f\n", ctas, chunk / 1024, stages, l2hit, nsub, ms, tot / ms / 1e6, tot / ms / 1e6 / ctas);
                    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                }
    return 0;
}
