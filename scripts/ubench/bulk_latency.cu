// Latency / throughput of cp.async.bulk global -> shared as a function of how many copies one SM keeps in flight.
// W warps per CTA (lane 0 only), each with K private buffers of `size` bytes: wait for a buffer's copy, re-issue it at once
// (no consumer work, no cross-warp hand-off).  Reports the time per copy seen by a warp (= latency when K = 1) and GB/s per SM.
//   src 0: HBM (every warp walks its own region, nothing is re-read)    src 1: L2 (all CTAs walk the same 8 MB)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_latency bulk_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__global__ void __launch_bounds__(544, 1) k(const unsigned char* src, size_t warp_stride, size_t region, int W, int K, int size, int iters, unsigned long long* cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);           // [W][K]
    unsigned char* buf = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < W * K; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= W || lane != 0) return;
    const unsigned char* base = src + ((size_t)blockIdx.x * W + warp) * warp_stride;
    const int per = (int)(region / size);
    auto issue = [&](int kbuf, int i) {
        uint64_t* b = &bars[warp * K + kbuf];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(size) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(buf + (size_t)(warp * K + kbuf) * size)), "l"(base + (size_t)(i % per) * size), "r"(size), "r"(s32(b)) : "memory");
    };
    for (int kb = 0; kb < K; ++kb) issue(kb, kb);
    const long long t0 = clock64();
    for (int i = K; i < iters + K; ++i) {
        const int kb = i % K; const uint32_t use = (uint32_t)(i / K) - 1;
        while (!try_wait(&bars[warp * K + kb], use & 1)) {}
        issue(kb, i);
    }
    const long long t1 = clock64();
    for (int kb = 0; kb < K; ++kb) { const int i = iters + K + kb; const int kk = i % K; while (!try_wait(&bars[warp * K + kk], ((uint32_t)(i / K) - 1) & 1)) {} }
    cyc[blockIdx.x * 17 + warp] = (unsigned long long)(t1 - t0);
}
int main() {
    const size_t total = (size_t)24 << 30;
    unsigned char* src; if (cudaMalloc(&src, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(src, 1, total);
    unsigned long long* cyc; cudaMalloc(&cyc, 148 * 17 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("src ctas W K sizeKB | us per copy per warp | in flight per SM KB | GB/s per SM | TB/s total\n");
    for (int l2 : {0, 1})
        for (int ctas : {8, 104})
            for (int size : {8192, 16384, 32768})
                for (int W : {1, 5, 15})
                    for (int K : {1, 2}) {
                        if ((size_t)W * K * size > 190 * 1024) continue;
                        const int iters = 400;
                        size_t warp_stride = l2 ? 0 : (size_t)14 << 20, region = l2 ? ((size_t)8 << 20) : ((size_t)14 << 20);
                        if (!l2 && (size_t)ctas * W * warp_stride > total) continue;
                        float best = 1e30f; double us_copy = 0;
                        for (int rep = 0; rep < 2; ++rep) {
                            k<<<ctas, 544, 1024 + W * K * size>>>(src, warp_stride, region, W, K, size, iters, cyc);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                            static unsigned long long h[148 * 17]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * ctas * 17, cudaMemcpyDeviceToHost);
                            double mean = 0; for (int c = 0; c < ctas; ++c) for (int w = 0; w < W; ++w) mean += (double)h[c * 17 + w];
                            mean /= (double)ctas * W;
                            us_copy = mean / iters / 1965.0;
                            if (us_copy < best) best = (float)us_copy;
                        }
                        const double gbs = (double)W * size / best / 1e3;
                        printf("%s %3d %2d %d %2d | %6.3f | %4d | %6.1f | %5.2f\n", l2 ? "L2 " : "HBM", ctas, W, K, size / 1024, best, W * K * size / 1024, gbs, gbs * ctas / 1e3);
                    }
    return 0;
}
