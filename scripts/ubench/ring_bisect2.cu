// Bisect 2: producer style x consumer style, independently.  8 MB shared L2 source, 104 CTAs, 544 threads, producer = warp 16,
// 16 consumer warps.
//   pstyle: 0 lane 0 alone in the loop   1 whole warp loops, lane 0 waits/issues, syncwarp   2 whole warp loops, all lanes poll
//   cstyle: 0 lane 0 alone               1 warp: lane 0 try_wait + syncwarp                   2 warp: all lanes try_wait
//           3 warp: all lanes test_wait spin   4 warp: lane 0 try_wait, result broadcast by shfl (no divergent spin)
//           5 warp: all lanes try_wait, then 64 dependent FMAs of "work" per chunk (keeps the warp busy between waits)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ bool test_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
struct Cfg { int pstyle, cstyle, chunk, stages, ncons; };
__global__ void __launch_bounds__(544, 1) k(const unsigned char* src, size_t bytes, Cfg c, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;
    unsigned char* ring = smem + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < c.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(c.ncons));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nchunks = (int)(bytes / c.chunk);
    const int nwin = (8 << 20) / c.chunk;
    auto issue = [&](int i, int s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(c.chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(ring + (size_t)s * c.chunk)), "l"(src + (size_t)(i % nwin) * c.chunk), "r"(c.chunk), "r"(s32(&full[s])) : "memory");
    };
    if (warp == 16) {
        if (c.pstyle == 0) {
            if (lane == 0)
                for (int i = 0; i < nchunks; ++i) {
                    const int s = i % c.stages; const uint32_t use = i / c.stages;
                    if (use > 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {}
                    issue(i, s);
                }
        } else {
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % c.stages; const uint32_t use = i / c.stages;
                if (use > 0) { if (c.pstyle == 2) { while (!try_wait(&empty[s], (use & 1) ^ 1)) {} } else if (lane == 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {} }
                if (lane == 0) issue(i, s);
                __syncwarp();
            }
        }
    } else if (warp < c.ncons) {
        float acc = (float)tid;
        if (c.cstyle == 0) {
            if (lane == 0)
                for (int i = 0; i < nchunks; ++i) {
                    const int s = i % c.stages; const uint32_t use = i / c.stages;
                    while (!try_wait(&full[s], use & 1)) {}
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
                }
        } else {
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % c.stages; const uint32_t use = i / c.stages;
                if (c.cstyle == 1) { if (lane == 0) while (!try_wait(&full[s], use & 1)) {} }
                else if (c.cstyle == 2 || c.cstyle == 5) { while (!try_wait(&full[s], use & 1)) {} }
                else if (c.cstyle == 3) { while (!test_wait(&full[s], use & 1)) {} }
                else { for (;;) { int ok = 0; if (lane == 0) ok = try_wait(&full[s], use & 1); if (__shfl_sync(0xffffffffu, ok, 0)) break; } }
                __syncwarp();
                if (c.cstyle == 5) {
                    const float w = reinterpret_cast<const float*>(ring + (size_t)s * c.chunk)[tid];
#pragma unroll
                    for (int q = 0; q < 64; ++q) acc = fmaf(acc, 1.0001f, w);
                    __syncwarp();
                }
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
            }
        }
        if (acc == 123.456f) sink[tid] = acc;
    }
    __syncthreads();
}
int main() {
    unsigned char* src; cudaMalloc(&src, 64u << 20); cudaMemset(src, 0, 64u << 20);
    float* sink; cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const size_t bytes = 64u << 20;
    for (int chunk : {32768, 16384})
        for (int ps : {0, 1, 2})
            for (int cs : {0, 1, 2, 3, 4, 5}) {
                Cfg c{ps, cs, chunk, chunk == 32768 ? 5 : 10, 16};
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    k<<<104, 544, 256 + c.chunk * c.stages>>>(src, bytes, c, sink);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (rep > 0 && ms < best) best = ms;
                }
                cudaError_t e = cudaGetLastError();
                printf("chunk=%2dK x%2d pstyle=%d cstyle=%d : %.3f us/chunk %6.1f GB/s per SM %s\n", chunk / 1024, c.stages, ps, cs, best * 1e3 / (bytes / chunk), bytes / best / 1e6,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
