// Bisect: which structural feature makes the 544-thread ring 3x slower per chunk than bulk_bw.cu's minimal loop?
// Source: 8 MB region shared by all CTAs (L2 hits), 32 KB x 4 stages, 104 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
struct Cfg { int prod_warp, ncons, warp_loop, chunk, stages, idle_exit, all_poll; };
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(const unsigned char* src, size_t bytes, Cfg c) {
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
    // consumer warps: the first ncons warps that are not the producer warp
    const int cidx = warp < c.prod_warp ? warp : warp - 1;
    if (warp == c.prod_warp) {
        if (c.warp_loop) {
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % c.stages; const uint32_t use = i / c.stages;
                if (use > 0) { if (c.all_poll) { while (!try_wait(&empty[s], (use & 1) ^ 1)) {} } else if (lane == 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {} }
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(c.chunk) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(ring + (size_t)s * c.chunk)), "l"(src + (size_t)(i % nwin) * c.chunk), "r"(c.chunk), "r"(s32(&full[s])) : "memory");
                }
                __syncwarp();
            }
        } else if (lane == 0) {
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % c.stages; const uint32_t use = i / c.stages;
                if (use > 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(c.chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(ring + (size_t)s * c.chunk)), "l"(src + (size_t)(i % nwin) * c.chunk), "r"(c.chunk), "r"(s32(&full[s])) : "memory");
            }
        }
    } else if (cidx < c.ncons) {
        if (c.warp_loop) {
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % c.stages; const uint32_t use = i / c.stages;
                if (c.all_poll) { while (!try_wait(&full[s], use & 1)) {} } else { if (lane == 0) while (!try_wait(&full[s], use & 1)) {} }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
            }
        } else if (lane == 0) {
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % c.stages; const uint32_t use = i / c.stages;
                while (!try_wait(&full[s], use & 1)) {}
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
            }
        }
    } else if (c.idle_exit) return;
    __syncthreads();
}
template <int THREADS>
void run(const unsigned char* src, Cfg c, int ctas) {
    const size_t bytes = 64u << 20;
    cudaFuncSetAttribute(k<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<THREADS><<<ctas, THREADS, 256 + c.chunk * c.stages>>>(src, bytes, c);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("T=%3d ctas=%3d prod_warp=%2d ncons=%2d warp_loop=%d all_poll=%d idle_exit=%d chunk=%2dK x%d : %.3f us/chunk %.1f GB/s per SM %s\n", THREADS, ctas, c.prod_warp, c.ncons,
           c.warp_loop, c.all_poll, c.idle_exit, c.chunk / 1024, c.stages, best * 1e3 / (bytes / c.chunk), bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    unsigned char* src; cudaMalloc(&src, 64u << 20); cudaMemset(src, 1, 64u << 20);
    for (int ctas : {104}) {
        for (int chunk : {32768, 16384}) {
            const int st = chunk == 32768 ? 4 : 8;
            run<128>(src, Cfg{0, 3, 0, chunk, st, 0, 0}, ctas);       // bulk_bw.cu
            run<128>(src, Cfg{3, 3, 0, chunk, st, 0, 0}, ctas);       // producer in the last warp
            run<128>(src, Cfg{0, 3, 1, chunk, st, 0, 0}, ctas);       // warp loops with syncwarp
            run<128>(src, Cfg{0, 3, 1, chunk, st, 0, 1}, ctas);       // ... all lanes poll
            run<544>(src, Cfg{0, 3, 0, chunk, st, 0, 0}, ctas);       // 544 threads, rest parked at the barrier
            run<544>(src, Cfg{0, 3, 0, chunk, st, 1, 0}, ctas);       // 544 threads, rest exit
            run<544>(src, Cfg{16, 3, 0, chunk, st, 0, 0}, ctas);      // producer = warp 16
            run<544>(src, Cfg{16, 3, 1, chunk, st, 0, 0}, ctas);      // pipe_rtt flags=0
            run<544>(src, Cfg{16, 3, 1, chunk, st, 0, 1}, ctas);      // pipe_rtt flags=8-ish
            run<544>(src, Cfg{16, 16, 0, chunk, st, 0, 0}, ctas);     // 16 consumers, lane-0 loops
            run<544>(src, Cfg{16, 16, 1, chunk, st, 0, 1}, ctas);     // decode kernel style
            run<544>(src, Cfg{0, 16, 0, chunk, st, 0, 0}, ctas);      // producer = warp 0, 16 consumers lane-0 loops
        }
    }
    return 0;
}
