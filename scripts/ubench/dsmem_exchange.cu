// Cost of one all-to-all exchange round inside an 8-CTA cluster (the decode kernel does 40 per decoder step):
// every CTA pushes `items` 16-byte chunks to each of the 8 CTAs (itself included), then waits until the chunks of all 8 have landed.
//   mode 0: st.async (data + complete_tx on the receiver's mbarrier), all 512 threads, peer-major item order (FFN2 reduce-scatter pattern)
//   mode 1: st.async, items threads x 8 peers each (y3 pattern)
//   mode 2: plain st.shared::cluster.v4, peer-major, then bar.sync + 8 remote mbarrier.arrive.release.cluster (receiver counts 8 arrivals)
//   mode 3: plain st.shared::cluster.v4, then every pushing warp: __syncwarp + lanes 0..7 remote arrive (receiver counts warps x 8)
//   mode 4: st.async, peer-minor order (consecutive lanes -> different peers)
// wait: 0 = warp 0 polls the mbarrier, bar.sync releases the rest; 1 = all 16 warps poll
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_exchange dsmem_exchange.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void st_async(uint32_t addr, uint4 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_plain(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint32_t bar) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory"); }

__global__ void __launch_bounds__(544, 1) k(int mode, int items, int wait_mode, int iters, long long* out, const unsigned char* gsrc, int bg) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ uint64_t rfull[5];
    __shared__ volatile int stopflag;
    __shared__ __align__(16) uint4 src[128];
    __shared__ __align__(16) uint4 dst[8][128];
    __shared__ uint64_t bars[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < 128; i += 544) src[i] = make_uint4(i, rank, 0, 0);
    const bool tx = mode == 0 || mode == 1 || mode == 4;
    const int pushing_warps = mode == 1 ? (items + 31) / 32 : 16;
    const uint32_t count = tx ? 1u : (mode == 2 ? 8u : 8u * (uint32_t)pushing_warps);
    if (tid == 0) {
        stopflag = 0;
        for (int b = 0; b < 5; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&rfull[b])));
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[b])), "r"(count));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tx) for (int b = 0; b < 2; ++b) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[b])), "r"(items * 16 * 8) : "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 16) {
        if (lane == 0 && bg) {                          // background stream: keep 5 x 32 KB copies in flight from an L2-resident region
            const unsigned char* base = gsrc + (size_t)(blockIdx.x % 8) * (6u << 20);
            unsigned n = 0;
            for (int s5 = 0; s5 < 5; ++s5, ++n) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&rfull[s5])), "r"(32768) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + s5 * 32768)), "l"(base + (size_t)(n % 160) * 32768), "r"(32768), "r"(s32(&rfull[s5])) : "memory");
            }
            unsigned i = 0;
            for (;; ++i) {
                const int s5 = i % 5;
                while (!try_wait(&rfull[s5], (i / 5) & 1)) {}
                if (stopflag) break;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&rfull[s5])), "r"(32768) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + s5 * 32768)), "l"(base + (size_t)((n++) % 160) * 32768), "r"(32768), "r"(s32(&rfull[s5])) : "memory");
            }
            for (unsigned r = i + 1; r < i + 5; ++r) while (!try_wait(&rfull[r % 5], (r / 5) & 1)) {}   // drain
            out[8 + blockIdx.x] = n;
        }
        return;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint64_t* bar = &bars[it & 1];
        const uint32_t barl = s32(bar);
        if (mode == 0 || mode == 2 || mode == 3) {
            for (int i = tid; i < items * 8; i += 512) {
                const int peer = i / items, j = i - peer * items;
                const uint4 v = src[j];
                if (mode == 0) st_async(mapa(s32(&dst[rank][j]), peer), v, mapa(barl, peer));
                else st_plain(mapa(s32(&dst[rank][j]), peer), v);
            }
        } else if (mode == 4) {
            for (int i = tid; i < items * 8; i += 512) {
                const int peer = i & 7, j = i >> 3;
                st_async(mapa(s32(&dst[rank][j]), peer), src[j], mapa(barl, peer));
            }
        } else if (tid < items) {
            const uint4 v = src[tid];
#pragma unroll
            for (int peer = 0; peer < 8; ++peer) st_async(mapa(s32(&dst[rank][tid]), peer), v, mapa(barl, peer));
        }
        if (mode == 2) {
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid < 8) arrive_remote(mapa(barl, tid));
        } else if (mode == 3) {
            __syncwarp();
            if (lane < 8) arrive_remote(mapa(barl, lane));
        }
        const uint32_t par = (it >> 1) & 1;
        if (wait_mode == 0) {
            if (warp == 0) {
                while (!try_wait(bar, par)) {}
                if (tx && lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barl), "r"(items * 16 * 8) : "memory");
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
        } else {
            while (!try_wait(bar, par)) {}
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tx && tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barl), "r"(items * 16 * 8) : "memory");
        }
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (tid == 0) stopflag = 1;
    if (dst[0][0].x == 0xdeadbeef) out[1] = 1;
    // keep every CTA alive until all peers are done pushing into it
    for (int b = 0; b < 1; ++b) {}
}
int main() {
    long long* out; cudaMalloc(&out, 2048);
    unsigned char* gsrc; cudaMalloc(&gsrc, 48u << 20); cudaMemset(gsrc, 1, 48u << 20);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 32768);
    printf("mode items wait | cycles/round  us/round (1.965 GHz) | bytes out per CTA\n");
    for (int bg : {0, 1})
    for (int mode : {0, 1})
        for (int items : {20, 40, 80})
            for (int wm : {0}) {
                if (mode == 1 && items > 512) continue;
                cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(8 * 13); cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = 5 * 32768;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                const int iters = 2000;
                cudaLaunchKernelEx(&cfg, k, mode, items, wm, iters, out, (const unsigned char*)gsrc, bg);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[9]; cudaMemcpy(h, out, 72, cudaMemcpyDeviceToHost);
                printf("bg=%d stream %.0f GB/s | %d %3d %d | %8.0f %6.3f | %d %s\n", bg, bg ? (double)h[8] * 32768 / ((double)h[0] / 1.965) : 0.0, mode, items, wm, (double)h[0] / iters, (double)h[0] / iters / 1965.0, items * 16 * 8, e == cudaSuccess ? "" : cudaGetErrorString(e));
                if (e != cudaSuccess) return 1;
            }
    return 0;
}
