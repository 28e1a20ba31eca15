// Does st.async (DSMEM store + complete_tx on the receiver's mbarrier) work towards the own CTA and towards a peer?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(2, 1, 1) k(int mode, int* out) {
    __shared__ __align__(16) uint32_t buf[2][128 * 4];
    __shared__ __align__(8) uint64_t bar;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t expect = (mode == 0 ? 1 : mode == 1 ? 1 : 2) * 128 * 16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(expect) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int peer = 0; peer < 2; ++peer) {
        const bool self = peer == (int)rank;
        if ((mode == 0 && self) || (mode == 1 && !self)) continue;     // mode 0: peer only, 1: self only, 2: both
        uint32_t daddr, baddr;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(daddr) : "r"(s32(&buf[rank][threadIdx.x * 4])), "r"(peer));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(baddr) : "r"(s32(&bar)), "r"(peer));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                     ::"r"(daddr), "r"(rank * 1000 + threadIdx.x), "r"(1u), "r"(2u), "r"(3u), "r"(baddr) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (threadIdx.x == 5) out[blockIdx.x] = buf[mode == 1 ? rank : rank ^ 1][5 * 4];
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
    int* out; cudaMalloc(&out, 8);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(out, 0, 8);
        k<<<2, 128>>>(mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        int h[2]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %s  out = %d %d\n", mode, cudaGetErrorString(e), h[0], h[1]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
