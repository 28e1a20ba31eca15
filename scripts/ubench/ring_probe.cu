// What bounds the per-chunk cost of the decode kernel's shared-memory ring (producer warp -> cp.async.bulk -> full mbarrier ->
// 16 consumer warps -> empty mbarrier -> producer)?  Realistic configuration (544 threads, 1 CTA / SM, 104 or 120 CTAs, L2- or
// HBM-resident source, 32 KB x 5 stages), wait primitives toggled one by one, plus per-chunk latency stamps of one CTA.
//   cwait: how consumers wait on `full`   0 try_wait all lanes   1 try_wait lane 0 + syncwarp   2 test_wait spin all lanes
//                                          3 test_wait spin lane 0 + syncwarp                     4 try_wait with a 20 ns suspend hint
//   pwait: how the producer waits on `empty` (same codes; lane 0 = the issuing lane)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_probe ring_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ bool try_wait_hint(uint64_t* b, uint32_t par, uint32_t ns) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par), "r"(ns) : "memory");
    return ok;
}
__device__ __forceinline__ bool test_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void wait_mode(int mode, uint64_t* b, uint32_t par, int lane) {
    switch (mode) {
    case 0: while (!try_wait(b, par)) {} break;
    case 1: if (lane == 0) while (!try_wait(b, par)) {} __syncwarp(); break;
    case 2: while (!test_wait(b, par)) {} break;
    case 3: if (lane == 0) while (!test_wait(b, par)) {} __syncwarp(); break;
    default: while (!try_wait_hint(b, par, 20)) {} break;
    }
}

struct Args {
    const unsigned char* src; size_t region, cta_stride; int rank_mod, stage, stages, nchunks, cwait, pwait, consumers;
    unsigned long long* stamps;     // [nchunks][3] of CTA 0: issue, consumer-0 wake, producer sees empty (globaltimer ns)
};

__global__ void __launch_bounds__(544, 1) ring_kernel(Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;
    unsigned char* ring = smem + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(a.consumers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t base = (size_t)(a.rank_mod ? blockIdx.x % a.rank_mod : blockIdx.x) * a.cta_stride;
    const int per_pass = (int)(a.region / a.stage);
    const bool stamper = a.stamps && blockIdx.x == 0;
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (warp == 16) {
        for (int i = 0; i < a.nchunks; ++i) {
            const int s = i % a.stages; const uint32_t use = i / a.stages;
            if (use > 0) {
                if (a.pwait == 0 || a.pwait == 2 || a.pwait == 4) wait_mode(a.pwait, &empty[s], (use & 1) ^ 1, lane);   // all lanes
                else wait_mode(a.pwait, &empty[s], (use & 1) ^ 1, lane);
            }
            if (lane == 0) {
                if (stamper && use > 0) a.stamps[(size_t)(i - a.stages) * 3 + 2] = now();
                if (stamper) a.stamps[(size_t)i * 3] = now();
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(a.stage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(ring + (size_t)s * a.stage)), "l"(a.src + base + (size_t)(i % per_pass) * a.stage), "r"(a.stage), "r"(s32(&full[s])) : "memory");
            }
            __syncwarp();
        }
    } else if (warp < a.consumers) {
        for (int i = 0; i < a.nchunks; ++i) {
            const int s = i % a.stages; const uint32_t use = i / a.stages;
            wait_mode(a.cwait, &full[s], use & 1, lane);
            if (stamper && tid == 0) a.stamps[(size_t)i * 3 + 1] = now();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
}

int main() {
    const size_t total = (size_t)5 << 30;
    unsigned char* src; if (cudaMalloc(&src, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(src, 1, total);
    unsigned long long* stamps; cudaMalloc(&stamps, 4096 * 3 * 8);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("src ctas stageKB stages cons cwait pwait | us/chunk GB/s per SM | CTA0 ns: issue->wake  wake->producer-sees-empty  issue->issue\n");
    for (int l2 : {1, 0})
        for (int ctas : {104, 120})
            for (int stage : {32768, 16384})
                for (int cons : {16, 1})
                    for (int cw : {0, 1, 2, 3, 4})
                        for (int pw : {0, 2, 3}) {
                            if (cons == 1 && !(cw == 0 || cw == 2)) continue;
                            if (stage == 16384 && !(cw == 0 || cw == 2)) continue;
                            if (ctas == 120 && !(cw == 0 || cw == 2)) continue;
                            Args a{}; a.src = src; a.stage = stage; a.stages = stage == 32768 ? 5 : 10; a.nchunks = l2 ? 4096 : 1024;
                            a.consumers = cons; a.cwait = cw; a.pwait = pw; a.stamps = stamps;
                            if (l2) { a.region = (size_t)5570560 / stage * stage; a.cta_stride = 6u << 20; a.rank_mod = 8; }
                            else { a.region = 32u << 20; a.cta_stride = 32u << 20; a.rank_mod = 0; }
                            cudaMemset(stamps, 0, 4096 * 3 * 8);
                            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                            float best = 1e30f;
                            for (int rep = 0; rep < 3; ++rep) {
                                cudaEventRecord(e0);
                                ring_kernel<<<ctas, 544, 256 + a.stage * a.stages>>>(a);
                                cudaEventRecord(e1); cudaEventSynchronize(e1);
                                float ms; cudaEventElapsedTime(&ms, e0, e1);
                                if (rep > 0 && ms < best) best = ms;
                            }
                            cudaError_t e = cudaGetLastError();
                            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                            static unsigned long long h[4096 * 3];
                            cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost);
                            double d1 = 0, d2 = 0, d3 = 0; int n = 0;
                            for (int i = a.nchunks / 2; i < a.nchunks - 2 * a.stages; ++i) {
                                d1 += (double)(h[i * 3 + 1] - h[i * 3]); d2 += (double)(h[i * 3 + 2] - h[i * 3 + 1]); d3 += (double)(h[(i + 1) * 3] - h[i * 3]); ++n;
                            }
                            const double us = best * 1e3 / a.nchunks;
                            printf("%s %4d %3d %2d %2d %d %d | %6.3f %7.1f | %7.0f %7.0f %7.0f\n", l2 ? "L2 " : "HBM", ctas, stage / 1024, a.stages, cons, cw, pw, us,
                                   a.stage / us / 1e3, d1 / n, d2 / n, d3 / n);
                        }
    return 0;
}
