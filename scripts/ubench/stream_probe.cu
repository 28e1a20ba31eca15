// Microbenchmark behind the round-2 decode kernel (profiles/r02_stream_probe.md):
//   (1) how many 4 / 8 / 16-CTA clusters of a 1-CTA-per-SM kernel (227 KB shared memory) can be co-resident on this GPU;
//   (2) per-SM streaming rate of global -> shared bulk copies as a function of the PIECE size (one mbarrier-tracked stage of
//       32 KB is filled by 32 KB / piece separate cp.async.bulk copies taken from regions `gap` bytes apart, like the K/V
//       rows of different (utterance, head) pairs), for HBM-resident data;
//   (3) the same for the decode kernel's weight stream: every "rank" r = cta % 8 re-reads its own 5.45 MB slice (L2 hits),
//       16 consumer warps each arriving on the stage's empty barrier (as the decode kernel's consumers do).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
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
__device__ __forceinline__ uint64_t policy(int kind) {
    uint64_t p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct Args {
    const unsigned char* src;
    size_t region;        // bytes streamed per pass by one CTA
    size_t cta_stride;    // distance between the regions of CTAs (0: shared by all CTAs of the same rank)
    int rank_mod;         // region index = cta % rank_mod when cta_stride == 0 semantics are wanted via rank regions
    int stage, stages, piece;
    size_t gap;           // distance between the source regions of the pieces of one stage
    int passes, hint;     // hint: 0 none, 1 evict_last, 2 evict_first
    int consumers;        // consumer warps that must arrive per stage
};

__global__ void __launch_bounds__(544, 1) stream_kernel(Args a) {
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
    const int npieces = a.stage / a.piece;
    // a pass streams `region` bytes: piece j of stage i comes from base + j * gap + i * piece
    const int nstages_pass = (int)(a.region / a.stage);
    const int total = nstages_pass * a.passes;
    if (warp == 16) {
        if (lane == 0) {
            const uint64_t pol = policy(a.hint);
            for (int i = 0; i < total; ++i) {
                const int s = i % a.stages; const uint32_t use = i / a.stages;
                if (use > 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(a.stage) : "memory");
                const int ip = i % nstages_pass;
                for (int j = 0; j < npieces; ++j) {
                    const unsigned char* src = a.src + base + (size_t)j * a.gap + (size_t)ip * a.piece;
                    unsigned char* dst = ring + (size_t)s * a.stage + (size_t)j * a.piece;
                    if (a.hint)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                     ::"r"(s32(dst)), "l"(src), "r"(a.piece), "r"(s32(&full[s])), "l"(pol) : "memory");
                    else
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(s32(dst)), "l"(src), "r"(a.piece), "r"(s32(&full[s])) : "memory");
                }
            }
        }
    } else if (warp < a.consumers) {
        for (int i = 0; i < total; ++i) {
            const int s = i % a.stages; const uint32_t use = i / a.stages;
            while (!try_wait(&full[s], use & 1)) {}
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
}

__global__ void __launch_bounds__(544, 1) occ_kernel(int* p) { extern __shared__ unsigned char sm[]; if (p && threadIdx.x == 9999) p[0] = sm[0]; }

static float run(const Args& a, int ctas, int smem) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<ctas, 544, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); exit(1); }
    return best;
}

int main() {
    // ---- (1) cluster co-residency
    cudaFuncSetAttribute(occ_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(occ_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = 227 * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, occ_kernel, &cfg);
        printf("cluster_size %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
        cudaGetLastError();
    }
    const size_t total = (size_t)6 << 30;
    unsigned char* src; if (cudaMalloc(&src, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(src, 1, total);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // ---- (2) HBM-resident pieces: each CTA owns 32 MB, split into stage/piece sub-regions `gap` apart
    printf("\nHBM stream, 32 KB stages x 5, piece-size sweep (pieces of one stage come from regions 1 MB apart)\n");
    printf("ctas piece_KB | ms | TB/s total | GB/s per SM\n");
    for (int ctas : {32, 104, 128, 148})
        for (int piece : {2048, 4096, 8192, 16384, 32768}) {
            Args a{}; a.src = src; a.region = 16u << 20; a.cta_stride = 32u << 20; a.rank_mod = 0;
            a.stage = 32768; a.stages = 5; a.piece = piece; a.gap = 1u << 20; a.passes = 1; a.hint = 2; a.consumers = 16;
            const float ms = run(a, ctas, 256 + a.stage * a.stages);
            const double tot = (double)a.region * ctas;
            printf("%4d %5d | %7.3f | %6.2f | %6.1f\n", ctas, piece / 1024, ms, tot / ms / 1e9, tot / ms / 1e6 / ctas);
        }
    printf("\nHBM stream, 16 KB stages x 10, single piece\n");
    for (int ctas : {104, 128}) {
        Args a{}; a.src = src; a.region = 16u << 20; a.cta_stride = 32u << 20; a.stage = 16384; a.stages = 10; a.piece = 16384; a.gap = 0; a.passes = 1; a.hint = 2; a.consumers = 16;
        const float ms = run(a, ctas, 256 + a.stage * a.stages);
        const double tot = (double)a.region * ctas;
        printf("%4d 16 | %7.3f | %6.2f | %6.1f\n", ctas, ms, tot / ms / 1e9, tot / ms / 1e6 / ctas);
    }
    // ---- (3) weight stream: rank r = cta % 8 re-reads its own 5.45 MB slice, 40 passes
    printf("\nL2 weight stream: 8 rank slices of 5.45 MB re-read by every cluster, 5 stages\n");
    printf("ctas stage_KB hint consumers | ms | TB/s total | GB/s per SM\n");
    for (int ctas : {104, 128})
        for (int stage : {16384, 32768})
            for (int hint : {0, 1})
                for (int cons : {1, 16}) {
                    Args a{}; a.src = src; a.region = (size_t)5570560 / stage * stage; a.cta_stride = 6u << 20; a.rank_mod = 8;
                    a.stage = stage; a.stages = stage == 16384 ? 10 : 5; a.piece = stage; a.gap = 0; a.passes = 40; a.hint = hint; a.consumers = cons;
                    const float ms = run(a, ctas, 256 + a.stage * a.stages);
                    const double tot = (double)a.region * a.passes * ctas;
                    printf("%4d %5d %d %2d | %7.3f | %6.2f | %6.1f\n", ctas, stage / 1024, hint, cons, ms, tot / ms / 1e9, tot / ms / 1e6 / ctas);
                }
    return 0;
}
