// Gate for a tcgen05 version of the decode GEMMs: weights stream through a 5 x 32 KB shared-memory ring (bulk copies of tiles
// that are ALREADY in the 128-byte-swizzled K-major UMMA layout in global memory), one thread issues tcgen05.mma
// (M = 128 weight rows, N = 16 batch columns, K = 16) straight from the ring, tcgen05.commit hands the slot back: no LDS of the
// weights, no ldmatrix of the activations.  Reports GB/s per SM (104 CTAs, rank slices re-read from L2) and checks tile 0
// against a host reference; also dumps where the rows of an M = 64 accumulator land in TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_stream umma_stream.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
// byte offset of element (r, k) in a K-major SWIZZLE_128B block [rows][64 k]
__host__ __device__ inline int sw128(int r, int k) { return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2; }

constexpr int STAGES = 5, STAGE = 32768;
// M = 128: tile = [128 rows][512 k] = 4 stages of [128][128 k] (two 16 KB blocks).  M = 64: tile = [64][512 k] = 2 stages of [64][256 k] (four 8 KB blocks)
__global__ void __launch_bounds__(192, 1) k(const unsigned char* src, size_t slice_bytes, size_t slice_stride, int ntiles, int M, const bf16* act, float* dump, unsigned* sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;                          // 5 x 32 KB
    unsigned char* bact = smem + STAGES * STAGE;         // [8 k-blocks][16 rows][128 B] = 16 KB
    uint64_t* full = reinterpret_cast<uint64_t*>(bact + 16384);
    uint64_t* empty = full + STAGES;
    uint64_t* accf = empty + STAGES;                     // [2] accumulator full
    uint64_t* acce = accf + 2;                           // [2] accumulator drained
    uint32_t* tslot = reinterpret_cast<uint32_t*>(acce + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s]))); }
        for (int a = 0; a < 2; ++a) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&accf[a]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(s32(&acce[a]))); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // activations: act[16][512] row-major in global -> swizzled blocks
    for (int i = threadIdx.x; i < 16 * 512; i += 192) {
        const int m = i / 512, kk = i % 512;
        *reinterpret_cast<bf16*>(bact + (kk >> 6) * 2048 + sw128(m, kk & 63)) = act[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tslot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    const int spt = M == 128 ? 4 : 2;                    // stages per tile
    const unsigned char* base = src + (size_t)(blockIdx.x % 8) * slice_stride;
    const int stages_per_slice = (int)(slice_bytes / STAGE);
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < ntiles; ++t)
                for (int i = 0; i < spt; ++i, ++it) {
                    const int s = it % STAGES; const uint32_t use = it / STAGES;
                    if (use > 0) while (!try_wait(&empty[s], (use & 1) ^ 1)) {}
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(STAGE) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(ring + s * STAGE)), "l"(base + (size_t)(it % stages_per_slice) * STAGE), "r"(STAGE), "r"(s32(&full[s])) : "memory");
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t ID = M == 128 ? idesc(128, 16) : idesc(64, 16);
            for (int t = 0; t < ntiles; ++t) {
                const uint32_t acc = t & 1, ause = t >> 1;
                if (ause > 0) while (!try_wait(&acce[acc], (ause & 1) ^ 1)) {}
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + acc * 16;
                for (int i = 0; i < spt; ++i, ++it) {
                    const int s = it % STAGES; const uint32_t use = it / STAGES;
                    while (!try_wait(&full[s], use & 1)) {}
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int nblk = M == 128 ? 2 : 4, blk_bytes = M == 128 ? 16384 : 8192;
                    for (int b = 0; b < nblk; ++b) {
                        const int kb = i * nblk + b;                        // 64-wide k block of the tile
                        const uint32_t a_addr = s32(ring + s * STAGE + b * blk_bytes), b_addr = s32(bact + kb * 2048);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t da = smem_desc(a_addr + kk * 32), db = smem_desc(b_addr + kk * 32);
                            const uint32_t accum = (kb | kk) != 0;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(d), "l"(da), "l"(db), "r"(ID), "r"(accum) : "memory");
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[s])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&accf[acc])) : "memory");
            }
        }
    } else {                                             // warps 2..5: epilogue, lanes 32*(warp&3)..
        const int lg = warp & 3;
        unsigned chk = 0;
        for (int t = 0; t < ntiles; ++t) {
            const uint32_t acc = t & 1;
            while (!try_wait(&accf[acc], (t >> 1) & 1)) {}
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[16];
            const uint32_t taddr = tmem + acc * 16 + ((uint32_t)(lg * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&acce[acc])) : "memory");
            if (t == 0 && blockIdx.x == 0 && dump)
                for (int j = 0; j < 16; ++j) dump[(lg * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
            chk ^= v[0] ^ v[5];
        }
        if (chk == 0x12345u) sink[threadIdx.x] = chk;
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
    }
}
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
int main() {
    const size_t slice = 5505024;                        // 168 stages of 32 KB (~ one rank's weight stream)
    const size_t stride = 6u << 20;
    std::vector<unsigned char> h(8 * stride, 0);
    // tile 0 of slice 0, both layouts are generated from W[row][k] = f(row, k)
    auto wval = [](int r, int kk) { return (float)(((r * 131 + kk * 7) % 61) - 30) / 64.f; };
    std::vector<uint16_t> act(16 * 512);
    for (int m = 0; m < 16; ++m) for (int kk = 0; kk < 512; ++kk) act[m * 512 + kk] = f2bf((float)(((m * 17 + kk * 3) % 41) - 20) / 32.f);
    unsigned char* d_src; cudaMalloc(&d_src, h.size());
    bf16* d_act; cudaMalloc(&d_act, act.size() * 2); cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice);
    float* d_dump; cudaMalloc(&d_dump, 128 * 16 * 4);
    unsigned* sink; cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int M : {128, 64}) {
        // fill every slice with small random-ish bf16 (finite), then lay tile 0 of slice 0 out properly
        for (size_t i = 0; i < h.size(); i += 2) { const uint16_t v = f2bf((float)((int)((i / 2) * 2654435761u >> 27) - 16) / 64.f); h[i] = v & 255; h[i + 1] = v >> 8; }
        const int rows = M, nblk_stage = M == 128 ? 2 : 4, blk_bytes = M == 128 ? 16384 : 8192, spt = M == 128 ? 4 : 2;
        for (int i = 0; i < spt; ++i)
            for (int b = 0; b < nblk_stage; ++b) {
                const int kb = i * nblk_stage + b;
                for (int r = 0; r < rows; ++r)
                    for (int kk = 0; kk < 64; ++kk) {
                        const uint16_t v = f2bf(wval(r, kb * 64 + kk));
                        const size_t off = (size_t)i * STAGE + (size_t)b * blk_bytes + sw128(r, kk);
                        h[off] = v & 255; h[off + 1] = v >> 8;
                    }
            }
        cudaMemcpy(d_src, h.data(), h.size(), cudaMemcpyHostToDevice);
        cudaMemset(d_dump, 0, 128 * 16 * 4);
        const int ntiles = M == 128 ? 42 * 40 : 84 * 40;       // 40 passes over the slice
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k<<<104, 192, STAGES * STAGE + 16384 + 1024 + 256>>>(d_src, slice, stride, ntiles, M, d_act, d_dump, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        const double bytes = (double)ntiles * (M == 128 ? 4 : 2) * STAGE;
        printf("M=%3d: %.3f ms, %.1f GB/s per SM (104 CTAs, %.2f TB/s) %s\n", M, best, bytes / best / 1e6, bytes * 104 / best / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
        std::vector<float> dump(128 * 16);
        cudaMemcpy(dump.data(), d_dump, dump.size() * 4, cudaMemcpyDeviceToHost);
        // reference
        double maxerr = 0; int bad = 0;
        std::vector<int> lane_of_row(rows, -1);
        for (int r = 0; r < rows; ++r) {
            float ref[16];
            for (int m = 0; m < 16; ++m) { double a = 0; for (int kk = 0; kk < 512; ++kk) a += (double)bf2f(f2bf(wval(r, kk))) * bf2f(act[m * 512 + kk]); ref[m] = (float)a; }
            // find the TMEM lane that holds this row
            for (int ln = 0; ln < 128; ++ln) {
                double err = 0; for (int m = 0; m < 16; ++m) err = fmax(err, fabs(dump[ln * 16 + m] - ref[m]));
                if (err < 2e-2) { lane_of_row[r] = ln; break; }
            }
            if (lane_of_row[r] < 0) ++bad;
            else for (int m = 0; m < 16; ++m) maxerr = fmax(maxerr, fabs(dump[lane_of_row[r] * 16 + m] - ref[m]));
        }
        printf("   rows matched %d / %d, max err %.4f; row -> lane: ", rows - bad, rows, maxerr);
        for (int r = 0; r < rows; r += 8) printf("%d:%d ", r, lane_of_row[r]);
        printf("\n");
    }
    return 0;
}
