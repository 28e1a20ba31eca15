// What does one 32-row attention unit of the decode kernel cost on an SM?  (scores = K . q and out += V^T . p on mma.sync
// m16n8k16, online softmax in registers; data already in shared memory.)  W warps run `iters` units back to back.
//   mode 0: the full unit   1: HMMA only (operands loaded once)   2: loads only (no HMMA)   3: full, but 64-row units (4 sub-chunks per softmax round)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o attn_unit attn_unit.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }

template <int NSC>   // sub-chunks of 16 rows per softmax round
__device__ __forceinline__ void unit(const unsigned char* kb, const unsigned char* vb, int kc0, const uint32_t (&qb0)[4], const uint32_t (&qb1)[4],
                                     float& m, float& l, float (&o)[4][4], int mode) {
    float v[NSC][2][2];
#pragma unroll
    for (int s2 = 0; s2 < NSC; ++s2) {
        float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        uint4 kr[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            kr[j][0] = *reinterpret_cast<const uint4*>(kb + s2 * 2048 + j * 256 + kc0);
            kr[j][1] = *reinterpret_cast<const uint4*>(kb + s2 * 2048 + j * 256 + (kc0 ^ 16));
        }
        if (mode != 2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t a[4] = {qb0[ks], qb0[ks ^ 2], qb1[ks], qb1[ks ^ 2]};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint4 w = kr[j][ks >> 1];
                    mma(sc[j], a, (ks & 1) ? w.z : w.x, (ks & 1) ? w.w : w.y);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) { sc[j][0] = __uint_as_float(kr[j][0].x ^ kr[j][1].y) * 1e-30f; sc[j][3] = __uint_as_float(kr[j][0].z ^ kr[j][1].w) * 1e-30f; }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) { v[s2][j][0] = sc[j][0]; v[s2][j][1] = sc[j][3]; }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < NSC; ++s2)
#pragma unroll
        for (int j = 0; j < 2; ++j) mx = fmaxf(mx, fmaxf(v[s2][j][0], v[s2][j][1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float mnew = fmaxf(m, mx);
    if (mnew != m) {
        const float scale = (m == -INFINITY) ? 0.f : ex2(m - mnew);
        l *= scale;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= scale; o[dt][1] *= scale; o[dt][2] *= scale; o[dt][3] *= scale; }
        m = mnew;
    }
#pragma unroll
    for (int s2 = 0; s2 < NSC; ++s2) {
        const float p0 = ex2(v[s2][0][0] - mnew), p1 = ex2(v[s2][0][1] - mnew), p2 = ex2(v[s2][1][0] - mnew), p3 = ex2(v[s2][1][1] - mnew);
        l += (p0 + p1) + (p2 + p3);
        const uint32_t b0 = pack(p0, p1), b1 = pack(p2, p3);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
            const uint4 w = *reinterpret_cast<const uint4*>(vb + s2 * 2048 + dt * 512);
            const uint32_t a[4] = {w.x, w.y, w.z, w.w};
            if (mode != 2) mma(o[dt], a, b0, b1);
            else o[dt][0] += __uint_as_float(w.x ^ w.y ^ w.z ^ w.w) * 1e-30f * __uint_as_float(b0 ^ b1);
        }
    }
}

__global__ void __launch_bounds__(544, 1) k(float* out, unsigned long long* cyc, int W, int iters, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < 160 * 1024 / 4; i += 544) {
        const float a = (float)((i * 2654435761u) >> 20 & 255) / 256.f - 0.5f, b = (float)((i * 40503u) >> 8 & 255) / 256.f - 0.5f;
        reinterpret_cast<uint32_t*>(smem)[i] = pack(a, b);
    }
    __syncthreads();
    if (warp >= W) return;
    uint32_t qb0[4], qb1[4];
    for (int ks = 0; ks < 4; ++ks) { qb0[ks] = pack(0.1f * (t4 + ks), -0.05f * ks); qb1[ks] = pack(0.02f * g, 0.03f); }
    float m = -INFINITY, l = 0.f, o[4][4] = {};
    const int gi = warp / 3;
    const unsigned char* base = smem + gi * 32768;
    const int krow_off = (4 * (g >> 1) + (g & 1)) * 128 + t4 * 32, kc0 = (g & 1) << 4;
    const long long t0 = clock64();
    if (mode == 3) {
        for (int it = 0; it < iters / 2; ++it) {
            const unsigned char* st = base + (it & 1) * 16384;
            unit<4>(st + krow_off, st + 8192 + lane * 16, kc0, qb0, qb1, m, l, o, 0);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            const unsigned char* st = base + ((it >> 1) & 1) * 16384 + (it & 1) * 4096;
            unit<2>(st + krow_off, st + 8192 + lane * 16, kc0, qb0, qb1, m, l, o, mode);
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 17 + warp] = (unsigned long long)(t1 - t0);
    float acc = l + m;
    for (int dt = 0; dt < 4; ++dt) acc += o[dt][0] + o[dt][1] + o[dt][2] + o[dt][3];
    if (acc == 123.456f) out[tid] = acc;
}

int main() {
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 148 * 17 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int iters = 2000;
    printf("mode warps | cycles per 32-row unit per warp | 32-row units per 1000 cycles per SM | GB/s per SM at 1.965 GHz (8 KB per unit)\n");
    for (int mode = 0; mode < 4; ++mode)
        for (int W : {1, 3, 4, 8, 15}) {
            cudaMemset(cyc, 0, 148 * 17 * 8);
            k<<<104, 544, 160 * 1024>>>(out, cyc, W, iters, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h[17]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double mx = 0; for (int w = 0; w < W; ++w) mx = h[w] > mx ? (double)h[w] : mx;
            const double per = mx / iters;
            printf("%d %2d | %7.1f | %6.2f | %6.1f\n", mode, W, per, 1000.0 * W / per, 8192.0 * W / per * 1.965);
        }
    return 0;
}
