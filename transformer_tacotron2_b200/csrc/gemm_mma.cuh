// Large-M bf16 GEMM with fused epilogue:  C[M,N] = epi( sum_tap A[m + tap - pad, :] . W[tap][n, :] )
// Used by the encoder, the cross-K/V projection, the teacher-forced decoder and the postnet
// (SURVEY.md 8(a) rows a3, a4, a10).  conv1d(k=5, pad=2) is 5 row-shifted GEMMs accumulated into one
// tile; rows that fall outside [0, T) of their own utterance are zero-filled by the loader.
// Warp-level mma.sync m16n8k16 mainloop, cp.async 3-stage pipeline, 128x128x32 CTA tile.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace tts {

enum GemmAct { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };
enum GemmScatter { SC_NONE = 0, SC_CROSS_KV = 1, SC_HEAD = 2, SC_CROSS_KV_VT = 3 };

struct GemmParams {
    // operands
    const bf16* A; int lda;          // [M][lda] bf16, K valid columns (K % 32 == 0)
    const bf16* W; int ldw;          // [taps][Nw][ldw] bf16, Nw = N rounded up to 128 (zero rows)
    int M, N, K, taps, Nw;
    int T;                           // rows per utterance (row m -> b = m / T, t = m % T); T = M if unused
    // epilogue
    const float* bias;               // [N] or null
    int act;
    const bf16* resid_bf16; const float* resid_f32; int ldr;
    const float* pe; float alpha;    // + alpha * pe[t][n]   (pe row stride 512)
    const int* lens;                 // zero rows with t >= lens[b]
    int drop_site; uint64_t seed; int utt_offset;   // p = 0.5 bit dropout after the activation (site < 0: off)
    float* out_f32; bf16* out_bf16; int ldo;
    int scatter;                     // GemmScatter
    // SC_CROSS_KV: out_bf16 = cache [layers][2][B][H][S][64], N = layers * 1024, T = S
    // SC_HEAD    : out_f32 = mel_before [M][80], out2_f32 = stop_logits [M]
    float* out2_f32; int B;
    int Lpad;                        // SC_CROSS_KV*: row capacity of the cache per (layer, kv, b, h), multiple of 16
};

constexpr int GB_M = 128, GB_N = 128, GB_K = 32, G_STAGES = 3, G_LDS = GB_K + 8;   // smem row = 40 bf16 = 80 B
constexpr int G_SMEM_BYTES = G_STAGES * (GB_M + GB_N) * G_LDS * 2;

TTS_D void gemm_store(const GemmParams& p, int m, int n, float v0, float v1) {
    // (m, n) and (m, n+1); n is even
    if (m >= p.M || n >= p.N) return;
    const bool has1 = (n + 1) < p.N;
    const int b = m / p.T, t = m - b * p.T;
    if (p.bias) { v0 += p.bias[n]; if (has1) v1 += p.bias[n + 1]; }
    if (p.resid_bf16) {
        v0 += __bfloat162float(p.resid_bf16[(size_t)m * p.ldr + n]);
        if (has1) v1 += __bfloat162float(p.resid_bf16[(size_t)m * p.ldr + n + 1]);
    }
    if (p.resid_f32) {
        v0 += p.resid_f32[(size_t)m * p.ldr + n];
        if (has1) v1 += p.resid_f32[(size_t)m * p.ldr + n + 1];
    }
    if (p.pe) {
        v0 += p.alpha * p.pe[(size_t)t * kDModel + n];
        if (has1) v1 += p.alpha * p.pe[(size_t)t * kDModel + n + 1];
    }
    if (p.act == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
    else if (p.act == ACT_TANH) { v0 = tanhf(v0); v1 = tanhf(v1); }
    if (p.drop_site >= 0) {
        v0 = keep_bit(p.seed, p.drop_site, t, p.utt_offset + b, n) ? 2.f * v0 : 0.f;
        if (has1) v1 = keep_bit(p.seed, p.drop_site, t, p.utt_offset + b, n + 1) ? 2.f * v1 : 0.f;
    }
    if (p.lens && t >= p.lens[b]) { v0 = 0.f; v1 = 0.f; }
    if (p.scatter == SC_CROSS_KV || p.scatter == SC_CROSS_KV_VT) {
        // n = layer * 1024 + kv * 512 + h * 64 + d  ->  cache [layer][kv][B][H][Lpad * 64]; K rows are row-major, V is
        // row-major too (SC_CROSS_KV, teacher-forced flash attention) or in transposed 16-row blocks
        // [Lpad/16][64 d][16 rows] (SC_CROSS_KV_VT, decode_cluster.cuh)
        const int lkv = n >> 9, h = (n >> 6) & 7, d = n & 63;
        const size_t base = (((size_t)lkv * p.B + b) * kHeads + h) * (size_t)p.Lpad * kDHead;
        if (p.scatter == SC_CROSS_KV_VT && (lkv & 1)) {
            bf16* blk = p.out_bf16 + base + (size_t)(t >> 4) * 1024 + (t & 15);
            blk[d * 16] = __float2bfloat16(v0);
            blk[(d + 1) * 16] = __float2bfloat16(v1);
        } else {
            *reinterpret_cast<uint32_t*>(p.out_bf16 + base + (size_t)t * kDHead + d) = pack_bf16x2(v0, v1);
        }
        return;
    }
    if (p.scatter == SC_HEAD) {
        if (n < 80) { p.out_f32[(size_t)m * 80 + n] = v0; if (n + 1 < 80) p.out_f32[(size_t)m * 80 + n + 1] = v1; }
        if (n == 80) p.out2_f32[m] = v0;
        return;
    }
    if (p.out_f32) {
        p.out_f32[(size_t)m * p.ldo + n] = v0;
        if (has1) p.out_f32[(size_t)m * p.ldo + n + 1] = v1;
    }
    if (p.out_bf16) {
        if (has1 && ((p.ldo & 1) == 0)) *reinterpret_cast<uint32_t*>(p.out_bf16 + (size_t)m * p.ldo + n) = pack_bf16x2(v0, v1);
        else {
            p.out_bf16[(size_t)m * p.ldo + n] = __float2bfloat16(v0);
            if (has1) p.out_bf16[(size_t)m * p.ldo + n + 1] = __float2bfloat16(v1);
        }
    }
}

__global__ void __launch_bounds__(256) gemm_mma_kernel(const GemmParams p) {
    extern __shared__ __align__(16) unsigned char g_smem[];
    bf16* As = reinterpret_cast<bf16*>(g_smem);                       // [stages][128][40]
    bf16* Bs = As + G_STAGES * GB_M * G_LDS;                          // [stages][128][40]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;                          // 2 x 4 warps -> 64 x 32 warp tile
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    const int kpt = p.K / GB_K;                                       // k-blocks per tap
    const int nkb = kpt * p.taps;
    const int pad = p.taps >> 1;

    // loader assignment: 2 x 16-byte chunks of A and of B per thread per stage
    int ld_row[2], ld_chunk[2], a_t[2]; bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int c = tid + i * 256;
        ld_row[i] = c >> 2; ld_chunk[i] = c & 3;
        int m = m0 + ld_row[i];
        a_ok[i] = m < p.M;
        a_t[i] = a_ok[i] ? (m % p.T) : 0;
    }
    auto load_stage = [&](int stage, int kb) {
        const int tap = kb / kpt, kc = (kb - tap * kpt) * GB_K;
        const int shift = tap - pad;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int tt = a_t[i] + shift;
            const bool ok = a_ok[i] && tt >= 0 && tt < p.T;
            const bf16* src = p.A + (size_t)(ok ? (m0 + ld_row[i] + shift) : 0) * p.lda + kc + ld_chunk[i] * 8;
            cp_async_16(As + (stage * GB_M + ld_row[i]) * G_LDS + ld_chunk[i] * 8, src, ok);
            const bf16* wsrc = p.W + ((size_t)tap * p.Nw + n0 + ld_row[i]) * p.ldw + kc + ld_chunk[i] * 8;
            cp_async_16(Bs + (stage * GB_N + ld_row[i]) * G_LDS + ld_chunk[i] * 8, wsrc, true);
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        {   // prefetch k-block kb + STAGES - 1 into the slot freed at iteration kb - 1
            const int nk = kb + G_STAGES - 1;
            if (nk < nkb) load_stage(nk % G_STAGES, nk);
            cp_async_commit();
        }
        const bf16* as = As + (kb % G_STAGES) * GB_M * G_LDS;
        const bf16* bs = Bs + (kb % G_STAGES) * GB_N * G_LDS;
#pragma unroll
        for (int ks = 0; ks < GB_K; ks += 16) {
            uint32_t af[4][4], bfr[2][4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
                ldmatrix_x4(af[mi], as + (wm * 64 + mi * 16 + (lane & 15)) * G_LDS + ks + (lane >> 4) * 8);
#pragma unroll
            for (int nj = 0; nj < 2; ++nj)
                ldmatrix_x4(bfr[nj], bs + (wn * 32 + nj * 16 + (lane & 7) + (lane >> 4) * 8) * G_LDS + ks + ((lane >> 3) & 1) * 8);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    mma_bf16_16816(acc[mi][ni], af[mi], bfr[ni >> 1][(ni & 1) * 2], bfr[ni >> 1][(ni & 1) * 2 + 1]);
        }
    }
    cp_async_wait<0>();

    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int m = m0 + wm * 64 + mi * 16 + g;
            const int n = n0 + wn * 32 + ni * 8 + t4 * 2;
            gemm_store(p, m, n, acc[mi][ni][0], acc[mi][ni][1]);
            gemm_store(p, m + 8, n, acc[mi][ni][2], acc[mi][ni][3]);
        }
}

inline cudaError_t launch_gemm(const GemmParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((p.N + GB_N - 1) / GB_N, (p.M + GB_M - 1) / GB_M);
    gemm_mma_kernel<<<grid, 256, G_SMEM_BYTES, stream>>>(p);
    ++launch_counter();
    return cudaGetLastError();
}

}  // namespace tts
