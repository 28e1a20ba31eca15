// Flash-style attention forward, dh = 64, for the sequence-parallel (teacher-forced / encoder) path
// (SURVEY.md 8(a) row a5): encoder self-attention (key-padding mask), causal decoder self-attention
// (causal + key-padding) and encoder-decoder cross-attention (key-padding).  Online softmax in fp32,
// exp2 with the 1/sqrt(64) scale folded in; bf16 operands on warp-level mma.sync m16n8k16.
// One CTA = 64 query rows of one (b, h); 4 warps x 16 rows; K/V streamed in 64-row tiles (cp.async
// double buffer).
#pragma once
#include "common.cuh"

namespace tts {

struct AttnParams {
    const bf16 *Q, *K, *V; bf16* O;
    // element strides: batch, head, row
    long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
    int B, H, Lq, Lk;
    const int* klens;     // keys >= klens[b] are masked (null: all Lk valid)
    int causal;
    float scale_log2;     // (1/sqrt(dh)) * log2(e)
};

constexpr int FA_BM = 64, FA_BN = 64, FA_LDS = 72;   // padded smem row (bf16 elements)

__global__ void __launch_bounds__(128) flash_attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(16) bf16 Qs[FA_BM * FA_LDS];
    __shared__ __align__(16) bf16 Ks[2][FA_BN * FA_LDS];
    __shared__ __align__(16) bf16 Vs[2][FA_BN * FA_LDS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
    const bf16* Qg = p.Q + b * p.q_bs + h * p.q_hs;
    const bf16* Kg = p.K + b * p.k_bs + h * p.k_hs;
    const bf16* Vg = p.V + b * p.v_bs + h * p.v_hs;
    const int klen = p.klens ? min(p.klens[b], p.Lk) : p.Lk;
    int nkt = (klen + FA_BN - 1) / FA_BN;
    if (p.causal) nkt = min(nkt, (min(q0 + FA_BM, p.Lq) + FA_BN - 1) / FA_BN);

    // 64 rows x 8 chunks of 16 B = 512 chunks; 4 per thread
    auto load_tile = [&](bf16* dst, const bf16* src, long rs, int r0, int rmax) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int c = tid + i * 128, r = c >> 3, ch = c & 7;
            bool ok = (r0 + r) < rmax;
            cp_async_16(dst + r * FA_LDS + ch * 8, src + (long)(ok ? (r0 + r) : 0) * rs + ch * 8, ok);
        }
    };
    load_tile(Qs, Qg, p.q_rs, q0, p.Lq);
    if (nkt > 0) { load_tile(Ks[0], Kg, p.k_rs, 0, klen); load_tile(Vs[0], Vg, p.v_rs, 0, klen); }
    cp_async_commit();

    float o_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t qf[4][4];
    const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) {
            load_tile(Ks[buf ^ 1], Kg, p.k_rs, (kt + 1) * FA_BN, klen);
            load_tile(Vs[buf ^ 1], Vg, p.v_rs, (kt + 1) * FA_BN, klen);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                ldmatrix_x4(qf[ks], Qs + (warp * 16 + (lane & 15)) * FA_LDS + ks * 16 + (lane >> 4) * 8);
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                uint32_t kf[4];
                ldmatrix_x4(kf, Ks[buf] + (nj * 16 + (lane & 7) + (lane >> 4) * 8) * FA_LDS + ks * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16_16816(s[nj * 2], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[nj * 2 + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // mask + online softmax
        const int kbase = kt * FA_BN;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kj = kbase + ni * 8 + t4 * 2 + (e & 1);
                const int r = e >> 1;
                float v = s[ni][e] * p.scale_log2;
                if (kj >= klen || (p.causal && kj > qrow[r])) v = -INFINITY;
                s[ni][e] = v;
                mx[r] = fmaxf(mx[r], v);
            }
        float corr[2], msafe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(m_run[r], mx[r]);
            msafe[r] = (mnew == -INFINITY) ? 0.f : mnew;
            corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] - msafe[r]);
            m_run[r] = mnew;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            const float p0 = exp2f(s[ni][0] - msafe[0]), p1 = exp2f(s[ni][1] - msafe[0]);
            const float p2 = exp2f(s[ni][2] - msafe[1]), p3 = exp2f(s[ni][3] - msafe[1]);
            rs[0] += p0 + p1; rs[1] += p2 + p3;
            const int j = ni >> 1;
            if ((ni & 1) == 0) { pf[j][0] = pack_bf16x2(p0, p1); pf[j][1] = pack_bf16x2(p2, p3); }
            else               { pf[j][2] = pack_bf16x2(p0, p1); pf[j][3] = pack_bf16x2(p2, p3); }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            o_acc[ni][0] *= corr[0]; o_acc[ni][1] *= corr[0];
            o_acc[ni][2] *= corr[1]; o_acc[ni][3] *= corr[1];
        }
        // O += P V
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) {
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, Vs[buf] + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * FA_LDS + nd * 16 + (lane >> 4) * 8);
                mma_bf16_16816(o_acc[nd * 2], pf[j], vf[0], vf[1]);
                mma_bf16_16816(o_acc[nd * 2 + 1], pf[j], vf[2], vf[3]);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    // finalize
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    bf16* Og = p.O + b * p.o_bs + h * p.o_hs;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (qrow[r] < p.Lq) {
            const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                uint32_t v = pack_bf16x2(o_acc[ni][r * 2] * inv, o_acc[ni][r * 2 + 1] * inv);
                *reinterpret_cast<uint32_t*>(Og + (long)qrow[r] * p.o_rs + ni * 8 + t4 * 2) = v;
            }
        }
    }
}

inline cudaError_t launch_flash_attn(const AttnParams& p, cudaStream_t stream) {
    dim3 grid((p.Lq + FA_BM - 1) / FA_BM, p.H, p.B);
    flash_attn_fwd_kernel<<<grid, 128, 0, stream>>>(p);
    ++launch_counter();
    return cudaGetLastError();
}

}  // namespace tts
