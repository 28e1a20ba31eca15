// Flash attention backward on tcgen05 / TMEM / TMA, dh = 64 (training path, SURVEY.md 8(a) rows a5 + a12).
// Given Q, K, V, dO and the forward's per-row log-sum-exp L (log2 domain) and D = rowsum(dO * O):
//     P = 2^(S * scale_log2 - L),  S = Q K^T        dP = dO V^T        dS = P * (dP - D) / sqrt(dh)
//     dV = P^T dO        dK = dS^T Q        dQ = dS K
// One CTA = one 128-key tile of one (b, h), looping over the 128-query tiles that see it (causal: i >= j).
//   warp 0      TMA producer: K_j, V_j once; Q_i, dO_i in a 3-stage ring (4-D maps, 128B swizzle, zero fill past the end)
//   warp 1      MMA issuer:   S, dP        M128(q) N128(keys) K64,  operands K-major from shared memory  -> TMEM
//                             dV += P^T dO, dK += dS^T Q   M128(keys) N64 K128(q): A = the bf16 P / dS tile in shared memory
//                                                          read as an MN-major operand, B = dO_i / Q_i as MN-major
//                             dQ_i = dS K_j               M128(q) N64 K128(keys): A = dS K-major, B = K_j MN-major
// Pipelining: S is double-buffered in TMEM and dQ_i reuses dP_i's columns (all 512 columns are in use), so S_{i+1} exists
// before tile i's three MMAs are issued and the exponentials of tile i + 1 run underneath them.
//   warps 2..9  thread = query row x half of the columns: S -> P (bf16, swizzled store), dP -> dS, then dQ_i from TMEM -> fp32 red.add into the
//               dQ accumulator (several key tiles add into the same rows); at the end dK_j, dV_j -> bf16.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "attention_tc.cuh"

namespace tts {

struct AttnBwdParams {
    const bf16 *Q, *K, *V, *dO;           // strides as in AttnParams (dO uses the o_* strides)
    long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
    const float *lse, *dsum;              // [B][H][Lq]
    float* dQ; long dq_bs, dq_hs, dq_rs;  // fp32 accumulator (zeroed by the caller), red.add
    bf16* dQ16; long dq16_bs, dq16_hs, dq16_rs;   // if set AND there is a single key tile (Lk <= 128): dQ is complete after one tile and
                                                  // is written directly as bf16 here (no zeroing, no fp32 pass)
    bf16 *dK, *dV; long dk_bs, dk_hs, dk_rs, dv_bs, dv_hs, dv_rs;
    int B, H, Lq, Lk;
    const int* klens;
    int causal;
    float scale_log2, scale;              // log2(e)/sqrt(dh), 1/sqrt(dh)
};

struct AttnBwdTcParams {
    alignas(64) CUtensorMap tm_q, tm_k, tm_v, tm_do;
    AttnBwdParams a;
};

constexpr int FB_CWARPS = 8;                                        // compute warps: two per TMEM lane quarter, half of the columns each
constexpr int FB_THREADS = 64 + 32 * FB_CWARPS;
constexpr int FB_TILE = 128 * 64 * 2;                               // 16 KB
constexpr int FB_QSTAGES = 3;                                       // Q / dO ring: tile i + 2 is loaded while tile i is in the MMAs
constexpr int FB_SMEM_BYTES = (6 + 2 * FB_QSTAGES) * FB_TILE + 256;   // K, V, Q[3], dO[3], P (2 halves), dS (2 halves)
constexpr int FB_COL_S = 0, FB_COL_DP = 256, FB_COL_DQ = 256, FB_COL_DV = 384, FB_COL_DK = 448;   // S[2] | dP (dQ aliases it) | dV | dK = 512 columns
// kind::f16 instruction descriptors (see gemm_tc.cuh): bit 15 = A is MN-major, bit 16 = B is MN-major
constexpr uint32_t FB_IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t FB_IDESC_KV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t FB_IDESC_Q = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// 128B-swizzled tile whose MN extent spans two 64-element blocks `lbo_bytes` apart (MN-major operand)
TTS_D uint64_t fb_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
TTS_D void fb_red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(FB_THREADS, 1) flash_attn_bwd_tc_kernel(const __grid_constant__ AttnBwdTcParams p) {
    extern __shared__ __align__(1024) unsigned char fb_smem[];
    unsigned char* sK = fb_smem;
    unsigned char* sV = fb_smem + FB_TILE;
    unsigned char* sQ = fb_smem + 2 * FB_TILE;                          // [FB_QSTAGES]
    unsigned char* sdO = fb_smem + (2 + FB_QSTAGES) * FB_TILE;          // [FB_QSTAGES]
    unsigned char* sP = fb_smem + (2 + 2 * FB_QSTAGES) * FB_TILE;       // [q 128][keys 0..63], [q 128][keys 64..127]
    unsigned char* sdS = fb_smem + (4 + 2 * FB_QSTAGES) * FB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(fb_smem + (6 + 2 * FB_QSTAGES) * FB_TILE);
    uint64_t *kv_full = bars, *qd_full = bars + 1, *qd_empty = bars + 4, *s_full = bars + 7, *s_empty = bars + 9;
    uint64_t *dp_full = bars + 11, *dpq_free = bars + 12, *ds_full = bars + 13, *mma3_done = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const AttnBwdParams& a = p.a;
    const int j = blockIdx.z, h = blockIdx.x, b = blockIdx.y;       // key tile slowest: with a causal mask tile 0 has the most work, and runs first
    const int k0 = j * 128;
    const int klen = a.klens ? min(a.klens[b], a.Lk) : a.Lk;
    const int nq = (a.Lq + 127) / 128;
    const int i0 = a.causal ? j : 0;
    const int ni = (k0 < klen && i0 < nq) ? nq - i0 : 0;             // query tiles that see this key tile

    if (threadIdx.x == 0) {
        if (tc_smem_u32(fb_smem) & 1023) __trap();
        tc_mbar_init(kv_full, 1);
        for (int s = 0; s < FB_QSTAGES; ++s) { tc_mbar_init(&qd_full[s], 1); tc_mbar_init(&qd_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc_mbar_init(&s_full[s], 1); tc_mbar_init(&s_empty[s], FB_CWARPS); }
        tc_mbar_init(dp_full, 1); tc_mbar_init(dpq_free, FB_CWARPS); tc_mbar_init(ds_full, FB_CWARPS); tc_mbar_init(mma3_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (ni > 0) {                                    // producer thread: K / V and the first Q / dO tile go out under the TMEM allocation
            tc_mbar_expect_tx(kv_full, 2 * FB_TILE);
            ft_tma_4d(sK, &p.tm_k, 0, k0, h, b, kv_full);
            ft_tma_4d(sV, &p.tm_v, 0, k0, h, b, kv_full);
            tc_mbar_expect_tx(&qd_full[0], 2 * FB_TILE);
            ft_tma_4d(sQ, &p.tm_q, 0, i0 * 128, h, b, &qd_full[0]);
            ft_tma_4d(sdO, &p.tm_do, 0, i0 * 128, h, b, &qd_full[0]);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && ni > 0) {                       // ---------------- TMA producer (K, V and tile 0 were issued in the prologue)
            for (int it = 1; it < ni; ++it) {
                const int s = it % FB_QSTAGES; const uint32_t use = it / FB_QSTAGES;
                if (use > 0) tc_mbar_wait(&qd_empty[s], (use & 1) ^ 1);
                tc_mbar_expect_tx(&qd_full[s], 2 * FB_TILE);
                ft_tma_4d(sQ + s * FB_TILE, &p.tm_q, 0, (i0 + it) * 128, h, b, &qd_full[s]);
                ft_tma_4d(sdO + s * FB_TILE, &p.tm_do, 0, (i0 + it) * 128, h, b, &qd_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && ni > 0) {                       // ---------------- MMA issuer
            const uint32_t ka = tc_smem_u32(sK), va = tc_smem_u32(sV), pa = tc_smem_u32(sP), dsa = tc_smem_u32(sdS);
            // S_it -> TMEM S[it & 1] (double-buffered: the exponentials of tile it + 1 run while tile it's three MMAs do)
            auto issue_s = [&](int it) {
                const int s = it & 1, qs = it % FB_QSTAGES;
                tc_mbar_wait(&qd_full[qs], (it / FB_QSTAGES) & 1);
                if (it >= 2) tc_mbar_wait(&s_empty[s], ((it >> 1) - 1) & 1);     // S_{it-2} has been read
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t qa = tc_smem_u32(sQ + qs * FB_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k) ft_mma(tmem_base + FB_COL_S + s * 128, tc_smem_desc(qa + k * 32), tc_smem_desc(ka + k * 32), FB_IDESC_S, k != 0);
                ft_commit(&s_full[s]);
            };
            // dP_it -> the dP/dQ columns (free once dQ_{it-1}, which aliases them, has been read)
            auto issue_dp = [&](int it) {
                const int qs = it % FB_QSTAGES;
                tc_mbar_wait(&qd_full[qs], (it / FB_QSTAGES) & 1);
                if (it > 0) tc_mbar_wait(dpq_free, (it - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t da = tc_smem_u32(sdO + qs * FB_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k) ft_mma(tmem_base + FB_COL_DP, tc_smem_desc(da + k * 32), tc_smem_desc(va + k * 32), FB_IDESC_S, k != 0);
                ft_commit(dp_full);
            };
            tc_mbar_wait(kv_full, 0);
            issue_s(0);
            issue_dp(0);
            if (ni > 1) issue_s(1);
            for (int it = 0; it < ni; ++it) {
                const int s = it % FB_QSTAGES;
                tc_mbar_wait(ds_full, it & 1);           // P_it and dS_it are in shared memory
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t qa = tc_smem_u32(sQ + s * FB_TILE), da = tc_smem_u32(sdO + s * FB_TILE);
#pragma unroll
                for (int k = 0; k < 8; ++k) {            // 16 queries per MMA: rows 16k.. of the P / dS / dO / Q tiles
                    ft_mma(tmem_base + FB_COL_DV, fb_desc_mn(pa + k * 2048, FB_TILE), fb_desc_mn(da + k * 2048, 16), FB_IDESC_KV, (it | k) != 0);
                    ft_mma(tmem_base + FB_COL_DK, fb_desc_mn(dsa + k * 2048, FB_TILE), fb_desc_mn(qa + k * 2048, 16), FB_IDESC_KV, (it | k) != 0);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)              // dQ_it = dS K_j into the columns dP_it occupied (it has been consumed)
                    ft_mma(tmem_base + FB_COL_DQ, tc_smem_desc(dsa + (k >> 2) * FB_TILE + (k & 3) * 32), fb_desc_mn(ka + k * 2048, 16), FB_IDESC_Q, k != 0);
                ft_commit(&qd_empty[s]);
                ft_commit(mma3_done);
                if (it + 1 < ni) issue_dp(it + 1);       // waits until the compute warps have taken dQ_it out
                if (it + 2 < ni) issue_s(it + 2);        // its Q stage was freed one tile ago and has been refilled meanwhile
            }
        }
    } else if (ni > 0) {                                 // ---------------- thread = query row of the tile, half of the columns
        const int lg = warp & 3, r = lg * 32 + lane, ch = (warp - 2) >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
        auto take_dq = [&](int it) {                     // dQ of tile `it` (this key tile's contribution) -> fp32 red.add
            const int qi = (i0 + it) * 128 + r;
            tc_mbar_wait(mma3_done, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            ft_ld32_nowait(lane_addr + FB_COL_DQ + ch * 32, v);
            ft_ld_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ft_arrive(dpq_free);          // the dP / dQ columns may be overwritten by dP of the next tile
            if (qi < a.Lq) {
                if (a.dQ16 && a.Lk <= 128) {
                    bf16* dq = a.dQ16 + b * a.dq16_bs + h * a.dq16_hs + (long)qi * a.dq16_rs + ch * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 8)
                        *reinterpret_cast<uint4*>(dq + i) = make_uint4(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                                                       pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                                                                       pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                                                                       pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
                } else {
                    float* dq = a.dQ + b * a.dq_bs + h * a.dq_hs + (long)qi * a.dq_rs + ch * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        fb_red4(dq + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                }
            }
        };
        for (int it = 0; it < ni; ++it) {
            const int s = it & 1;
            const int qi = (i0 + it) * 128 + r;
            const bool qvalid = qi < a.Lq;
            const long rowi = ((long)b * a.H + h) * a.Lq + qi;
            const float L = qvalid ? a.lse[rowi] : INFINITY;
            const float Dr = qvalid ? a.dsum[rowi] : 0.f;
            const int kmax = min(klen, a.causal ? qi + 1 : klen) - k0;       // keys [0, kmax) of this tile are visible to this row
            // P_it in registers: runs while the tensor pipe works on tile it - 1
            tc_mbar_wait(&s_full[s], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t pr[32];                                                 // P (this warp's 64 keys) as bf16 pairs
#pragma unroll
            for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 32) {
                uint32_t v[32];
                ft_ld32_nowait(lane_addr + FB_COL_S + s * 128 + c0, v);
                ft_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = (c0 + i < kmax) ? fast_exp2(__uint_as_float(v[i]) * a.scale_log2 - L) : 0.f;
                    const float p1 = (c0 + i + 1 < kmax) ? fast_exp2(__uint_as_float(v[i + 1]) * a.scale_log2 - L) : 0.f;
                    pr[((c0 & 63) + i) >> 1] = pack_bf16x2(p0, p1);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ft_arrive(&s_empty[s]);
            if (it > 0) take_dq(it - 1);                 // also: the MMAs that read the previous P / dS tiles are complete
#pragma unroll
            for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 32) {
                unsigned char* half = sP + (c0 >> 6) * FB_TILE + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ((c0 & 63) >> 3) + q, w = ((c0 & 63) >> 1) + q * 4;
                    *reinterpret_cast<uint4*>(half + ((chunk ^ (r & 7)) << 4)) = make_uint4(pr[w], pr[w + 1], pr[w + 2], pr[w + 3]);
                }
            }
            tc_mbar_wait(dp_full, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 32) {
                uint32_t v[32];
                ft_ld32_nowait(lane_addr + FB_COL_DP + c0, v);
                ft_ld_wait();
                uint32_t ds[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 pp = unpack_bf16x2(pr[((c0 & 63) + i) >> 1]);
                    ds[i >> 1] = pack_bf16x2(pp.x * (__uint_as_float(v[i]) - Dr) * a.scale, pp.y * (__uint_as_float(v[i + 1]) - Dr) * a.scale);
                }
                unsigned char* half = sdS + (c0 >> 6) * FB_TILE + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ((c0 & 63) >> 3) + q;
                    *reinterpret_cast<uint4*>(half + ((chunk ^ (r & 7)) << 4)) = make_uint4(ds[q * 4], ds[q * 4 + 1], ds[q * 4 + 2], ds[q * 4 + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) ft_arrive(ds_full);
        }
        take_dq(ni - 1);
        // dK_j, dV_j: TMEM lane = key row
        const int ki = k0 + r;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            uint32_t o[32];
            const uint32_t col = (which ? FB_COL_DK : FB_COL_DV) + ch * 32;
            ft_ld32_nowait(lane_addr + col, o);
            ft_ld_wait();
            if (ki < a.Lk) {
                bf16* g = (which ? a.dK + b * a.dk_bs + h * a.dk_hs + (long)ki * a.dk_rs : a.dV + b * a.dv_bs + h * a.dv_hs + (long)ki * a.dv_rs) + ch * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 8)
                    *reinterpret_cast<uint4*>(g + i) = make_uint4(pack_bf16x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])),
                                                                  pack_bf16x2(__uint_as_float(o[i + 2]), __uint_as_float(o[i + 3])),
                                                                  pack_bf16x2(__uint_as_float(o[i + 4]), __uint_as_float(o[i + 5])),
                                                                  pack_bf16x2(__uint_as_float(o[i + 6]), __uint_as_float(o[i + 7])));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {                                             // key tile that no query sees (all keys padded): zero gradients
        const int r = (warp & 3) * 32 + lane, ki = k0 + r, ch = (warp - 2) >> 2;
        if (ki < a.Lk) {
            bf16* gk = a.dK + b * a.dk_bs + h * a.dk_hs + (long)ki * a.dk_rs + ch * 32;
            bf16* gv = a.dV + b * a.dv_bs + h * a.dv_hs + (long)ki * a.dv_rs + ch * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) { *reinterpret_cast<uint4*>(gk + i) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(gv + i) = make_uint4(0, 0, 0, 0); }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// D[b][h][q] = sum_d dO * O (both bf16, o_* strides)
__global__ void attn_dsum_kernel(const bf16* O, const bf16* dO, long o_bs, long o_hs, long o_rs, float* dsum, int B, int H, int Lq) {
    const int idx = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);        // 8 threads per (b, h, q)
    const int sub = threadIdx.x & 7;
    float acc = 0.f;
    if (idx < B * H * Lq) {
        const int q = idx % Lq, bh = idx / Lq, hh = bh % H, bb = bh / H;
        const long off = bb * o_bs + hh * o_hs + (long)q * o_rs + sub * 8;
        const uint4 x = *reinterpret_cast<const uint4*>(O + off), y = *reinterpret_cast<const uint4*>(dO + off);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 u = unpack_bf16x2(xs[i]), w = unpack_bf16x2(ys[i]); acc += u.x * w.x + u.y * w.y; }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2); acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0 && idx < B * H * Lq) dsum[idx] = acc;
}

inline cudaError_t launch_flash_attn_bwd_tc(const AttnBwdParams& a, cudaStream_t stream) {
    static PerDevice pd;
    {
        const cudaError_t e = per_device_once(pd, nullptr, [] {
            return cudaFuncSetAttribute(flash_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES);
        });
        if (e != cudaSuccess) return e;
    }
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return cudaErrorInvalidValue;
    AttnBwdTcParams p;
    p.a = a;
    auto make = [&](CUtensorMap* tm, const bf16* base, long bs, long hs, long rs, int L) -> bool {
        if ((bs & 7) || (hs & 7) || (rs & 7)) return false;
        const cuuint64_t dims[4] = {64, (cuuint64_t)L, (cuuint64_t)a.H, (cuuint64_t)a.B};
        const cuuint64_t strides[3] = {(cuuint64_t)rs * 2, (cuuint64_t)hs * 2, (cuuint64_t)bs * 2};
        const cuuint32_t box[4] = {64, 128, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!make(&p.tm_q, a.Q, a.q_bs, a.q_hs, a.q_rs, a.Lq) || !make(&p.tm_k, a.K, a.k_bs, a.k_hs, a.k_rs, a.Lk) ||
        !make(&p.tm_v, a.V, a.v_bs, a.v_hs, a.v_rs, a.Lk) || !make(&p.tm_do, a.dO, a.o_bs, a.o_hs, a.o_rs, a.Lq))
        return cudaErrorInvalidValue;
    dim3 grid(a.H, a.B, (a.Lk + 127) / 128);
    flash_attn_bwd_tc_kernel<<<grid, FB_THREADS, FB_SMEM_BYTES, stream>>>(p);
    ++launch_counter();
    return cudaGetLastError();
}

}  // namespace tts
