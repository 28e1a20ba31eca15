// Flash attention forward on tcgen05 / TMEM / TMA, dh = 64 (component C10): encoder
// self-attention (key padding), causal decoder self-attention, encoder-decoder cross-attention.  Online softmax in
// fp32 with the 1/sqrt(64) scale folded into exp2.
// One CTA = 128 query rows of one (b, h); K/V streamed in 128-row tiles; two CTAs per SM (256 TMEM columns, 80 KB
// shared memory each) so one CTA's tensor work overlaps the other's softmax.
//   warp 0      TMA producer: Q once, K / V tiles double-buffered (4-D tensor maps {64, L, H, B}, 128B swizzle,
//               out-of-range rows zero-filled)
//   warp 1      MMA issuer (one lane):  S_j = Q K_j^T   tcgen05.mma M128 N128 K16 x4 (both operands from shared memory)
//                                       O  += P_j V_j   tcgen05.mma M128 N64  K16 x8, A = P_j read from TMEM, B = the
//                                                       row-major V tile used as an MN-major operand; O accumulates in TMEM
//   warps 2..5  softmax (thread = query row = TMEM lane): S_j is pulled into 128 registers in one pass (which frees the
//               S columns for S_{j+1} at once), row max via 3-input max, p = 2^(s*scale - m) with packed f32x2 FMAs and
//               the SFU (every 4th pair of scores takes a polynomial 2^x on the FMA pipe instead), P_j stored to TMEM as bf16 pairs.  The reference max m is moved (and O, l rescaled in TMEM)
//               only when some row of the warp exceeds it by more than 2^8, so most tiles never touch O.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_tc.cuh"         // mbarrier / TMA / descriptor helpers

namespace tts {

struct AttnParams {
    const bf16 *Q, *K, *V; bf16* O;
    // element strides: batch, head, row (all multiples of 8 elements; bases 16-byte aligned)
    long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
    int B, H, Lq, Lk;
    const int* klens;     // keys >= klens[b] are masked (null: all Lk valid)
    int causal;
    float scale_log2;     // (1/sqrt(dh)) * log2(e)
    float* lse;           // optional [B][H][Lq]: log2-domain log-sum-exp of the scaled scores (kept for the backward pass)
};

constexpr int FT_BM = 128, FT_BN = 128, FT_THREADS = 192;
constexpr int FT_TILE_BYTES = 128 * 64 * 2;                    // one [128][64] bf16 tile = 16 KB
constexpr int FT_SMEM_BYTES = 5 * FT_TILE_BYTES + 256;         // Q, K[2], V[2] + barriers: two CTAs per SM
constexpr int FT_TMEM_COLS = 256, FT_COL_S = 0, FT_COL_O = 128, FT_COL_P = 192;
constexpr float FT_RESCALE_LOG2 = 8.f;                         // running max may lag the true max by 2^8 before O is rescaled

struct AttnTcParams {
    alignas(64) CUtensorMap tm_q, tm_k, tm_v;
    AttnParams a;
};

constexpr uint32_t FT_IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t FT_IDESC_T = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

TTS_D void ft_tma_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(tc_smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(tc_smem_u32(bar)) : "memory");
}
TTS_D void ft_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
TTS_D void ft_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
TTS_D void ft_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accum) {   // A operand from TMEM
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
TTS_D void ft_ld32_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
TTS_D void ft_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
TTS_D void ft_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
TTS_D void ft_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
TTS_D void ft_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
TTS_D float ft_max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
TTS_D uint64_t ft_pack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
TTS_D void ft_unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
TTS_D uint64_t ft_fma2(uint64_t a, uint64_t b, uint64_t c) {     // two fp32 FMAs per issue slot
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
TTS_D uint64_t ft_add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 2^x for a packed pair on the FMA pipe (the SFU's 16 ex2 / clk / SM is the bound of this kernel at dh = 64): Cody-Waite
// range reduction with the 1.5 * 2^23 rounding constant, degree-3 polynomial for 2^f on [-0.5, 0.5] (max relative error
// 7.7e-5, far below the bf16 rounding of P), exponent spliced in with an integer shift-add.  x is clamped to >= -126.
// Measured at L = 4096: never 746 TFLOP/s, every 8th pair 788, every 4th pair 799, every 2nd pair 717 (issue-bound).
constexpr int FT_POLY_EVERY = 4;
TTS_D void ft_exp2_poly2(uint64_t x2, float& p_lo, float& p_hi) {
    float xl, xh;
    ft_unpack2(x2, xl, xh);
    const uint64_t xc = ft_pack2(fmaxf(xl, -126.f), fmaxf(xh, -126.f));
    const uint64_t magic = ft_pack2(12582912.f, 12582912.f), nmagic = ft_pack2(-12582912.f, -12582912.f), none = ft_pack2(-1.f, -1.f);
    const uint64_t fr = ft_add2(xc, magic);                    // low mantissa bits = round(x)
    const uint64_t f = ft_fma2(ft_add2(fr, nmagic), none, xc);  // x - round(x)
    uint64_t p = ft_fma2(f, ft_pack2(0.05508868396282196f, 0.05508868396282196f), ft_pack2(0.24260404706001282f, 0.24260404706001282f));
    p = ft_fma2(p, f, ft_pack2(0.6932762265205383f, 0.6932762265205383f));
    p = ft_fma2(p, f, ft_pack2(0.9999289512634277f, 0.9999289512634277f));
    float fl, fh, pl, ph;
    ft_unpack2(fr, fl, fh); ft_unpack2(p, pl, ph);
    p_lo = __uint_as_float(__float_as_uint(pl) + (__float_as_uint(fl) << 23));
    p_hi = __uint_as_float(__float_as_uint(ph) + (__float_as_uint(fh) << 23));
}
TTS_D void ft_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory"); }

template <int kPoly>
__global__ void __launch_bounds__(FT_THREADS, 2) flash_attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
    extern __shared__ __align__(1024) unsigned char ft_smem[];
    unsigned char* sQ = ft_smem;
    unsigned char* sK = ft_smem + FT_TILE_BYTES;          // [2]
    unsigned char* sV = ft_smem + 3 * FT_TILE_BYTES;      // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ft_smem + 5 * FT_TILE_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 7;
    uint64_t *s_full = bars + 9, *s_empty = bars + 10, *p_full = bars + 11, *t_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const AttnParams& a = p.a;
    // query tile slowest and reversed: with a causal mask the last tile sees the most keys, so the longest CTAs start first
    const int q0 = ((int)gridDim.z - 1 - (int)blockIdx.z) * FT_BM, h = blockIdx.x, b = blockIdx.y;
    const int klen = a.klens ? min(a.klens[b], a.Lk) : a.Lk;
    int nt = (klen + FT_BN - 1) / FT_BN;
    if (a.causal) nt = min(nt, (min(q0 + FT_BM, a.Lq) + FT_BN - 1) / FT_BN);

    if (threadIdx.x == 0) {
        if (tc_smem_u32(ft_smem) & 1023) __trap();        // the 128B-swizzle atoms need a 1024-byte aligned base
        tc_mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { tc_mbar_init(&k_full[i], 1); tc_mbar_init(&k_empty[i], 1); tc_mbar_init(&v_full[i], 1); tc_mbar_init(&v_empty[i], 1); }
        tc_mbar_init(s_full, 1); tc_mbar_init(s_empty, 4); tc_mbar_init(p_full, 4); tc_mbar_init(t_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // this thread is also the TMA producer: the first loads go out now, underneath the TMEM allocation and the block-wide
        // barrier (at L = 800 a CTA lives for ~4 tiles, so its prologue is a third of its life)
        if (nt > 0) {
            tc_mbar_expect_tx(q_full, FT_TILE_BYTES);
            ft_tma_4d(sQ, &p.tm_q, 0, q0, h, b, q_full);
            tc_mbar_expect_tx(&k_full[0], FT_TILE_BYTES);
            ft_tma_4d(sK, &p.tm_k, 0, 0, h, b, &k_full[0]);
            tc_mbar_expect_tx(&v_full[0], FT_TILE_BYTES);
            ft_tma_4d(sV, &p.tm_v, 0, 0, h, b, &v_full[0]);
        }
    }
    if (warp == 1) {                                     // TMEM: S cols 0..127 (fp32), O 128..191 (fp32), P 192..255 (bf16 pairs)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(FT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && nt > 0) {                       // ---------------- TMA producer (Q and tile 0 were issued in the prologue)
            for (int j = 1; j < nt; ++j) {
                const int s = j & 1; const uint32_t use = j >> 1;
                if (use > 0) tc_mbar_wait(&k_empty[s], (use & 1) ^ 1);
                tc_mbar_expect_tx(&k_full[s], FT_TILE_BYTES);
                ft_tma_4d(sK + s * FT_TILE_BYTES, &p.tm_k, 0, j * FT_BN, h, b, &k_full[s]);
                if (use > 0) tc_mbar_wait(&v_empty[s], (use & 1) ^ 1);
                tc_mbar_expect_tx(&v_full[s], FT_TILE_BYTES);
                ft_tma_4d(sV + s * FT_TILE_BYTES, &p.tm_v, 0, j * FT_BN, h, b, &v_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nt > 0) {                       // ---------------- MMA issuer
            auto issue_s = [&](int j) {
                const int s = j & 1;
                tc_mbar_wait(&k_full[s], (j >> 1) & 1);
                if (j > 0) tc_mbar_wait(s_empty, (j - 1) & 1);               // the softmax warps hold S_{j-1} in registers
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t qa = tc_smem_u32(sQ), ka = tc_smem_u32(sK + s * FT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ft_mma(tmem_base + FT_COL_S, tc_smem_desc(qa + k * 32), tc_smem_desc(ka + k * 32), FT_IDESC_S, k != 0);
                ft_commit(&k_empty[s]);
                ft_commit(s_full);
            };
            tc_mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nt; ++j) {
                if (j + 1 < nt) issue_s(j + 1);          // S_{j+1} runs on the tensor pipe while the softmax of tile j runs
                const int s = j & 1;
                tc_mbar_wait(&v_full[s], (j >> 1) & 1);
                tc_mbar_wait(p_full, j & 1);             // P_j is written and O has been rescaled if it had to be
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t va = tc_smem_u32(sV + s * FT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 8; ++k) {            // 16 keys per MMA; V rows 16k.. are the MN-major B operand
                    ft_mma_ts(tmem_base + FT_COL_O, tmem_base + FT_COL_P + k * 8, tc_smem_desc(va + k * 2048), FT_IDESC_T, (j | k) != 0);
                }
                ft_commit(&v_empty[s]);
                ft_commit(t_full);
            }
        }
    } else if (nt > 0) {                                 // ---------------- softmax: thread = query row = TMEM lane
        const int lg = warp & 3, r = lg * 32 + lane;
        const int qi = q0 + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
        float m = -INFINITY, l = 0.f;                    // m: the max the accumulators are scaled to (may lag the true max)
        for (int j = 0; j < nt; ++j) {
            tc_mbar_wait(s_full, j & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[128];
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) ft_ld32_nowait(lane_addr + FT_COL_S + c0, v + c0);
            ft_ld_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ft_arrive(s_empty);           // S is in registers: the issuer may overwrite it with S_{j+1}
            const int kmax = min(klen, a.causal ? qi + 1 : klen) - j * FT_BN;      // keys [0, kmax) of this tile are visible to this row
            if (kmax < 128) {                            // edge / diagonal tile: hide the masked keys
#pragma unroll
                for (int i = 0; i < 128; ++i) if (i >= kmax) v[i] = 0xff800000u;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                mx0 = ft_max3(mx0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                mx1 = ft_max3(mx1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                mx2 = ft_max3(mx2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                mx3 = ft_max3(mx3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            mx *= a.scale_log2;
            if (__any_sync(0xffffffffu, mx > m + FT_RESCALE_LOG2)) {        // warp-uniform: move the reference max, rescale O and l
                const float mnew = fmaxf(m, mx);
                const float alpha = (m == -INFINITY) ? 0.f : fast_exp2(m - mnew);
                if (j > 0) {
                    tc_mbar_wait(t_full, (j - 1) & 1);                       // O holds tiles 0..j-1
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t o[32];
                        ft_ld32_nowait(lane_addr + FT_COL_O + c0, o);
                        ft_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        ft_st32(lane_addr + FT_COL_O + c0, o);
                    }
                    ft_st_wait();
                }
                l *= alpha;
                m = mnew;
            }
            const float msafe = (m == -INFINITY) ? 0.f : m;
            if (j > 0) tc_mbar_wait(t_full, (j - 1) & 1);                    // P_{j-1} has been consumed
            const uint64_t sc2 = ft_pack2(a.scale_log2, a.scale_log2), nm2 = ft_pack2(-msafe, -msafe);
            uint64_t rs0 = 0, rs1 = 0;                   // (0.f, 0.f)
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float x0, x1, p2, p3;
                    ft_unpack2(ft_fma2(ft_pack2(__uint_as_float(v[c0 + i]), __uint_as_float(v[c0 + i + 1])), sc2, nm2), x0, x1);
                    const uint64_t xb = ft_fma2(ft_pack2(__uint_as_float(v[c0 + i + 2]), __uint_as_float(v[c0 + i + 3])), sc2, nm2);
                    const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
                    if (kPoly >= 2 && ((i >> 2) % (kPoly >= 2 ? kPoly / 2 : 1)) == 0) {      // every kPoly-th pair of scores
                        ft_exp2_poly2(xb, p2, p3);               // FMA pipe
                    } else {
                        float x2, x3;
                        ft_unpack2(xb, x2, x3);
                        p2 = fast_exp2(x2); p3 = fast_exp2(x3);  // SFU
                    }
                    rs0 = ft_add2(rs0, ft_pack2(p0, p1));
                    rs1 = ft_add2(rs1, ft_pack2(p2, p3));
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                    pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
                }
                ft_st16(lane_addr + FT_COL_P + (c0 >> 1), pk);
            }
            float ra, rb, rc, rd;
            ft_unpack2(rs0, ra, rb); ft_unpack2(rs1, rc, rd);
            const float rs = (ra + rb) + (rc + rd);
            l += rs;
            ft_st_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ft_arrive(p_full);
        }
        tc_mbar_wait(t_full, (nt - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t o[64];
        ft_ld32_nowait(lane_addr + FT_COL_O, o); ft_ld32_nowait(lane_addr + FT_COL_O + 32, o + 32);
        ft_ld_wait();
        if (qi < a.Lq) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            if (a.lse) a.lse[((long)b * a.H + h) * a.Lq + qi] = l > 0.f ? m + log2f(l) : INFINITY;
            bf16* og = a.O + b * a.o_bs + h * a.o_hs + (long)qi * a.o_rs;
#pragma unroll
            for (int i = 0; i < 64; i += 8)
                *reinterpret_cast<uint4*>(og + i) = make_uint4(pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv),
                                                               pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv),
                                                               pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv),
                                                               pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {                                             // no visible key at all (klen == 0): zeros
        const int r = (warp & 3) * 32 + lane, qi = q0 + r;
        if (qi < a.Lq) {
            if (a.lse) a.lse[((long)b * a.H + h) * a.Lq + qi] = INFINITY;
            bf16* og = a.O + b * a.o_bs + h * a.o_hs + (long)qi * a.o_rs;
#pragma unroll
            for (int i = 0; i < 64; i += 8) *reinterpret_cast<uint4*>(og + i) = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FT_TMEM_COLS) : "memory");
    }
}

inline cudaError_t launch_flash_attn_tc(const AttnParams& a, cudaStream_t stream) {
    static PerDevice pd;
    {
        const cudaError_t e = per_device_once(pd, nullptr, [] {
            return cudaFuncSetAttribute(flash_attn_tc_kernel<FT_POLY_EVERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES);
        });
        if (e != cudaSuccess) return e;
    }
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return cudaErrorInvalidValue;
    AttnTcParams p;
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
        !make(&p.tm_v, a.V, a.v_bs, a.v_hs, a.v_rs, a.Lk))
        return cudaErrorInvalidValue;
    dim3 grid(a.H, a.B, (a.Lq + FT_BM - 1) / FT_BM);
    flash_attn_tc_kernel<FT_POLY_EVERY><<<grid, FT_THREADS, FT_SMEM_BYTES, stream>>>(p);
    ++launch_counter();
    return cudaGetLastError();
}

}  // namespace tts
