// Autoregressive decode step over the KV cache (the north-star hot
// loop).  One step = 52 dependent phases (3 prenet GEMMs, 6 x [QKV, self-attn, O+res, crossQ,
// cross-attn, O+res, FFN1, FFN2+res], heads).  The same phase engine runs either
//   * one phase per launch (debug / bring-up), or
//   * as ONE persistent cooperative kernel (one 512-thread CTA per SM) that loops over steps and
//     phases with a grid barrier between phases -- no launch latency, no host round trips.
// GEMM phases are weight-streaming skinny GEMMs (M = B <= 64 per 16-row tile): weights are stored
// pre-swizzled in mma.sync B-fragment order so every lane issues 16-byte coalesced loads; the
// activation tile goes through shared memory (LayerNorm of the previous sub-layer is applied while
// it is staged -- post-LN order).  Attention phases stream K/V rows with 16-byte
// loads, 4 rows per warp request, work split evenly over ALL warps of the grid in flat (pair,row)
// space, partial (m, l, acc) merged by the last-arriving warp (deterministic order).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace tts {

constexpr int kDecThreads = 512;
constexpr int kDecWarps = 16;
constexpr int kMaxParts = 64;            // partial slots per (b, h) pair
constexpr int kDecMaxK = 2048;
constexpr int kDecSmemBytes = 16 * (kDecMaxK + 8) * 2 + kDecWarps * 16 * 8 * 4;

enum PhaseType : int { PH_GEMM = 0, PH_ATTN = 1 };
enum AKind : int { A_BF16 = 0, A_F32 = 1, A_F32_LN = 2, A_FRAME = 3 };
enum DecEpi : int { EPI_QKV = 0, EPI_F32 = 1, EPI_RESID_F32 = 2, EPI_RELU_BF16 = 3, EPI_DROP_BF16 = 4, EPI_PE_F32 = 5, EPI_HEAD = 6 };

struct PhaseDesc {
    int type;
    // ---- GEMM phase: out[b, n] = epi( A[b, :] . W[n, :] + bias[n] )
    int N, K, Kreal, nt;           // N real columns; K padded (mult of 32); Kreal valid A columns; nt n-tiles/item
    int a_kind, epi, lda, ldo, site, layer;
    const void* a;
    const float *ln_g, *ln_b;      // A_F32_LN: LayerNorm affine applied to A (K == 512)
    float* xres_out;               // A_F32_LN: normalised rows written back (the residual stream), [B][512]
    const uint4* w;                // packed B fragments [Npad/8][K/32][32] x uint4
    const float* bias;
    const float* resid;            // EPI_RESID_F32
    float* out_f32; bf16* out_bf16;
    // ---- attention phase
    const float* q;                // [B][512] fp32 (unscaled)
    const bf16 *kc, *vc;           // [B*H][Lmax][64]
    int Lmax, L_fixed;             // L_fixed == 0: self-attention over rows 0..t
    const int* lens;               // per-utterance valid rows (cross-attention key padding) or null
    bf16* attn_out;                // [B][512]
};

struct DecodeParams {
    const PhaseDesc* phases; int n_phases;
    int B, Tmax, S;
    uint64_t seed; int utt_offset;
    float dec_alpha; const float* pe;
    bf16* self_kv;                 // [layers][2][B][H][Tmax][64]
    float* mel_before;             // [B][Tmax][80]
    float* stop_logits;            // [B][Tmax]
    int* lens; int* finished; int* n_finished; int* t_done;
    float* part_acc; float* part_ml; unsigned* part_cnt;
    unsigned* barrier;
    unsigned long long* ts;        // optional [steps][64] globaltimer stamps (debug / profiling)
};

// ------------------------------------------------------------------------------------------------
TTS_D void dec_epilogue(const PhaseDesc& d, const DecodeParams& p, int t, int b, int n, float v) {
    if (d.bias) v += __ldg(d.bias + n);
    switch (d.epi) {
    case EPI_QKV:
        if (n < kDModel) d.out_f32[b * kDModel + n] = v;
        else {
            const int kv = n >= 2 * kDModel, c = n - kDModel - kv * kDModel, h = c >> 6, dd = c & 63;
            size_t idx = ((((size_t)(d.layer * 2 + kv) * p.B + b) * kHeads + h) * p.Tmax + t) * kDHead + dd;
            p.self_kv[idx] = __float2bfloat16(v);
        }
        break;
    case EPI_F32: d.out_f32[b * d.ldo + n] = v; break;
    case EPI_RESID_F32: d.out_f32[b * d.ldo + n] = v + ld_cg_f(d.resid + b * d.ldo + n); break;
    case EPI_RELU_BF16: d.out_bf16[b * d.ldo + n] = __float2bfloat16(fmaxf(v, 0.f)); break;
    case EPI_DROP_BF16: {
        v = fmaxf(v, 0.f);
        const bool keep = keep_bit(p.seed, (uint32_t)d.site, (uint32_t)t, (uint32_t)(p.utt_offset + b), (uint32_t)n);
        d.out_bf16[b * d.ldo + n] = __float2bfloat16(keep ? 2.f * v : 0.f);
    } break;
    case EPI_PE_F32: d.out_f32[b * d.ldo + n] = v + p.dec_alpha * __ldg(p.pe + (size_t)t * kDModel + n); break;
    case EPI_HEAD:
        if (n < 80) p.mel_before[((size_t)b * p.Tmax + t) * 80 + n] = v;
        else if (n == 80) {
            p.stop_logits[(size_t)b * p.Tmax + t] = v;
            if (v > 0.f && p.finished[b] == 0) {            // P10: fp32 logit > 0; firing frame counted
                p.finished[b] = 1; p.lens[b] = t + 1; atomicAdd(p.n_finished, 1);
            }
        }
        break;
    }
}

template <int NT>
__device__ __noinline__ void dec_gemm_item(const PhaseDesc& d, const DecodeParams& p, int t, int mt, int slab, unsigned char* smem) {
    constexpr int KS = kDecWarps / NT;
    const int K = d.K, lds = K + 8;
    bf16* As = reinterpret_cast<bf16*>(smem);
    float* red = reinterpret_cast<float*>(smem + 16 * (kDecMaxK + 8) * 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = mt * 16;

    // ---------------- stage the 16-row activation tile as bf16 (optionally LayerNorm'ed)
    if (d.a_kind == A_F32_LN) {                       // K == 512: warp w owns row w, lane owns 16 columns
        const int b = row0 + warp;
        float v[16];
        if (b < p.B) {
            const float* src = reinterpret_cast<const float*>(d.a) + (size_t)b * d.lda;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 x = ld_cg_f4(src + i * 128 + lane * 4);
                v[i * 4] = x.x; v[i * 4 + 1] = x.y; v[i * 4 + 2] = x.z; v[i * 4 + 3] = x.w;
            }
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) s += v[i];
            const float mean = warp_sum(s) * (1.f / 512.f);
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float dlt = v[i] - mean; ss += dlt * dlt; }
            const float rstd = rsqrtf(warp_sum(ss) * (1.f / 512.f) + 1e-5f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = i * 128 + lane * 4;
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(d.ln_g + c));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(d.ln_b + c));
                v[i * 4] = (v[i * 4] - mean) * rstd * g4.x + b4.x;
                v[i * 4 + 1] = (v[i * 4 + 1] - mean) * rstd * g4.y + b4.y;
                v[i * 4 + 2] = (v[i * 4 + 2] - mean) * rstd * g4.z + b4.z;
                v[i * 4 + 3] = (v[i * 4 + 3] - mean) * rstd * g4.w + b4.w;
            }
            if (d.xres_out) {                         // each item writes a disjoint share of the columns
                const int nslabs = (d.N + NT * 8 - 1) / (NT * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int grp = i * 32 + lane;    // float4 group index 0..127
                    if (grp % nslabs == slab)         // nslabs > 128: slabs 0..127 cover all 128 groups
                        *reinterpret_cast<float4*>(d.xres_out + (size_t)b * kDModel + grp * 4) =
                            make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint2 u = make_uint2(pack_bf16x2(v[i * 4], v[i * 4 + 1]), pack_bf16x2(v[i * 4 + 2], v[i * 4 + 3]));
            *reinterpret_cast<uint2*>(As + warp * lds + i * 128 + lane * 4) = u;
        }
    } else if (d.a_kind == A_BF16) {                  // cp.async.cg: L2-coherent, all chunks in flight at once
        const int cpr = K >> 3;                       // 16-byte chunks per row
        const bf16* src = reinterpret_cast<const bf16*>(d.a);
        for (int c = tid; c < 16 * cpr; c += kDecThreads) {
            const int r = c / cpr, ch = c - r * cpr, b = row0 + r;
            const bool ok = b < p.B && ch * 8 < d.Kreal;
            cp_async_16(As + r * lds + ch * 8, src + (ok ? ((size_t)b * d.lda + ch * 8) : 0), ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
    } else {                                          // A_F32 / A_FRAME: fp32 rows -> bf16 (K <= 512: <= 4 chunks/thread)
        const int cpr = K >> 2;                       // float4 chunks per row
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                 // issue every load before the first use
            const int c = tid + i * kDecThreads;
            const int r = c / cpr, ch = c - r * cpr, b = row0 + r;
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < 16 * cpr && b < p.B && ch * 4 < d.Kreal) {
                if (d.a_kind == A_FRAME) {            // previous frame (fp32 feedback, P8); zero go-frame at t = 0
                    if (t > 0) x[i] = ld_cg_f4(p.mel_before + ((size_t)b * p.Tmax + (t - 1)) * 80 + ch * 4);
                } else {
                    x[i] = ld_cg_f4(reinterpret_cast<const float*>(d.a) + (size_t)b * d.lda + ch * 4);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * kDecThreads;
            const int r = c / cpr, ch = c - r * cpr;
            if (c < 16 * cpr)
                *reinterpret_cast<uint2*>(As + r * lds + ch * 4) = make_uint2(pack_bf16x2(x[i].x, x[i].y), pack_bf16x2(x[i].z, x[i].w));
        }
    }
    __syncthreads();

    // ---------------- MMA: warp (nti, ks) owns n-tile nti of the slab and K-slice ks
    const int nti = warp % NT, ks = warp / NT;
    const int kp_total = K >> 5, kp_per = kp_total / KS;          // k-pairs (32 columns) per warp
    const int ntile_g = slab * NT + nti;
    const uint4* wp = d.w + ((size_t)ntile_g * kp_total + (size_t)ks * kp_per) * 32 + lane;
    const bf16* arow = As + (lane & 15) * lds + (lane >> 4) * 8 + ks * kp_per * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const uint64_t wpol = l2_policy_evict_last();
    for (int i0 = 0; i0 < kp_per; i0 += 8) {
        uint4 wreg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j < kp_per) wreg[j] = ld_weight_u4(wp + (size_t)(i0 + j) * 32, wpol);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j < kp_per) {
                uint32_t af[4];
                ldmatrix_x4(af, arow + (i0 + j) * 32);
                mma_bf16_16816(acc, af, wreg[j].x, wreg[j].y);
                ldmatrix_x4(af, arow + (i0 + j) * 32 + 16);
                mma_bf16_16816(acc, af, wreg[j].z, wreg[j].w);
            }
    }
    {
        const int g = lane >> 2, t4 = lane & 3;
        float* r = red + warp * 128;
        *reinterpret_cast<float2*>(r + g * 8 + t4 * 2) = make_float2(acc[0], acc[1]);
        *reinterpret_cast<float2*>(r + (g + 8) * 8 + t4 * 2) = make_float2(acc[2], acc[3]);
    }
    __syncthreads();
    constexpr int NS = NT * 8;
    for (int o = tid; o < 16 * NS; o += kDecThreads) {
        const int r = o / NS, c = o - r * NS, ntl = c >> 3, cc = c & 7;
        float v = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < KS; ++k2) v += red[((k2 * NT + ntl) * 16 + r) * 8 + cc];   // fixed order: deterministic
        const int n = slab * NS + c, b = row0 + r;
        if (b < p.B && n < d.N) dec_epilogue(d, p, t, b, n, v);
    }
    __syncthreads();
}

TTS_D void dec_gemm_phase(const PhaseDesc& d, const DecodeParams& p, int t, unsigned char* smem) {
    const int mtiles = (p.B + 15) >> 4;
    const int nslabs = (d.N + d.nt * 8 - 1) / (d.nt * 8);
    const int items = mtiles * nslabs;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int mt = item % mtiles, slab = item / mtiles;
        switch (d.nt) {
        case 1: dec_gemm_item<1>(d, p, t, mt, slab, smem); break;
        case 2: dec_gemm_item<2>(d, p, t, mt, slab, smem); break;
        case 4: dec_gemm_item<4>(d, p, t, mt, slab, smem); break;
        case 8: dec_gemm_item<8>(d, p, t, mt, slab, smem); break;
        default: dec_gemm_item<16>(d, p, t, mt, slab, smem); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void dec_attn_phase(const PhaseDesc& d, const DecodeParams& p, int t) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = gridDim.x * kDecWarps, gw = blockIdx.x * kDecWarps + warp;
    const int P = p.B * kHeads;
    const int L = d.L_fixed ? d.L_fixed : t + 1;
    const long total = (long)P * L;
    int per = (int)((total + W - 1) / W);
    per = max(per, 16);
    per = max(per, (L + kMaxParts - 3) / (kMaxParts - 2));
    per = (per + 3) & ~3;
    long f0 = (long)gw * per;
    const long f1 = min(f0 + (long)per, total);
    const int g4 = lane >> 3, sub = lane & 7;
    const float qscale = 0.125f * kLog2e;
    const uint64_t kvpol = l2_policy_evict_first();
    while (f0 < f1) {
        const int pr = (int)(f0 / L), row = (int)(f0 - (long)pr * L);
        const int n = min(L - row, (int)(f1 - f0));
        const int b = pr / kHeads;
        const int vlen = d.lens ? min(L, __ldg(d.lens + b)) : L;
        const int rend = min(row + n, vlen);
        float q[8];
        {
            const float* qp = d.q + (size_t)pr * kDHead + sub * 8;
            const float4 a = ld_cg_f4(qp), c = ld_cg_f4(qp + 4);
            q[0] = a.x * qscale; q[1] = a.y * qscale; q[2] = a.z * qscale; q[3] = a.w * qscale;
            q[4] = c.x * qscale; q[5] = c.y * qscale; q[6] = c.z * qscale; q[7] = c.w * qscale;
        }
        const bf16* kb = d.kc + (size_t)pr * d.Lmax * kDHead + sub * 8;
        const bf16* vb = d.vc + (size_t)pr * d.Lmax * kDHead + sub * 8;
        float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int r0 = row; r0 < rend; r0 += 16) {
            uint4 kk[4], vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 4 + g4;
                if (r < rend) { kk[u] = ld_stream_u4(kb + (size_t)r * kDHead, kvpol); vv[u] = ld_stream_u4(vb + (size_t)r * kDHead, kvpol); }
                else { kk[u] = make_uint4(0, 0, 0, 0); vv[u] = make_uint4(0, 0, 0, 0); }
            }
            float s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 k0 = unpack_bf16x2(kk[u].x), k1 = unpack_bf16x2(kk[u].y);
                const float2 k2 = unpack_bf16x2(kk[u].z), k3 = unpack_bf16x2(kk[u].w);
                float ps = q[0] * k0.x + q[1] * k0.y + q[2] * k1.x + q[3] * k1.y + q[4] * k2.x + q[5] * k2.y + q[6] * k3.x + q[7] * k3.y;
                ps += __shfl_xor_sync(0xffffffffu, ps, 1);
                ps += __shfl_xor_sync(0xffffffffu, ps, 2);
                ps += __shfl_xor_sync(0xffffffffu, ps, 4);
                s[u] = (r0 + u * 4 + g4 < rend) ? ps : -INFINITY;
            }
            const float mnew = fmaxf(m, fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])));
            if (mnew > -INFINITY) {
                const float sc = (m == -INFINITY) ? 0.f : exp2f(m - mnew);
                l *= sc;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] *= sc;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float pu = exp2f(s[u] - mnew);
                    l += pu;
                    const float2 v0 = unpack_bf16x2(vv[u].x), v1 = unpack_bf16x2(vv[u].y);
                    const float2 v2 = unpack_bf16x2(vv[u].z), v3 = unpack_bf16x2(vv[u].w);
                    acc[0] += pu * v0.x; acc[1] += pu * v0.y; acc[2] += pu * v1.x; acc[3] += pu * v1.y;
                    acc[4] += pu * v2.x; acc[5] += pu * v2.y; acc[6] += pu * v3.x; acc[7] += pu * v3.y;
                }
                m = mnew;
            }
        }
        // merge the 4 row-groups of the warp (butterfly: every lane ends with the full result)
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, m, off), lo = __shfl_xor_sync(0xffffffffu, l, off);
            const float mn = fmaxf(m, mo);
            const float e1 = (m == -INFINITY) ? 0.f : exp2f(m - mn);
            const float e2 = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
            l = l * e1 + lo * e2;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float ao = __shfl_xor_sync(0xffffffffu, acc[j], off);
                acc[j] = acc[j] * e1 + ao * e2;
            }
            m = mn;
        }
        const int w_lo = (int)(((long)pr * L) / per), w_hi = (int)((((long)pr + 1) * L - 1) / per);
        const int count = w_hi - w_lo + 1;
        if (count == 1) {
            if (lane < 8) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
                uint4 o = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                                     pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
                *reinterpret_cast<uint4*>(d.attn_out + (size_t)pr * kDHead + sub * 8) = o;
            }
        } else {
            const int slot = pr * kMaxParts + (gw - w_lo);
            if (lane < 8) {
                float* pa = p.part_acc + (size_t)slot * kDHead + sub * 8;
                *reinterpret_cast<float4*>(pa) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                *reinterpret_cast<float4*>(pa + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                if (lane == 0) { p.part_ml[slot * 2] = m; p.part_ml[slot * 2 + 1] = l; }
            }
            __threadfence();
            __syncwarp();
            unsigned old = 0;
            if (lane == 0) old = atomicAdd(p.part_cnt + pr, 1u);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == (unsigned)(count - 1)) {        // last arriver merges, in slot order
                __threadfence();
                const int base = pr * kMaxParts;
                float mm = -INFINITY;
                for (int i = lane; i < count; i += 32) mm = fmaxf(mm, ld_cg_f(p.part_ml + (base + i) * 2));
                mm = warp_max(mm);
                float ls = 0.f, ox = 0.f, oy = 0.f;
                for (int i = 0; i < count; ++i) {
                    const float2 ml = ld_cg_f2(p.part_ml + (base + i) * 2);
                    const float e = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - mm);
                    const float2 a = ld_cg_f2(p.part_acc + (size_t)(base + i) * kDHead + lane * 2);
                    ls += ml.y * e; ox += a.x * e; oy += a.y * e;
                }
                const float inv = ls > 0.f ? 1.f / ls : 0.f;
                *reinterpret_cast<uint32_t*>(d.attn_out + (size_t)pr * kDHead + lane * 2) = pack_bf16x2(ox * inv, oy * inv);
                if (lane == 0) p.part_cnt[pr] = 0u;
            }
        }
        f0 += n;
    }
}

// ------------------------------------------------------------------------------------------------
TTS_D void grid_barrier(unsigned* bar, unsigned& target) {
    __syncthreads();                                   // every thread's writes happen-before thread 0's release
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// persistent == 1: launched cooperatively (all CTAs co-resident), barrier between phases.
// persistent == 0: one phase per launch (ph_begin + 1 == ph_end, n_steps == 1); stream order is the barrier.
__global__ void __launch_bounds__(kDecThreads, 1)
decode_kernel(const DecodeParams p, int t0, int n_steps, int ph_begin, int ph_end, int persistent) {
    extern __shared__ __align__(16) unsigned char dec_smem[];
    unsigned target = 0;
    for (int step = 0; step < n_steps; ++step) {
        const int t = t0 + step;
        for (int ph = ph_begin; ph < ph_end; ++ph) {
            const PhaseDesc& d = p.phases[ph];
            if (d.type == PH_GEMM) dec_gemm_phase(d, p, t, dec_smem);
            else dec_attn_phase(d, p, t);
            if (persistent) {
                grid_barrier(p.barrier, target);
                if (p.ts && blockIdx.x == 0 && threadIdx.x == 0) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                    p.ts[(size_t)step * 64 + ph] = now;
                }
            }
        }
        if (persistent) {
            if (blockIdx.x == 0 && threadIdx.x == 0) *p.t_done = t + 1;
            if (ld_cg_i(p.n_finished) >= p.B) break;      // uniform: written before the last barrier
        }
    }
}

}  // namespace tts
