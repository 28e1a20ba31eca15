// Cluster-partitioned autoregressive decode: the north-star hot loop.
//
// Utterances are independent end to end, so the batch is cut into groups of 8 utterances and each
// group is decoded by ONE thread-block cluster of 16 CTAs (one cluster per GPC, 8 clusters = 128 SMs
// for B = 64) with no grid-wide synchronisation at all:
//   * every GEMM of the step is split over the 16 CTAs along N (FFN2 along K); the 8 utterances are
//     the MMA n = 8 dimension ("swap-AB": weights are the 16x16 A operand of mma.sync.m16n8k16, the
//     activations the 16x8 B operand), so each weight byte is read once per cluster;
//   * results are exchanged through distributed shared memory (st.shared::cluster) and ordered by
//     barrier.cluster (release/acquire) -- about 0.2 us instead of a multi-us grid barrier;
//   * LayerNorm, residuals, dropout, PE and the stop test run on the gathered rows in shared memory;
//   * ALL global traffic of a CTA -- its 1/16 slice of the weights and the K/V rows of its 4
//     (utterance, head) pairs -- is one ordered stream of cp.async.bulk copies into a 4 x 32 KB
//     shared-memory ring guarded by mbarriers, issued up to 3 chunks ahead of use, so HBM/L2 latency
//     is hidden across the dependent phases of the step.
// Weights are packed per CTA rank in exactly the order the step consumes them (tts_b200.cu:
// pack_cluster_weights), each 16(n) x 32(k) block in A-fragment order.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace tts {

constexpr int CL_SIZE = 16, CL_THREADS = 512, CL_WARPS = 16, CL_G = 8;
constexpr int CL_STAGES = 4, CL_STAGE_BYTES = 32768, CL_KV_ROWS = 128, CL_BLOCK_BYTES = 1024;

// byte sizes of one rank's packed weight segments
constexpr int CLW_FC1 = 3 * 1024, CLW_FC2 = 8 * 1024, CLW_PROJ = 16 * 1024;
constexpr int CLW_QKV = 96 * 1024, CLW_O = 32 * 1024, CLW_W1 = 128 * 1024, CLW_W2 = 128 * 1024, CLW_HEAD = 16 * 1024;
constexpr int CLW_LAYER = CLW_QKV + 3 * CLW_O + CLW_W1 + CLW_W2;
constexpr size_t CLW_RANK_BYTES = (size_t)CLW_FC1 + CLW_FC2 + CLW_PROJ + 6 * (size_t)CLW_LAYER + CLW_HEAD;

// shared memory carve-up (bytes)
constexpr int SM_RING = 0;
constexpr int SM_XRES = SM_RING + CL_STAGES * CL_STAGE_BYTES;   // f32 [8][512]
constexpr int SM_YBUF = SM_XRES + 16384;                        // f32 [8][512]   (gathered)
constexpr int SM_RECV = SM_YBUF + 16384;                        // f32 [16][8][32] (FFN2 reduce-scatter)
constexpr int SM_RED = SM_RECV + 16384;                         // f32 [16][128]
constexpr int SM_XA = SM_RED + 8192;                            // bf16 [8][520]
constexpr int SM_ABUF = SM_XA + 8320;                           // bf16 [8][520]  (gathered)
constexpr int SM_QKV = SM_ABUF + 8320;                          // f32 [8][192]
constexpr int SM_HBUF = SM_QKV + 6144;                          // bf16 [8][136]
constexpr int SM_H1 = SM_HBUF + 2176;                           // bf16 [8][264]  (gathered)
constexpr int SM_H2 = SM_H1 + 4224;                             // bf16 [8][264]  (gathered)
constexpr int SM_FBUF = SM_H2 + 4224;                           // bf16 [8][104]  (gathered)
constexpr int SM_AMERGE = SM_FBUF + 1664;                       // f32 [16][68]
constexpr int SM_MISC = SM_AMERGE + 4352;                       // mbarriers + flags
constexpr int CL_SMEM_BYTES = SM_MISC + 128;
constexpr int LDX512 = 520, LDX256 = 264, LDX128 = 136, LDX96 = 104;

struct ClusterLayerParams {
    const float *bqkv, *bo, *bq2, *bo2, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b, *ln3g, *ln3b;
};
struct ClusterParams {
    int B, Tmax, S, ngroups;
    uint64_t seed; int utt_offset; float dec_alpha; const float* pe;
    const unsigned char* wpack;                  // [16][CLW_RANK_BYTES]
    const float *b_fc1, *b_fc2, *b_proj, *b_head;
    ClusterLayerParams layer[6];
    bf16* self_kv;                               // [6][2][B][8][Tmax][64]
    const bf16* cross_kv;                        // [6][2][B][8][S][64]
    const int* plens;
    float* mel_before; float* stop_logits; int* lens; int* finished; int* n_finished;
};

// ---------------------------------------------------------------- PTX helpers
TTS_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TTS_D uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TTS_D uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
TTS_D uint32_t cluster_nid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
TTS_D void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TTS_D uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
TTS_D void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
TTS_D void st_cluster_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
TTS_D void st_cluster_b16(uint32_t addr, bf16 v) {
    unsigned short u = *reinterpret_cast<unsigned short*>(&v);
    asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(u) : "memory");
}
TTS_D void red_cluster_add_u32(uint32_t addr, uint32_t v) { asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
TTS_D void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
TTS_D void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
TTS_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TTS_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
TTS_D void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
TTS_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ---------------------------------------------------------------- per-CTA context
struct ClCtx {
    unsigned char* smem;
    uint64_t *full, *empty;
    int rank, head, half, tid, warp, lane;
    int b0, G, r_lo, np;                 // group base utterance, rows in group, my attention rows [r_lo, r_lo+np)
    uint32_t consumed;                   // chunks consumed (uniform over the CTA)
    // producer state (meaningful in thread 0 only)
    uint32_t issued; int pt, pseg, pi; size_t pwoff; int t_end;
    uint64_t pol_w, pol_kv;
};

// The ordered stream of chunks of one step for this rank.  seg: 0 fc1, 1 fc2, 2 proj,
// 3+8l+{0 qkv, 1 self-KV, 2 o, 3 q2, 4 cross-KV, 5 o2, 6 w1, 7 w2}, 51 head.
TTS_D int seg_chunks(const ClusterParams& p, const ClCtx& c, int seg, int t) {
    if (seg < 3) return 1;
    if (seg == 51) return c.rank < 6 ? 1 : 0;
    switch ((seg - 3) & 7) {
    case 0: return 3;
    case 1: return c.np * ((t + CL_KV_ROWS - 1) / CL_KV_ROWS);
    case 4: return c.np * ((p.S + CL_KV_ROWS - 1) / CL_KV_ROWS);
    case 6: case 7: return 4;
    default: return 1;
    }
}
TTS_D uint32_t seg_weight_bytes(int seg) {          // bytes per chunk of a weight segment
    if (seg == 0) return CLW_FC1;
    if (seg == 1) return CLW_FC2;
    if (seg == 2) return CLW_PROJ;
    if (seg == 51) return CLW_HEAD;
    return CL_STAGE_BYTES;
}

// thread 0: issue chunks while a ring slot is free (never blocks unless `must` chunks are required)
TTS_D void cl_pump(const ClusterParams& p, ClCtx& c, uint32_t need_issued) {
    while (c.pt < c.t_end) {
        if (c.issued >= c.consumed + CL_STAGES) break;
        const int stage = c.issued % CL_STAGES;
        const uint32_t use = c.issued / CL_STAGES;
        if (use > 0) {                                   // slot's previous chunk must be released by all 16 warps
            if (c.issued < need_issued) mbar_wait(&c.empty[stage], (use & 1) ^ 1);
            else if (!mbar_try_wait(&c.empty[stage], (use & 1) ^ 1)) break;
        }
        // skip empty segments / advance to the next step
        while (c.pi >= seg_chunks(p, c, c.pseg, c.pt)) {
            c.pi = 0;
            if (++c.pseg > 51) { c.pseg = 0; c.pwoff = 0; ++c.pt; }
            if (c.pt >= c.t_end) return;
        }
        unsigned char* dst = c.smem + SM_RING + stage * CL_STAGE_BYTES;
        const int seg = c.pseg, sub = seg < 3 || seg == 51 ? -1 : ((seg - 3) & 7);
        if (sub == 1 || sub == 4) {                      // K rows + V rows of one (utterance, head) pair
            const int l = (seg - 3) >> 3;
            const int L = sub == 1 ? c.pt : p.S, Lmax = sub == 1 ? p.Tmax : p.S;
            const int nck = (L + CL_KV_ROWS - 1) / CL_KV_ROWS;
            const int pr = c.pi / nck, ci = c.pi - pr * nck;
            const int b = c.b0 + c.r_lo + pr;
            const int row0 = ci * CL_KV_ROWS, nrows = min(CL_KV_ROWS, L - row0);
            const bf16* base = sub == 1 ? p.self_kv : p.cross_kv;
            const size_t kidx = ((((size_t)(l * 2) * p.B + b) * kHeads + c.head) * Lmax + row0) * kDHead;
            const size_t vidx = ((((size_t)(l * 2 + 1) * p.B + b) * kHeads + c.head) * Lmax + row0) * kDHead;
            const uint32_t bytes = (uint32_t)nrows * 128u;
            asm volatile("fence.proxy.async;" ::: "memory");   // rows written with st.global earlier in this launch
            mbar_expect_tx(&c.full[stage], 2 * bytes);
            bulk_g2s(dst, base + kidx, bytes, &c.full[stage], c.pol_kv);
            bulk_g2s(dst + CL_KV_ROWS * 128, base + vidx, bytes, &c.full[stage], c.pol_kv);
        } else {
            const uint32_t bytes = seg_weight_bytes(seg);
            mbar_expect_tx(&c.full[stage], bytes);
            bulk_g2s(dst, p.wpack + (size_t)c.rank * CLW_RANK_BYTES + c.pwoff, bytes, &c.full[stage], c.pol_w);
            c.pwoff += bytes;
        }
        ++c.pi;
        ++c.issued;
    }
}

TTS_D unsigned char* cl_acquire(const ClusterParams& p, ClCtx& c) {
    if (c.tid == 0) cl_pump(p, c, c.consumed + 1);
    const int stage = c.consumed % CL_STAGES;
    mbar_wait(&c.full[stage], (c.consumed / CL_STAGES) & 1);
    return c.smem + SM_RING + stage * CL_STAGE_BYTES;
}
TTS_D void cl_release(ClCtx& c) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[c.consumed % CL_STAGES]);
    ++c.consumed;
}

// ---------------------------------------------------------------- GEMM over ring chunks
// out[col = tile*16 + n][m] = sum_k W[col][k] * X[m][k];  nblocks blocks of 16(n) x 32(k) per chunk in
// [tile][kp] order, KP k-pairs per tile, X rows are bf16 with stride ldx, first k-pair of the segment kp0.
template <class Epi>
TTS_D void cl_gemm(const ClusterParams& p, ClCtx& c, int nchunks, int nblocks, int KP, const bf16* X, int ldx, Epi epi) {
    float* red = reinterpret_cast<float*>(c.smem + SM_RED);
    const int tpc = nblocks / KP;                                   // tiles per chunk
    for (int ch = 0; ch < nchunks; ++ch) {
        const unsigned char* st = cl_acquire(p, c);
        const int bi0 = c.warp * 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int bi = bi0 + j;
            if (bi < nblocks) {
                const int kp = bi % KP;
                const uint4 w0 = reinterpret_cast<const uint4*>(st)[(bi * 2) * 32 + c.lane];
                const uint4 w1 = reinterpret_cast<const uint4*>(st)[(bi * 2 + 1) * 32 + c.lane];
                uint32_t bfrag[4];
                ldmatrix_x4(bfrag, X + (c.lane & 7) * ldx + kp * 32 + (c.lane >> 3) * 8);
                const uint32_t a0[4] = {w0.x, w0.y, w0.z, w0.w}, a1[4] = {w1.x, w1.y, w1.z, w1.w};
                mma_bf16_16816(acc, a0, bfrag[0], bfrag[1]);
                mma_bf16_16816(acc, a1, bfrag[2], bfrag[3]);
            }
        }
        cl_release(c);
        if (bi0 < nblocks) {
            const int g = c.lane >> 2, t4 = c.lane & 3;
            float* r = red + c.warp * 128;
            *reinterpret_cast<float2*>(r + g * 8 + t4 * 2) = make_float2(acc[0], acc[1]);
            *reinterpret_cast<float2*>(r + (g + 8) * 8 + t4 * 2) = make_float2(acc[2], acc[3]);
        }
        __syncthreads();
        for (int o = c.tid; o < tpc * 128; o += CL_THREADS) {
            const int ti = o >> 7, n = (o >> 3) & 15, m = o & 7;
            const int w_lo = (ti * KP) >> 1, w_hi = ((ti + 1) * KP - 1) >> 1;
            float v = 0.f;
            for (int w = w_lo; w <= w_hi; ++w) v += red[w * 128 + n * 8 + m];   // fixed order: deterministic
            if (m < c.G) epi(ch * tpc + ti, n, m, v);
        }
        __syncthreads();
    }
}

// write a value into the same shared-memory location of every CTA of the cluster
TTS_D void bcast_f32(uint32_t local_addr, float v) {
#pragma unroll
    for (int r = 0; r < CL_SIZE; ++r) st_cluster_f32(map_to_rank(local_addr, r), v);
}
TTS_D void bcast_b16(uint32_t local_addr, bf16 v) {
#pragma unroll
    for (int r = 0; r < CL_SIZE; ++r) st_cluster_b16(map_to_rank(local_addr, r), v);
}

// LayerNorm of the gathered rows: ybuf -> xres (f32) + xa (bf16); warp m < G owns row m.
TTS_D void cl_layernorm(ClCtx& c, const float* g, const float* b) {
    if (c.warp < c.G) {
        const float* y = reinterpret_cast<const float*>(c.smem + SM_YBUF) + c.warp * 512;
        float* xr = reinterpret_cast<float*>(c.smem + SM_XRES) + c.warp * 512;
        bf16* xa = reinterpret_cast<bf16*>(c.smem + SM_XA) + c.warp * LDX512;
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(y + i * 128 + c.lane * 4);
            v[i * 4] = x.x; v[i * 4 + 1] = x.y; v[i * 4 + 2] = x.z; v[i * 4 + 3] = x.w;
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += v[i];
        const float mean = warp_sum(s) * (1.f / 512.f);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; ss += d * d; }
        const float rstd = rsqrtf(warp_sum(ss) * (1.f / 512.f) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int col = i * 128 + c.lane * 4;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + col)), b4 = __ldg(reinterpret_cast<const float4*>(b + col));
            const float o0 = (v[i * 4] - mean) * rstd * g4.x + b4.x, o1 = (v[i * 4 + 1] - mean) * rstd * g4.y + b4.y;
            const float o2 = (v[i * 4 + 2] - mean) * rstd * g4.z + b4.z, o3 = (v[i * 4 + 3] - mean) * rstd * g4.w + b4.w;
            *reinterpret_cast<float4*>(xr + col) = make_float4(o0, o1, o2, o3);
            *reinterpret_cast<uint2*>(xa + col) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
    }
    __syncthreads();
}

// Attention of this CTA's pairs (rows r_lo..r_lo+np of head `head`) over K/V chunks from the ring.
// self: rows 0..t-1 from the cache + the newest row (k_t, v_t) from qkvbuf.  cross: rows 0..len-1.
TTS_D void cl_attention(const ClusterParams& p, ClCtx& c, bool self, int t) {
    const float* qkv = reinterpret_cast<const float*>(c.smem + SM_QKV);
    float* am = reinterpret_cast<float*>(c.smem + SM_AMERGE);
    const int g4 = c.lane >> 3, sub = c.lane & 7;
    const float qscale = 0.125f * kLog2e;
    const int L = self ? t : p.S;
    const int nck = (L + CL_KV_ROWS - 1) / CL_KV_ROWS;
    for (int pi = 0; pi < c.np; ++pi) {
        const int gi = c.r_lo + pi;
        const int vlen = self ? L : min(L, __ldg(p.plens + c.b0 + gi));
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = qkv[gi * 192 + sub * 8 + j] * qscale;
        float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int ci = 0; ci < nck; ++ci) {
            const unsigned char* st = cl_acquire(p, c);
            const int nrows = min(CL_KV_ROWS, vlen - ci * CL_KV_ROWS);
            uint4 kk[2], vv[2];
            float s[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = c.warp * 8 + u * 4 + g4;
                if (r < nrows) {
                    kk[u] = *reinterpret_cast<const uint4*>(st + r * 128 + sub * 16);
                    vv[u] = *reinterpret_cast<const uint4*>(st + CL_KV_ROWS * 128 + r * 128 + sub * 16);
                } else { kk[u] = make_uint4(0, 0, 0, 0); vv[u] = make_uint4(0, 0, 0, 0); }
            }
            cl_release(c);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float2 k0 = unpack_bf16x2(kk[u].x), k1 = unpack_bf16x2(kk[u].y), k2 = unpack_bf16x2(kk[u].z), k3 = unpack_bf16x2(kk[u].w);
                float ps = q[0] * k0.x + q[1] * k0.y + q[2] * k1.x + q[3] * k1.y + q[4] * k2.x + q[5] * k2.y + q[6] * k3.x + q[7] * k3.y;
                ps += __shfl_xor_sync(0xffffffffu, ps, 1);
                ps += __shfl_xor_sync(0xffffffffu, ps, 2);
                ps += __shfl_xor_sync(0xffffffffu, ps, 4);
                s[u] = (c.warp * 8 + u * 4 + g4 < nrows) ? ps : -INFINITY;
            }
            const float mnew = fmaxf(m, fmaxf(s[0], s[1]));
            if (mnew > -INFINITY) {
                const float sc = (m == -INFINITY) ? 0.f : exp2f(m - mnew);
                l *= sc;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] *= sc;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float pu = exp2f(s[u] - mnew);
                    l += pu;
                    const float2 v0 = unpack_bf16x2(vv[u].x), v1 = unpack_bf16x2(vv[u].y), v2 = unpack_bf16x2(vv[u].z), v3 = unpack_bf16x2(vv[u].w);
                    acc[0] += pu * v0.x; acc[1] += pu * v0.y; acc[2] += pu * v1.x; acc[3] += pu * v1.y;
                    acc[4] += pu * v2.x; acc[5] += pu * v2.y; acc[6] += pu * v3.x; acc[7] += pu * v3.y;
                }
                m = mnew;
            }
        }
        if (self && c.warp == 0 && g4 == 0) {            // newest row: bf16-rounded k_t, v_t as the cache stores them
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) ps += q[j] * __bfloat162float(__float2bfloat16(qkv[gi * 192 + 64 + sub * 8 + j]));
            ps += __shfl_xor_sync(0x000000ffu, ps, 1);
            ps += __shfl_xor_sync(0x000000ffu, ps, 2);
            ps += __shfl_xor_sync(0x000000ffu, ps, 4);
            const float mnew = fmaxf(m, ps);
            const float sc = (m == -INFINITY) ? 0.f : exp2f(m - mnew), pu = exp2f(ps - mnew);
            l = l * sc + pu;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = acc[j] * sc + pu * __bfloat162float(__float2bfloat16(qkv[gi * 192 + 128 + sub * 8 + j]));
            m = mnew;
        }
        // merge the 4 row-groups of the warp, then the 16 warps through shared memory
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, m, off), lo = __shfl_xor_sync(0xffffffffu, l, off);
            const float mn = fmaxf(m, mo);
            const float e1 = (m == -INFINITY) ? 0.f : exp2f(m - mn), e2 = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
            l = l * e1 + lo * e2;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float ao = __shfl_xor_sync(0xffffffffu, acc[j], off); acc[j] = acc[j] * e1 + ao * e2; }
            m = mn;
        }
        if (c.lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) am[c.warp * 68 + sub * 8 + j] = acc[j];
            if (c.lane == 0) { am[c.warp * 68 + 64] = m; am[c.warp * 68 + 65] = l; }
        }
        __syncthreads();
        if (c.tid < 64) {
            float mm = -INFINITY;
#pragma unroll
            for (int w = 0; w < CL_WARPS; ++w) mm = fmaxf(mm, am[w * 68 + 64]);
            float ls = 0.f, o = 0.f;
#pragma unroll
            for (int w = 0; w < CL_WARPS; ++w) {
                const float mw = am[w * 68 + 64];
                const float e = (mw == -INFINITY) ? 0.f : exp2f(mw - mm);
                ls += am[w * 68 + 65] * e; o += am[w * 68 + c.tid] * e;
            }
            const float out = ls > 0.f ? o / ls : 0.f;
            bf16* ab = reinterpret_cast<bf16*>(c.smem + SM_ABUF) + gi * LDX512 + c.head * 64 + c.tid;
            bcast_b16(smem_u32(ab), __float2bfloat16(out));
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(CL_THREADS, 1) decode_cluster_kernel(const ClusterParams p, int t0, int n_steps) {
    extern __shared__ __align__(128) unsigned char cl_smem[];
    ClCtx c;
    c.smem = cl_smem;
    c.full = reinterpret_cast<uint64_t*>(cl_smem + SM_MISC);
    c.empty = c.full + CL_STAGES;
    volatile int* flags = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 64);   // [0] finished utterances of the group
    c.rank = (int)cluster_ctarank(); c.head = c.rank >> 1; c.half = c.rank & 1;
    c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    c.pol_w = l2_policy_evict_last(); c.pol_kv = l2_policy_evict_first();
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_nid_x();
    const uint32_t partner = (uint32_t)(c.rank ^ 1);

    float* xres = reinterpret_cast<float*>(cl_smem + SM_XRES);
    float* ybuf = reinterpret_cast<float*>(cl_smem + SM_YBUF);
    float* recv = reinterpret_cast<float*>(cl_smem + SM_RECV);
    bf16* xa = reinterpret_cast<bf16*>(cl_smem + SM_XA);
    bf16* abuf = reinterpret_cast<bf16*>(cl_smem + SM_ABUF);
    float* qkvb = reinterpret_cast<float*>(cl_smem + SM_QKV);
    bf16* hbuf = reinterpret_cast<bf16*>(cl_smem + SM_HBUF);
    bf16* h1 = reinterpret_cast<bf16*>(cl_smem + SM_H1);
    bf16* h2 = reinterpret_cast<bf16*>(cl_smem + SM_H2);
    bf16* fbuf = reinterpret_cast<bf16*>(cl_smem + SM_FBUF);

    for (int grp = cid; grp < p.ngroups; grp += ncl) {
        c.b0 = grp * CL_G; c.G = min(CL_G, p.B - c.b0);
        const int hsplit = (c.G + 1) >> 1;
        c.r_lo = c.half ? hsplit : 0; c.np = c.half ? c.G - hsplit : hsplit;
        // ---- (re)initialise the pipeline and the activation buffers
        for (int i = c.tid; i < (CL_SMEM_BYTES - SM_XRES) / 4; i += CL_THREADS) reinterpret_cast<uint32_t*>(cl_smem + SM_XRES)[i] = 0u;
        __syncthreads();
        if (c.tid == 0) {
            for (int s = 0; s < CL_STAGES; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], CL_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            int nf = 0;
            for (int m = 0; m < c.G; ++m) nf += p.finished[c.b0 + m];
            flags[0] = nf;
        }
        if (t0 > 0) {                                   // resume: previous frame from global memory (fp32 -> bf16)
            for (int i = c.tid; i < c.G * 80; i += CL_THREADS) {
                const int m = i / 80, col = i - m * 80;
                fbuf[m * LDX96 + col] = __float2bfloat16(p.mel_before[((size_t)(c.b0 + m) * p.Tmax + (t0 - 1)) * 80 + col]);
            }
        }
        c.consumed = 0; c.issued = 0; c.pt = t0; c.pseg = 0; c.pi = 0; c.pwoff = 0; c.t_end = t0 + n_steps;
        __syncthreads();
        cluster_sync_all();                             // peers' buffers are initialised before any DSMEM write
        if (flags[0] >= c.G) { cluster_sync_all(); continue; }

        for (int t = t0; t < t0 + n_steps; ++t) {
            // ================= decoder prenet (dropout always on, P7) =================
            cl_gemm(p, c, 1, 3, 3, fbuf, LDX96, [&](int, int n, int m, float v) {
                const int col = c.rank * 16 + n;
                v = fmaxf(v + __ldg(p.b_fc1 + col), 0.f);
                v = keep_bit(p.seed, SITE_DEC_PRENET_FC1, (uint32_t)t, (uint32_t)(p.utt_offset + c.b0 + m), (uint32_t)col) ? 2.f * v : 0.f;
                bcast_b16(smem_u32(h1 + m * LDX256 + col), __float2bfloat16(v));
            });
            cluster_sync_all();
            cl_gemm(p, c, 1, 8, 8, h1, LDX256, [&](int, int n, int m, float v) {
                const int col = c.rank * 16 + n;
                v = fmaxf(v + __ldg(p.b_fc2 + col), 0.f);
                v = keep_bit(p.seed, SITE_DEC_PRENET_FC2, (uint32_t)t, (uint32_t)(p.utt_offset + c.b0 + m), (uint32_t)col) ? 2.f * v : 0.f;
                bcast_b16(smem_u32(h2 + m * LDX256 + col), __float2bfloat16(v));
            });
            cluster_sync_all();
            cl_gemm(p, c, 1, 16, 8, h2, LDX256, [&](int ti, int n, int m, float v) {
                const int col = c.rank * 32 + ti * 16 + n;
                v += __ldg(p.b_proj + col) + p.dec_alpha * __ldg(p.pe + (size_t)t * kDModel + col);
                bcast_f32(smem_u32(xres + m * 512 + col), v);
                bcast_b16(smem_u32(xa + m * LDX512 + col), __float2bfloat16(v));
            });
            cluster_sync_all();

            for (int l = 0; l < 6; ++l) {
                const ClusterLayerParams& W = p.layer[l];
                // ---- QKV: 96 columns = 32 dims of q, k, v of head `head` (this rank's half)
                cl_gemm(p, c, 3, 32, 16, xa, LDX512, [&](int ti, int n, int m, float v) {
                    const int cc = ti * 16 + n, part = cc >> 5, dd = cc & 31;
                    const int gcol = part * 512 + c.head * 64 + c.half * 32 + dd;
                    v += __ldg(W.bqkv + gcol);
                    float* dst = qkvb + m * 192 + part * 64 + c.half * 32 + dd;
                    *dst = v;
                    st_cluster_f32(map_to_rank(smem_u32(dst), partner), v);
                    if (part > 0) {
                        const size_t idx = ((((size_t)(l * 2 + part - 1) * p.B + c.b0 + m) * kHeads + c.head) * p.Tmax + t) * kDHead + c.half * 32 + dd;
                        p.self_kv[idx] = __float2bfloat16(v);
                    }
                });
                cluster_sync_all();
                cl_attention(p, c, true, t);
                cluster_sync_all();
                // ---- O projection + residual, gathered -> LayerNorm 1
                cl_gemm(p, c, 1, 32, 16, abuf, LDX512, [&](int ti, int n, int m, float v) {
                    const int col = c.rank * 32 + ti * 16 + n;
                    v += __ldg(W.bo + col) + xres[m * 512 + col];
                    bcast_f32(smem_u32(ybuf + m * 512 + col), v);
                });
                cluster_sync_all();
                cl_layernorm(c, W.ln1g, W.ln1b);
                // ---- cross-attention query (this rank's 32 dims of head `head`)
                cl_gemm(p, c, 1, 32, 16, xa, LDX512, [&](int ti, int n, int m, float v) {
                    const int cc = ti * 16 + n, col = c.rank * 32 + cc;
                    v += __ldg(W.bq2 + col);
                    float* dst = qkvb + m * 192 + c.half * 32 + cc;
                    *dst = v;
                    st_cluster_f32(map_to_rank(smem_u32(dst), partner), v);
                });
                cluster_sync_all();
                cl_attention(p, c, false, t);
                cluster_sync_all();
                cl_gemm(p, c, 1, 32, 16, abuf, LDX512, [&](int ti, int n, int m, float v) {
                    const int col = c.rank * 32 + ti * 16 + n;
                    v += __ldg(W.bo2 + col) + xres[m * 512 + col];
                    bcast_f32(smem_u32(ybuf + m * 512 + col), v);
                });
                cluster_sync_all();
                cl_layernorm(c, W.ln2g, W.ln2b);
                // ---- FFN: hidden slice [128 rank, +128) stays local; FFN2 is split along K
                cl_gemm(p, c, 4, 32, 16, xa, LDX512, [&](int ti, int n, int m, float v) {
                    const int cc = ti * 16 + n;
                    v = fmaxf(v + __ldg(W.b1 + c.rank * 128 + cc), 0.f);
                    hbuf[m * LDX128 + cc] = __float2bfloat16(v);
                });
                cl_gemm(p, c, 4, 32, 4, hbuf, LDX128, [&](int ti, int n, int m, float v) {
                    const int col = ti * 16 + n;                          // partial sum over this rank's 128 hidden units
                    float* dst = recv + (c.rank * 8 + m) * 32 + (col & 31);
                    st_cluster_f32(map_to_rank(smem_u32(dst), (uint32_t)(col >> 5)), v);
                });
                cluster_sync_all();
                if (c.tid < 256) {                                        // reduce-scatter result: my 32 columns
                    const int m = c.tid >> 5, cc = c.tid & 31, col = c.rank * 32 + cc;
                    if (m < c.G) {
                        float v = __ldg(W.b2 + col) + xres[m * 512 + col];
#pragma unroll
                        for (int r = 0; r < CL_SIZE; ++r) v += recv[(r * 8 + m) * 32 + cc];    // fixed order
                        bcast_f32(smem_u32(ybuf + m * 512 + col), v);
                    }
                }
                cluster_sync_all();
                cl_layernorm(c, W.ln3g, W.ln3b);
            }
            // ================= [mel | stop] heads: ranks 0..5 own 16 of the 81(+15) columns =================
            if (c.rank < 6) {
                cl_gemm(p, c, 1, 16, 16, xa, LDX512, [&](int, int n, int m, float v) {
                    const int col = c.rank * 16 + n, b = c.b0 + m;
                    if (col < 80) {
                        v += __ldg(p.b_head + col);
                        p.mel_before[((size_t)b * p.Tmax + t) * 80 + col] = v;       // fp32 feedback (P8)
                        bcast_b16(smem_u32(fbuf + m * LDX96 + col), __float2bfloat16(v));
                    } else if (col == 80) {
                        v += __ldg(p.b_head + col);
                        p.stop_logits[(size_t)b * p.Tmax + t] = v;
                        if (v > 0.f && p.finished[b] == 0) {                         // P10
                            p.finished[b] = 1; p.lens[b] = t + 1; atomicAdd(p.n_finished, 1);
                            const uint32_t fa = smem_u32(const_cast<int*>(flags));
#pragma unroll
                            for (int r = 0; r < CL_SIZE; ++r) red_cluster_add_u32(map_to_rank(fa, r), 1u);
                        }
                    }
                });
            }
            cluster_sync_all();
            if (flags[0] >= c.G) break;                                   // every utterance of the group has fired
        }
        // ---- drain copies that were issued ahead but never consumed, then make the ring reusable
        if (c.tid == 0) {
            for (uint32_t i = c.consumed; i < c.issued; ++i) mbar_wait(&c.full[i % CL_STAGES], (i / CL_STAGES) & 1);
        }
        __syncthreads();
        cluster_sync_all();
    }
}

}  // namespace tts
