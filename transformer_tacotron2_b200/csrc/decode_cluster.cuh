// Cluster-partitioned autoregressive decode: the north-star hot loop.
//
// Utterances are independent end to end, so the batch is cut into groups of up to 8 utterances and
// each group is decoded by ONE thread-block cluster of 8 CTAs (CTA r owns attention head r) with no
// grid-wide synchronisation at all:
//   * every GEMM of the step is split over the 8 CTAs along N (FFN2 along K); the <= 8 utterances
//     are the MMA n = 8 dimension ("swap-AB": weights are the 16x16 A operand of
//     mma.sync.m16n8k16, activations the 16x8 B operand), so a weight byte is read once per cluster
//     and the accumulators of a warp stay in registers across the whole K loop;
//   * results are exchanged through distributed shared memory (16-byte st.shared::cluster pushes)
//     and ordered by an mbarrier-based cluster barrier (remote arrive.release / local
//     try_wait.acquire), ~40 per step instead of 52 grid barriers;
//   * LayerNorm, residuals, dropout, PE and the stop test run on the gathered rows in shared memory;
//   * ALL global traffic of a CTA -- its 1/8 slice of the weights and the K/V rows of its 8
//     (utterance, head) pairs -- is one ordered stream of cp.async.bulk copies into a 4 x 32 KB
//     shared-memory ring (mbarrier full/empty), issued by a dedicated producer warp that keeps 96 KB
//     in flight per SM (profiles/r01_bulk_bw_ubench.md: what one SM needs to pull > 100 GB/s).
// Weights are packed per CTA rank in exactly the order the step consumes them (tts_b200.cu), each
// 16(n) x 32(k) block in A-fragment order, blocks ordered [chunk][warp][2].
// KV-cache layout (P18, adapted to the tensor-core decode attention): per (layer, b, h) and padded length
// Lp = roundup(L, 16):  K row-major [Lp][64];  V in 16-row blocks stored transposed [Lp/16][64 d][16 rows],
// so that a 16-row chunk of either is one contiguous 2 KB bulk copy and both are A operands of
// mma.sync.m16n8k16 straight from shared memory with conflict-free 16-/8-byte loads (the k index of an
// MMA may be permuted freely as long as A and B agree).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "philox.cuh"

namespace tts {

constexpr int CL_SIZE = 8, CL_CONSUMERS = 512, CL_THREADS = 544, CL_WARPS = 16;
// Row capacity of a cluster.  5 rather than the MMA's 8: the activation buffers shrink by 35 KB, which buys a fifth
// 32 KB ring stage (128 KB instead of 96 KB in flight per SM); B = 64 still fits one pass (13 clusters of <= 15).
constexpr int CL_G = 5;
constexpr int CL_STAGES = 5, CL_STAGE_BYTES = 32768, CL_KV_ROWS = 16;   // K/V chunk: 16 rows of every pair of the CTA
constexpr int CL_NS = 512 / CL_SIZE;                 // 64: columns of a 512-wide output owned by one rank

// bytes of one rank's packed weight segments (stream order: fc1 fc2 proj | 6 x (qkv o q2 o2 w1 w2) | head)
constexpr int CLW_FC1 = 8 * 1024, CLW_FC2 = 16 * 1024, CLW_PROJ = 32 * 1024;
constexpr int CLW_QKV = 192 * 1024, CLW_O = 64 * 1024, CLW_W1 = 256 * 1024, CLW_W2 = 256 * 1024, CLW_HEAD = 16 * 1024;
constexpr int CLW_LAYER = CLW_QKV + 3 * CLW_O + CLW_W1 + CLW_W2;
constexpr size_t CLW_RANK_BYTES = (size_t)CLW_FC1 + CLW_FC2 + CLW_PROJ + 6 * (size_t)CLW_LAYER + CLW_HEAD;

// shared memory carve-up (bytes)
constexpr int SM_RING = 0;
constexpr int SM_XRES = SM_RING + CL_STAGES * CL_STAGE_BYTES;   // f32 [G][512]  residual stream
constexpr int SM_YBUF = SM_XRES + CL_G * 2048;                  // f32 [G][512]  gathered pre-LN sums / FFN2 partial staging
constexpr int SM_RECV = SM_YBUF + CL_G * 2048;                  // f32 [8 ranks][G][64] FFN2 reduce-scatter; epilogue staging otherwise
constexpr int SM_RED = SM_RECV + CL_G * 2048;                   // f32 [16][128] K-split partials; [4096, 8192): LayerNorm gamma | beta
constexpr int SM_XA = SM_RED + 8192;                            // bf16 [G][520] LN output as MMA operand
constexpr int SM_ABUF = SM_XA + CL_G * 1040;                    // bf16 [G][520] gathered attention outputs (prenet: h2 [G][264])
constexpr int SM_QKV = SM_ABUF + CL_G * 1040;                   // f32 [G][192]  q | k_t | v_t of this head
constexpr int SM_HBUF = SM_QKV + CL_G * 768;                    // bf16 [G][264] FFN hidden slice (local)
constexpr int SM_H1 = SM_HBUF + CL_G * 528;                     // bf16 [G][264] gathered prenet activations
constexpr int SM_H2 = SM_ABUF;                                  // aliases abuf (idle during the prenet)
constexpr int SM_FBUF = SM_H1 + CL_G * 528;                     // bf16 [G][136] previous frame (K padded to 128)
constexpr int SM_AMERGE = SM_FBUF + CL_G * 272;                 // f32 [16][68]  attention warp scratch
constexpr int SM_MISC = SM_AMERGE + 4352;                       // mbarriers + flags + head records
constexpr int CL_SMEM_BYTES = SM_MISC + 320;
static_assert(CL_SMEM_BYTES <= 232448, "decode kernel shared memory exceeds 227 KB");
// (MMA B fragments are loaded with ldmatrix over 8 rows: rows >= G read whatever follows the buffer -- finite or not, they
//  only feed output columns m >= G, which are never used.)
constexpr int LDX512 = 520, LDX256 = 264, LDX128 = 136;

struct ClusterLayerParams {
    const float *bqkv, *bo, *bq2, *bo2, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b, *ln3g, *ln3b;
};
struct ClusterParams {
    // TMA descriptors of the two K/V caches viewed as 4-D bf16 tensors [12*B (layer,kv,b)][8 h][Lpad rows][64]:
    // one box {64, 16, 1, G} = the 16-row chunk of all G utterances of a cluster for one head, in one copy
    alignas(64) CUtensorMap tm_self;
    alignas(64) CUtensorMap tm_cross;
    int B, Tmax, S, G, ngroups;                  // G utterances per cluster (<= 8)
    int Tpad, Spad;                              // cache row capacities, multiples of 16
    uint64_t seed; int utt_offset; float dec_alpha; float ln_eps; const float* pe;
    const unsigned char* wpack;                  // [8][CLW_RANK_BYTES]
    const float *b_fc1, *b_fc2, *b_proj, *b_head;
    ClusterLayerParams layer[6];
    bf16* self_kv;                               // [6][2][B][8][Tpad*64]  (K row-major, V blocked-transposed)
    const bf16* cross_kv;                        // [6][2][B][8][Spad*64]
    const int* plens;
    float* mel_before; float* stop_logits; int* lens; int* finished; int* n_finished;
    unsigned long long* ts;                      // optional [Tmax][64] %globaltimer stamps (cluster 0, rank 0)
};

// ---------------------------------------------------------------- PTX helpers
TTS_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TTS_D uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TTS_D uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
TTS_D uint32_t cluster_nid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
TTS_D void hw_cluster_sync() {                   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TTS_D void consumer_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }     // the 16 consumer warps only
TTS_D uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
TTS_D void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
TTS_D void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
TTS_D void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
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
TTS_D void tma_g2s_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
TTS_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ---------------------------------------------------------------- per-CTA context (consumer side)
struct ClCtx {
    unsigned char* smem;
    uint64_t *full, *empty, *gsync;           // ring full/empty, data-carrying gather barriers [2]
    int rank, tid, warp, lane;
    int b0, G;                           // group base utterance, rows in this group
    uint32_t consumed;                   // chunks consumed (uniform over the consumer warps)
    uint32_t gphase;                     // gather phase counter (40 per decoder step)
};

// The ordered stream of chunks of one step for one rank.  seg: 0 fc1, 1 fc2, 2 proj,
// 3+8l+{0 qkv, 1 self-KV, 2 o, 3 q2, 4 cross-KV, 5 o2, 6 w1, 7 w2}, 51 head.
TTS_D int seg_chunks(int seg, int t, int S, int G, int rank) {
    if (seg < 3) return 1;
    if (seg == 51) return rank < 6 ? 1 : 0;
    switch ((seg - 3) & 7) {
    case 0: case 6: case 7: return 8;
    case 1: return (t + CL_KV_ROWS - 1) / CL_KV_ROWS;
    case 4: return (S + CL_KV_ROWS - 1) / CL_KV_ROWS;
    default: return 2;
    }
}
// consumer warps that read a chunk of segment `seg` (the producer arrives on the slot's empty barrier for the others,
// so warps that do not need a chunk never touch it and a slot is free again as soon as its readers have loaded it)
TTS_D int seg_readers(int seg, int G) {
    if (seg == 0) return 4;
    if (seg == 1 || seg == 51) return 8;
    if (seg == 2) return 16;
    switch ((seg - 3) & 7) {
    case 0: return 12;
    case 1: case 4: return G;                      // one warp per (utterance, head) pair and chunk
    default: return 16;
    }
}
TTS_D uint32_t seg_weight_bytes(int seg) {          // bytes per chunk of a weight segment
    if (seg == 0) return CLW_FC1;
    if (seg == 1) return CLW_FC2;
    if (seg == 51) return CLW_HEAD;
    if (seg >= 3 && ((seg - 3) & 7) == 0) return CLW_QKV / 8;
    return CL_STAGE_BYTES;
}

// Producer warp: issue the whole chunk stream of steps [t0, t_end) in order, each chunk as soon as its ring
// slot has been released by all 16 consumer warps.  Weight chunks are one bulk copy; a K/V chunk is two TMA
// tensor copies (the 16-row K box and the V block box of all G pairs): separate <= 2 KB bulk copies cost ~50 ns
// each in the copy engine (profiles/r01_bulk_bw_ubench.md) and capped the stream at ~40 GB/s per SM.  Stops early when the
// consumers raise flags[1].  The control flow is warp-uniform.
TTS_D void cl_producer(const ClusterParams& p, unsigned char* smem, uint64_t* full, uint64_t* empty, volatile int* flags,
                       int rank, int b0, int G, int t0, int t_end, int lane) {
    // ONE lane runs the whole loop alone (the other 31 go straight to the group-end barrier): a warp-wide loop with
    // lane-0-predicated copies and a __syncwarp per chunk costs ~0.35 us per chunk, a lone lane ~0.16 us
    // (scripts/ubench/ring_bisect2.cu, profiles/r02_ring_ubench.md) -- the difference between 90 and 200 GB/s per SM.
    if (lane != 0) return;
    const uint64_t pol_w = l2_policy_evict_last(), pol_kv = l2_policy_evict_first();
    uint32_t issued = 0;
    const unsigned char* wbase = p.wpack + (size_t)rank * CLW_RANK_BYTES;
    bool stopped = false;
    for (int t = t0; t < t_end && !stopped; ++t) {
        size_t woff = 0;
        for (int seg = 0; seg <= 51 && !stopped; ++seg) {
            const int n = seg_chunks(seg, t, p.S, G, rank);
            const int sub = (seg < 3 || seg == 51) ? -1 : ((seg - 3) & 7);
            for (int i = 0; i < n; ++i) {
                const int stage = issued % CL_STAGES;
                const uint32_t use = issued / CL_STAGES;
                if (use > 0) {
                    while (!mbar_try_wait(&empty[stage], (use & 1) ^ 1)) {
                        if (flags[1]) { stopped = true; break; }
                    }
                    if (stopped) break;
                }
                unsigned char* dst = smem + SM_RING + stage * CL_STAGE_BYTES;
                if (sub == 1 || sub == 4) {                      // 16 K rows + one V block of every pair (b0.., head rank): 2 TMA boxes
                    const int l = (seg - 3) >> 3;
                    const CUtensorMap* tm = sub == 1 ? &p.tm_self : &p.tm_cross;
                    mbar_expect_tx(&full[stage], 2u * (uint32_t)p.G * 2048u);
                    tma_g2s_4d(dst, tm, 0, i * CL_KV_ROWS, rank, (l * 2) * p.B + b0, &full[stage], pol_kv);
                    tma_g2s_4d(dst + 16384, tm, 0, i * CL_KV_ROWS, rank, (l * 2 + 1) * p.B + b0, &full[stage], pol_kv);
                } else {
                    const uint32_t bytes = seg_weight_bytes(seg);
                    mbar_expect_tx(&full[stage], bytes);
                    bulk_g2s(dst, wbase + woff, bytes, &full[stage], pol_w);
                    woff += bytes;
                }
                {
                    const int nskip = CL_WARPS - seg_readers(seg, G);
                    if (nskip > 0) mbar_arrive_n(&empty[stage], (uint32_t)nskip);
                }
                ++issued;
            }
        }
    }
    // wait until the consumers are done with the group, then drain copies that were issued but never consumed
    while (!flags[1]) {}
    __threadfence_block();
    const uint32_t final_consumed = (uint32_t)flags[2];
    for (uint32_t i = final_consumed; i < issued; ++i) mbar_wait(&full[i % CL_STAGES], (i / CL_STAGES) & 1);
}

TTS_D unsigned char* cl_acquire(ClCtx& c) {
    const int stage = c.consumed % CL_STAGES;
    // every lane polls: a single polling lane with the rest parked at __syncwarp wakes up ~2x slower
    // (scripts/ubench/pipe_rtt.cu: 0.57 -> 0.32 us per 32 KB chunk)
    mbar_wait(&c.full[stage], (c.consumed / CL_STAGES) & 1);
    return c.smem + SM_RING + stage * CL_STAGE_BYTES;
}
TTS_D void cl_release(ClCtx& c) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[c.consumed % CL_STAGES]);
    ++c.consumed;
}

// ---------------------------------------------------------------- GEMM over ring chunks
// out[col = tile*16 + n][m] = sum_k W[col][k] * X[m][k].  Each of the first NT*KSPLIT warps owns tile
// (w % NT) and K-slice (w / NT) and consumes blocks 2w, 2w+1 of every chunk (k-pairs kq*2*nchunks + 2c + j);
// typeB (FFN2): warp w owns tiles 2w, 2w+1 and chunk c carries k-pair c of both.  Accumulators stay in
// registers across chunks; K-split partials are reduced once per segment through shared memory.
template <class BiasFn, class Epi>
TTS_D void cl_gemm(ClCtx& c, int nchunks, int NT, int KSPLIT, bool typeB, const bf16* X, int ldx, BiasFn biasf, Epi epi) {
    const int nactive = typeB ? CL_WARPS : NT * KSPLIT;
    const bool active = c.warp < nactive;
    const int kq = typeB ? 0 : c.warp / NT;
    const int g = c.lane >> 2, t4 = c.lane & 3;
    const bool direct = typeB || KSPLIT == 1;            // complete sums end up in registers
    // biases (global memory) are fetched before the first chunk arrives, off the critical path
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (direct) {
        if (active) {
            const int tile0 = typeB ? c.warp * 2 : c.warp;
            bias[0] = biasf(tile0, g); bias[1] = biasf(tile0, g + 8);
            if (typeB) { bias[2] = biasf(tile0 + 1, g); bias[3] = biasf(tile0 + 1, g + 8); }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int o = c.tid + k * CL_CONSUMERS;
            if (o < NT * 128) bias[k] = biasf(o >> 7, (o >> 3) & 15);
        }
    }
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    if (!active) c.consumed += nchunks;                  // the producer released these chunks on my behalf
    else
    for (int ch = 0; ch < nchunks; ++ch) {
        const unsigned char* st = cl_acquire(c);
        {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int bi = c.warp * 2 + j;
                const int kp = typeB ? ch : kq * 2 * nchunks + 2 * ch + j;
                const uint4 w0 = reinterpret_cast<const uint4*>(st)[(bi * 2) * 32 + c.lane];
                const uint4 w1 = reinterpret_cast<const uint4*>(st)[(bi * 2 + 1) * 32 + c.lane];
                uint32_t bfrag[4];
                ldmatrix_x4(bfrag, X + (c.lane & 7) * ldx + kp * 32 + (c.lane >> 3) * 8);
                const uint32_t a0[4] = {w0.x, w0.y, w0.z, w0.w}, a1[4] = {w1.x, w1.y, w1.z, w1.w};
                if (typeB && j == 1) { mma_bf16_16816(acc1, a0, bfrag[0], bfrag[1]); mma_bf16_16816(acc1, a1, bfrag[2], bfrag[3]); }
                else { mma_bf16_16816(acc0, a0, bfrag[0], bfrag[1]); mma_bf16_16816(acc0, a1, bfrag[2], bfrag[3]); }
            }
        }
        cl_release(c);
    }
    if (direct) {
        if (active) {
            const int tile0 = typeB ? c.warp * 2 : c.warp;
            const int m0 = t4 * 2;
            if (m0 < c.G) { epi(tile0, g, m0, acc0[0] + bias[0]); epi(tile0, g + 8, m0, acc0[2] + bias[1]); }
            if (m0 + 1 < c.G) { epi(tile0, g, m0 + 1, acc0[1] + bias[0]); epi(tile0, g + 8, m0 + 1, acc0[3] + bias[1]); }
            if (typeB) {
                if (m0 < c.G) { epi(tile0 + 1, g, m0, acc1[0] + bias[2]); epi(tile0 + 1, g + 8, m0, acc1[2] + bias[3]); }
                if (m0 + 1 < c.G) { epi(tile0 + 1, g, m0 + 1, acc1[1] + bias[2]); epi(tile0 + 1, g + 8, m0 + 1, acc1[3] + bias[3]); }
            }
        }
        consumer_bar();
    } else {
        float* red = reinterpret_cast<float*>(c.smem + SM_RED);
        if (active) {
            float* r = red + c.warp * 128;
            *reinterpret_cast<float2*>(r + g * 8 + t4 * 2) = make_float2(acc0[0], acc0[1]);
            *reinterpret_cast<float2*>(r + (g + 8) * 8 + t4 * 2) = make_float2(acc0[2], acc0[3]);
        }
        consumer_bar();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int o = c.tid + k * CL_CONSUMERS;
            if (o < NT * 128) {
                const int ti = o >> 7, n = (o >> 3) & 15, m = o & 7;
                float v = bias[k];
                for (int q = 0; q < KSPLIT; ++q) v += red[(q * NT + ti) * 128 + n * 8 + m];     // fixed order: deterministic
                if (m < c.G) epi(ti, n, m, v);
            }
        }
        consumer_bar();
    }
}

// ---- DSMEM pushes that carry their own completion ---------------------------------------------------------------
// Every exchange of the step ("gather phase") is a set of 16-byte st.async stores into the peers' shared memory, each
// of which completes 16 bytes of a transaction count on the RECEIVER's mbarrier: the receiver waits for the expected
// byte count of the phase instead of a separate arrive/wait round after the data (one DSMEM latency instead of
// two-and-a-half).  Two barriers alternate (a peer can be at most one phase ahead, because every rank contributes to
// every phase); barrier k&1 is re-armed for phase k+2 the moment phase k completes, so data never precedes its arming.
TTS_D void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c2, uint32_t d, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 ::"r"(addr), "r"(a), "r"(b), "r"(c2), "r"(d), "r"(remote_bar) : "memory");
}
// expected bytes of gather phase ph (0..39 within a step) at every receiver
TTS_D uint32_t gather_bytes(int ph, int G) {
    if (ph < 2) return 512u * G;                         // prenet fc1 / fc2: 8 ranks x G x 32 cols bf16
    if (ph == 2) return 3072u * G;                       // prenet proj: 8 x G x 64 x (f32 + bf16)
    if (ph == 39) return 160u * G + 128u;                // head: 5 ranks x G x 16 cols bf16 + one 16-byte record per rank
    const int k = (ph - 3) % 6;
    return (k == 0 || k == 2) ? 1024u * G : 2048u * G;   // attention outputs (bf16) : f32 slices (O, O2, reduce-scatter, y3)
}
TTS_D uint32_t gather_bar(const ClCtx& c) { return smem_u32(&c.gsync[c.gphase & 1]); }
TTS_D void gather_wait(ClCtx& c) {
    uint64_t* bar = &c.gsync[c.gphase & 1];
    mbar_wait(bar, (c.gphase >> 1) & 1);
    if (c.tid == 0) mbar_expect_tx(bar, gather_bytes((int)((c.gphase + 2) % 40), c.G));
    ++c.gphase;
}
// dst (f32, row stride dld) of every peer <- stage[m][0..ncols)
TTS_D void push_f32_all(const ClCtx& c, const float* stage, int ld, float* dst, int dld, int ncols) {
    const int ppr = ncols >> 2, per_peer = c.G * ppr;
    const uint32_t bar = gather_bar(c);
    for (int i = c.tid; i < per_peer * CL_SIZE; i += CL_CONSUMERS) {
        const int peer = i / per_peer, j = i % per_peer, m = j / ppr, pc = j - m * ppr;
        const float4 v = *reinterpret_cast<const float4*>(stage + m * ld + pc * 4);
        st_async_v4(map_to_rank(smem_u32(dst + m * dld + pc * 4), (uint32_t)peer),
                    __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w), map_to_rank(bar, (uint32_t)peer));
    }
}
// dst (bf16, row stride dld) of every peer <- bf16(stage[m][0..ncols)), ncols % 8 == 0
TTS_D void push_bf16_all(const ClCtx& c, const float* stage, int ld, bf16* dst, int dld, int ncols) {
    const int ppr = ncols >> 3, per_peer = c.G * ppr;
    const uint32_t bar = gather_bar(c);
    for (int i = c.tid; i < per_peer * CL_SIZE; i += CL_CONSUMERS) {
        const int peer = i / per_peer, j = i % per_peer, m = j / ppr, pc = j - m * ppr;
        const float4 v0 = *reinterpret_cast<const float4*>(stage + m * ld + pc * 8);
        const float4 v1 = *reinterpret_cast<const float4*>(stage + m * ld + pc * 8 + 4);
        st_async_v4(map_to_rank(smem_u32(dst + m * dld + pc * 8), (uint32_t)peer),
                    pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w), map_to_rank(bar, (uint32_t)peer));
    }
}

// LayerNorm of the gathered rows: ybuf -> xres (f32) + xa (bf16); warp m < G owns row m.  The affine parameters
// (global memory) are fetched asynchronously into shared memory by ln_prefetch() BEFORE the exchange the rows are
// waited on (buffer: the upper half of the K-split scratch, idle between GEMMs).
TTS_D void ln_prefetch(const ClCtx& c, const float* g, const float* b) {
    float* dst = reinterpret_cast<float*>(c.smem + SM_RED + 4096);        // [0,512) gamma, [512,1024) beta
    if (c.tid < 256) {
        const float* src = (c.tid < 128 ? g : b) + (c.tid & 127) * 4;
        cp_async_16(dst + c.tid * 4, src, true);
    }
    cp_async_commit();
}
TTS_D void cl_layernorm(ClCtx& c, float ln_eps) {
    cp_async_wait<0>();
    consumer_bar();
    if (c.warp < c.G) {
        const float* y = reinterpret_cast<const float*>(c.smem + SM_YBUF) + c.warp * 512;
        const float* gb = reinterpret_cast<const float*>(c.smem + SM_RED + 4096);
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
        const float rstd = rsqrtf(warp_sum(ss) * (1.f / 512.f) + ln_eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int col = i * 128 + c.lane * 4;
            const float4 g4 = *reinterpret_cast<const float4*>(gb + col), b4 = *reinterpret_cast<const float4*>(gb + 512 + col);
            const float o0 = (v[i * 4] - mean) * rstd * g4.x + b4.x, o1 = (v[i * 4 + 1] - mean) * rstd * g4.y + b4.y;
            const float o2 = (v[i * 4 + 2] - mean) * rstd * g4.z + b4.z, o3 = (v[i * 4 + 3] - mean) * rstd * g4.w + b4.w;
            *reinterpret_cast<float4*>(xr + col) = make_float4(o0, o1, o2, o3);
            *reinterpret_cast<uint2*>(xa + col) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
    }
    consumer_bar();
}

// Attention of this CTA's pairs (utterance w of the group, head = rank): warp w < G owns pair w and runs a
// flash-style online softmax over 16-row chunks on the tensor cores:
//   scores[16 rows] = K_chunk[16 x 64] . q        4 x mma.m16n8k16 (A = K rows, B = q in column 0)
//   out[64 d]      += V_chunk^T[64 x 16] . p      4 x mma.m16n8k16 (A = V^T block, B = p in column 0)
// self: rows 0..t-1 from the cache chunks + the newest row (k_t, v_t) from qkvbuf; cross: rows 0..len-1.
// Output rows are staged in `stage` [8][64] and pushed to abuf[m][rank*64 ..] of every CTA.
TTS_D void cl_attention(const ClusterParams& p, ClCtx& c, bool self, int t, float* stage) {
    const float* qkv = reinterpret_cast<const float*>(c.smem + SM_QKV);
    float* pscr = reinterpret_cast<float*>(c.smem + SM_AMERGE) + c.warp * 68;      // 16 probabilities of the current chunk
    const int g = c.lane >> 2, t4 = c.lane & 3;
    const float qs = 0.125f * kLog2e;
    const int L = self ? t : p.S;
    const int nck = (L + CL_KV_ROWS - 1) / CL_KV_ROWS;
    const int gi = c.warp >> 1, par2 = c.warp & 1;         // warps 2p, 2p+1 share pair p: even / odd chunks
    const bool active = gi < c.G;
    int vlen = L;
    uint32_t qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0};
    if (active) {
        if (!self) vlen = min(L, __ldg(p.plens + c.b0 + gi));
        // B fragments of q, replicated in all 8 columns (every lane then holds valid scores: the row max needs only
        // 3 shuffles); k slots <-> dims 16 t4 + 4 ks + {0..3}
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float* qp = qkv + gi * 192 + 16 * t4 + 4 * ks;
            qb0[ks] = pack_bf16x2(qp[0] * qs, qp[1] * qs);
            qb1[ks] = pack_bf16x2(qp[2] * qs, qp[3] * qs);
        }
    }
    float m = -INFINITY, l = 0.f, o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    for (int ci = 0; ci < nck; ++ci) {
        const bool mine = active && (ci & 1) == par2;
        if (!mine) { ++c.consumed; continue; }           // not my chunk: the producer / its owner release it
        const unsigned char* st = cl_acquire(c);
        uint32_t kr[8], kr8[8];
        uint2 vf[4][2];
        {
            const unsigned char* kb = st + gi * 2048;
            const unsigned char* vb = st + 16384 + gi * 2048;
            const int par = g & 1;                       // XOR-ordered 16-byte loads: conflict-free at a 128 B row stride
            const uint4 ka = *reinterpret_cast<const uint4*>(kb + g * 128 + ((2 * t4 + par) << 4));
            const uint4 kbb = *reinterpret_cast<const uint4*>(kb + g * 128 + ((2 * t4 + 1 - par) << 4));
            const uint4 kc = *reinterpret_cast<const uint4*>(kb + (g + 8) * 128 + ((2 * t4 + par) << 4));
            const uint4 kd = *reinterpret_cast<const uint4*>(kb + (g + 8) * 128 + ((2 * t4 + 1 - par) << 4));
            const uint4 lo0 = par ? kbb : ka, hi0 = par ? ka : kbb, lo8 = par ? kd : kc, hi8 = par ? kc : kd;
            kr[0] = lo0.x; kr[1] = lo0.y; kr[2] = lo0.z; kr[3] = lo0.w; kr[4] = hi0.x; kr[5] = hi0.y; kr[6] = hi0.z; kr[7] = hi0.w;
            kr8[0] = lo8.x; kr8[1] = lo8.y; kr8[2] = lo8.z; kr8[3] = lo8.w; kr8[4] = hi8.x; kr8[5] = hi8.y; kr8[6] = hi8.z; kr8[7] = hi8.w;
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) {             // V^T block [64 d][16 rows]: rows 4 t4 .. 4 t4 + 3 of d = 16 dt + g (+8)
                vf[dt][0] = *reinterpret_cast<const uint2*>(vb + (dt * 16 + g) * 32 + 8 * t4);
                vf[dt][1] = *reinterpret_cast<const uint2*>(vb + (dt * 16 + g + 8) * 32 + 8 * t4);
            }
        }
        cl_release(c);
        {
            float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t a[4] = {kr[2 * ks], kr8[2 * ks], kr[2 * ks + 1], kr8[2 * ks + 1]};
                mma_bf16_16816(sc, a, qb0[ks], qb1[ks]);
            }
            const int r0 = ci * CL_KV_ROWS + g;
            const float s0 = r0 < vlen ? sc[0] : -INFINITY;
            const float s8 = r0 + 8 < vlen ? sc[2] : -INFINITY;
            float mx = fmaxf(s0, s8);                    // all columns are equal: reduce over the 8 row-groups only
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
            const float mnew = fmaxf(m, mx);
            if (mnew > -INFINITY) {                      // warp-uniform
                const float p0 = fast_exp2(s0 - mnew), p8 = fast_exp2(s8 - mnew);
                if (t4 == 0) { pscr[g] = p0; pscr[g + 8] = p8; }
                if (mnew != m) {                         // warp-uniform: rescale only when the running max moved
                    const float scale = (m == -INFINITY) ? 0.f : fast_exp2(m - mnew);
                    l *= scale;
#pragma unroll
                    for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= scale; o[dt][1] *= scale; o[dt][2] *= scale; o[dt][3] *= scale; }
                }
                l += (t4 == 0) ? p0 + p8 : 0.f;          // lane-partial sum, reduced once at the end
                __syncwarp();
                const float4 pv = *reinterpret_cast<const float4*>(pscr + 4 * t4);
                __syncwarp();
                const uint32_t b0 = pack_bf16x2(pv.x, pv.y), b1 = pack_bf16x2(pv.z, pv.w);     // p replicated in all columns too
#pragma unroll
                for (int dt = 0; dt < 4; ++dt) {
                    const uint32_t a[4] = {vf[dt][0].x, vf[dt][1].x, vf[dt][0].y, vf[dt][1].y};
                    mma_bf16_16816(o[dt], a, b0, b1);
                }
                m = mnew;
            }
        }
    }
    if (active) {
        if (self && par2 == 0) {                         // newest row: bf16-rounded q, k_t, v_t exactly as the MMA path sees them
            const float* qp = qkv + gi * 192 + 2 * c.lane;
            const float2 qd = unpack_bf16x2(pack_bf16x2(qp[0] * qs, qp[1] * qs));
            const float2 kd = unpack_bf16x2(pack_bf16x2(qp[64], qp[65]));
            const float st = warp_sum(qd.x * kd.x + qd.y * kd.y);
            const float mnew = fmaxf(m, st);
            const float scale = (m == -INFINITY) ? 0.f : fast_exp2(m - mnew), pt = fast_exp2(st - mnew);
            l = l * scale + (c.lane == 0 ? pt : 0.f);
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) {
                const float v0 = __bfloat162float(__float2bfloat16(qkv[gi * 192 + 128 + dt * 16 + g]));
                const float v8 = __bfloat162float(__float2bfloat16(qkv[gi * 192 + 128 + dt * 16 + g + 8]));
                o[dt][0] = o[dt][0] * scale + pt * v0;
                o[dt][2] = o[dt][2] * scale + pt * v8;
            }
            m = mnew;
        }
        // partial (m, l, o[64]) of this warp -> scratch; the even warp of the pair merges the two halves
        const float ls = warp_sum(l);
        if (t4 == 0) {
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) { pscr[dt * 16 + g] = o[dt][0]; pscr[dt * 16 + g + 8] = o[dt][2]; }
        }
        if (c.lane == 0) { pscr[64] = m; pscr[65] = ls; }
    }
    consumer_bar();
    if (active && par2 == 0) {
        const float* pa = pscr;
        const float* pb = pscr + 68;
        const float ma = pa[64], mb = pb[64], mm = fmaxf(ma, mb);
        const float ea = (ma == -INFINITY) ? 0.f : fast_exp2(ma - mm), eb = (mb == -INFINITY) ? 0.f : fast_exp2(mb - mm);
        const float ls = pa[65] * ea + pb[65] * eb;
        const float inv = ls > 0.f ? 1.f / ls : 0.f;
        stage[gi * 64 + c.lane] = (pa[c.lane] * ea + pb[c.lane] * eb) * inv;
        stage[gi * 64 + 32 + c.lane] = (pa[32 + c.lane] * ea + pb[32 + c.lane] * eb) * inv;
    }
    consumer_bar();
    push_bf16_all(c, stage, 64, reinterpret_cast<bf16*>(c.smem + SM_ABUF) + c.rank * 64, LDX512, 64);
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(CL_THREADS, 1) decode_cluster_kernel(const __grid_constant__ ClusterParams p, int t0, int n_steps) {
    extern __shared__ __align__(128) unsigned char cl_smem[];
    ClCtx c;
    c.smem = cl_smem;
    c.full = reinterpret_cast<uint64_t*>(cl_smem + SM_MISC);
    c.empty = c.full + CL_STAGES;
    c.gsync = c.empty + CL_STAGES;
    // flags[0] finished utterances of the group (rank 5 counts), [1] consumers done (producer stop), [2] final consumed count
    // hrec: one 16-byte record per rank, pushed in the head phase (rank 5: finished count)
    volatile int* flags = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 128);
    volatile int* hrec = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 128 + 32);     // [8 ranks][4]
    c.rank = (int)cluster_ctarank();
    c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_nid_x();
    const bool is_producer = c.warp == CL_WARPS;

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
    float* stg = recv;                                   // epilogue staging (free outside the FFN2 exchange)
    const bool stamper = p.ts != nullptr && cid == 0 && c.rank == 0 && c.tid == 0;
    auto stamp = [&](int t, int idx) {
        if (stamper) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.ts[(size_t)t * 64 + idx] = now;
        }
    };

    for (int grp = cid; grp < p.ngroups; grp += ncl) {
        c.b0 = grp * p.G; c.G = min(p.G, p.B - c.b0);
        // ---- (re)initialise the ring and the activation buffers
        if (!is_producer)
            for (int i = c.tid; i < (SM_MISC - SM_XRES) / 4; i += CL_CONSUMERS) reinterpret_cast<uint32_t*>(cl_smem + SM_XRES)[i] = 0u;
        if (c.tid == 0) {
            for (int s = 0; s < CL_STAGES; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], CL_WARPS); }
            mbar_init(&c.gsync[0], 1); mbar_init(&c.gsync[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&c.gsync[0], gather_bytes(0, c.G));          // phases 0 and 1 are armed before the start barrier
            mbar_expect_tx(&c.gsync[1], gather_bytes(1, c.G));
            int nf = 0;
            for (int m = 0; m < c.G; ++m) nf += p.finished[c.b0 + m];
            flags[0] = nf; flags[1] = 0; flags[2] = 0;
        }
        __syncthreads();
        if (!is_producer && t0 > 0) {                    // resume: previous frame from global memory (fp32 -> bf16)
            for (int i = c.tid; i < c.G * 80; i += CL_CONSUMERS) {
                const int m = i / 80, col = i - m * 80;
                fbuf[m * LDX128 + col] = __float2bfloat16(p.mel_before[((size_t)(c.b0 + m) * p.Tmax + (t0 - 1)) * 80 + col]);
            }
        }
        c.consumed = 0; c.gphase = 0;
        const bool skip = flags[0] >= c.G;               // every utterance of the group already finished
        __syncthreads();

        // Group boundaries use the hardware cluster barrier (all 544 threads of all 8 CTAs): every peer's buffers and
        // mbarriers are initialised before any DSMEM push reaches them, and nobody re-initialises them while a peer is
        // still inside the group.  (`skip` is uniform over the cluster: every CTA reads the same finished[] flags.)
        hw_cluster_sync();
        if (skip) {
        } else if (is_producer) {
            cl_producer(p, cl_smem, c.full, c.empty, flags, c.rank, c.b0, c.G, t0, t0 + n_steps, c.lane);
            __syncwarp();                                // lanes 1..31 park here: the cluster barrier below is .aligned
        } else {
            for (int t = t0; t < t0 + n_steps; ++t) {
                // ================= decoder prenet (dropout always on, P7) =================
                cl_gemm(c, 1, 2, 2, false, fbuf, LDX128,
                        [&](int ti, int n) { return __ldg(p.b_fc1 + c.rank * 32 + ti * 16 + n); },
                        [&](int ti, int n, int m, float v) {
                            const int cc = ti * 16 + n, col = c.rank * 32 + cc;
                            v = fmaxf(v, 0.f);
                            stg[m * 32 + cc] = keep_bit(p.seed, SITE_DEC_PRENET_FC1, (uint32_t)t, (uint32_t)(p.utt_offset + c.b0 + m), (uint32_t)col) ? 2.f * v : 0.f;
                        });
                push_bf16_all(c, stg, 32, h1 + c.rank * 32, LDX256, 32);
                gather_wait(c);
                stamp(t, 0);
                cl_gemm(c, 1, 2, 4, false, h1, LDX256,
                        [&](int ti, int n) { return __ldg(p.b_fc2 + c.rank * 32 + ti * 16 + n); },
                        [&](int ti, int n, int m, float v) {
                            const int cc = ti * 16 + n, col = c.rank * 32 + cc;
                            v = fmaxf(v, 0.f);
                            stg[m * 32 + cc] = keep_bit(p.seed, SITE_DEC_PRENET_FC2, (uint32_t)t, (uint32_t)(p.utt_offset + c.b0 + m), (uint32_t)col) ? 2.f * v : 0.f;
                        });
                push_bf16_all(c, stg, 32, h2 + c.rank * 32, LDX256, 32);
                gather_wait(c);
                stamp(t, 1);
                cl_gemm(c, 1, 4, 4, false, h2, LDX256,
                        [&](int ti, int n) {
                            const int col = c.rank * CL_NS + ti * 16 + n;
                            return __ldg(p.b_proj + col) + p.dec_alpha * __ldg(p.pe + (size_t)t * kDModel + col);
                        },
                        [&](int ti, int n, int m, float v) { stg[m * CL_NS + ti * 16 + n] = v; });
                push_f32_all(c, stg, CL_NS, xres + c.rank * CL_NS, 512, CL_NS);
                push_bf16_all(c, stg, CL_NS, xa + c.rank * CL_NS, LDX512, CL_NS);
                gather_wait(c);
                stamp(t, 2);

                for (int l = 0; l < 6; ++l) {
                    const ClusterLayerParams& W = p.layer[l];
                    // ---- q, k, v of head `rank` for every row of the group (local; k_t, v_t appended to the cache)
                    cl_gemm(c, 8, 12, 1, false, xa, LDX512,
                            [&](int ti, int n) { const int cc = ti * 16 + n; return __ldg(W.bqkv + (cc >> 6) * 512 + c.rank * 64 + (cc & 63)); },
                            [&](int ti, int n, int m, float v) {
                                const int cc = ti * 16 + n, part = cc >> 6, dd = cc & 63;
                                qkvb[m * 192 + cc] = v;
                                if (part > 0) {               // append to the cache: K row-major, V in transposed 16-row blocks
                                    const size_t pbase = (((size_t)(l * 2 + part - 1) * p.B + c.b0 + m) * kHeads + c.rank) * p.Tpad * kDHead;
                                    const size_t off = part == 1 ? (size_t)t * kDHead + dd : (size_t)(t >> 4) * 1024 + dd * 16 + (t & 15);
                                    p.self_kv[pbase + off] = __float2bfloat16(v);
                                }
                            });
                    // the appended rows are read by the async proxy (TMA copies) from the next step on: order the
                    // generic-proxy stores before them here, on the writer side (a fence next to every copy would
                    // serialise the producer's copies); the ring's empty/full mbarriers carry the rest of the ordering
                    asm volatile("fence.proxy.async;" ::: "memory");
                    stamp(t, 3 + 8 * l);
                    cl_attention(p, c, true, t, stg);
                    gather_wait(c);
                    stamp(t, 4 + 8 * l);
                    // ---- O projection + residual, gathered -> LayerNorm 1
                    cl_gemm(c, 2, 4, 4, false, abuf, LDX512,
                            [&](int ti, int n) { return __ldg(W.bo + c.rank * CL_NS + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) {
                                const int cc = ti * 16 + n;
                                stg[m * CL_NS + cc] = v + xres[m * 512 + c.rank * CL_NS + cc];
                            });
                    ln_prefetch(c, W.ln1g, W.ln1b);
                    push_f32_all(c, stg, CL_NS, ybuf + c.rank * CL_NS, 512, CL_NS);
                    gather_wait(c);
                    cl_layernorm(c, p.ln_eps);
                    stamp(t, 5 + 8 * l);
                    // ---- cross-attention query of head `rank` (local)
                    cl_gemm(c, 2, 4, 4, false, xa, LDX512,
                            [&](int ti, int n) { return __ldg(W.bq2 + c.rank * CL_NS + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) { qkvb[m * 192 + ti * 16 + n] = v; });
                    stamp(t, 6 + 8 * l);
                    cl_attention(p, c, false, t, stg);
                    gather_wait(c);
                    stamp(t, 7 + 8 * l);
                    cl_gemm(c, 2, 4, 4, false, abuf, LDX512,
                            [&](int ti, int n) { return __ldg(W.bo2 + c.rank * CL_NS + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) {
                                const int cc = ti * 16 + n;
                                stg[m * CL_NS + cc] = v + xres[m * 512 + c.rank * CL_NS + cc];
                            });
                    ln_prefetch(c, W.ln2g, W.ln2b);
                    push_f32_all(c, stg, CL_NS, ybuf + c.rank * CL_NS, 512, CL_NS);
                    gather_wait(c);
                    cl_layernorm(c, p.ln_eps);
                    stamp(t, 8 + 8 * l);
                    // ---- FFN: hidden slice [256 rank, +256) stays local (bf16); FFN2 is split along K
                    cl_gemm(c, 8, 16, 1, false, xa, LDX512,
                            [&](int ti, int n) { return __ldg(W.b1 + c.rank * 256 + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) { hbuf[m * LDX256 + ti * 16 + n] = __float2bfloat16(fmaxf(v, 0.f)); });
                    stamp(t, 9 + 8 * l);
                    // partial sums over this rank's 256 hidden units, staged in ybuf (free between LN2 and the y3 gather)
                    cl_gemm(c, 8, 32, 1, true, hbuf, LDX256, [&](int, int) { return 0.f; },
                            [&](int ti, int n, int m, float v) { ybuf[m * 512 + ti * 16 + n] = v; });
                    {
                        const uint32_t bar = gather_bar(c);
                        for (int i = c.tid; i < CL_SIZE * c.G * 16; i += CL_CONSUMERS) {   // reduce-scatter: 64 columns to each peer
                            const int peer = i / (c.G * 16), j = i % (c.G * 16), m = j >> 4, pc = j & 15;
                            const float4 v = *reinterpret_cast<const float4*>(ybuf + m * 512 + peer * CL_NS + pc * 4);
                            st_async_v4(map_to_rank(smem_u32(recv + (c.rank * CL_G + m) * CL_NS + pc * 4), (uint32_t)peer),
                                        __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w),
                                        map_to_rank(bar, (uint32_t)peer));
                        }
                    }
                    const float b2v = __ldg(W.b2 + c.rank * CL_NS + (c.tid & 63));
                    ln_prefetch(c, W.ln3g, W.ln3b);
                    gather_wait(c);
                    {
                        float* st2 = reinterpret_cast<float*>(c.smem + SM_RED);    // [8][64] staging of my reduced columns
                        const int m = c.tid >> 6, cc = c.tid & 63, col = c.rank * CL_NS + cc;
                        if (m < c.G) {
                            float v = b2v + xres[m * 512 + col];
#pragma unroll
                            for (int r = 0; r < CL_SIZE; ++r) v += recv[(r * CL_G + m) * CL_NS + cc];    // fixed order
                            st2[m * CL_NS + cc] = v;
                        }
                        consumer_bar();
                        push_f32_all(c, st2, CL_NS, ybuf + c.rank * CL_NS, 512, CL_NS);
                    }
                    gather_wait(c);
                    cl_layernorm(c, p.ln_eps);
                    stamp(t, 10 + 8 * l);
                }
                // ================= [mel | stop] heads: ranks 0..5 own 16 of the 81(+15) columns =================
                if (c.rank < 6) {
                    cl_gemm(c, 1, 1, 8, false, xa, LDX512,
                            [&](int, int n) { const int col = c.rank * 16 + n; return col <= 80 ? __ldg(p.b_head + col) : 0.f; },
                            [&](int, int n, int m, float v) {
                                const int col = c.rank * 16 + n, b = c.b0 + m;
                                if (col < 80) p.mel_before[((size_t)b * p.Tmax + t) * 80 + col] = v;   // fp32 feedback (P8)
                                else if (col == 80) {
                                    p.stop_logits[(size_t)b * p.Tmax + t] = v;
                                    if (v > 0.f && p.finished[b] == 0) {                         // P10
                                        p.finished[b] = 1; p.lens[b] = t + 1; atomicAdd(p.n_finished, 1);
                                        atomicAdd(const_cast<int*>(flags), 1);                   // rank 5 keeps the group's count
                                    }
                                }
                                stg[m * 16 + n] = v;
                            });
                    if (c.rank < 5) push_bf16_all(c, stg, 16, fbuf + c.rank * 16, LDX128, 16);   // the frame = next step's prenet input
                }
                if (c.tid < CL_SIZE) {                    // every rank contributes one record (rank 5: finished count) to every peer
                    const uint32_t bar = gather_bar(c);
                    st_async_v4(map_to_rank(smem_u32(const_cast<int*>(hrec) + c.rank * 4), (uint32_t)c.tid),
                                (uint32_t)flags[0], 0u, 0u, 0u, map_to_rank(bar, (uint32_t)c.tid));
                }
                gather_wait(c);
                stamp(t, 51);
                if (hrec[5 * 4] >= c.G) break;                               // every utterance of the group has fired
            }
            // ---- tell the producer we are done (it drains the copies that were issued but never consumed)
            consumer_bar();
            if (c.tid == 0) { flags[2] = (int)c.consumed; __threadfence_block(); flags[1] = 1; }
        }
        hw_cluster_sync();                               // peers have finished the group; no CTA exits (or re-initialises) while a
                                                         // peer may still touch its shared memory
    }
}

}  // namespace tts
