// Cluster-partitioned autoregressive decode: the north-star hot loop.
//
// Utterances are independent end to end, so the batch is cut into groups of up to 5 utterances and each group is decoded
// by ONE thread-block cluster of 8 CTAs (CTA r owns attention head r) with no grid-wide synchronisation at all:
//   * every GEMM of the step is split over the 8 CTAs along N (FFN2 along K); the <= 5 utterances are the MMA n = 8
//     dimension ("swap-AB": weights are the 16x16 A operand of mma.sync.m16n8k16, activations the 16x8 B operand);
//   * results are exchanged through distributed shared memory: 16-byte st.async pushes that complete a transaction
//     count on the RECEIVER's mbarrier (one DSMEM latency per exchange, no separate arrive / wait round);
//   * LayerNorm, residuals, dropout, PE and the stop test run on the gathered rows in shared memory;
//   * ALL global traffic of a CTA -- its 1/8 slice of the weights and the K/V rows of its (utterance, head) pairs -- is
//     one ordered stream of cp.async.bulk copies into a 5 x 32 KB shared-memory ring (mbarrier full / empty).
//
// Round 2 (profiles/r02_ring_ubench.md): the ring is bound by a FIXED COST PER CHUNK HAND-OFF (producer ~0.16 us per copy
// when one lane runs the loop alone, ~0.35 us when a whole warp loops; a consumer warp ~0.28 us per wait -> release round),
// not by bytes.  So:
//   * the producer is ONE elected lane;
//   * a stage is read only by the warps that own its contents.  Wide GEMMs (QKV, FFN1, FFN2): warp w owns output tile(s) w
//     over the whole local K; its weights are contiguous runs, streamed in two K halves (all warps' first halves first, so
//     that slots are handed back after half an MMA loop).  Narrow GEMMs (O, cross-Q, O2: four tiles per rank): 4 tiles x 4
//     K quarters on all 16 warps, the quarters summed in fixed order through shared memory (cl_gemm_ksplit) -- with one
//     warp per tile the MMA loop is a long dependent chain.  Attention: a stage is up to 128 cache rows of ONE
//     (utterance, head) pair, read by that pair's 3 warps, one online-softmax round per warp and stage;
//   * every stage is ONE bulk copy (the copy engine needs ~44 ns per separate piece: the round-1 K/V chunk of 2 x G pieces
//     of 2 KB ran at 41 GB/s per SM).
//
// KV-cache layout (P18, adapted; kv_k_elem / kv_v_elem in common.cuh): per (layer, utterance, head) a run of 64-row blocks of
// 16 KB = four 16-row sub-chunks [ K 16 rows row-major [16][64] bf16 | V 16 rows in mma.m16n8k16 A-fragment order of V^T ].
// Rows [0, 16 n) of a pair are one contiguous range.  K rows are the B operand of the score MMAs (A = q), V fragments the A
// operand of the output MMAs (B = p): the probabilities go from the score registers straight into the next MMA.
// Weights are packed per CTA rank in exactly the order the step consumes them (tts_b200.cu: pack_cluster_segment), each
// 16(n) x 32(k) block in A-fragment order.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "philox.cuh"

namespace tts {

constexpr int CL_SIZE = 8, CL_CONSUMERS = 512, CL_THREADS = 544, CL_WARPS = 16;
// Row capacity of a cluster.  5: at most 15 8-CTA clusters of this kernel are co-resident on a B200 (120 SMs), so B = 64 is
// 13 groups of <= 5; the activation buffers of 5 rows leave room for five 32 KB ring stages.
constexpr int CL_G = 5;
constexpr int CL_STAGES = 5, CL_STAGE_BYTES = 32768;
// `full` barriers: two per ring slot, used alternately (stage idx -> barrier idx % 10, parity (idx / 10) & 1).  A warp waits only
// on the stages it reads, so it can reach its wait before the PREVIOUS use of the slot has landed; with one barrier per slot
// the one-bit parity test would then pass on the use before that (same parity) and the warp would read stale bytes.  With two,
// the test can only be fooled by a warp >= 10 stages ahead of the landed data, and the block-wide barriers after every
// segment (and the in-order waits inside an attention segment: G + 5 <= 10) keep every reader closer than that.
constexpr int CL_FULL_BARS = 2 * CL_STAGES;
constexpr int CL_TS_COLS = 128;                      // %globaltimer stamps per decoder step (profiling aid): 0..51 phases, 52.. fine-grained
constexpr int CL_NS = 512 / CL_SIZE;                 // 64: columns of a 512-wide output owned by one rank
constexpr int KV_BLOCK_ROWS = 64, KV_BLOCK_ELEMS = 8192;     // one cache block: 64 rows = four 16-row sub-chunks [K | V] of 4 KB (common.cuh)
constexpr int KV_STAGE_ROWS = 128;                   // rows of one pair per ring stage (two blocks)
constexpr int ATT_WPP = 3;                           // warps per (utterance, head) pair: sub-chunk c of 16 rows -> warp c % 3

// bytes of one rank's packed weight segments (stream order: fc1 fc2 proj | 6 x (qkv o q2 o2 w1 w2) | head)
constexpr int CLW_FC1 = 8 * 1024, CLW_FC2 = 16 * 1024, CLW_PROJ = 32 * 1024;
constexpr int CLW_QKV = 192 * 1024, CLW_O = 64 * 1024, CLW_W1 = 256 * 1024, CLW_W2 = 256 * 1024, CLW_HEAD = 16 * 1024;
constexpr int CLW_LAYER = CLW_QKV + 3 * CLW_O + CLW_W1 + CLW_W2;
constexpr size_t CLW_RANK_BYTES = (size_t)CLW_FC1 + CLW_FC2 + CLW_PROJ + 6 * (size_t)CLW_LAYER + CLW_HEAD;

// shared memory carve-up (bytes)
constexpr int SM_RING = 0;
constexpr int SM_XRES = SM_RING + CL_STAGES * CL_STAGE_BYTES;   // f32 [G][512]  residual stream
constexpr int SM_YBUF = SM_XRES + CL_G * 2048;                  // f32 [G][512]  gathered pre-LN sums / FFN2 partial sums
constexpr int SM_RECV = SM_YBUF + CL_G * 2048;                  // f32 [8 ranks][G][64] FFN2 reduce-scatter
constexpr int SM_RED = SM_RECV + CL_G * 2048;                   // [0, 4096): attention output staging; [4096, 8192): LayerNorm gamma | beta
constexpr int SM_XA = SM_RED + 8192;                            // bf16 [G][520] LN output as MMA operand
constexpr int SM_ABUF = SM_XA + CL_G * 1040;                    // bf16 [G][520] gathered attention outputs (prenet: h2 [G][264])
constexpr int SM_QKV = SM_ABUF + CL_G * 1040;                   // f32 [G][192]  q | k_t | v_t of this head
constexpr int SM_HBUF = SM_QKV + CL_G * 768;                    // bf16 [G][264] FFN hidden slice (local)
constexpr int SM_H1 = SM_HBUF + CL_G * 528;                     // bf16 [G][264] gathered prenet activations
constexpr int SM_H2 = SM_ABUF;                                  // aliases abuf (idle during the prenet)
constexpr int SM_FBUF = SM_H1 + CL_G * 528;                     // bf16 [G][136] previous frame (K padded to 128)
constexpr int SM_AMERGE = SM_FBUF + CL_G * 272;                 // f32 [16][68]  attention partials (m, l, o[64]) per warp
constexpr int SM_WST = SM_AMERGE + 4352;                        // f32 [4 warps][G][16] epilogue tiles of the narrow GEMMs before their push
constexpr int SM_MISC = SM_WST + 4 * CL_G * 64;                 // mbarriers + flags + head records + group lengths
constexpr int CL_SMEM_BYTES = SM_MISC + 448;                    // 17 mbarriers (136 B) | flags @192 | hrec @224 | glens @352 | guids @384 | gtlens @416
static_assert(CL_SMEM_BYTES <= 232448, "decode kernel shared memory exceeds 227 KB");
// (MMA B fragments are loaded with ldmatrix over 8 rows: rows >= G read whatever follows the buffer -- finite or not, they
//  only feed output columns m >= G, which are never used.)
constexpr int LDX512 = 520, LDX256 = 264, LDX128 = 136;

struct ClusterLayerParams {
    const float *bqkv, *bo, *bq2, *bo2, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b, *ln3g, *ln3b;
};
struct ClusterParams {
    int B, Tmax, S, G, ngroups;                  // G utterances per cluster (<= CL_G)
    int nblk_self, nblk_cross;                   // 64-row blocks per (layer, utterance, head) in the two caches
    uint64_t seed; int utt_offset; float dec_alpha; float ln_eps; const float* pe;
    const unsigned char* wpack;                  // [8][CLW_RANK_BYTES]
    const float *b_fc1, *b_fc2, *b_proj, *b_head;
    ClusterLayerParams layer[6];
    bf16* self_kv;                               // [6][B][8][nblk_self][KV_BLOCK_ELEMS]
    const bf16* cross_kv;                        // [6][B][8][nblk_cross][KV_BLOCK_ELEMS]
    const int* plens;
    const int* utt_ids;                          // optional [B]: global utterance id of every row (dropout key); null -> utt_offset + b
    const int* tlens;                            // optional [B]: per-utterance frame budget (an utterance also stops at tlens[b]); null -> Tmax
    int* group_queue;                            // optional device counter: clusters draw their next utterance group from it (work
                                                 // stealing: a cluster whose group has stopped takes the next one); null -> static round-robin
    float* mel_before; float* stop_logits; int* lens; int* finished; int* n_finished;
    unsigned long long* ts;                      // optional [Tmax][CL_TS_COLS] %globaltimer stamps (cluster 0, rank 0)
    int dbg_rank;
    float* dbg;                                  // optional debug dump (cluster 0, rank dbg_rank, first step of the launch): slots of 2560 floats
};
// element offset of cache block `blk` of (layer, utterance, head)
TTS_HD size_t kv_block_offset(int layer, int B, int b, int h, int nblk, int blk) {
    return ((((size_t)layer * B + b) * kHeads + h) * nblk + blk) * KV_BLOCK_ELEMS;
}
// element offsets of row r (0..63) inside a block: kv_k_elem / kv_v_elem (common.cuh)

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
TTS_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// %globaltimer read placed behind a shared-memory load (see the stamp lambda of the kernel)
TTS_D unsigned long long timer_after_lds(const void* smem_word) {
    unsigned long long now; uint32_t dummy;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dummy) : "r"(smem_u32(smem_word)) : "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "r"(dummy) : "memory");                // input operand: issues after the load returned
    return now;
}

// ---------------------------------------------------------------- per-CTA context (consumer side)
struct ClCtx {
    unsigned char* smem;
    uint64_t *full, *empty, *gsync;           // ring full/empty, data-carrying gather barriers [2]
    int rank, tid, warp, lane;
    int b0, G;                           // group base utterance, rows in this group
    uint32_t consumed;                   // ring stages consumed or skipped so far (uniform over the consumer warps)
    uint32_t gphase;                     // gather phase counter (40 per decoder step)
};

// ---------------------------------------------------------------- the stage stream of one step (shared by producer and consumers)
// seg: 0 fc1, 1 fc2, 2 proj, 3+8l+{0 qkv, 1 self-KV, 2 o, 3 q2, 4 cross-KV, 5 o2, 6 w1, 7 w2}, 51 head.
// A weight segment is `total` bytes for `na` warps, `bw` bytes each, 32768 / bw warps per stage.
struct WSeg { int total, bw, na, nh; };             // nh: the K range of every warp is streamed in nh parts (1 or 2), part-major
TTS_D WSeg wseg(int seg, int rank) {
    // (prenet and heads stay on one warp per tile: spreading these tiny GEMMs over K parts measured slower)
    if (seg == 0) return {CLW_FC1, 4096, 2, 1};
    if (seg == 1) return {CLW_FC2, 8192, 2, 1};
    if (seg == 2) return {CLW_PROJ, 8192, 4, 1};
    if (seg == 51) return {rank < 6 ? CLW_HEAD : 0, 16384, rank < 6 ? 1 : 0, 1};
    switch ((seg - 3) & 7) {
    case 0: return {CLW_QKV, 8192, 12, 2};               // the wide GEMMs stream every warp's K range in two halves (cl_gemm<.., 2>)
    case 6: return {CLW_W1, 8192, 16, 2};
    case 7: return {CLW_W2, 8192, 16, 2};
    default: return {CLW_O, 4096, 16, 1};                 // O / cross-Q / O2: 4 tiles x 4 K quarters, one per warp (cl_gemm_ksplit)
    }
}

// Producer: ONE lane issues the whole stage stream of steps [t0, t_end) in order, each stage as soon as its ring slot has
// been released (the other 31 lanes of the warp park at the group-end barrier).  Every stage is ONE bulk copy: 32 KB of
// weights, or up to 128 rows of one pair's K/V (16-row granularity; the cache layout keeps them contiguous).  It arrives
// on the slot's empty barrier on behalf of the warps that do not read the stage.  Stops early when the consumers raise flags[1].
TTS_D void cl_producer(const ClusterParams& p, unsigned char* smem, uint64_t* full, uint64_t* empty, volatile int* flags,
                       const volatile int* glens, int rank, int b0, int G, int t0, int t_end) {
    const uint64_t pol_w = l2_policy_evict_last(), pol_kv = l2_policy_evict_first();
    uint32_t issued = 0;
    const unsigned char* wbase = p.wpack + (size_t)rank * CLW_RANK_BYTES;
    bool stopped = false;
    uint64_t* fbar = nullptr;                            // full barrier of the stage being issued
    auto slot = [&](uint32_t& stage) -> bool {           // wait until ring slot issued % STAGES is free
        stage = issued % CL_STAGES;
        fbar = &full[issued % CL_FULL_BARS];
        const uint32_t use = issued / CL_STAGES;
        if (use > 0)
            while (!mbar_try_wait(&empty[stage], (use & 1) ^ 1))
                if (flags[1]) { stopped = true; return false; }
        return true;
    };
    for (int t = t0; t < t_end && !stopped; ++t) {
        size_t woff = 0;
        for (int seg = 0; seg <= 51 && !stopped; ++seg) {
            const int sub = (seg < 3 || seg == 51) ? -1 : ((seg - 3) & 7);
            if (sub == 1 || sub == 4) {                  // K/V rows of every pair (b0 + pair, head rank), 128 rows per stage
                const int l = (seg - 3) >> 3;
                const bool self = sub == 1;
                const int L = self ? t : p.S, nblk = self ? p.nblk_self : p.nblk_cross;
                const bf16* cache = self ? p.self_kv : p.cross_kv;
                const int nj = (L + KV_STAGE_ROWS - 1) / KV_STAGE_ROWS;
                for (int j = 0; j < nj && !stopped; ++j)
                    for (int pair = 0; pair < G; ++pair) {
                        uint32_t stage;
                        if (!slot(stage)) break;
                        unsigned char* dst = smem + SM_RING + stage * CL_STAGE_BYTES;
                        const int Lb = self ? t : min(L, glens[pair]);
                        const int rows = max(0, min(KV_STAGE_ROWS, Lb - j * KV_STAGE_ROWS));
                        const int r16 = (rows + 15) & ~15;
                        const bf16* src = cache + kv_block_offset(l, p.B, b0 + pair, rank, nblk, 2 * j);
                        if (r16 == 0) mbar_arrive(fbar);
                        else {                           // rows [128 j, 128 j + r16) of the pair are one contiguous range
                            mbar_expect_tx(fbar, (uint32_t)r16 * 256u);
                            bulk_g2s(dst, src, (uint32_t)r16 * 256u, fbar, pol_kv);
                        }
                        mbar_arrive_n(&empty[stage], (uint32_t)(CL_WARPS - ATT_WPP));
                        ++issued;
                    }
            } else {
                const WSeg ws = wseg(seg, rank);
                const int wps = CL_STAGE_BYTES / ws.bw, part = ws.total / ws.nh;
                for (int done = 0, w0 = 0; done < ws.total && !stopped; done += CL_STAGE_BYTES, w0 += wps) {
                    if (done == part) w0 = 0;            // second half: the warps again, from warp 0 (a part is a whole number of stages)
                    uint32_t stage;
                    if (!slot(stage)) break;
                    const uint32_t bytes = (uint32_t)min(CL_STAGE_BYTES, (done < part ? part : ws.total) - done);
                    const int readers = min(wps, ws.na - w0);
                    mbar_expect_tx(fbar, bytes);
                    bulk_g2s(smem + SM_RING + stage * CL_STAGE_BYTES, wbase + woff, bytes, fbar, pol_w);
                    woff += bytes;
                    mbar_arrive_n(&empty[stage], (uint32_t)(CL_WARPS - readers));
                    ++issued;
                }
            }
        }
    }
    // wait until the consumers are done with the group, then drain copies that were issued but never consumed
    while (!flags[1]) {}
    __threadfence_block();
    const uint32_t final_consumed = (uint32_t)flags[2];
    for (uint32_t i = final_consumed; i < issued; ++i) mbar_wait(&full[i % CL_FULL_BARS], (i / CL_FULL_BARS) & 1);
}

// ring stage `idx` (absolute index in the stream): wait until it has landed / hand it back
TTS_D const unsigned char* cl_acquire(const ClCtx& c, uint32_t idx) {
    const uint32_t stage = idx % CL_STAGES;
    mbar_wait(&c.full[idx % CL_FULL_BARS], (idx / CL_FULL_BARS) & 1);    // all lanes poll: the fastest wake-up for a whole warp (r02_ring_ubench.md)
    return c.smem + SM_RING + stage * CL_STAGE_BYTES;
}
TTS_D void cl_release(const ClCtx& c, uint32_t idx) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[idx % CL_STAGES]);
}

// ---------------------------------------------------------------- GEMM: one warp = TW output tiles x the whole local K
// out[col = tile*16 + n][m] = sum_k W[col][k] * X[m][k] for tiles warp*TW .. warp*TW + TW-1, K = 32*KP.  The warp's weights are one
// contiguous run of TW*KP KB ([kp][j] blocks of 1 KB) inside a ring stage shared with its neighbours; the accumulators never leave
// registers (NCH independent MMA chains per tile).  epi(tile, n, m, value) receives complete sums (+ bias).  No block-level
// barrier inside: callers synchronise where the results are consumed.
template <int KP, int TW, int NH = 1, class BiasFn, class Epi>
TTS_D void cl_gemm(ClCtx& c, int na, const bf16* X, int ldx, BiasFn biasf, Epi epi, unsigned long long* tstamp = nullptr) {
    // NH = 2: the warp's K range arrives in two halves, all warps' first halves first (BW bytes per warp and half, 32768 / BW
    // warps per stage): a slot is handed back after half an MMA loop, so the stages beyond the ring's depth are re-issued (and
    // land) while the first halves are still being multiplied.
    constexpr int KH = KP / NH, BW = TW * KH * 1024, WPS = CL_STAGE_BYTES / BW;
    constexpr int NCH = (TW == 1 && KP >= 16) ? 4 : 2;
    static_assert(KH % 4 == 0, "k-steps per part must be a multiple of the unroll factor");
    const int nst = (na + WPS - 1) / WPS;                // stages per part
    if (c.warp < na) {
        const int g = c.lane >> 2, t4 = c.lane & 3;
        float bias[TW][2];
#pragma unroll
        for (int j = 0; j < TW; ++j) { bias[j][0] = biasf(c.warp * TW + j, g); bias[j][1] = biasf(c.warp * TW + j, g + 8); }
        float acc[TW][NCH][4];
#pragma unroll
        for (int j = 0; j < TW; ++j)
#pragma unroll
            for (int q = 0; q < NCH; ++q) { acc[j][q][0] = acc[j][q][1] = acc[j][q][2] = acc[j][q][3] = 0.f; }
        const bf16* xrow = X + (c.lane & 7) * ldx + (c.lane >> 3) * 8;
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            const uint32_t idx = c.consumed + (uint32_t)(h * nst + c.warp / WPS);
            if (tstamp && h == 0) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) :: "memory"); tstamp[0] = now; }
            const uint4* wp = reinterpret_cast<const uint4*>(cl_acquire(c, idx) + (c.warp % WPS) * BW);
            if (tstamp && h == 0) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "l"(wp) : "memory"); tstamp[1] = now; }
#pragma unroll 1
            for (int kq = 0; kq < KH / 4; ++kq) {        // four k-steps per iteration: a fully unrolled K loop in every GEMM makes
#pragma unroll                                           // the kernel too large for the instruction cache
                for (int ku = 0; ku < 4; ++ku) {
                    const int kl = kq * 4 + ku;          // k-pair inside this part
                    uint32_t bfrag[4];
                    ldmatrix_x4(bfrag, xrow + (h * KH + kl) * 32);
#pragma unroll
                    for (int j = 0; j < TW; ++j) {
                        const uint4 w0 = wp[((kl * TW + j) * 2) * 32 + c.lane];
                        const uint4 w1 = wp[((kl * TW + j) * 2 + 1) * 32 + c.lane];
                        const uint32_t a0[4] = {w0.x, w0.y, w0.z, w0.w}, a1[4] = {w1.x, w1.y, w1.z, w1.w};
                        mma_bf16_16816(acc[j][ku % NCH], a0, bfrag[0], bfrag[1]);
                        mma_bf16_16816(acc[j][ku % NCH], a1, bfrag[2], bfrag[3]);
                    }
                }
            }
            if (h + 1 < NH) cl_release(c, idx);
            else {
                if (tstamp) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "f"(acc[0][0][0]), "f"(acc[TW - 1][NCH - 1][3]) : "memory"); tstamp[2] = now; }
                cl_release(c, idx);
            }
        }
        const int m0 = t4 * 2;
#pragma unroll
        for (int j = 0; j < TW; ++j) {
            float s[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = acc[j][0][e];
#pragma unroll
                for (int q = 1; q < NCH; ++q) v += acc[j][q][e];           // fixed order: deterministic
                s[e] = v;
            }
            const int tile = c.warp * TW + j;
            if (m0 < c.G) { epi(tile, g, m0, s[0] + bias[j][0]); epi(tile, g + 8, m0, s[2] + bias[j][1]); }
            if (m0 + 1 < c.G) { epi(tile, g, m0 + 1, s[1] + bias[j][0]); epi(tile, g + 8, m0 + 1, s[3] + bias[j][1]); }
        }
    }
    c.consumed += (uint32_t)(NH * nst);
}

// The narrow GEMMs of a layer (NT <= 4 output tiles of this rank: O, cross-Q, O2) spread over NT x NQ warps: warp w multiplies
// K part w / NT of tile w % NT (a piece of the tile's contiguous run -- no repacking), parts 1.. park their sums in shared
// memory, the tile's NQ warps meet at a named barrier and warp `tile` (part 0) adds them in fixed order and runs the epilogue.
// With one warp per tile the MMA loop was a 2 KP-deep dependent chain (0.5 us for K = 512); now it is 2 KP / NQ MMAs per warp.
template <int NT, int KP, int NQ, class BiasFn, class Epi>
TTS_D void cl_gemm_ksplit(ClCtx& c, const bf16* X, int ldx, BiasFn biasf, Epi epi, unsigned long long* tstamp = nullptr) {
    constexpr int KQ = KP / NQ, RUN = KQ * 1024, RPS = CL_STAGE_BYTES / RUN;     // k-pairs per warp, bytes per warp, warps per stage
    static_assert(KQ >= 1 && KQ <= 4 && NT * NQ <= CL_WARPS && NQ <= 4, "cl_gemm_ksplit geometry");
    constexpr int NST = (NT * KP * 1024 + CL_STAGE_BYTES - 1) / CL_STAGE_BYTES;
    if (c.warp < NT * NQ) {
        const int tile = c.warp % NT, kq = c.warp / NT, run = tile * NQ + kq;
        const int g = c.lane >> 2, t4 = c.lane & 3;
        float bias0 = 0.f, bias1 = 0.f;
        if (kq == 0) { bias0 = biasf(tile, g); bias1 = biasf(tile, g + 8); }
        float acc[KQ][4];
#pragma unroll
        for (int q = 0; q < KQ; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
        const uint32_t idx = c.consumed + (uint32_t)(run / RPS);
        if (tstamp) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) :: "memory"); tstamp[0] = now; }
        const uint4* wp = reinterpret_cast<const uint4*>(cl_acquire(c, idx) + (run % RPS) * RUN);
        if (tstamp) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "l"(wp) : "memory"); tstamp[1] = now; }
        const bf16* xrow = X + (c.lane & 7) * ldx + (c.lane >> 3) * 8 + kq * KQ * 32;
#pragma unroll
        for (int ku = 0; ku < KQ; ++ku) {
            uint32_t bfrag[4];
            ldmatrix_x4(bfrag, xrow + ku * 32);
            const uint4 w0 = wp[(ku * 2) * 32 + c.lane];
            const uint4 w1 = wp[(ku * 2 + 1) * 32 + c.lane];
            const uint32_t a0[4] = {w0.x, w0.y, w0.z, w0.w}, a1[4] = {w1.x, w1.y, w1.z, w1.w};
            mma_bf16_16816(acc[ku], a0, bfrag[0], bfrag[1]);
            mma_bf16_16816(acc[ku], a1, bfrag[2], bfrag[3]);
        }
        if (tstamp) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "f"(acc[0][0]), "f"(acc[KQ - 1][3]) : "memory"); tstamp[2] = now; }
        cl_release(c, idx);
        float4 sum = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
#pragma unroll
        for (int q = 1; q < KQ; ++q) { sum.x += acc[q][0]; sum.y += acc[q][1]; sum.z += acc[q][2]; sum.w += acc[q][3]; }    // fixed order
        float4* part = reinterpret_cast<float4*>(c.smem + SM_RECV);       // [NQ - 1][NT][32 lanes]; idle outside the FFN2 reduce-scatter
        if (NQ > 1) {
            if (kq > 0) part[((kq - 1) * NT + tile) * 32 + c.lane] = sum;
            asm volatile("bar.sync %0, %1;" ::"r"(8 + tile), "n"(NQ * 32) : "memory");      // the NQ warps of the tile
        }
        if (kq == 0) {
#pragma unroll
            for (int q = 0; q < NQ - 1; ++q) {                            // fixed order: deterministic
                const float4 o = part[(q * NT + tile) * 32 + c.lane];
                sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
            }
            const int m0 = t4 * 2;
            if (m0 < c.G) { epi(tile, g, m0, sum.x + bias0); epi(tile, g + 8, m0, sum.z + bias1); }
            if (m0 + 1 < c.G) { epi(tile, g, m0 + 1, sum.y + bias0); epi(tile, g + 8, m0 + 1, sum.w + bias1); }
        }
    }
    c.consumed += (uint32_t)NST;
}

// ---- DSMEM pushes that carry their own completion ---------------------------------------------------------------
// Every exchange of the step ("gather phase") is a set of 16-byte st.async stores into the peers' shared memory, each
// of which completes 16 bytes of a transaction count on the RECEIVER's mbarrier: the receiver waits for the expected
// byte count of the phase instead of a separate arrive/wait round after the data.  Two barriers alternate (a peer can be
// at most one phase ahead, because every rank contributes to every phase); barrier k&1 is re-armed for phase k+2 the moment
// phase k completes.  (Data that overtakes the re-arming only drives the transaction count negative for a moment: the phase
// cannot complete before the arming arrive.)
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
// Block-wide completion of the current gather phase: warp 0 waits on the mbarrier (one warp polling instead of sixteen),
// re-arms it for phase + 2, and the named barrier releases everybody.
TTS_D void gather_sync(ClCtx& c) {
    if (c.warp == 0) {
        uint64_t* bar = &c.gsync[c.gphase & 1];
        mbar_wait(bar, (c.gphase >> 1) & 1);
        if (c.lane == 0) mbar_expect_tx(bar, gather_bytes((int)((c.gphase + 2) % 40), c.G));
    }
    consumer_bar();
    ++c.gphase;
}
// One warp pushes rows [0, G) x 16 fp32 columns of its private staging tile (row stride 16 floats) to dst[m][0..16)
// (row stride dld) of every peer.  Lanes 4 p .. 4 p + 3 serve peer p (one 16-byte column chunk each, all rows): the remote
// addresses are mapped once per lane and the loop has no index arithmetic.
TTS_D void push_tile_f32(const ClCtx& c, const float* wst, float* dst, int dld) {
    const uint32_t peer = (uint32_t)c.lane >> 2; const int pc = c.lane & 3;
    const uint32_t bar = map_to_rank(gather_bar(c), peer);
    const uint32_t rdst = map_to_rank(smem_u32(dst + pc * 4), peer);
    for (int m = 0; m < c.G; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(wst + m * 16 + pc * 4);
        st_async_v4(rdst + (uint32_t)(m * dld * 4), __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w), bar);
    }
}
// same, as bf16 (16 columns = two 16-byte chunks per row; the 2 G chunks of a peer go round its four lanes)
TTS_D void push_tile_bf16(const ClCtx& c, const float* wst, bf16* dst, int dld) {
    const uint32_t peer = (uint32_t)c.lane >> 2;
    const uint32_t bar = map_to_rank(gather_bar(c), peer);
    const uint32_t rdst = map_to_rank(smem_u32(dst), peer);
    for (int j = c.lane & 3; j < c.G * 2; j += 4) {
        const int m = j >> 1, pc = j & 1;
        const float4 v0 = *reinterpret_cast<const float4*>(wst + m * 16 + pc * 8);
        const float4 v1 = *reinterpret_cast<const float4*>(wst + m * 16 + pc * 8 + 4);
        st_async_v4(rdst + (uint32_t)((m * dld + pc * 8) * 2), pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w), bar);
    }
}

// LayerNorm of the gathered rows: ybuf -> xres (f32) + xa (bf16); warp m < G owns row m.  The affine parameters
// (global memory) are fetched asynchronously into shared memory by ln_prefetch() BEFORE the exchange the rows are
// waited on; the caller's gather_sync() orders both (cp.async.wait_group precedes its barrier).
TTS_D void ln_prefetch(const ClCtx& c, const float* g, const float* b) {
    float* dst = reinterpret_cast<float*>(c.smem + SM_RED + 4096);        // [0,512) gamma, [512,1024) beta
    if (c.tid < 256) {
        const float* src = (c.tid < 128 ? g : b) + (c.tid & 127) * 4;
        cp_async_16(dst + c.tid * 4, src, true);
    }
    cp_async_commit();
}
TTS_D void cl_layernorm(ClCtx& c, float ln_eps) {
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

// Attention of this CTA's pairs (utterance gi of the group, head = rank).  Warps 3 gi .. 3 gi + 2 own pair gi; the cache rows
// are cut into 16-row sub-chunks and sub-chunk c belongs to warp c % 3 of the pair -- a partition that depends on nothing but
// the row index, so an utterance's result is bit-identical whatever the group geometry.  Per sub-chunk, on the tensor cores:
//   scores[16 rows] = q . K_chunk^T               8 x mma.m16n8k16 (A = q in every row, B = K rows; two 8-row tiles)
//   out[64 d]      += V_chunk^T[64 x 16] . p      4 x mma.m16n8k16 (A = V^T sub-block in fragment order, B = p in every column)
// with one online-softmax round per warp and 128-row stage.  self: cache rows 0..t-1 plus the newest row (k_t, v_t) from qkvbuf; cross: rows 0..len-1.
// The pair's first warp merges the three partials (+ the newest row) in fixed order and pushes the pair's 64 outputs as bf16
// into abuf[gi][rank*64 ..] of every CTA.
TTS_D void cl_attention(const ClusterParams& p, ClCtx& c, bool self, int t, const volatile int* glens, unsigned long long* stamp_row) {
    const float* qkv = reinterpret_cast<const float*>(c.smem + SM_QKV);
    float* part = reinterpret_cast<float*>(c.smem + SM_AMERGE) + c.warp * 68;
    const int g = c.lane >> 2, t4 = c.lane & 3;
    const float qs = 0.125f * kLog2e;
    const int L = self ? t : p.S;
    const int nj = (L + KV_STAGE_ROWS - 1) / KV_STAGE_ROWS;
    const int gi = c.warp / ATT_WPP, wv = c.warp - gi * ATT_WPP;
    const bool active = gi < c.G;
    const int vlen = active ? (self ? t : min(L, glens[gi])) : 0;
    const int nsub = (vlen + 15) >> 4;
    if (active) {
        float m = -INFINITY, l = 0.f, o[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
        // B fragments of q, replicated in all 8 columns (every lane then holds valid scores: the row max needs only
        // 3 shuffles); k slots <-> dims 16 t4 + 4 ks + {0..3}
        uint32_t qb0[4], qb1[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float* qp = qkv + gi * 192 + 16 * t4 + 4 * ks;
            qb0[ks] = pack_bf16x2(qp[0] * qs, qp[1] * qs);
            qb1[ks] = pack_bf16x2(qp[2] * qs, qp[3] * qs);
        }
        // K row of this lane inside a 16-row sub-chunk (tile 0; tile 1 is two rows further) and its first 16-byte chunk
        const int krow_off = (4 * (g >> 1) + (g & 1)) * 128 + t4 * 32, kc0 = (g & 1) << 4;
        for (int j = 0; j < nj; ++j) {
            const uint32_t idx = c.consumed + (uint32_t)(j * c.G + gi);
            const unsigned char* st = cl_acquire(c, idx);
            // This warp's sub-chunks of the stage: c8 = c0, c0 + 3, c0 + 6 (< 8) with (8 j + c0) % 3 == wv; ONE online-softmax round
            // over all of them (the per-round cost -- shuffles, exp2 chain, rescale -- is what bounds a warp: scripts/ubench/attn_unit.cu).
            const int c0 = (wv + ATT_WPP - (8 * j) % ATT_WPP) % ATT_WPP;
            const int navail = min(8, nsub - 8 * j);     // sub-chunks of the stage that hold valid rows
            int nsc = 0;                                 // how many of them are this warp's (warp-uniform)
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) nsc += (c0 + 3 * s2 < navail) ? 1 : 0;
            if (nsc > 0) {
                // Scores: A = q, B = K rows.  Column n of tile jt <-> row 4 (n >> 1) + 2 jt + (n & 1) of the sub-chunk, so this lane ends
                // up with the scores of rows 4 t4 .. 4 t4 + 3 -- the rows its V fragments hold: no data movement between the two
                // products.  Odd columns load their two 16-byte chunks in swapped order (bank conflicts), i.e. see the k-steps in the
                // order 2, 3, 0, 1: A rows 8..15 carry q in that order; an odd column's score is read from row g + 8 (c3), an even
                // column's from row g (c0).
                float v[3][2][2];                        // [s2][tile][q] <-> row 128 j + 16 (c0 + 3 s2) + 4 t4 + 2 tile + q
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2) {
                    if (s2 < nsc) {
                        const int c8 = c0 + 3 * s2;
                        const unsigned char* kb = st + c8 * 4096 + krow_off;
                        float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                        uint4 kr[2][2];
#pragma unroll
                        for (int jt = 0; jt < 2; ++jt) {
                            kr[jt][0] = *reinterpret_cast<const uint4*>(kb + jt * 256 + kc0);
                            kr[jt][1] = *reinterpret_cast<const uint4*>(kb + jt * 256 + (kc0 ^ 16));
                        }
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t a[4] = {qb0[ks], qb0[ks ^ 2], qb1[ks], qb1[ks ^ 2]};
#pragma unroll
                            for (int jt = 0; jt < 2; ++jt) {
                                const uint4 w = kr[jt][ks >> 1];
                                mma_bf16_16816(sc[jt], a, (ks & 1) ? w.z : w.x, (ks & 1) ? w.w : w.y);
                            }
                        }
#pragma unroll
                        for (int jt = 0; jt < 2; ++jt) { v[s2][jt][0] = sc[jt][0]; v[s2][jt][1] = sc[jt][3]; }
                    } else {
#pragma unroll
                        for (int jt = 0; jt < 2; ++jt) { v[s2][jt][0] = v[s2][jt][1] = -INFINITY; }
                    }
                }
                if ((j + 1) * KV_STAGE_ROWS > vlen) {    // ragged tail (warp-uniform): mask the rows past the end
#pragma unroll
                    for (int s2 = 0; s2 < 3; ++s2)
#pragma unroll
                        for (int jt = 0; jt < 2; ++jt)
#pragma unroll
                            for (int q = 0; q < 2; ++q)
                                if (j * KV_STAGE_ROWS + (c0 + 3 * s2) * 16 + 4 * t4 + 2 * jt + q >= vlen) v[s2][jt][q] = -INFINITY;
                }
                float mx = -INFINITY;
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2) mx = fmaxf(mx, fmaxf(fmaxf(v[s2][0][0], v[s2][0][1]), fmaxf(v[s2][1][0], v[s2][1][1])));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float mnew = fmaxf(m, mx);         // finite: the warp's first sub-chunk holds at least one valid row
                if (mnew != m) {                         // warp-uniform: rescale only when the running max moved
                    const float scale = (m == -INFINITY) ? 0.f : fast_exp2(m - mnew);
                    l *= scale;
#pragma unroll
                    for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= scale; o[dt][1] *= scale; o[dt][2] *= scale; o[dt][3] *= scale; }
                    m = mnew;
                }
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2) {
                    if (s2 < nsc) {
                        const int c8 = c0 + 3 * s2;
                        const unsigned char* vb = st + c8 * 4096 + 2048 + c.lane * 16;
                        const float p0 = fast_exp2(v[s2][0][0] - mnew), p1 = fast_exp2(v[s2][0][1] - mnew);
                        const float p2 = fast_exp2(v[s2][1][0] - mnew), p3 = fast_exp2(v[s2][1][1] - mnew);
                        l += (p0 + p1) + (p2 + p3);      // this lane's rows only (the other 7 lanes of the column group hold copies)
                        const uint32_t b0 = pack_bf16x2(p0, p1), b1 = pack_bf16x2(p2, p3);   // B = p in every column: k slots <-> rows 4 t4 + {0..3}
#pragma unroll
                        for (int dt = 0; dt < 4; ++dt) { // A = V^T sub-block, stored fragment-major: one 16-byte load per lane and tile
                            const uint4 w = *reinterpret_cast<const uint4*>(vb + dt * 512);
                            const uint32_t a[4] = {w.x, w.y, w.z, w.w};
                            mma_bf16_16816(o[dt], a, b0, b1);
                        }
                    }
                }
            }
            cl_release(c, idx);
        }
        // partial (m, l, o[64]) of this warp -> scratch
        float ls = l + __shfl_xor_sync(0xffffffffu, l, 1);                     // the 4 lanes t4 = 0..3 hold disjoint rows
        ls += __shfl_xor_sync(0xffffffffu, ls, 2);
        if (t4 == 0) {
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) { part[dt * 16 + g] = o[dt][0]; part[dt * 16 + g + 8] = o[dt][2]; }
        }
        if (c.lane == 0) { part[64] = m; part[65] = ls; }
    }
    c.consumed += (uint32_t)(nj * c.G);
    if (stamp_row) stamp_row[self ? 61 : 64] = timer_after_lds(c.smem + SM_MISC + 192);
    // the pair's three warps only (named barrier 2 + gi): a pair that is done merges and pushes while the others still stream
    if (active) asm volatile("bar.sync %0, 96;" ::"r"(2 + gi) : "memory");
    if (stamp_row) stamp_row[self ? 62 : 65] = timer_after_lds(c.smem + SM_MISC + 192);
    if (active && wv == 0) {                             // merge the pair's three partials (+ the newest row) in fixed order
        const float* pa = part;
        float mm = fmaxf(fmaxf(pa[64], pa[68 + 64]), pa[136 + 64]);
        float st_new = -INFINITY;
        if (self) {                                      // newest row: bf16-rounded q, k_t, v_t exactly as the MMA path would see them
            const float* qp = qkv + gi * 192 + 2 * c.lane;
            const float2 qd = unpack_bf16x2(pack_bf16x2(qp[0] * qs, qp[1] * qs));
            const float2 kd = unpack_bf16x2(pack_bf16x2(qp[64], qp[65]));
            st_new = warp_sum(qd.x * kd.x + qd.y * kd.y);
            mm = fmaxf(mm, st_new);
        }
        float lsum = 0.f, v0 = 0.f, v1 = 0.f;
        if (mm > -INFINITY) {
#pragma unroll
            for (int q = 0; q < ATT_WPP; ++q) {
                const float mq = pa[q * 68 + 64];
                const float e = (mq == -INFINITY) ? 0.f : fast_exp2(mq - mm);
                lsum += pa[q * 68 + 65] * e;
                v0 += pa[q * 68 + c.lane] * e; v1 += pa[q * 68 + 32 + c.lane] * e;
            }
            if (self) {
                const float pt = fast_exp2(st_new - mm);
                lsum += pt;
                v0 += pt * __bfloat162float(__float2bfloat16(qkv[gi * 192 + 128 + c.lane]));
                v1 += pt * __bfloat162float(__float2bfloat16(qkv[gi * 192 + 128 + 32 + c.lane]));
            }
        }
        const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
        float* ost = reinterpret_cast<float*>(c.smem + SM_RED) + gi * 64;
        ost[c.lane] = v0 * inv; ost[32 + c.lane] = v1 * inv;
        __syncwarp();
        bf16* dst = reinterpret_cast<bf16*>(c.smem + SM_ABUF) + gi * LDX512 + c.rank * 64;
        const uint32_t bar = gather_bar(c);
#pragma unroll
        for (int k = 0; k < 2; ++k) {                    // 8 peers x 8 chunks of 8 bf16
            const int i = c.lane + 32 * k, peer = i >> 3, ch = i & 7;
            const float4 a = *reinterpret_cast<const float4*>(ost + ch * 8), b = *reinterpret_cast<const float4*>(ost + ch * 8 + 4);
            st_async_v4(map_to_rank(smem_u32(dst + ch * 8), (uint32_t)peer), pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w),
                        pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w), map_to_rank(bar, (uint32_t)peer));
        }
    }
    if (stamp_row) stamp_row[self ? 63 : 66] = timer_after_lds(c.smem + SM_MISC + 192);
}

// ---------------------------------------------------------------- the kernel
// DBG = false is the product kernel: the %globaltimer stamps and the debug dumps are compiled out (the kernel is bound by
// instruction fetch at every phase boundary -- profiles/r02_decode_summary.md -- so code that never runs still costs time).
template <bool DBG>
__global__ void __launch_bounds__(CL_THREADS, 1) decode_cluster_kernel(const __grid_constant__ ClusterParams p, int t0, int n_steps) {
    extern __shared__ __align__(128) unsigned char cl_smem[];
    ClCtx c;
    c.smem = cl_smem;
    c.full = reinterpret_cast<uint64_t*>(cl_smem + SM_MISC);
    c.empty = c.full + CL_FULL_BARS;
    c.gsync = c.empty + CL_STAGES;
    // flags[0] finished utterances of the group (rank 5 counts), [1] consumers done (producer stop), [2] final consumed count
    // hrec: one 16-byte record per rank, pushed in the head phase (rank 5: finished count); glens: phoneme lengths of the group
    volatile int* flags = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 192);
    volatile int* hrec = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 192 + 32);             // [8 ranks][4]
    volatile int* glens = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 192 + 32 + 128);      // [CL_G]
    volatile int* guids = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 384);                 // [CL_G] global utterance ids (dropout key)
    volatile int* gtlens = reinterpret_cast<volatile int*>(cl_smem + SM_MISC + 416);                // [CL_G] frame budgets; [7] = next group (queue mode)
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
    float* wst = reinterpret_cast<float*>(cl_smem + SM_WST) + (c.warp & 3) * (CL_G * 16);    // private epilogue tile [G][16] of warps 0..3 (the narrow GEMMs)
    const bool stamper = DBG && p.ts != nullptr && cid == 0 && c.rank == 0 && c.tid == 0;
    auto stamp = [&](int t, int idx) {
        if (stamper) {
            // BAR.SYNC does not block at issue (the warp stalls at the next instruction that touches barrier-protected state): a
            // shared-memory load in front of the timer read makes the stamp the barrier's RELEASE time, not its issue time
            unsigned long long now; uint32_t dummy;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dummy) : "r"(smem_u32(cl_smem + SM_MISC + 192)) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "r"(dummy) : "memory");      // input operand: issues after the load returned
            p.ts[(size_t)t * CL_TS_COLS + idx] = now;
        }
    };

    // debug dump: slot `slot` <- rows [0, G) x n values (row stride ld) of a shared-memory buffer, widened to fp32
    const bool dumper = DBG && p.dbg != nullptr && cid == 0 && c.rank == p.dbg_rank && !is_producer;
    auto dbg_dump = [&](int t, int slot, const void* src, int ld, int n, bool is_bf16) {
        if (dumper && t == t0) {
            for (int i = c.tid; i < c.G * n; i += CL_CONSUMERS) {
                const int m = i / n, k = i - m * n;
                p.dbg[(size_t)slot * 2560 + i] = is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(src)[m * ld + k]) : reinterpret_cast<const float*>(src)[m * ld + k];
            }
        }
    };

    for (int grp = cid;; grp += ncl) {
        if (p.group_queue != nullptr) {                  // work stealing: CTA 0 draws the cluster's next group and tells its peers
            if (c.rank == 0 && c.tid == 0) {
                const int gq = atomicAdd(p.group_queue, 1);
                for (int r = 0; r < CL_SIZE; ++r)
                    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(map_to_rank(smem_u32(const_cast<int*>(gtlens) + 7), (uint32_t)r)), "r"(gq) : "memory");
            }
            hw_cluster_sync();
            grp = gtlens[7];
        }
        if (grp >= p.ngroups) break;
        c.b0 = grp * p.G; c.G = min(p.G, p.B - c.b0);
        // ---- (re)initialise the ring and the activation buffers
        if (!is_producer)
            for (int i = c.tid; i < (SM_MISC - SM_XRES) / 4; i += CL_CONSUMERS) reinterpret_cast<uint32_t*>(cl_smem + SM_XRES)[i] = 0u;
        if (c.tid == 0) {
            for (int s = 0; s < CL_FULL_BARS; ++s) mbar_init(&c.full[s], 1);
            for (int s = 0; s < CL_STAGES; ++s) mbar_init(&c.empty[s], CL_WARPS);
            mbar_init(&c.gsync[0], 1); mbar_init(&c.gsync[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&c.gsync[0], gather_bytes(0, c.G));          // phases 0 and 1 are armed before the start barrier
            mbar_expect_tx(&c.gsync[1], gather_bytes(1, c.G));
            int nf = 0;
            for (int m = 0; m < c.G; ++m) {
                nf += p.finished[c.b0 + m]; glens[m] = p.plens[c.b0 + m];
                guids[m] = p.utt_ids ? p.utt_ids[c.b0 + m] : p.utt_offset + c.b0 + m;
                gtlens[m] = p.tlens ? min(p.Tmax, p.tlens[c.b0 + m]) : p.Tmax;
            }
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
            if (c.lane == 0) cl_producer(p, cl_smem, c.full, c.empty, flags, glens, c.rank, c.b0, c.G, t0, t0 + n_steps);
            __syncwarp();                                // lanes 1..31 park here: the cluster barrier below is .aligned
        } else {
            for (int t = t0; t < t0 + n_steps; ++t) {
                // ================= decoder prenet (dropout always on, P7) =================
                cl_gemm<4, 1>(c, 2, fbuf, LDX128,
                        [&](int ti, int n) { return __ldg(p.b_fc1 + c.rank * 32 + ti * 16 + n); },
                        [&](int ti, int n, int m, float v) {
                            const int col = c.rank * 32 + ti * 16 + n;
                            v = fmaxf(v, 0.f);
                            wst[m * 16 + n] = keep_bit(p.seed, SITE_DEC_PRENET_FC1, (uint32_t)t, (uint32_t)guids[m], (uint32_t)col) ? 2.f * v : 0.f;
                        });
                if (c.warp < 2) { __syncwarp(); push_tile_bf16(c, wst, h1 + c.rank * 32 + c.warp * 16, LDX256); }
                gather_sync(c);
                stamp(t, 0);
                cl_gemm<8, 1>(c, 2, h1, LDX256,
                        [&](int ti, int n) { return __ldg(p.b_fc2 + c.rank * 32 + ti * 16 + n); },
                        [&](int ti, int n, int m, float v) {
                            const int col = c.rank * 32 + ti * 16 + n;
                            v = fmaxf(v, 0.f);
                            wst[m * 16 + n] = keep_bit(p.seed, SITE_DEC_PRENET_FC2, (uint32_t)t, (uint32_t)guids[m], (uint32_t)col) ? 2.f * v : 0.f;
                        });
                if (c.warp < 2) { __syncwarp(); push_tile_bf16(c, wst, h2 + c.rank * 32 + c.warp * 16, LDX256); }
                gather_sync(c);
                stamp(t, 1);
                cl_gemm<8, 1>(c, 4, h2, LDX256,
                        [&](int ti, int n) {
                            const int col = c.rank * CL_NS + ti * 16 + n;
                            return __ldg(p.b_proj + col) + p.dec_alpha * __ldg(p.pe + (size_t)t * kDModel + col);
                        },
                        [&](int, int n, int m, float v) { wst[m * 16 + n] = v; });
                if (c.warp < 4) {
                    __syncwarp();
                    push_tile_f32(c, wst, xres + c.rank * CL_NS + c.warp * 16, 512);
                    push_tile_bf16(c, wst, xa + c.rank * CL_NS + c.warp * 16, LDX512);
                }
                gather_sync(c);
                stamp(t, 2);
                dbg_dump(t, 0, xa, LDX512, 512, true);

                for (int l = 0; l < 6; ++l) {
                    const ClusterLayerParams& W = p.layer[l];
                    // ---- q, k, v of head `rank` for every row of the group (local; k_t, v_t appended to the cache)
                    cl_gemm<16, 1, 2>(c, 12, xa, LDX512,
                            [&](int ti, int n) { const int cc = ti * 16 + n; return __ldg(W.bqkv + (cc >> 6) * 512 + c.rank * 64 + (cc & 63)); },
                            [&](int ti, int n, int m, float v) {
                                const int cc = ti * 16 + n, part = cc >> 6, dd = cc & 63;
                                qkvb[m * 192 + cc] = v;
                                if (part > 0) {               // append row t to the cache block t / 64
                                    bf16* blk = p.self_kv + kv_block_offset(l, p.B, c.b0 + m, c.rank, p.nblk_self, t >> 6);
                                    blk[part == 1 ? kv_k_elem(t & 63, dd) : kv_v_elem(t & 63, dd)] = __float2bfloat16(v);
                                }
                            });
                    consumer_bar();
                    stamp(t, 3 + 8 * l);
                    if (l == 0) dbg_dump(t, 1, qkvb, 192, 192, false);
                    cl_attention(p, c, true, t, glens, (stamper && l == 0) ? p.ts + (size_t)t * CL_TS_COLS : nullptr);
                    gather_sync(c);
                    stamp(t, 4 + 8 * l);
                    if (l == 0) dbg_dump(t, 2, abuf, LDX512, 512, true);
                    // ---- O projection + residual, gathered -> LayerNorm 1
                    ln_prefetch(c, W.ln1g, W.ln1b);
                    cl_gemm_ksplit<4, 16, 4>(c, abuf, LDX512,
                            [&](int ti, int n) { return __ldg(W.bo + c.rank * CL_NS + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) { wst[m * 16 + n] = v + xres[m * 512 + c.rank * CL_NS + ti * 16 + n]; },
                            (stamper && l == 0) ? p.ts + (size_t)t * CL_TS_COLS + 73 : nullptr);
                    if (l == 0) stamp(t, 52);
                    if (c.warp < 4) { __syncwarp(); push_tile_f32(c, wst, ybuf + c.rank * CL_NS + c.warp * 16, 512); }
                    if (l == 0) stamp(t, 53);
                    cp_async_wait<0>();
                    gather_sync(c);
                    if (l == 0) stamp(t, 54);
                    cl_layernorm(c, p.ln_eps);
                    stamp(t, 5 + 8 * l);
                    if (l == 0) dbg_dump(t, 3, xa, LDX512, 512, true);
                    // ---- cross-attention query of head `rank` (local)
                    cl_gemm_ksplit<4, 16, 4>(c, xa, LDX512,
                            [&](int ti, int n) { return __ldg(W.bq2 + c.rank * CL_NS + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) { qkvb[m * 192 + ti * 16 + n] = v; });
                    consumer_bar();
                    stamp(t, 6 + 8 * l);
                    cl_attention(p, c, false, t, glens, (stamper && l == 0) ? p.ts + (size_t)t * CL_TS_COLS : nullptr);
                    gather_sync(c);
                    stamp(t, 7 + 8 * l);
                    if (l == 0) dbg_dump(t, 4, abuf, LDX512, 512, true);
                    ln_prefetch(c, W.ln2g, W.ln2b);
                    cl_gemm_ksplit<4, 16, 4>(c, abuf, LDX512,
                            [&](int ti, int n) { return __ldg(W.bo2 + c.rank * CL_NS + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) { wst[m * 16 + n] = v + xres[m * 512 + c.rank * CL_NS + ti * 16 + n]; });
                    if (c.warp < 4) { __syncwarp(); push_tile_f32(c, wst, ybuf + c.rank * CL_NS + c.warp * 16, 512); }
                    cp_async_wait<0>();
                    gather_sync(c);
                    cl_layernorm(c, p.ln_eps);
                    stamp(t, 8 + 8 * l);
                    if (l == 0) dbg_dump(t, 5, xa, LDX512, 512, true);
                    // ---- FFN: hidden slice [256 rank, +256) stays local (bf16); FFN2 is split along K
                    cl_gemm<16, 1, 2>(c, 16, xa, LDX512,
                            [&](int ti, int n) { return __ldg(W.b1 + c.rank * 256 + ti * 16 + n); },
                            [&](int ti, int n, int m, float v) { hbuf[m * LDX256 + ti * 16 + n] = __float2bfloat16(fmaxf(v, 0.f)); },
                            (stamper && l == 0) ? p.ts + (size_t)t * CL_TS_COLS + 76 : nullptr);
                    consumer_bar();
                    stamp(t, 9 + 8 * l);
                    if (l == 0) dbg_dump(t, 6, hbuf, LDX256, 256, true);
                    // partial sums over this rank's 256 hidden units, staged in ybuf (free between LN2 and the y3 gather)
                    ln_prefetch(c, W.ln3g, W.ln3b);
                    cl_gemm<8, 2, 2>(c, 16, hbuf, LDX256, [&](int, int) { return 0.f; },
                            [&](int ti, int n, int m, float v) { ybuf[m * 512 + ti * 16 + n] = v; },
                            (stamper && l == 0) ? p.ts + (size_t)t * CL_TS_COLS + 79 : nullptr);
                    if (l == 0) stamp(t, 55);
                    {   // reduce-scatter, warp by warp: warp w's 32 columns all belong to rank w / 2 -- no block barrier, and the
                        // pushes of the early warps overlap the weight stream of the late ones
                        __syncwarp();
                        const uint32_t bar = gather_bar(c);
                        const int peer = c.warp >> 1;
                        for (int i = c.lane; i < c.G * 8; i += 32) {
                            const int m = i >> 3, pc = (c.warp & 1) * 8 + (i & 7);
                            const float4 v = *reinterpret_cast<const float4*>(ybuf + m * 512 + peer * CL_NS + pc * 4);
                            st_async_v4(map_to_rank(smem_u32(recv + (c.rank * CL_G + m) * CL_NS + pc * 4), (uint32_t)peer),
                                        __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w),
                                        map_to_rank(bar, (uint32_t)peer));
                        }
                    }
                    if (l == 0) stamp(t, 57);
                    // my 64 columns: 8 partials (fixed order) + bias + residual -> every peer's ybuf.  Four threads per (row, 4-column
                    // chunk), two peers each; bias and residual are fetched before the wait
                    const bool reducer = c.tid < c.G * 64;
                    const int rm = c.tid >> 6, rpc = (c.tid >> 2) & 15, rcol = c.rank * CL_NS + rpc * 4, rq = c.tid & 3;
                    float4 racc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (reducer) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(W.b2 + rcol));
                        const float4 xr = *reinterpret_cast<const float4*>(xres + rm * 512 + rcol);
                        racc = make_float4(bb.x + xr.x, bb.y + xr.y, bb.z + xr.z, bb.w + xr.w);
                    }
                    gather_sync(c);
                    if (l == 0) stamp(t, 58);
                    if (l == 0) dbg_dump(t, 10, recv, 2560, 2560, false);
                    if (reducer) {
                        float4 v = racc;
#pragma unroll
                        for (int r = 0; r < CL_SIZE; ++r) {
                            const float4 q = *reinterpret_cast<const float4*>(recv + (r * CL_G + rm) * CL_NS + rpc * 4);
                            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                        }
                        const uint32_t bar = gather_bar(c), dsta = smem_u32(ybuf + rm * 512 + rcol);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t peer = (uint32_t)(2 * rq + k);
                            st_async_v4(map_to_rank(dsta, peer), __float_as_uint(v.x), __float_as_uint(v.y),
                                        __float_as_uint(v.z), __float_as_uint(v.w), map_to_rank(bar, peer));
                        }
                    }
                    if (l == 0) stamp(t, 59);
                    cp_async_wait<0>();
                    gather_sync(c);
                    if (l == 0) stamp(t, 60);
                    if (l == 0) dbg_dump(t, 9, ybuf, 512, 512, false);
                    cl_layernorm(c, p.ln_eps);
                    stamp(t, 10 + 8 * l);
                    if (l == 0) dbg_dump(t, 7, xa, LDX512, 512, true);
                }
                // The cache rows appended in this step are read by the async proxy (bulk copies) from the next step on: order the
                // generic-proxy stores before them here, once per step and long after the stores were issued (a fence next to every
                // append would wait for the stores of that very moment).
                asm volatile("fence.proxy.async;" ::: "memory");
                // ================= [mel | stop] heads: ranks 0..5 own 16 of the 81(+15) columns =================
                if (c.rank < 6) {
                    cl_gemm<16, 1>(c, 1, xa, LDX512,
                            [&](int, int n) { const int col = c.rank * 16 + n; return col <= 80 ? __ldg(p.b_head + col) : 0.f; },
                            [&](int, int n, int m, float v) {
                                const int col = c.rank * 16 + n, b = c.b0 + m;
                                if (col < 80) p.mel_before[((size_t)b * p.Tmax + t) * 80 + col] = v;   // fp32 feedback (P8)
                                else if (col == 80) {
                                    p.stop_logits[(size_t)b * p.Tmax + t] = v;
                                    if ((v > 0.f || t + 1 >= gtlens[m]) && p.finished[b] == 0) { // P10 (or the utterance's frame budget is used up)
                                        p.finished[b] = 1; p.lens[b] = t + 1; atomicAdd(p.n_finished, 1);
                                        atomicAdd(const_cast<int*>(flags), 1);                   // rank 5 keeps the group's count
                                    }
                                }
                                wst[m * 16 + n] = v;
                            });
                }
                if (c.warp == 0) {
                    __syncwarp();
                    if (c.rank < 5) push_tile_bf16(c, wst, fbuf + c.rank * 16, LDX128);           // the frame = next step's prenet input
                    if (c.lane < CL_SIZE) {              // every rank contributes one record (rank 5: finished count) to every peer
                        const uint32_t bar = gather_bar(c);
                        st_async_v4(map_to_rank(smem_u32(const_cast<int*>(hrec) + c.rank * 4), (uint32_t)c.lane),
                                    (uint32_t)flags[0], 0u, 0u, 0u, map_to_rank(bar, (uint32_t)c.lane));
                    }
                }
                gather_sync(c);
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
