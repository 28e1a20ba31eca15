/* tts_b200.h -- C ABI of libtts_b200.so: the B200-native (sm_100a) Transformer-TTS hot path.
 *
 * What this boundary replaces.  The reference repository (keonlee9420/Transformer-tacotron2)
 * ships NO code -- /root/reference/README.md:1-3 is its whole content (title, one sentence and
 * the link to Li et al., AAAI 2019 at README.md:3).  BASELINE.json `north_star` therefore defines
 * the interface to keep as a torch.nn.Module:
 *
 *     TransformerTTS.forward(phonemes, phoneme_lens, mels, mel_lens)   -> mel_before, mel_after, stop_logits
 *     TransformerTTS.inference(phonemes, phoneme_lens, max_len, seed)  -> mel_after, mel_lens, stop_logits
 *     + the training loop around them: model.train(); loss = tts_loss(model(...)); loss.backward(); all-reduce; Adam.step()
 *
 * (restated executable in oracle/transformer_tts.py:TransformerTTS).  The entry
 * points below are exactly what a Python binding of those two methods calls (the ctypes stub is in
 * INTEGRATION.md and, live, in transformer_tacotron2_b200/_lib.py).  Each declaration cites the
 * interface it stands behind.
 *
 * Conventions: plain C types only; raw DEVICE pointers + explicit sizes + a cudaStream_t passed as
 * void*; the caller owns every buffer including the workspace (sized by tts_workspace_bytes); the
 * handle owns only the packed weights.  All work is enqueued on the caller's stream; calls marked
 * [sync] synchronise that stream.  Return value: 0 = ok, < 0 = argument error (TTS_E_*),
 * > 0 = cudaError_t.  tts_last_error_string() gives detail.  A handle is bound to one device and is
 * not thread-safe.  There is no CPU fallback: without an sm_100 device tts_create fails.
 */
#ifndef TTS_B200_H_
#define TTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTS_E_ARG (-1)       /* bad argument / shape */
#define TTS_E_STATE (-2)     /* call order (weights not finalised, decode not begun, ...) */
#define TTS_E_WEIGHT (-3)    /* unknown / missing / mis-sized weight */
#define TTS_E_DEVICE (-4)    /* no sm_100 device */

typedef struct TtsHandle TtsHandle;

/* oracle/transformer_tts.py:TTSConfig (SURVEY.md section 5 "Config / flags").  Only the base model
 * (d_model 512, 8 heads, d_ff 2048, 80 mels, prenet 256, kernel 5) is accepted. */
typedef struct TtsConfig {
    uint32_t struct_size;
    int32_t n_vocab, d_model, n_heads, n_enc_layers, n_dec_layers, d_ff, n_mels, d_prenet;
    int32_t enc_conv_layers, conv_kernel, postnet_channels, postnet_layers, max_pos;
    float ln_eps, bn_eps;
} TtsConfig;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* TransformerTTS.__init__ (oracle/transformer_tts.py:TransformerTTS.__init__). */
int tts_create(const TtsConfig* cfg, int device, TtsHandle** out);
int tts_destroy(TtsHandle* h);
const char* tts_last_error_string(TtsHandle* h);
/* Library / build identification ("tts_b200 <version> sm_100a"). */
const char* tts_version(void);
/* Kernel launches issued by the library since it was loaded (bench.py: gpu_launches). */
unsigned long long tts_launch_count(void);

/* ---- weights: nn.Module.load_state_dict (same keys as the oracle's state_dict) --------------- */
/* Stage one fp32 tensor (HOST pointer, `numel` elements, C-contiguous) under its state_dict key. */
int tts_load_weight(TtsHandle* h, const char* name, const float* host_data, int64_t numel);
/* Pack everything staged so far: bf16 cast, BatchNorm folded into the convs (P5), Q/K/V and
 * [mel|stop] concatenated, decode-step weights swizzled into MMA fragment order; uploads. [sync] */
int tts_finalize_weights(TtsHandle* h);

/* ---- options ---------------------------------------------------------------------------------- */
/* "cluster_group": utterances per 8-CTA cluster of the decode kernel, 1..5 (0 = auto: as few as the number of
 *   co-resident clusters allows);  "decode_timestamps": 1 record per-phase timestamps (tts_debug_phase_timestamps);
 *   "decode_debug": 1 dump layer-0 activations of the first step (tts_debug_read_dump);
 *   "train_graph": 1 (default) replay tts_train_step from a CUDA graph captured on the second step of a shape, 0 eager;
 *   "print_info": print device / cluster geometry to stderr. */
int tts_set_option(TtsHandle* h, const char* key, int64_t value);

/* ---- workspace ---------------------------------------------------------------------------------- */
/* Bytes of caller-owned device scratch (KV caches, activations) for batch B, S phonemes, T frames. */
size_t tts_workspace_bytes(TtsHandle* h, int B, int S, int T);

/* ---- TransformerTTS.inference (oracle/transformer_tts.py:TransformerTTS.inference) ---------- */
/* Encoder + hoisted cross-K/V projection.  phonemes [B][S] int64, phoneme_lens [B] int32 (device).
 * T = the frame count the workspace was sized for (max_len of the decode that follows).
 * memory_out: optional fp32 [B][S][512] copy of the encoder output (may be NULL). */
int tts_encode(TtsHandle* h, void* ws, const int64_t* phonemes, const int32_t* phoneme_lens,
               int B, int S, int T, float* memory_out, void* stream);
/* Reset the AR state (KV cache cursor, lengths, stop flags).  utt_offset = global id of utterance 0
 * (dropout masks are keyed by global utterance id so a sharded batch reproduces the unsharded one). */
int tts_decode_begin(TtsHandle* h, void* ws, int B, int S, int max_len, uint64_t seed, int utt_offset, void* stream);
/* Run up to n_steps decoder steps (stops early, on the device, once every utterance has fired). */
int tts_decode_steps(TtsHandle* h, void* ws, int n_steps, void* stream);
/* [sync] steps completed so far and number of finished utterances. */
int tts_decode_status(TtsHandle* h, void* ws, int* t_done, int* n_finished, void* stream);
/* Mask past the stop frame, run the postnet, add the residual.  T_out = t_done.  Outputs (device):
 * mel_after [B][T_out][80] f32, mel_lens [B] i32, stop_logits [B][T_out] f32, mel_before (optional). */
int tts_decode_end(TtsHandle* h, void* ws, int T_out, float* mel_after, int32_t* mel_lens,
                   float* stop_logits, float* mel_before, void* stream);
/* The whole of .inference() with HOST buffers (pageable or pinned): H2D, encode, decode loop,
 * postnet, D2H.  Outputs are written compactly as [B][T_out]...; *T_out returns the frames run.
 * `ws` is device scratch of tts_workspace_bytes(B, S, max_len). [sync] */
int tts_infer_host(TtsHandle* h, void* ws, const int64_t* phonemes, const int32_t* phoneme_lens, int B, int S,
                   int max_len, uint64_t seed, int utt_offset, float* mel_after, int32_t* mel_lens,
                   float* stop_logits, int* T_out, void* stream);

/* ---- TransformerTTS.forward, eval mode (oracle/transformer_tts.py:TransformerTTS.forward) --- */
/* Teacher-forced forward.  Device pointers: phonemes [B][S] i64, lens i32, mels [B][T][80] f32,
 * outputs mel_before / mel_after [B][T][80] f32 and stop_logits [B][T] f32 (zero past mel_lens). */
int tts_forward(TtsHandle* h, void* ws, const int64_t* phonemes, const int32_t* phoneme_lens,
                const float* mels, const int32_t* mel_lens, int B, int S, int T, uint64_t seed, int utt_offset,
                float* mel_before, float* mel_after, float* stop_logits, void* stream);

/* Optional per-utterance controls of a decode session (between tts_decode_begin and the first tts_decode_steps; device int32 [B],
 * either may be NULL):
 *   utt_ids   global utterance id of every row = its dropout key (default utt_offset + row).  With ids the batch may be decoded in
 *             ANY order -- e.g. sorted by length so that the utterances of a cluster group stop together -- with bit-identical
 *             per-utterance results;
 *   max_lens  per-utterance frame budget <= max_len: the utterance also stops (length = budget) when it is used up;
 *   work_stealing != 0: the clusters draw their utterance groups from a device-side queue instead of a static round-robin, so a
 *             cluster whose group has stopped immediately starts the next one (8(f)-3: continuous batching at group granularity). */
int tts_decode_set_batch(TtsHandle* h, void* ws, const int32_t* utt_ids, const int32_t* max_lens, int work_stealing, void* stream);

/* Forced ("step-locked") decoding: overwrite frame t (< frames decoded so far) of the session's mel_before buffer with
 * frames [B][80] fp32 (device memory).  The next tts_decode_steps() call resumes from frame t_done - 1, so writing frame
 * t_done - 1 between single-step calls feeds the decoder somebody else's trajectory (the parity tests feed the oracle's). */
int tts_decode_set_frame(TtsHandle* h, void* ws, int t, const float* frames, void* stream);
/* Read frame t of the session back: frames [B][80] fp32 (mel_before) and, if not NULL, stop_logits [B] (device memory). */
int tts_decode_get_frame(TtsHandle* h, void* ws, int t, float* frames, float* stop_logits, void* stream);

/* Profiling aid: with option "decode_timestamps" = 1 the persistent decode kernel stamps %globaltimer
 * after every phase; this copies [n_steps][n_phases] stamps (ns) to the host and returns n_phases. [sync] */
int tts_debug_phase_timestamps(TtsHandle* h, void* ws, unsigned long long* out, int n_steps, void* stream);
/* Debugging aid: with option "decode_debug" = 1 cluster 0 / CTA 0 of the decode kernel dumps the activations of the group's
 * rows after the prenet and after every sub-layer of decoder layer 0 at the first step of each launch (slots of 2560 floats,
 * see decode_cluster.cuh: dbg_dump); this copies `n` floats starting at float `offset` of that dump to the host. [sync] */
int tts_debug_read_dump(TtsHandle* h, void* ws, int64_t offset, int64_t n, float* out_host, void* stream);

/* Layout of the decode kernel's K/V cache (host arithmetic only, no device work): element index (0 .. 8191) of K[row][dim]
 * (which = 0) or V[row][dim] (which = 1) inside a 64-row cache block; row in [0, 64), dim in [0, 64).  -1 on bad arguments.
 * The layout is part of the kernel contract (a ring stage must be one contiguous copy, a V fragment one 16-byte load), so the
 * CPU test suite checks it without a GPU. */
int tts_debug_kv_index(int row, int dim, int which);
/* The decode kernel's weight stream (host arithmetic only): packs rows `rows[0 .. nrows)` (nrows a multiple of 16 * TW; -1 = zero
 * row) x k-pairs [kp_base, kp_base + KP) (32 columns each) of the fp32 matrix w[N][K] exactly as tts_finalize_weights does for one
 * segment of one rank -- per warp a contiguous run of [k-pair][tile] blocks of 1 KB in mma.m16n8k16 A-fragment order, bf16.
 * Writes nrows / 16 * KP KB to `out` (returns the byte count, or the count needed if out is NULL / too small; -1 on bad
 * arguments).  The CPU tests re-read the stream with the kernel's addressing. */
int64_t tts_debug_pack_segment(const float* w, int N, int K, const int32_t* rows, int nrows, int kp_base, int KP, int TW,
                               unsigned char* out, int64_t out_bytes);

/* ---- training step (oracle: TransformerTTS.forward in .train() mode + tts_loss + autograd + torch.optim.Adam;
 *      oracle/transformer_tts.py:forward, :tts_loss, :masked_batchnorm) ---------------- */
/* Build the training state from the weights loaded so far: one flat fp32 parameter buffer (+ gradients, Adam
 * moments), BatchNorm running statistics, bf16 operand copies.  [sync] */
int tts_train_begin(TtsHandle* h);
int tts_train_end(TtsHandle* h);
size_t tts_train_workspace_bytes(TtsHandle* h, int B, int S, int T);
/* Train-mode forward (batch-statistics BatchNorm, Philox dropout at every site), loss, full backward.  All pointers
 * are DEVICE pointers.  Afterwards the gradient of every parameter is in the flat gradient buffer (tts_train_grads),
 * loss_out[0] holds the scalar loss, and BatchNorm running statistics have been updated. */
int tts_train_step(TtsHandle* h, void* workspace, const int64_t* phonemes, const int32_t* phoneme_lens, const float* mels,
                   const int32_t* mel_lens, int B, int S, int T, uint64_t seed, int utt_offset, double p_residual,
                   float pos_weight, float* loss_out, void* stream);
/* Forward outputs of the last tts_train_step ([B][T][80], [B][T][80], [B][T] fp32, device pointers, any may be null). */
int tts_train_outputs(TtsHandle* h, void* workspace, int B, int S, int T, float* mel_before, float* mel_after,
                      float* stop_logits, void* stream);
/* The flat fp32 gradient buffer (device pointer, numel): the data-parallel exchange is ONE all-reduce over it. */
int tts_train_grads(TtsHandle* h, float** grads_dev, int64_t* numel);
/* Adam over the flat buffers (g <- grad * grad_scale, e.g. 1 / world_size after a sum all-reduce), then refresh the
 * bf16 operand copies. */
int tts_train_adam(TtsHandle* h, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* Data-parallel optimiser step fused with its collective: CUDA-IPC handles (64 bytes each) of this rank's
 * flat parameter and gradient buffers; tts_train_set_peers maps every rank's buffers (handles_* = [world][64] bytes, in rank
 * order; world <= 8, one node).  tts_train_adam_peers then runs ONE kernel that, for this rank's 1/world shard, sums the
 * gradient over all ranks by NVLink peer loads (reduce-scatter), applies Adam with shard-local moments (ZeRO-1), and stores
 * the new parameters into every rank's buffer by peer stores (all-gather).  The caller brackets it with two cross-rank
 * barriers on the stream and calls tts_train_repack afterwards. */
int tts_train_ipc_handles(TtsHandle* h, void* handle_P_64, void* handle_G_64);
int tts_train_set_peers(TtsHandle* h, int rank, int world, const void* handles_P, const void* handles_G);
int tts_train_adam_peers(TtsHandle* h, float lr, float beta1, float beta2, float eps, void* stream);
int tts_train_repack(TtsHandle* h, void* stream);
/* Enumerate the tensors inside the flat buffers: state_dict name, offset, numel; is_buffer = 1 for running stats. */
int tts_train_num_tensors(TtsHandle* h);
int tts_train_tensor_info(TtsHandle* h, int index, const char** name, int64_t* offset, int64_t* numel, int* is_buffer);
/* Copy a range of the parameter (which = 0), gradient (1) or running-statistics (2) buffer to the host. [sync] */
int tts_train_read(TtsHandle* h, int which, int64_t offset, int64_t numel, float* host_out);
/* Overwrite a range of the parameter (which = 0) or running-statistics (2) buffer from the host; parameters take effect after
 * tts_train_repack() (which refreshes the bf16 operand copies). [sync] */
int tts_train_write(TtsHandle* h, int which, int64_t offset, int64_t numel, const float* host_in);

/* ---- autograd bridge (the module's train-mode forward(): oracle `model.train(); out = model(...); loss.backward()`) ----------
 * tts_train_forward: the train-mode forward of tts_train_step alone (batch-statistics BatchNorm, every dropout site on); the
 *   activations stay in the workspace, the outputs are read with tts_train_outputs().
 * tts_train_backward: back-propagates caller-supplied output gradients d_before [B][T][80], d_after [B][T][80], d_stop [B][T]
 *   (fp32, device memory: dLoss/d(mel_before), dLoss/d(mel_after), dLoss/d(stop_logits) of ANY loss) through that forward into the
 *   flat gradient buffer (tts_train_grads).  Same workspace, shape, seed, utt_offset and p_residual as the forward it follows. */
int tts_train_forward(TtsHandle* h, void* workspace, const int64_t* phonemes, const int32_t* phoneme_lens, const float* mels,
                      const int32_t* mel_lens, int B, int S, int T, uint64_t seed, int utt_offset, double p_residual, void* stream);
int tts_train_backward(TtsHandle* h, void* workspace, int B, int S, int T, int utt_offset, double p_residual,
                       const float* d_before, const float* d_after, const float* d_stop, void* stream);

/* ---- per-kernel entry points (tests/test_gpu_kernels.py; not part of the drop-in surface) ---- */
/* C[M][N] = act(A[M][K] . W[N][K]^T + bias) ; bf16 in, fp32 out; act 0 none / 1 relu / 2 tanh (tcgen05 kernel, gemm_tc.cuh). */
int tts_k_gemm(const void* A_bf16, const void* W_bf16, const float* bias, float* C, int M, int N, int K, int act, void* stream);
/* conv1d(k=5, pad=2) over [B][T][Cin] with W [5][Cout][Cin] bf16 -> fp32 [B][T][Cout]; rows t >= lens[b] zeroed. */
int tts_k_conv5(const void* X_bf16, const void* W_bf16, const float* bias, const int32_t* lens, float* Y,
                int B, int T, int Cin, int Cout, int act, void* stream);
/* Attention core: Q [B][Lq][H*64], K/V [B][Lk][H*64] bf16 -> O [B][Lq][H*64] bf16. */
int tts_k_attention(const void* Q, const void* K, const void* V, void* O, const int32_t* klens,
                    int B, int H, int Lq, int Lk, int causal, void* stream);
/* Attention core with the per-row log-sum-exp kept for the backward pass: lse [B][H][Lq] fp32 (log2 domain). */
int tts_k_attention_lse(const void* Q, const void* K, const void* V, void* O, float* lse, const int32_t* klens,
                        int B, int H, int Lq, int Lk, int causal, void* stream);
/* Attention backward (training path): Q/K/V/O/dO as above + lse -> dQ/dK/dV bf16 in the same layouts.
 * scratch: fp32 [B*Lq*H*64 + B*H*Lq] device buffer. */
int tts_k_attention_bwd(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse,
                        const int32_t* klens, void* dQ, void* dK, void* dV, float* scratch,
                        int B, int H, int Lq, int Lk, int causal, void* stream);
/* LayerNorm over rows of 512: fp32 in -> bf16 out. */
int tts_k_layernorm(const float* X, const float* gamma, const float* beta, void* Y_bf16, int M, float eps, void* stream);
/* Keep-bits of a p = 0.5 dropout site: out[t][b][c] (uint8) for t < T, b < B, c < C. */
int tts_k_philox_bits(uint64_t seed, int site, int T, int B, int C, int utt_offset, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTS_B200_H_ */
