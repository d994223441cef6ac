/* speechcatcher_b200 -- C ABI of the B200-native multi-stream streaming decode path.
 *
 * Drop-in boundary for speechcatcher's `Speech2TextStreaming` native decoder.  The reference is pure
 * Python (no FFI of its own), so every entry point below cites the Python interface it replaces
 * (paths relative to the reference checkout).  A maintainer binds these with ctypes; the binding the
 * repo ships is speechcatcher_b200/_lib.py and the reference-side stub is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only, no torch types; `stream` is a cudaStream_t passed as void*
 *   - the caller owns every device buffer (including the engine workspace); the library allocates no
 *     device memory (every table, including the frontend's window / mel matrix, lives in the engine workspace) and keeps
 *     no global mutable state besides a cache of TMA descriptors and per-device "attribute set" flags
 *   - every function returns 0 on success and a negative code on failure; sc_last_error() explains
 *   - functions never throw across the boundary; one driver thread per engine handle
 */
#ifndef SPEECHCATCHER_B200_H
#define SPEECHCATCHER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_ERR_CUDA (-1)
#define SC_ERR_ARG (-2)
#define SC_ERR_CAPACITY (-3)
#define SC_ERR_STATE (-4)

/* Architecture + capacity of one engine (one GPU, many streams).
 * Mirrors what Speech2TextStreaming.__init__/_load_model read from config.yaml
 * (speech2text_streaming.py:43-51, 210-232) plus batching capacities. */
typedef struct ScConfig {
  int32_t d_model;       /* encoder_conf.output_size (256) */
  int32_t enc_heads;     /* encoder_conf.attention_heads */
  int32_t enc_layers;    /* encoder_conf.num_blocks */
  int32_t dec_heads;     /* decoder_conf.attention_heads */
  int32_t dec_layers;    /* decoder_conf.num_blocks */
  int32_t vocab;         /* rows of decoder.embed.0.weight (1024) */
  int32_t ffn;           /* linear_units (2048; not configurable in the reference) */
  int32_t n_streams;     /* concurrent streams S held by this engine */
  int32_t beam;          /* beam_size (<= 20) */
  int32_t max_chunk;     /* largest number of samples one stream may push per call */
  int32_t max_frames;    /* capacity of the per-stream encoder buffer (frames of 40 ms) */
  int32_t use_bbd;       /* Speech2TextStreaming(use_bbd=...) */
  int32_t precision;     /* 0 = fp32 on the CUDA cores, 1 = bf16 tensor-core GEMMs, 2 = fp32 results on the tensor
                          * cores: every Linear as a split-fp16 (hi + lo 2^-11) tcgen05 GEMM, everything else as 0 */
  float ctc_weight;      /* 0.3; decoder weight is 1 - ctc_weight (speech2text_streaming.py:143-150) */
} ScConfig;

typedef struct ScPushStats {
  int32_t n_feature_frames;   /* feature frames emitted by the frontend over all streams */
  int32_t n_encoder_blocks;   /* encoder blocks processed */
  int32_t n_encoder_frames;   /* encoder output frames appended */
  int32_t n_decode_steps;     /* search iterations launched (max over streams, incl. the probe step) */
  int32_t n_kernel_launches;  /* CUDA kernels launched by this call */
  int32_t reserved[3];
} ScPushStats;

typedef struct ScStreamPlan {     /* host-side shape plan of one stream for one push (debug / tests) */
  int32_t called;             /* 1 if process_block would be called (frontend emitted features) */
  int32_t n_feat;             /* emitted feature frames */
  int32_t n_sub;              /* new sub-sampled frames */
  int32_t n_blocks;           /* encoder blocks formed */
  int32_t n_enc_out;          /* encoder output frames appended */
  int32_t enc_len;            /* encoder buffer length after the push */
  int32_t n_decode_blocks;    /* decode blocks queued */
  int32_t last_T;             /* memory length of the last queued decode block (0 if none) */
} ScStreamPlan;

const char* sc_version(void);
const char* sc_last_error(void);

/* ---- engine (replaces Speech2TextStreaming + BlockwiseSynchronousBeamSearch state, batched) ---- */
/* Bytes of caller-owned device workspace the engine needs for `cfg`. */
int sc_engine_workspace_bytes(const ScConfig* cfg, size_t* bytes);
/* Create an engine over `workspace` (device memory, 256-byte aligned, zero-initialised by the caller). */
int sc_engine_create(const ScConfig* cfg, void* workspace, size_t bytes, void** handle);
int sc_engine_destroy(void* handle);
/* Register one packed weight tensor by name (device pointer, caller keeps it alive).
 * Names/layouts: see speechcatcher_b200/weights.py (built from the reference state_dict keys). */
int sc_engine_set_weight(void* handle, const char* name, const void* dev_ptr, size_t n_elem);
/* Host tables: hann(400), mel_fb[257*80] (model.frontend buffers, stft_frontend.py:68-85) and the
 * global MVN statistics mean/std[80] in fp64 (checkpoint_loader.py:210-237); mean may be NULL. */
int sc_engine_set_frontend(void* handle, const float* window400, const float* mel_fb, const double* mean,
                           const double* std_);
int sc_engine_finalize(void* handle);
/* Speech2TextStreaming.reset() for the listed streams (speech2text_streaming.py:252-263). */
int sc_engine_reset(void* handle, const int32_t* streams, int32_t n, void* stream);
/* One Speech2TextStreaming.__call__ per listed stream (speech2text_streaming.py:402-464):
 * wave_dev[i*ld_wave ...] holds n_samples[i] new samples of stream streams[i]. */
int sc_engine_push(void* handle, const float* wave_dev, int32_t ld_wave, const int32_t* streams,
                   const int32_t* n_samples, const int32_t* is_final, int32_t n, void* stream,
                   ScPushStats* stats);
/* Same for pre-computed feature frames (the 2-D / 3-D input of Speech2TextStreaming.__call__,
 * speech2text_streaming.py:438-450): feats_dev[i*ld_feats ...] holds n_frames[i] rows of 80 floats for stream
 * streams[i], already normalised by the caller ((x - mean) / std for 2-D input, untouched for 3-D input, like the
 * reference); the frontend is skipped and process_block runs for every listed stream.  At most
 * (max_chunk + 400) / 160 + 2 frames per stream and push (what the largest waveform chunk would produce). */
int sc_engine_push_features(void* handle, const float* feats_dev, int32_t ld_feats, const int32_t* streams,
                            const int32_t* n_frames, const int32_t* is_final, int32_t n, void* stream,
                            ScPushStats* stats);
/* Current beam of one stream (beam_state.hypotheses): copies to host buffers and synchronises.
 * yseq/xpos: [beam][max_len] int32, score: [beam] fp64. */
int sc_engine_read_beam(void* handle, int32_t stream_id, int32_t max_len, int32_t* n_hyp, int32_t* len,
                        int32_t* process_idx, int32_t* yseq, int32_t* xpos, double* score, void* stream);
/* Beams of ALL streams in four bulk copies (results of a whole batch): ctl16 [S][16] int32 control words
 * (0 = current ping-pong buffer, 1 = n_hyp, 2 = len, 3 = process_idx), yseq / xpos [2][S][beam][token_capacity] int32,
 * score [2][S][beam] fp64.  Host buffers should be pinned.  Synchronises the stream. */
int sc_engine_read_all(void* handle, int32_t* ctl16, int32_t* yseq, int32_t* xpos, double* score, void* stream);
int sc_engine_token_capacity(void* handle, int32_t* token_capacity);
/* Host plan of the last push for stream `stream_id`. */
int sc_engine_last_plan(void* handle, int32_t stream_id, ScStreamPlan* plan);
/* Named internal device buffer (tests / debugging): pointer, element count and row pitch. */
int sc_engine_buffer(void* handle, const char* name, void** ptr, size_t* n_elem);

/* Engine options.  "lazy_threshold" = n > 0: a push stops iterating the beam search once fewer than n streams
 * are still active and leaves the stragglers' decode blocks queued on the device (they continue during later
 * pushes; any final call drains everything), which trades per-call completeness of non-final beams for fewer,
 * fuller search iterations.  0 (default) = strict: every push fully decodes its blocks like the reference.
 * "overlap" = 0/1 (default 1): in deferred mode run a push's frontend/encoder on a second CUDA stream while the
 * caller's stream keeps iterating the search for blocks queued by earlier pushes.
 * "encoder_sms" = n (default 0 = off): run the frontend + encoder part of every push in a green context that owns ~n SMs
 * (CUDA partition granularity: 8), leaving the other SMs to the search chain; counter "encoder_sms" tells what was
 * provisioned (0 when the driver refused and the plain stream is used).
 * "mma_attention" = 0/1/2: CUDA-core or tensor-core attention in the bf16 mode (1: beam <= 16; 2: also beam 17..32 with
 * the hypotheses tiled over two m16 tiles -- not yet run on a device).  "pdl" = 0/1: programmatic dependent
 * launch of the decode-step kernel chain (process-wide).  "fuse_layernorm" = 0/1: experimental LN-in-epilogue GEMMs.
 * "ln_prologue" / "ln_prologue_decoder" = 0/1 (default 0): compute LayerNorm inside the consuming GEMM (everywhere /
 * decode step only; the latter has not run on a device yet).
 * "graph_decode" / "graph_encoder" = 0/1 (default 0, experimental): replay one search iteration / the encoder stack of
 * a push as an instantiated CUDA graph instead of ~165 / ~7-per-layer individual launches (their arguments are
 * iteration-invariant); needs a non-default CUDA stream, falls back to plain launches when capture is refused and
 * while a kernel is being profiled. */
int sc_engine_set_option(void* handle, const char* name, int32_t value);

/* Step trace (direct parity tests of the decoder log-probs / CTC prefix scores / pre-beam against the reference's
 * batch_score_hypotheses, beam_search.py:71-185): the next `max_steps` search iterations each copy one fixed-size
 * record into the caller-owned device buffer, taken after score combination and before pruning.  Record fields
 * (byte offsets in offsets8): 0 n_rows int32[1], 1 row_sh int32[R] (row -> stream * beam + hyp), 2 logp float[R][V]
 * (decoder log-softmax), 3 pre_ids int32[R][40], 4 psi float[R][40] (CTC log_psi of the 40 candidates), 5 psi_eos
 * float[R], 6 ctc_s float[2][S][beam] (s_prev of both beam buffers), 7 ctl int32[S][16]; R = S * beam.
 * dev_buf = NULL switches the trace off.  Graph replay is suspended while a trace is being recorded. */
int sc_engine_trace_layout(void* handle, int64_t* offsets8, int64_t* record_bytes);
int sc_engine_set_trace(void* handle, void* dev_buf, size_t bytes, int32_t max_steps);
/* Named host-side counters: "graphs_replayed" (cudaGraphLaunch calls so far), "graph_failed", "trace_steps". */
int sc_engine_counter(void* handle, const char* name, int64_t* value);

/* Live kernel timing with CUDA-event pairs on the launching stream (bench.py roofline and step breakdown).
 * tag > 0: every launch of that kernel; tag = -1: every kernel, decode steps sampled every `stride` steps.
 * Tags: 1 ctc_prefix, 2 dec_self_attn, 3 dec_cross_attn, 4 dec_ffn1, 5 enc_ffn1, 6 prebeam, 7 enc_attn, 8 conv2,
 * 9 dec_ffn2, 10 enc_ffn2, 11 ctc_state_update, 12 frontend, 13 conv1, 14 sub_out, 15 block_assemble, 16 enc_ln,
 * 17 enc_qkv, 18 enc_o, 19 enc_handover, 20 stitch_norm, 21 ctc_head, 22 cross_kv, 23 dec_embed, 24 dec_ln,
 * 25 dec_qkv, 26 dec_self_o, 27 dec_cross_q, 28 dec_cross_o, 29 dec_out, 30 combine_topk, 31 beam_prune,
 * 32 step_finish, 33 decode_step_total, 34 encoder_total.  counters8: device-side algorithmic counters
 * (0 ctc bytes, 1 active rows, 2 cross-KV bytes, 3 self-KV bytes, 4 search iterations). */
int sc_engine_profile_begin(void* handle, int32_t tag, int32_t max_launches, int32_t stride);
int sc_engine_profile_end(void* handle, int32_t n_tags, int32_t* launches_per_tag, double* ms_per_tag,
                          double* host_flops, uint64_t* counters8);

/* ---- host-only shape planner (no CUDA calls; CPU-testable) ---- */
int sc_planner_create(int32_t n_streams, void** planner);
int sc_planner_destroy(void* planner);
int sc_planner_reset(void* planner, int32_t stream_id);
int sc_planner_push(void* planner, int32_t stream_id, int32_t n_samples, int32_t is_final, ScStreamPlan* plan);
int sc_planner_push_features(void* planner, int32_t stream_id, int32_t n_frames, int32_t is_final, ScStreamPlan* plan);

/* ---- offline segmentation of long files (replaces speechcatcher/simple_endpointing.py) ---- */
/* Parameters of the cut-point search = the reference's BeamSearch constructor (simple_endpointing.py:23-34);
 * segment_speech() fills them as beam 10, ideal 100 * average_segment_length, look-ahead 18000, min 2000, step 10,
 * len_reward_weight 12, energy_weight 1 (:124-130).  Lengths are in 10 ms frames. */
typedef struct ScSegmentParams {
  int32_t beam_size;
  int32_t ideal_segment_len;
  int32_t max_lookahead;
  int32_t min_len;
  int32_t step;
  int32_t reserved;
  double len_reward_weight;
  double energy_weight;
} ScSegmentParams;
/* Frames python_speech_features.logfbank(winlen=0.025, winstep=0.01) yields for n_samples at 16 kHz (host only). */
int sc_segment_num_frames(int64_t n_samples, int64_t* n_frames);
/* The 28 FFT-bin edges of the 26 mel triangles (python_speech_features get_filterbanks; host only, for tests). */
int sc_segment_filterbank_bins(double* bins28);
/* simple_endpointing.py:73-75 on the device, fp64: energy_dev[t] = sum_j log(fbank_j(frame t)) / 10 and
 * smoothed_dev = -gaussian_filter1d(energy, sigma) (scipy defaults: reflect, radius int(4 sigma + .5)).
 * pcm_dev: n_samples int16 samples (the raw file, 16 kHz mono); both outputs hold n_frames doubles. */
int sc_segment_energy(const int16_t* pcm_dev, int64_t n_samples, double* energy_dev, double* smoothed_dev,
                      int64_t n_frames, double sigma, void* stream);
/* BeamSearch.search (simple_endpointing.py:43-69) on the host over the smoothed curve (host memory): writes the
 * best cut list [0, c1, c2, ...] (segments are consecutive pairs); sequential, data dependent, microseconds. */
int sc_segment_search(const double* smoothed_host, int64_t n_frames, const ScSegmentParams* p, int64_t* cuts,
                      int32_t max_cuts, int32_t* n_cuts);

/* ---- single operators over raw device pointers (used by the parity tests) ---- */
/* K1, the frontend of one waveform slab (speech2text_streaming.py:350-358 + stft_frontend.py:87-154): centred STFT
 * (n_fft 512, hop 160, hann 400, reflect padding) -> power -> 257x80 mel -> log(clamp 1e-10) -> (x - mean) / std in fp64
 * (mean_dev may be NULL: no normalisation).  workspace_dev: sc_frontend_workspace_bytes() of caller-owned device memory,
 * initialised once with the host tables by sc_frontend_init.  feats_dev receives 1 + n_samples / 160 rows of 80 floats. */
size_t sc_frontend_workspace_bytes(void);
int sc_frontend_init(void* workspace_dev, const float* window400, const float* mel_fb);
int sc_frontend_fbank_mvn(void* workspace_dev, const float* wave_dev, int32_t n_samples, const double* mean_dev,
                          const double* std_dev, float* feats_dev, int32_t* n_frames, void* stream);
/* K7, one call of the batched CTC prefix scorer (ctc_prefix_score_full.py:88-291, CTCPrefixScoreTH.__call__) for n_hyp
 * hypotheses of equal length over explicit tensors: x_dev [t][v] emission store, r_prev_dev [n_hyp][t][2] forward
 * variables (non-blank, blank) of every hypothesis, last_tok_dev [n_hyp], prefix_len = len(yseq) - 1, cand_ids_dev
 * [n_hyp][40] pre-beam candidates.  Outputs: psi_dev [n_hyp][40] = log_psi of the candidates, psi_eos_dev [n_hyp] =
 * r_sum[t-1] (the <eos> entry), r_new_dev (may be NULL) [n_hyp][40][t][2] = the forward variables select_state hands on. */
int sc_ctc_prefix_step(const float* x_dev, int32_t t, int32_t v, const float* r_prev_dev, const int32_t* last_tok_dev,
                       int32_t prefix_len, const int32_t* cand_ids_dev, int32_t n_hyp, float* psi_dev, float* psi_eos_dev,
                       float* r_new_dev, void* stream);
/* LayerNorm eps=1e-12 (model/layers/normalization.py:23) */
int sc_layernorm_f32(const float* x, const float* w, const float* b, float* y, int32_t rows, int32_t d,
                     void* stream);
/* y = act(x W^T + bias) + residual  (torch.nn.functional.linear; fp32 CUDA-core path) */
int sc_linear_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                  int32_t m, int32_t n, int32_t k, int32_t relu, void* stream);
/* same contract (fp32 in, fp32 out, fp32-class accuracy) on the tcgen05 tensor cores: x is split into fp16 hi / lo
 * parts inside the kernel, w_planes_f16 = [2][n][k] fp16 (hi plane, then (w - hi) * 2^11) built once per weight
 * (speechcatcher_b200/weights.py:split_f16); three UMMAs per K step into two TMEM accumulators (precision 2) */
int sc_linear_x3(const float* x, const void* w_planes_f16, const float* bias, const float* residual, float* y,
                 int32_t m, int32_t n, int32_t k, int32_t relu, void* stream);
/* The engine's own form of that GEMM: the activation arrives as split fp16 planes written by its producer
 * (x_planes_f16: hi plane [x_rows][k], lo plane x_plane_elems further, x_rows = row capacity >= m) and is staged by
 * TMA like the weights; the result is written as fp32 rows (y, may be NULL) and / or as split planes for the next
 * Linear (y_planes_f16, may be NULL).  kernel: 0 = the engine's choice (the persistent A-resident kernel of
 * kernels_gemm_x3p.cu for dense products with n % 128 == 0, k = 256 or k % 128 == 0 >= 384, one output form and an
 * in-place residual; the per-tile kernel otherwise), 1 = per-tile kernel, 2 = persistent kernel (error if ineligible; k = 256 without a
 * residual takes the form with A in tensor memory, kernels_gemm_x3t.cu), 3 = persistent kernel, operands in shared memory. */
int sc_linear_x3_planes(const void* x_planes_f16, int64_t x_plane_elems, int32_t x_rows, const void* w_planes_f16,
                        const float* bias, const float* residual, float* y, void* y_planes_f16, int64_t y_plane_elems,
                        int32_t m, int32_t n, int32_t k, int32_t relu, int32_t kernel, void* stream);
/* y = act(LayerNorm(x) W^T + bias) for fp32 rows x [m][256]: the LayerNorm (eps 1e-12, the arithmetic of
 * sc_layernorm_split) is computed in the prologue of the persistent GEMM, which builds its resident A tile from it;
 * output as fp32 rows (y) or as split planes (y_planes_f16), exactly one of them.  n_rows_dev (may be NULL): device-side
 * count of active rows <= m (the decode step's compact row list); only row tiles holding active rows are computed.
 * Replaces normalization.py:23 + torch.nn.functional.linear (decoder_layer.py:80-132, contextual_block_encoder_layer.py). */
int sc_linear_x3_ln(const float* x, const float* ln_w, const float* ln_b, const void* w_planes_f16, const float* bias,
                    float* y, void* y_planes_f16, int64_t y_plane_elems, int32_t m, int32_t n, int32_t relu,
                    const int32_t* n_rows_dev, void* stream);
/* LayerNorm eps=1e-12 whose result is written as split fp16 planes (hi [rows][d], lo y_plane_elems further) */
int sc_layernorm_split(const float* x, const float* w, const float* b, void* y_planes_f16, int64_t y_plane_elems,
                       int32_t rows, int32_t d, void* stream);
/* same contract on the tcgen05 tensor-core path: bf16 operands, fp32 accumulate */
int sc_linear_bf16(const void* x_bf16, const void* w_bf16, const float* bias, const float* residual,
                   float* y_f32, void* y_bf16, int32_t m, int32_t n, int32_t k, int32_t relu, void* stream);

/* y = x W^T + bias + residual (N = 256) with the LayerNorm of every finished row fused into the epilogue:
 * ln_out = bf16(LayerNorm(y) * ln_w + ln_b), eps 1e-12 */
int sc_linear_bf16_ln(const void* x_bf16, const void* w_bf16, const float* bias, const float* residual, float* y_f32,
                      const float* ln_w, const float* ln_b, void* ln_out_bf16, int32_t m, int32_t k, void* stream);

/* y = act(LayerNorm(x) W^T + bias) with the LayerNorm of the fp32 rows x [m][256] computed inside the GEMM
 * (the normalised bf16 A tile is written directly into the swizzled shared-memory operand layout) */
int sc_linear_bf16_lnA(const float* x_f32, const float* ln_w, const float* ln_b, const void* w_bf16, const float* bias,
                       float* y_f32, void* y_bf16, int32_t m, int32_t n, int32_t relu, void* stream);

/* y[m][256] (+)= W2 ReLU(W1 x + b1) + b2: the encoder layer's position-wise feed-forward
 * (model/layers/feed_forward.py:41-50) as one tcgen05 kernel; x [m][256] bf16, w1 [f][256] bf16, w2 [256][f] bf16,
 * f a multiple of 128.  accumulate != 0: y already holds the residual and the result is added to it in place
 * (contextual_block_encoder_layer.py:243-251), else y is overwritten.  The hidden activation is rounded to bf16 after
 * bias + ReLU, as in the two-GEMM path.  splits > 1 (needs accumulate): the hidden dimension is divided over that
 * many CTAs per 128-row tile whose partial results are added at the L2 (small-m launches of the decode step). */
int sc_ffn_bf16(const void* x_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                float* y_f32, int32_t accumulate, int32_t m, int32_t f, int32_t splits, void* stream);

/* Same kernel with a device-side timeline: `stamps` (device, >= 1024 int64) receives clock64() stamps of the first
 * CTA's MMA-issue warp and of one epilogue thread per hidden chunk (performance diagnosis, scripts/ffn_timeline.py). */
int sc_ffn_bf16_timeline(const void* x_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                         float* y_f32, int32_t accumulate, int32_t m, int32_t f, int64_t* stamps, void* stream);

#ifdef __cplusplus
}
#endif
#endif
