// Launcher declarations for every CUDA kernel of the streaming decode path.
// All launchers return 0 on success, -1 on failure (message via scb::get_last_error()).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "x3_split.cuh"

namespace scb {

// ---------------------------------------------------------------- descriptors
// One entry per stream that receives samples in this push (frontend plan).
struct FrontendDesc {
  int stream;      // stream slot
  int n_prev;      // samples already buffered
  int n_new;       // samples in this chunk
  int slab;        // samples framed this call (>= n_prev + n_new only when zero padded)
  int n_frames;    // STFT frames of the slab (1 + slab / 160)
  int emit0;       // first emitted frame
  int emit1;       // one past the last emitted frame
  int feat_off;    // row in featbuf[stream] where the first emitted frame goes
  int new_buf;     // samples kept for the next call (0 when final)
  int frame_base;  // prefix sum of n_frames over descriptors (for flat frame indexing)
};

// One entry per encoder block formed in this push.
struct BlockDesc {
  int stream;
  int sub_start;     // first sub-sampled frame of the block inside subbuf[stream]
  int clen;          // frames present (40 except the last block of a final call)
  int pe_frame_off;  // positional offset of the block's first frame
  int pe_ctx_off;    // positional offset of its context vector
  int prev_blk;      // previous block of the same stream in this push, or -1
  int has_prev_addin;  // prev_blk < 0: 1 -> use the stream's stored prev_addin, 0 -> own addin
  int is_last;       // last block of this stream in this push
  int out_slot0;     // first slot emitted to the encoder buffer
  int out_count;     // number of slots emitted
  int out_t0;        // destination frame index in encbuf[stream]
  int n_rows;        // rows that take part in attention: 42, or T for the short path
  int short_path;    // 1: no mask, no context slots
  int has_past_ctx;  // stream carries past_encoder_ctx from an earlier call
  int out_row0;      // dense index (over all streams of the push) of the block's first emitted frame
  int pad_;
};

// ---------------------------------------------------------------- elementwise
// y = LayerNorm(x) (eps 1e-12) over rows of width D (D % 32 == 0, D <= 512).
int launch_layernorm(const float* x, int ldx, const float* w, const float* b, float* y, int ldy,
                     int rows, int D, const int* n_rows_dev, cudaStream_t st);
int launch_layernorm_bf16(const float* x, int ldx, const float* w, const float* b, __nv_bfloat16* y,
                          int ldy, int rows, int D, const int* n_rows_dev, cudaStream_t st);

// ---------------------------------------------------------------- GEMM (fp32 SIMT)
struct GemmArgs {
  const float* A = nullptr;       // [M][lda] unless a_row_off is given
  int lda = 0;
  const int64_t* a_row_off = nullptr;  // element offset of row m (gather-A)
  const int* a_seg_off = nullptr;      // element offset added per K segment (implicit-GEMM conv)
  int seg_len = 0;                     // K segment length (multiple of 16) when a_seg_off != null
  const float* W = nullptr;       // [N][K] row-major (torch Linear layout)
  const float* bias = nullptr;    // [N] or null
  const float* R = nullptr;       // residual [M][ldr] or null (may alias C)
  int ldr = 0;
  float* C = nullptr;             // [M][ldc] unless c_row_off is given
  int ldc = 0;
  const int64_t* c_row_off = nullptr;  // element offset of output row m
  int M = 0, N = 0, K = 0;
  int relu = 0;
  const int* n_rows_dev = nullptr;     // optional dynamic M (device scalar)
  __nv_bfloat16* Cb = nullptr;         // optional bf16 copy of the output, dense [M][ldcb]
  int ldcb = 0;
};
int launch_gemm_f32(const GemmArgs& g, cudaStream_t st);

// ---------------------------------------------------------------- GEMM (fp32 results on the tensor cores, split fp16)
// Extra operands of launch_gemm_x3 (kernels_gemm_x3.cu, x3_split.cuh): A and / or C as split fp16 planes.
struct X3Extra {
  const void* A2 = nullptr;   // A as planes: hi [a2_rows][K] (leading dimension GemmArgs::lda), lo at + a2_plane elements
  size_t a2_plane = 0;
  int a2_rows = 0;            // row capacity of the A buffer (>= M)
  void* C2 = nullptr;         // output as planes, dense rows of leading dimension ldc2; lo plane at + c2_plane elements
  size_t c2_plane = 0;
  int ldc2 = 0;
  const float* lnX = nullptr; // persistent kernel only: A = LayerNorm(lnX rows [M][256], leading dimension ldx), eps 1e-12,
  int ldx = 0;                // computed inside the GEMM (replaces launch_layernorm_split + the plane round trip)
  const float* ln_w = nullptr;
  const float* ln_b = nullptr;
  int kernel = 0;             // 0: pick (persistent kernels where eligible), 1: per-tile kernel, 2: persistent kernel or error,
                              // 3: persistent kernel with both operands in shared memory (never the A-in-TMEM form)
};
int launch_gemm_x3(const GemmArgs& g, const X3Extra& x, const void* W2, cudaStream_t st);
// persistent A-resident / chunk-accumulating form for encoder-sized products (kernels_gemm_x3p.cu)
bool gemm_x3p_eligible(const GemmArgs& g, const X3Extra& x);
int launch_gemm_x3p(const GemmArgs& g, const X3Extra& x, const void* W2, cudaStream_t st);
// persistent form with the A operand in tensor memory (K = 256, no residual; kernels_gemm_x3t.cu)
bool gemm_x3t_eligible(const GemmArgs& g, const X3Extra& x);
int launch_gemm_x3t(const GemmArgs& g, const X3Extra& x, const void* W2, cudaStream_t st);
// ---------------------------------------------------------------- row-local chains of the decode step (kernels_chain_x3.cu)
struct ChainStage {
  int type;                    // 0: GEMM stage, 1: LayerNorm stage (optionally folding split-K partial sums into x first)
  int n_ct, n_ks, kb_item;     // GEMM: column tiles of 128, K splits, K blocks of 64 per item (even)
  int map_a, map_w, map_o;     // indices into ChainParams::maps; hi plane, lo = + 1 (map_o: fp32 map, or planes hi / lo)
  int out_mode, relu;          // 0: fp32 store, 1: fp32 add into the output (residual in place), 2: split planes,
                               // 3: fp32 partial product of K split ks (row block ks * part_stride_rows)
  int wait_base, wait_stride, wait_ks, wait_target;   // dependency: ctr[base + rt * stride (+ ks)] >= target; base < 0: none
  int sig_base, sig_stride, sig_div;                  // completion: ctr[base + rt * stride + ct / div] += 1 per epilogue warp
  int n_part;                  // LayerNorm: number of partial products to fold into x (0: none)
  const float* bias;           // GEMM bias [N] or null
  float* x;                    // LayerNorm: residual stream rows [M][256]
  const float* part;           // LayerNorm: partial products [n_part][part_stride_rows][256]
  const float* pbias;          // LayerNorm: bias added with the partials
  const float* ln_w; const float* ln_b;
  __half* out_hi; size_t out_plane;                   // LayerNorm: split planes [M][256], lo plane out_plane elements further
};
struct ChainParams {
  ChainStage st[8];
  int n_stages;
  const CUtensorMap* maps;     // device array of tensor maps (64-byte aligned)
  int* ctr;                    // completion counters of THIS launch, zero on entry
  const int* n_rows_dev;       // device-side count of active rows (<= M)
  int M;                       // row capacity
  int part_stride_rows;        // rows between the partial products of consecutive K splits (multiple of 128)
};
int launch_chain_x3(const ChainParams& p, int max_items, cudaStream_t st);
// cached tensor maps (kernels_gemm_tc.cu): 16-bit row-major matrix, 64-column x box_rows box; fp32 matrix, 32-column box
int tc_get_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
int tc_get_map_f32(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out);

// LayerNorm (eps 1e-12, same arithmetic as launch_layernorm) whose result is written as split fp16 planes
int launch_layernorm_split(const float* x, int ldx, const float* w, const float* b, void* y2, size_t plane, int ldy,
                           int rows, int D, const int* n_rows_dev, cudaStream_t st);

// ---------------------------------------------------------------- frontend
struct FrontendTables {        // device-resident constants of the frontend, carved from the engine workspace
  float window[512];
  float2 twiddle[256];
  int2 melrange[80];
  float melfb[257 * 80];
};
int frontend_upload_tables(FrontendTables* dev, const float* window400, const float* mel_fb_257x80);
int launch_frontend(const FrontendTables* tab, const float* wave_in, int ld_wave, const float* wbuf, int ld_wbuf,
                    const FrontendDesc* desc, int n_desc, int total_frames, const double* mean, const double* std_,
                    float* featbuf, int feat_cap, cudaStream_t st);
int launch_wavebuf_update(const float* wave_in, int ld_wave, float* wbuf, int ld_wbuf,
                          const FrontendDesc* desc, int n_desc, cudaStream_t st);

// ---------------------------------------------------------------- encoder
struct SubDesc {       // one per stream that runs conv2d subsampling in this push
  int stream;
  int t_in;            // feature frames convolved (from featbuf[stream][0 : t_in])
  int t1;              // conv1 output frames
  int t2;              // conv2 output frames (= new sub-sampled frames)
  int row0;            // prefix sum of t2 (dense row of the first new frame)
  int sub_off;         // destination row in subbuf[stream]
  int carry_src;       // featbuf row where the carried frames start
  int carry_n;         // frames carried to the next call
};
int launch_conv1(const float* featbuf, int feat_cap, const float* w1, const float* b1, float* h1, int t1_cap,
                 const SubDesc* desc, int n_desc, int D, cudaStream_t st);
int launch_conv1_im2col_bf16(const float* featbuf, int feat_cap, const float* w1, const float* b1, __nv_bfloat16* A16,
                             int t2_cap, const SubDesc* desc, int n_desc, int D, cudaStream_t st);
int launch_conv2_rows(const SubDesc* desc, int n_desc, int t1_cap, int sub_cap, int D, int64_t* a_row_off,
                      int64_t* c_row_off, cudaStream_t st);
// move rows [src, src+n) of buf[stream] (row width `width`, `cap` rows per stream) to the front
int launch_carry_rows(float* buf, int cap, int width, const int* stream, const int* src, const int* n,
                      int n_desc, cudaStream_t st);
int launch_block_assemble(const float* subbuf, int sub_cap, const float* pe, const BlockDesc* blk, int n_blk,
                          float* addin, float* prev_addin, float* X, int D, cudaStream_t st);
// out (fp32 rows) or, when so.base is set, split fp16 planes (precise tensor-core mode)
int launch_enc_attention(const float* qkv, float* out, __nv_bfloat16* out16, const BlockDesc* blk, int n_blk,
                         int n_head, int d_model, cudaStream_t st, SplitOut so = SplitOut());
int launch_enc_attention_mma(const __nv_bfloat16* qkv16, float* out, __nv_bfloat16* out16, const BlockDesc* blk, int n_blk,
                             int n_head, int d_model, cudaStream_t st);
// precise mode: block attention on tensor cores with split-fp16 operands (kernels_attn_x3.cu); fp32 QKV in
int launch_enc_attention_x3(const float* qkv, float* out, const BlockDesc* blk, int n_blk, int n_head, int d_model,
                            SplitOut so, cudaStream_t st);
int launch_ctx_handover(float* X, float* enc_ctx, int layer, int n_layers, const BlockDesc* blk, int n_blk,
                        int D, const float* ln_w, const float* ln_b, __nv_bfloat16* nrm16, cudaStream_t st);
int launch_stitch_norm(const float* X, const BlockDesc* blk, int n_blk, const float* w, const float* b,
                       float* encbuf, int t_cap, int D, __nv_bfloat16* dense16, cudaStream_t st);

// ---------------------------------------------------------------- decoder / search
struct StreamCtl {     // per-stream search state that lives on the device
  int cur;             // beam buffer (0/1) holding the current hypotheses
  int n_hyp;           // hypotheses in the current beam (1 before the first step, then B)
  int len;             // tokens per hypothesis (all equal: block-synchronous)
  int process_idx;     // global step counter (beam_search.py:326)
  int has_ctc;         // hypotheses carry a CTC forward state
  int ctc_T;           // rows of that state
  int active;          // still iterating in the current block
  int Tb;              // memory length of the current block
  int is_final;        // current block is the final one
  int iters_done;      // iterations completed in the current block
  int blk_next;        // next entry of the per-push decode-block queue (SearchBuffers::blkq_*)
  int blk_count;       // entries in the queue
  int outcome;         // decision of the current iteration (set by beam_prune, applied by step_commit)
  int steps_total;     // statistics: iterations executed for this stream
  int pad_[2];
};

struct SearchBuffers {
  // configuration
  int S, B, V, D, H, Ld, Tcap, Lcap, use_bbd;
  float w_dec, w_ctc;
  // model memory
  const float* ctcx;      // [S][Tcap][V]
  float* xkv;             // [Ld][S][Tcap][2D]  cross-attention K|V (bf16 elements when kv_bf16)
  int kv_bf16;
  int kv_split;           // K|V caches hold split fp16 planes per row: [hi K|V][lo K|V] (kernels_attn_x3.cu)
  int attn_head_major;    // x3 decoder attention: grid (head, stream) instead of (stream, head)
  float* skv;             // [Ld][S][Lcap][B][2D] self-attention K|V (tree storage)
  // beam (ping-pong)
  int* yseq;              // [2][S][B][Lcap]
  int* xpos;              // [2][S][B][Lcap]
  unsigned char* anc;     // [2][S][B][Lcap] slot of the ancestor at each position
  double* score;          // [2][S][B]
  double* sc_dec;         // [2][S][B]
  double* sc_ctc;         // [2][S][B]
  float* ctc_r;           // [2][S][B][Tcap][2]
  float* ctc_s;           // [2][S][B]
  StreamCtl* ctl;         // [S]
  int* blkq_T;            // [S][qcap] memory length of each queued decode block
  int* blkq_final;        // [S][qcap] 1 for the final block
  int qcap;
  // per-step scratch (compact rows)
  int* row_sh;            // [S*B] compact row -> s*B+h
  int* row_base;          // [S] first compact row of the stream, -1 if inactive
  int* act_streams;       // [S] compact list of active streams
  int* n_rows;            // device scalar: active rows
  int* n_active;          // device scalar: active streams
  float* logp;            // [S*B][V] decoder log-probs
  int* pre_ids;           // [S*B][40]
  float* psi;             // [S*B][40]
  float* psi_eos;         // [S*B]
  float* cand_val;        // [S*B][B]
  int* cand_tok;          // [S*B][B]
  float* cand_dec;        // [S*B][B] unweighted decoder score of the candidate
  float* cand_ctc;        // [S*B][B] unweighted ctc score of the candidate
  float* cand_psi;        // [S*B][B] log_psi of the candidate (new CTC prefix score)
  int* cand_col;          // [S*B][B] token whose forward column is inherited (select_state)
  int* new_parent;        // [S][B] parent slot of each new hypothesis (this step)
  int* new_col;           // [S][B]
  int* upd_flag;          // [S] 1 -> the CTC state of the new beam must be written this step
  int* self_keys;         // [S][key_cap] per-step key list of the self-attention KV tree (see build_self_keys)
  int* self_nkeys;        // [S]
  int key_cap;
  unsigned long long* prof;   // [8] device counters: 0 ctc algorithmic bytes, 1 sum of active rows,
                              //     2 cross-attention KV bytes, 3 self-attention KV bytes, 4 search iterations
};

int launch_search_reset(const SearchBuffers& sb, const int* streams, int n, cudaStream_t st);
// q_T / q_final: [n_q][q_stride] entries of this push, appended to each stream's pending queue
int launch_search_begin(const SearchBuffers& sb, const int* q_stream, const int* q_n, const int* q_T,
                        const int* q_final, int q_stride, int n_q, cudaStream_t st);
int launch_dec_embed(const SearchBuffers& sb, const float* emb, const float* pe, float* x, const float* ln_w,
                     const float* ln_b, __nv_bfloat16* out16, cudaStream_t st);
int launch_dec_self_attention(const SearchBuffers& sb, int layer, const float* qkv, int ldq, float* out,
                              __nv_bfloat16* out16, cudaStream_t st, SplitOut so = SplitOut());
// Key list of the self-attention KV tree, built once per search iteration and shared by all layers and heads:
// key u < Lc = position u of the beam's common ancestor chain; later keys = (hypothesis, position) pairs of the
// divergent tail.  Packed as position | slot << 16 | (owner + 1) << 24 (owner 0 = visible to every hypothesis).
int launch_build_self_keys(const SearchBuffers& sb, cudaStream_t st);
int launch_dec_cross_attention(const SearchBuffers& sb, int layer, const float* q, int ldq, float* out,
                               __nv_bfloat16* out16, cudaStream_t st, SplitOut so = SplitOut());
// fp32 K|V caches, warp-per-head pipelines (kernels_attn_f32.cu).  mode 0: self-attention over the per-step key list
// (launch_build_self_keys must have run in this search iteration; also appends this step's K|V), 1: cross-attention.
int launch_dec_attention_f32(const SearchBuffers& sb, int mode, int layer, const float* q, int ldq, float* out,
                             SplitOut so, cudaStream_t st);
// split-plane K|V caches, tensor cores with fp32-class accuracy (kernels_attn_x3.cu); beam <= 16
int launch_dec_attention_x3(const SearchBuffers& sb, int mode, int layer, const float* q, int ldq, float* out,
                            SplitOut so, cudaStream_t st);
int launch_logsoftmax_prebeam(const SearchBuffers& sb, float* logits, cudaStream_t st);
int launch_ctc_prefix(const SearchBuffers& sb, cudaStream_t st);
int launch_ctc_prefix_op(const float* x, int T, int V, const float* r_prev, const int* last_tok, int L, const int* ids,
                         int n_hyp, float* psi, float* psi_eos, float* r_new, cudaStream_t st);
int launch_combine_topk(const SearchBuffers& sb, const float* logp, cudaStream_t st);
int launch_beam_prune(const SearchBuffers& sb, cudaStream_t st);
int launch_ctc_state_update(const SearchBuffers& sb, cudaStream_t st);
int launch_step_finish(const SearchBuffers& sb, cudaStream_t st);
// in-place log-softmax of the rows x + row_off[r] whose row_flag[r] != 0
int launch_logsoftmax_rows(float* x, const int64_t* row_off, const int* row_flag, int rows, int V, cudaStream_t st);

}  // namespace scb
