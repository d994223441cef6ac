// Host engine + C ABI: one engine per GPU holds the device state of S independent streams and
// advances all of them through frontend -> encoder -> block-synchronous beam search per push.
//
// Replaces (batched, on device): speechcatcher/speech2text_streaming.py:402-464 (__call__),
//   speechcatcher/beam_search/beam_search.py:507-653 (process_block) and everything below them.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/speechcatcher_b200.h"
#include "kernels.h"
#include "planner.h"

namespace scb {

int launch_gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, const float* R,
                     int ldr, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, int M, int N, int K, int relu,
                     const int* n_rows_dev, cudaStream_t st);
int launch_dec_attention_mma(const SearchBuffers& sb, int mode, int layer, const float* q, int ldq, float* out,
                             __nv_bfloat16* out16, cudaStream_t st);
int launch_enc_attention_mma(const __nv_bfloat16* qkv16, float* out, __nv_bfloat16* out16, const BlockDesc* blk, int n_blk,
                             int n_head, int d_model, cudaStream_t st);
int launch_ffn_fused_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W1, const float* b1,
                          const __nv_bfloat16* W2, const float* b2, float* C, int ldc, int accumulate, int M, int F,
                          int splits, const int* n_rows_dev, cudaStream_t st, long long* dbg = nullptr);
int launch_gemm_bf16_lnA(const float* X, int ldx, const float* lna_w, const float* lna_b, const __nv_bfloat16* W,
                         const float* bias, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, int M, int N, int relu,
                         const int* n_rows_dev, cudaStream_t st);
int launch_gemm_bf16_ln(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, const float* R,
                        int ldr, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, const int64_t* c_row_off, int M, int N,
                        int K, int relu, const int* n_rows_dev, const float* ln_w, const float* ln_b,
                        __nv_bfloat16* ln_out, cudaStream_t st);

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Caps {
  int feat_cap, t1_cap, t2_cap, sub_cap, nb_cap, qpush, qcap, Tcap, Lcap, nb_max, rows_max, sub_rows_max, R;
};

static Caps make_caps(const ScConfig& c) {
  Caps k;
  const int fmax = 1 + (c.max_chunk + 400) / 160 + 1;   // STFT frames of the largest slab
  const int tmax = 12 + fmax;                            // + carried feature frames
  k.feat_cap = tmax + 4;
  k.t1_cap = (tmax - 3) / 2 + 1;
  k.t2_cap = std::max(1, (k.t1_cap - 3) / 2 + 1);
  k.sub_cap = 40 + k.t2_cap + 1;
  k.nb_cap = std::max(1, (k.sub_cap - 24 + 15) / 16);
  k.qpush = k.nb_cap + 2;                                // decode blocks one push can queue per stream
  k.qcap = 16 * k.qpush;                                 // device queue: room for deferred blocks of earlier pushes
  k.Tcap = c.max_frames;
  // hypotheses grow by one token per completed iteration (<= max_length 500) plus one for every block whose
  // first iteration already ends in <eos> (kept without a rewind, beam_search.py:827): bound by the block count
  k.Lcap = ((kMaxLength + 2 + c.max_frames / kHopB + 8) + 31) / 32 * 32;
  k.nb_max = c.n_streams * k.nb_cap;
  k.rows_max = k.nb_max * kSlots;
  k.sub_rows_max = c.n_streams * k.t2_cap;
  k.R = c.n_streams * c.beam;
  return k;
}

struct EncLayerW { const float *ln1w, *ln1b, *qkvw, *qkvb, *ow, *ob, *ln2w, *ln2b, *f1w, *f1b, *f2w, *f2b;
                   const __nv_bfloat16 *qkvw16, *ow16, *f1w16, *f2w16; };
struct DecLayerW { const float *ln1w, *ln1b, *sqkvw, *sqkvb, *sow, *sob, *ln2w, *ln2b, *cqw, *cqb, *ckvw, *ckvb,
                   *cow, *cob, *ln3w, *ln3b, *f1w, *f1b, *f2w, *f2b;
                   const __nv_bfloat16 *sqkvw16, *sow16, *cqw16, *ckvw16, *cow16, *f1w16, *f2w16; };

struct Engine {
  ScConfig cfg;
  Caps cap;
  Planner planner;
  std::unordered_map<std::string, const void*> wmap;
  // precision 2 (fp32 results on the tensor cores): fp32 weight pointer -> its two fp16 planes (kernels_gemm_x3.cu)
  std::unordered_map<const float*, const void*> x3map;
  bool finalized = false;
  // weights
  const float *pe = nullptr, *c1w = nullptr, *c1b = nullptr, *c2w = nullptr, *c2b = nullptr, *eow = nullptr, *eob = nullptr;
  const float *eaw = nullptr, *eab = nullptr, *ctcw = nullptr, *ctcb = nullptr, *demb = nullptr, *daw = nullptr, *dab = nullptr,
              *doutw = nullptr, *doutb = nullptr;
  const __nv_bfloat16 *c2w16 = nullptr, *eow16 = nullptr, *ctcw16 = nullptr, *doutw16 = nullptr;
  std::vector<EncLayerW> enc;
  std::vector<DecLayerW> dec;
  // frontend stats (device, inside workspace)
  double *d_mean = nullptr, *d_std = nullptr;
  bool has_stats = false;
  FrontendTables* d_fe_tab = nullptr;       // window / twiddles / mel matrix (workspace)
  bool fe_tab_ready = false;
  // device buffers
  float *wbuf, *featbuf, *h1, *h2, *subbuf, *prev_addin, *enc_ctx, *addin, *X, *Nrm, *QKV, *Att, *FF, *encbuf, *ctcx;
  float *dx, *dn, *dqkv, *dq, *dattn, *dffn, *dlogp;
  __nv_bfloat16 *Nrm16, *Att16, *FF16, *h2_16, *dn16, *dattn16, *dffn16, *encnew16, *im2col16, *QKV16;
  SearchBuffers sb;
  // device descriptor arrays
  FrontendDesc* d_fd; SubDesc* d_sd; BlockDesc* d_blk;
  int *d_carry_f, *d_carry_s;       // 3 x S each: stream, src, n
  int64_t *d_c2_a, *d_c2_c, *d_en_a, *d_en_ctc, *d_en_kv;
  int *d_en_flag, *d_c2_seg, *d_q, *d_reset;
  // pinned host staging
  unsigned char* h_stage = nullptr; size_t h_stage_bytes = 0;
  int* h_flag = nullptr;            // [2] n_active probe
  cudaEvent_t ev[2];
  std::unordered_map<std::string, std::pair<void*, size_t>> named;
  std::vector<ScStreamPlan> last_plan;
  int launches = 0;
  bool attn_f32_rows = true;        // fp32 modes: warp-per-head decoder attention (kernels_attn_f32.cu) instead of the
                                    // first CTA-per-(stream, head) kernels (SCB_ATTN=cta / option "attn_f32_rows")
  bool enc_attn_x3 = true;          // precise mode: encoder block attention on tensor cores (split fp16, kernels_attn_x3.cu)
  // precise mode: LayerNorm computed in the prologue of the persistent GEMM that consumes it (kernels_gemm_x3p.cu) instead
  // of a LayerNorm kernel writing planes (SCB_X3_LN_FUSED = "0" / "enc" / "dec" / "1"; options "x3_ln_fused_encoder/_decoder")
  bool x3_ln_fused_enc = false;     // measured slower: the eight epilogue warps build A, so every row-tile switch stalls the MMAs
  bool x3_ln_fused_dec = false;     // measured slower too: the 225 KB CTAs keep the next kernels of the chain from becoming resident
  // precise mode: the row-local parts of the decode step as persistent chain kernels (kernels_chain_x3.cu): 2 launches per
  // layer instead of 9.  Correct (the golden suite passes with it) but measured SLOWER (6 071 -> 4 787 audio-s/s): with
  // programmatic dependent launch the separate kernels already cost only ~3-8 us each, while a chain pays a counter
  // round trip and an exposed TMA latency per stage and its 225 KB CTAs block the SMs.  Opt-in: SCB_X3_CHAIN=1 / "x3_chain".
  bool x3_chain = false;
  bool chain_ready = false;
  CUtensorMap* d_chain_maps = nullptr;      // workspace: tensor maps of the chains' operands
  int* d_chain_ctr = nullptr;               // workspace: completion counters, one region per chain launch of a step
  float* d_chain_part = nullptr;            // workspace: split-K partial products of FFN2 [4][Rp][D]
  int chain_ctr_ints = 0, chain_rp = 0;
  std::vector<ChainParams> chains;          // [0] LN1 -> QKV of layer 0; [1 + 2 l] after self-attention l; [2 + 2 l] after cross-attention l
  std::vector<int> chain_items;
  int x3_dec_persist = 0;           // decode-step projections with plane operands on the persistent kernel: 1 = K 256, 2 = also FFN2
  bool mma_attn = false;            // bf16 mode: tensor-core (mma.sync) decoder attention
  bool mma_enc = false;             // bf16 mode: tensor-core encoder block attention
  // bf16 mode, experimental: fuse every LayerNorm into the epilogue of the GEMM producing its input (BN = 256 tiles).
  // Correct (tests/test_gpu_gemm_tc.py) but currently slower than separate LayerNorm kernels: the one-row-per-thread
  // epilogue over 256 columns serialises too much and BN = 256 leaves few CTAs; off by default.
  bool fuse_ln = false;
  bool fuse_ln_dec = false;         // same for the decoder step (self-O -> norm2, cross-O -> norm3, FFN2 -> next norm1)
  // bf16 mode: compute every LayerNorm inside the GEMM that consumes it (LayerNorm-prologue GEMM, K = 256)
  bool ln_prologue = false;
  bool ln_prologue_dec = false;     // same, decode step only (QKV, cross-Q and output GEMMs; small M, untested on device)
  // bf16 mode: encoder FFN1 -> ReLU -> FFN2 as one kernel with the hidden activation kept on the SM
  // (kernels_ffn_fused.cu); used when a launch has at least `fused_ffn_min_rows` rows
  bool fused_ffn = false;
  int fused_ffn_min_rows = 1024;
  // hidden-dimension splits of the fused FFN (partial tiles are added at the L2): 0 = automatic (enough CTAs to
  // spread a small-M launch over the GPU; results then depend on the add order at fp32 rounding level), 1 = never split
  int ffn_splits = 0;
  bool fused_ffn_dec = false;       // decode step: LayerNorm -> fused FFN (split) instead of LayerNorm -> FFN1 -> FFN2
  // deferred decoding: a push stops iterating once fewer than `lazy_threshold` streams are active and leaves
  // the stragglers' blocks queued on the device; they continue during later pushes (0 = strict, drain every push)
  int lazy_threshold = 0;
  // deferred mode: the frontend/encoder part of a push runs on its own stream while the caller's stream keeps
  // iterating the search for blocks queued by earlier pushes (they touch disjoint rows of the append-only buffers)
  bool overlap = true;
  cudaStream_t st_enc = nullptr;
  // optional SM partition of the encoder stream (option "encoder_sms" / SCB_ENC_SMS): the frontend + encoder kernels of
  // a push run inside a green context that owns only that many SMs, so the remaining SMs are always free for the
  // search chain's small dependent kernels (otherwise each of them first waits for a wave of encoder GEMM CTAs to drain)
  CUgreenCtx enc_gctx = nullptr;
  int enc_sms = 0;                   // SMs of the partition actually provisioned (0 = no partition)
  cudaEvent_t ev_in = nullptr, ev_wave = nullptr, ev_enc = nullptr;
  bool enc_pending = false;
  std::vector<int> pending_bound;   // host upper bound of queued blocks per stream
  // live kernel timing (bench.py roofline): CUDA-event pairs around every launch of one tagged kernel
  int prof_tag = 0;                 // 0 off, >0 one tag, -1 every tag (decode steps sampled every prof_stride)
  int prof_stride = 1;
  long step_seq = 0;
  bool prof_sample = false;         // current decode step is sampled (all-tags mode)
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_ev_tag;
  int prof_used = 0;
  double prof_flops = 0.0;          // host-known algorithmic FLOPs of the tagged launches (encoder GEMMs)

  // CUDA graphs (opt-in, "graph_decode" / "graph_encoder"): a search iteration is ~165 kernel launches whose
  // arguments never change (every length lives in device memory), the encoder stack of a push ~7 launches per layer
  // whose arguments depend only on the number of blocks.  Capturing them once and replaying the instantiated graph
  // takes the per-launch driver cost off the host threads (4 shard threads x ~140k launches per pass otherwise).
  // Replay is skipped while a kernel is being profiled; a failed capture (e.g. on the legacy default stream) falls
  // back to plain launches for the lifetime of the engine.
  bool graph_decode = false, graph_encoder = false, graph_failed = false;
  int graph_warm = 0;                       // plain iterations before the first capture (lazy launch-state set-up)
  cudaGraphExec_t step_graph = nullptr;
  int step_graph_launches = 0;
  struct EncGraph { cudaGraphExec_t exec; int launches; int uses; };
  std::unordered_map<int, EncGraph> enc_graphs;    // n_blk -> captured encoder stack

  // step trace (tests): the first `trace_max` search iterations after sc_engine_set_trace copy their score tensors
  // into a caller-owned device buffer (one fixed-size record per iteration, layout: sc_engine_trace_layout)
  unsigned char* trace_buf = nullptr;
  int trace_max = 0, trace_used = 0;
  int graphs_replayed = 0;                  // cudaGraphLaunch calls since creation (tests assert that capture happened)

  explicit Engine(const ScConfig& c) : cfg(c), cap(make_caps(c)), planner(c.n_streams), pending_bound(c.n_streams, 0), last_plan(c.n_streams) {}
};

// ---------------------------------------------------------------- encoder stream (optionally SM-partitioned)
// (Re)creates e.st_enc.  want_sms > 0: a stream of a green context holding ~want_sms SMs (CUDA rounds to its partition
// granularity, 8 SMs on sm_90+); any failure falls back to a plain lowest-priority stream of the primary context.
// The driver API is reached through cudaGetDriverEntryPoint (no link-time dependency on libcuda: the library must load
// on a host without a driver, where only the symbol table and the host-side planner are exercised).
struct DriverApi {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  bool ok = false;
};
static const DriverApi& driver_api() {
  static const DriverApi api = [] {
    DriverApi a;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      *fn = nullptr;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && *fn != nullptr;
    };
    a.ok = get("cuDeviceGet", (void**)&a.DeviceGet) && get("cuDeviceGetDevResource", (void**)&a.DeviceGetDevResource) &&
           get("cuDevSmResourceSplitByCount", (void**)&a.DevSmResourceSplitByCount) &&
           get("cuDevResourceGenerateDesc", (void**)&a.DevResourceGenerateDesc) && get("cuGreenCtxCreate", (void**)&a.GreenCtxCreate) &&
           get("cuGreenCtxDestroy", (void**)&a.GreenCtxDestroy) && get("cuGreenCtxStreamCreate", (void**)&a.GreenCtxStreamCreate);
    if (!a.ok) cudaGetLastError();
    return a;
  }();
  return api;
}
static void destroy_green_ctx(Engine& e);

static void make_encoder_stream(Engine& e, int want_sms) {
  if (e.st_enc) { cudaStreamSynchronize(e.st_enc); cudaStreamDestroy(e.st_enc); e.st_enc = nullptr; }
  destroy_green_ctx(e);
  e.enc_sms = 0;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  const DriverApi& d = driver_api();
  if (want_sms > 0 && d.ok) {
    int dev = 0;
    cudaGetDevice(&dev);
    CUdevice cudev;
    CUdevResource all, part, rest;
    unsigned int n_groups = 1;
    CUdevResourceDesc desc = nullptr;
    CUstream cs = nullptr;
    bool ok = d.DeviceGet(&cudev, dev) == CUDA_SUCCESS &&
              d.DeviceGetDevResource(cudev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
              (unsigned)want_sms < all.sm.smCount &&
              d.DevSmResourceSplitByCount(&part, &n_groups, &all, &rest, 0, (unsigned)want_sms) == CUDA_SUCCESS && n_groups == 1 &&
              d.DevResourceGenerateDesc(&desc, &part, 1) == CUDA_SUCCESS &&
              d.GreenCtxCreate(&e.enc_gctx, desc, cudev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
              d.GreenCtxStreamCreate(&cs, e.enc_gctx, CU_STREAM_NON_BLOCKING, lo) == CUDA_SUCCESS;
    if (ok) { e.st_enc = (cudaStream_t)cs; e.enc_sms = (int)part.sm.smCount; return; }
    destroy_green_ctx(e);
    cudaGetLastError();
  }
  // lowest priority, so that the search chain (small dependent kernels on the caller's stream, which callers create
  // with a higher priority) is never queued behind a wave of encoder GEMM CTAs
  cudaStreamCreateWithPriority(&e.st_enc, cudaStreamNonBlocking, lo);
}

static void destroy_green_ctx(Engine& e) {
  if (e.enc_gctx && driver_api().GreenCtxDestroy) driver_api().GreenCtxDestroy(e.enc_gctx);
  e.enc_gctx = nullptr;
}

// ---------------------------------------------------------------- workspace carving
struct Carver {
  unsigned char* base; size_t off = 0; bool dry;
  Carver(void* b, bool d) : base((unsigned char*)b), dry(d) {}
  template <typename T> T* take(size_t n) {
    off = align_up(off);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

static void carve(Engine& e, Carver& cv) {
  const ScConfig& c = e.cfg; const Caps& k = e.cap;
  const size_t S = c.n_streams, D = c.d_model, F = c.ffn, V = c.vocab, B = c.beam, R = k.R;
  auto reg = [&](const char* name, void* p, size_t n) { if (!cv.dry) e.named[name] = {p, n}; };
  e.d_mean = cv.take<double>(80); e.d_std = cv.take<double>(80);
  e.d_fe_tab = cv.take<FrontendTables>(1);
  e.wbuf = cv.take<float>(S * 512); reg("wbuf", e.wbuf, S * 512);
  e.featbuf = cv.take<float>(S * k.feat_cap * 80); reg("featbuf", e.featbuf, S * k.feat_cap * 80);
  e.h1 = cv.take<float>(S * k.t1_cap * 39 * D);
  e.h2 = cv.take<float>((size_t)k.sub_rows_max * 19 * D);
  e.subbuf = cv.take<float>(S * k.sub_cap * D); reg("subbuf", e.subbuf, S * k.sub_cap * D);
  e.prev_addin = cv.take<float>(S * D);
  e.enc_ctx = cv.take<float>(S * c.enc_layers * D);
  e.addin = cv.take<float>((size_t)k.nb_max * D);
  e.X = cv.take<float>((size_t)k.rows_max * D); reg("X", e.X, (size_t)k.rows_max * D);
  e.Nrm = cv.take<float>((size_t)k.rows_max * D);
  e.QKV = cv.take<float>((size_t)k.rows_max * 3 * D);
  e.Att = cv.take<float>((size_t)k.rows_max * D);
  e.FF = cv.take<float>((size_t)k.rows_max * F);
  e.encbuf = cv.take<float>(S * k.Tcap * D); reg("encbuf", e.encbuf, S * k.Tcap * D);
  e.ctcx = cv.take<float>(S * k.Tcap * V); reg("ctcx", e.ctcx, S * k.Tcap * V);
  e.dx = cv.take<float>(R * D); e.dn = cv.take<float>(R * D); e.dqkv = cv.take<float>(R * 3 * D);
  e.dq = cv.take<float>(R * D); e.dattn = cv.take<float>(R * D); e.dffn = cv.take<float>(R * F);
  e.dlogp = cv.take<float>(R * V); reg("dlogp", e.dlogp, R * V);
  if (c.precision == 2) {
    const size_t rt = (R + 127) / 128;
    e.chain_rp = (int)(rt * 128);
    e.chain_ctr_ints = (int)((1 + 2 * c.dec_layers) * 8 * rt * 4);
    e.d_chain_maps = cv.take<CUtensorMap>(32 + 16 * (size_t)c.dec_layers);
    e.d_chain_ctr = cv.take<int>(e.chain_ctr_ints);
    e.d_chain_part = cv.take<float>(4 * rt * 128 * D);
  }
  if (c.precision == 1) {
    e.Nrm16 = cv.take<__nv_bfloat16>((size_t)k.rows_max * D);
    e.Att16 = cv.take<__nv_bfloat16>((size_t)k.rows_max * D);
    e.FF16 = cv.take<__nv_bfloat16>((size_t)k.rows_max * F);
    e.h2_16 = cv.take<__nv_bfloat16>((size_t)k.sub_rows_max * 19 * D);
    e.dn16 = cv.take<__nv_bfloat16>(R * D); e.dattn16 = cv.take<__nv_bfloat16>(R * D);
    e.dffn16 = cv.take<__nv_bfloat16>(R * F);
    e.encnew16 = cv.take<__nv_bfloat16>(S * k.sub_cap * D);
    e.im2col16 = cv.take<__nv_bfloat16>((size_t)k.sub_rows_max * 19 * 9 * D);
    e.QKV16 = cv.take<__nv_bfloat16>((size_t)k.rows_max * 3 * D);
  }
  SearchBuffers& sb = e.sb;
  sb.S = c.n_streams; sb.B = c.beam; sb.V = c.vocab; sb.D = c.d_model; sb.H = c.dec_heads; sb.Ld = c.dec_layers;
  sb.Tcap = k.Tcap; sb.Lcap = k.Lcap; sb.use_bbd = c.use_bbd; sb.qcap = k.qcap;
  sb.w_dec = (float)(1.0 - (double)c.ctc_weight); sb.w_ctc = c.ctc_weight;
  sb.ctcx = e.ctcx;
  sb.kv_bf16 = c.precision == 1;
  {
    // precise mode, beam <= 16: K|V caches as split fp16 planes for the tensor-core attention (kernels_attn_x3.cu);
    // SCB_ATTN = "rows" / "cta" keeps fp32 caches and the CUDA-core attention kernels
    const char* a = getenv("SCB_ATTN");
    { const char* hm = getenv("SCB_ATTN_HEAD_MAJOR"); sb.attn_head_major = hm ? atoi(hm) : 1; }
    sb.kv_split = c.precision == 2 && c.beam <= 32 && !(a && (strcmp(a, "rows") == 0 || strcmp(a, "cta") == 0));
  }
  if (sb.kv_bf16) sb.xkv = reinterpret_cast<float*>(cv.take<__nv_bfloat16>((size_t)c.dec_layers * S * k.Tcap * 2 * D));
  else sb.xkv = cv.take<float>((size_t)c.dec_layers * S * k.Tcap * 2 * D);
  if (sb.kv_bf16) sb.skv = reinterpret_cast<float*>(cv.take<__nv_bfloat16>((size_t)c.dec_layers * S * k.Lcap * B * 2 * D));
  else sb.skv = cv.take<float>((size_t)c.dec_layers * S * k.Lcap * B * 2 * D);
  sb.yseq = cv.take<int>(2 * S * B * k.Lcap); reg("yseq", sb.yseq, 2 * S * B * k.Lcap);
  sb.xpos = cv.take<int>(2 * S * B * k.Lcap);
  sb.anc = cv.take<unsigned char>(2 * S * B * k.Lcap);
  sb.score = cv.take<double>(2 * S * B); sb.sc_dec = cv.take<double>(2 * S * B); sb.sc_ctc = cv.take<double>(2 * S * B);
  sb.ctc_r = cv.take<float>(2 * S * B * (size_t)k.Tcap * 2);
  sb.ctc_s = cv.take<float>(2 * S * B);
  sb.ctl = cv.take<StreamCtl>(S); reg("ctl", sb.ctl, S * sizeof(StreamCtl) / sizeof(int));
  sb.blkq_T = cv.take<int>(S * k.qcap); sb.blkq_final = cv.take<int>(S * k.qcap);
  sb.row_sh = cv.take<int>(R); sb.row_base = cv.take<int>(S); sb.act_streams = cv.take<int>(S);
  sb.n_rows = cv.take<int>(1); sb.n_active = cv.take<int>(2);   // [1] = capacity error flag
  sb.logp = e.dlogp;
  sb.pre_ids = cv.take<int>(R * kPreBeam); reg("pre_ids", sb.pre_ids, R * kPreBeam);
  sb.psi = cv.take<float>(R * kPreBeam); reg("psi", sb.psi, R * kPreBeam);
  sb.psi_eos = cv.take<float>(R);
  sb.cand_val = cv.take<float>(R * B); sb.cand_tok = cv.take<int>(R * B); sb.cand_dec = cv.take<float>(R * B);
  sb.cand_ctc = cv.take<float>(R * B); sb.cand_psi = cv.take<float>(R * B); sb.cand_col = cv.take<int>(R * B);
  sb.new_parent = cv.take<int>(S * B); sb.new_col = cv.take<int>(S * B); sb.upd_flag = cv.take<int>(S);
  sb.prof = cv.take<unsigned long long>(8);
  sb.key_cap = (int)(B * k.Lcap);
  sb.self_keys = cv.take<int>(S * (size_t)sb.key_cap); sb.self_nkeys = cv.take<int>(S);
  // descriptors
  e.d_fd = cv.take<FrontendDesc>(S); e.d_sd = cv.take<SubDesc>(S); e.d_blk = cv.take<BlockDesc>(k.nb_max);
  e.d_carry_f = cv.take<int>(3 * S); e.d_carry_s = cv.take<int>(3 * S);
  e.d_c2_a = cv.take<int64_t>((size_t)k.sub_rows_max * 19); e.d_c2_c = cv.take<int64_t>(k.sub_rows_max);
  e.d_en_a = cv.take<int64_t>(S * k.sub_cap); e.d_en_ctc = cv.take<int64_t>(S * k.sub_cap);
  e.d_en_kv = cv.take<int64_t>(S * k.sub_cap); e.d_en_flag = cv.take<int>(S * k.sub_cap);
  e.d_c2_seg = cv.take<int>(16);
  e.d_q = cv.take<int>(2 * S + 2 * S * k.qpush);
  e.d_reset = cv.take<int>(S);
}

// ---------------------------------------------------------------- linear dispatch (fp32 SIMT / bf16 tcgen05)
struct Lin {
  const float* A; int lda; const __nv_bfloat16* A16; const float* W; const __nv_bfloat16* W16; const float* bias;
  const float* R; int ldr; float* C; int ldc; __nv_bfloat16* C16; int M, N, K, relu; const int* n_rows_dev;
  const int64_t* c_row_off = nullptr;
  const float* ln_w = nullptr; const float* ln_b = nullptr; __nv_bfloat16* ln_out = nullptr;   // bf16 mode: fused LayerNorm
};

static Lin with_ln(Lin l, const float* w, const float* b, __nv_bfloat16* out) { l.ln_w = w; l.ln_b = b; l.ln_out = out; return l; }

// fp32-result GEMM: CUDA-core kernel (precision 0) or the split-fp16 tensor-core kernel (precision 2)
static int gemm_fp32(Engine& e, const GemmArgs& g, cudaStream_t st, const X3Extra& x = X3Extra()) {
  if (e.cfg.precision == 2) {
    auto it = e.x3map.find(g.W);
    if (it == e.x3map.end()) { set_last_error("precise tensor-core mode: no fp16 planes registered for a %dx%dx%d linear", g.M, g.N, g.K); return -1; }
    return launch_gemm_x3(g, x, it->second, st);
  }
  return launch_gemm_f32(g, st);
}

// ---------------------------------------------------------------- precise tensor-core mode (precision 2)
// Activations that feed a Linear are handed over as split fp16 planes (x3_split.cuh) written by their producer
// (LayerNorm, attention, the FFN1 / conv2 epilogue); the fp32 scratch buffers of the fp32 mode are reused for the planes
// (2 planes x 2 bytes = the 4 bytes per element they were carved with).  The residual stream, Q|K|V and the KV caches
// stay fp32.
struct Planes { void* base; size_t plane; int rows; };
static inline SplitOut split_out(const Planes& p, int ld) { SplitOut so; so.base = (__half*)p.base; so.plane = p.plane; so.ld = ld; return so; }

// C (fp32, += R) and / or C2 (planes) = act(A2 * W^T + bias)
static int x3_linear(Engine& e, const Planes& a, int K, const float* W, const float* bias, const float* R, float* C, int ldc,
                     const Planes* c2, int M, int N, int relu, const int* n_rows_dev, cudaStream_t st,
                     const int64_t* c_row_off = nullptr) {
  e.launches++;
  GemmArgs g;
  g.lda = K; g.W = W; g.bias = bias; g.R = R; g.ldr = ldc; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.relu = relu;
  g.n_rows_dev = n_rows_dev; g.c_row_off = c_row_off;
  X3Extra x;
  x.A2 = a.base; x.a2_plane = a.plane; x.a2_rows = a.rows;
  if (c2) { x.C2 = c2->base; x.c2_plane = c2->plane; x.ldc2 = N; }
  if (n_rows_dev && !c_row_off && ((e.x3_dec_persist >= 1 && K == 256) || e.x3_dec_persist >= 2) && gemm_x3p_eligible(g, x)) x.kernel = 2;
  return gemm_fp32(e, g, st, x);
}

// C and / or C2 = act(LayerNorm(X) * W^T + bias), K = d_model = 256: one launch of the persistent kernel with the
// LayerNorm computed in its prologue (x3_ln_fused), else LayerNorm kernel -> planes `scratch` -> GEMM (two launches,
// the LayerNorm under ln_tag).  persistent: which GEMM kernel the two-launch form uses with a device-side row count.
static int x3_ln_linear(Engine& e, int ln_tag, int gemm_tag, bool dec, const float* X, int ldx, const float* ln_w,
                        const float* ln_b, const Planes& scratch, const float* W, const float* bias, float* C, int ldc,
                        const Planes* c2, int M, int N, int relu, const int* n_rows_dev, cudaStream_t st);

static int linear(Engine& e, const Lin& l, cudaStream_t st) {
  e.launches++;
  if (e.cfg.precision == 1) {
    if (!l.A16 || !l.W16) { set_last_error("bf16 mode: missing bf16 operand for a %dx%dx%d linear", l.M, l.N, l.K); return -1; }
    return launch_gemm_bf16_ln(l.A16, l.lda, l.W16, l.bias, l.R, l.ldr, l.C, l.ldc, l.C16, l.ldc, l.c_row_off, l.M, l.N,
                               l.K, l.relu, l.n_rows_dev, l.ln_w, l.ln_b, l.ln_out, st);
  }
  GemmArgs g;
  g.A = l.A; g.lda = l.lda; g.W = l.W; g.bias = l.bias; g.R = l.R; g.ldr = l.ldr; g.C = l.C; g.ldc = l.ldc;
  g.M = l.M; g.N = l.N; g.K = l.K; g.relu = l.relu; g.n_rows_dev = l.n_rows_dev; g.c_row_off = l.c_row_off;
  return gemm_fp32(e, g, st);
}

#define TRY(x) do { int _r = (x); if (_r != 0) return _r; } while (0)

enum ProfTag { T_NONE = 0, T_CTC_PREFIX = 1, T_DEC_SELF_ATTN = 2, T_DEC_CROSS_ATTN = 3, T_DEC_FFN1 = 4, T_ENC_FFN1 = 5,
               T_PREBEAM = 6, T_ENC_ATTN = 7, T_CONV2 = 8, T_DEC_FFN2 = 9, T_ENC_FFN2 = 10, T_CTC_UPDATE = 11,
               T_FRONTEND = 12, T_CONV1 = 13, T_SUBOUT = 14, T_BLOCK_ASM = 15, T_ENC_LN = 16, T_ENC_QKV = 17, T_ENC_O = 18,
               T_ENC_HANDOVER = 19, T_STITCH = 20, T_CTC_HEAD = 21, T_XKV = 22, T_DEC_EMBED = 23, T_DEC_LN = 24,
               T_DEC_QKV = 25, T_DEC_SO = 26, T_DEC_CQ = 27, T_DEC_CO = 28, T_DEC_OUT = 29, T_COMBINE = 30, T_PRUNE = 31,
               T_STEP_FINISH = 32, T_DEC_STEP_TOTAL = 33, T_ENC_TOTAL = 34, T_COUNT = 35 };

static inline bool prof_on(const Engine& e, int tag, bool decode_kernel) {
  if (e.prof_tag == 0 || e.prof_used + 2 > (int)e.prof_ev.size()) return false;
  if (e.prof_tag == tag) return true;
  return e.prof_tag == -1 && (!decode_kernel || e.prof_sample);
}
static inline void prof_mark(Engine& e, int tag, cudaStream_t st, bool begin) {
  cudaEventRecord(e.prof_ev[e.prof_used], st);
  e.prof_ev_tag[e.prof_used] = begin ? tag : -tag;
  e.prof_used++;
}
// wraps a launch in an event pair when its tag is being profiled (D = decode-step kernel, E = per-push kernel)
#define PROFX(tag, dec, expr)                                \
  do {                                                       \
    const bool _p = prof_on(e, (tag), (dec));                \
    if (_p) prof_mark(e, (tag), st, true);                   \
    TRY(expr);                                               \
    if (_p) prof_mark(e, (tag), st, false);                  \
  } while (0)
#define PD(tag, expr) PROFX(tag, true, expr)
#define PE(tag, expr) PROFX(tag, false, expr)

// LayerNorm(X) followed by a Linear.  bf16 mode with ln_prologue: one kernel (the GEMM normalises its own A tile);
// otherwise a LayerNorm kernel (fp32 or bf16 output into l.A / l.A16) followed by the GEMM, each under its own tag.
static int ln_linear(Engine& e, int ln_tag, int gemm_tag, bool dec, const float* X, int ldx, const float* ln_w,
                     const float* ln_b, int rows_max, const Lin& l, cudaStream_t st) {
  const bool tc = e.cfg.precision == 1;
  const bool fuse = dec ? e.fuse_ln_dec : e.fuse_ln;
  const bool lnp = e.ln_prologue || (dec && e.ln_prologue_dec);
  if (tc && lnp && !fuse) {
    e.launches++;
    PROFX(gemm_tag, dec, launch_gemm_bf16_lnA(X, ldx, ln_w, ln_b, l.W16, l.bias, l.C, l.ldc, l.C16, l.ldc, l.M, l.N, l.relu,
                                             l.n_rows_dev, st));
    return 0;
  }
  if (!(tc && fuse)) {
    e.launches++;
    if (tc) PROFX(ln_tag, dec, launch_layernorm_bf16(X, ldx, ln_w, ln_b, const_cast<__nv_bfloat16*>(l.A16), l.lda, rows_max, l.K, l.n_rows_dev, st));
    else PROFX(ln_tag, dec, launch_layernorm(X, ldx, ln_w, ln_b, const_cast<float*>(l.A), l.lda, rows_max, l.K, l.n_rows_dev, st));
  }
  PROFX(gemm_tag, dec, linear(e, l, st));
  return 0;
}

// splits of the fused FFN's hidden dimension so that a launch has roughly 64+ CTAs (one CTA per SM, 148 SMs)
static int ffn_auto_splits(const Engine& e, int rows, int F) {
  if (e.ffn_splits >= 1) return (F / 128) % e.ffn_splits == 0 ? e.ffn_splits : 1;
  const int tiles = (rows + 127) / 128, chunks = F / 128;
  int s = 1;
  while (tiles * s < 64 && s * 2 <= 8 && chunks % (s * 2) == 0) s *= 2;
  return s;
}

static int x3_ln_linear(Engine& e, int ln_tag, int gemm_tag, bool dec, const float* X, int ldx, const float* ln_w,
                        const float* ln_b, const Planes& scratch, const float* W, const float* bias, float* C, int ldc,
                        const Planes* c2, int M, int N, int relu, const int* n_rows_dev, cudaStream_t st) {
  const int K = e.cfg.d_model;
  const bool fuse = dec ? e.x3_ln_fused_dec : e.x3_ln_fused_enc;
  if (fuse && K == 256 && N % 128 == 0) {
    e.launches++;
    GemmArgs g;
    g.lda = K; g.W = W; g.bias = bias; g.C = C; g.ldc = ldc; g.ldr = ldc; g.M = M; g.N = N; g.K = K; g.relu = relu; g.n_rows_dev = n_rows_dev;
    X3Extra x;
    x.lnX = X; x.ldx = ldx; x.ln_w = ln_w; x.ln_b = ln_b;
    if (c2) { x.C2 = c2->base; x.c2_plane = c2->plane; x.ldc2 = N; }
    PROFX(gemm_tag, dec, gemm_fp32(e, g, st, x));
    return 0;
  }
  e.launches++;
  PROFX(ln_tag, dec, launch_layernorm_split(X, ldx, ln_w, ln_b, scratch.base, scratch.plane, K, M, K, n_rows_dev, st));
  PROFX(gemm_tag, dec, x3_linear(e, scratch, K, W, bias, nullptr, C, ldc, c2, M, N, relu, n_rows_dev, st));
  return 0;
}

// ---------------------------------------------------------------- encoder layers over all blocks of the push
static int run_encoder_layers_x3(Engine& e, int n_blk, cudaStream_t st) {
  const ScConfig& c = e.cfg; const int D = c.d_model, F = c.ffn, rows = n_blk * kSlots, cap = e.cap.rows_max;
  const Planes nrm{e.Nrm, (size_t)cap * D, cap}, att{e.Att, (size_t)cap * D, cap}, ff{e.FF, (size_t)cap * F, cap};
  for (int l = 0; l < c.enc_layers; ++l) {
    const EncLayerW& w = e.enc[l];
    if (e.prof_tag == T_ENC_QKV) e.prof_flops += 2.0 * n_blk * 41 * 3.0 * D * D;
    TRY(x3_ln_linear(e, T_ENC_LN, T_ENC_QKV, false, e.X, D, w.ln1w, w.ln1b, nrm, w.qkvw, w.qkvb, e.QKV, 3 * D, nullptr, rows, 3 * D, 0, nullptr, st));
    if (e.enc_attn_x3) PE(T_ENC_ATTN, launch_enc_attention_x3(e.QKV, nullptr, e.d_blk, n_blk, c.enc_heads, D, split_out(att, D), st));
    else PE(T_ENC_ATTN, launch_enc_attention(e.QKV, nullptr, nullptr, e.d_blk, n_blk, c.enc_heads, D, st, split_out(att, D)));
    PE(T_ENC_O, x3_linear(e, att, D, w.ow, w.ob, e.X, e.X, D, nullptr, rows, D, 0, nullptr, st));
    // algorithmic FLOPs of the roofline: 41 useful rows per block; 42 are executed
    if (e.prof_tag == T_ENC_FFN1 || e.prof_tag == T_ENC_FFN2) e.prof_flops += 2.0 * n_blk * 41 * (double)F * D;
    TRY(x3_ln_linear(e, T_ENC_LN, T_ENC_FFN1, false, e.X, D, w.ln2w, w.ln2b, nrm, w.f1w, w.f1b, nullptr, F, &ff, rows, F, 1, nullptr, st));
    PE(T_ENC_FFN2, x3_linear(e, ff, F, w.f2w, w.f2b, e.X, e.X, D, nullptr, rows, D, 0, nullptr, st));
    PE(T_ENC_HANDOVER, launch_ctx_handover(e.X, e.enc_ctx, l, c.enc_layers, e.d_blk, n_blk, D, nullptr, nullptr, nullptr, st));
    e.launches += 2;
  }
  return 0;
}

static int run_encoder_layers(Engine& e, int n_blk, cudaStream_t st) {
  if (e.cfg.precision == 2) return run_encoder_layers_x3(e, n_blk, st);
  const ScConfig& c = e.cfg; const int D = c.d_model, F = c.ffn, rows = n_blk * kSlots;
  const bool tc = c.precision == 1;
  const bool fl = tc && e.fuse_ln;
  // fused mode: only the first LayerNorm is a kernel of its own; every other norm rides in the epilogue of the GEMM
  // that produces its input (O-proj -> norm2, FFN2 -> next layer's norm1; the hand-over re-normalises slot 0)
  if (fl) PE(T_ENC_LN, launch_layernorm_bf16(e.X, D, e.enc[0].ln1w, e.enc[0].ln1b, e.Nrm16, D, rows, D, nullptr, st));
  for (int l = 0; l < c.enc_layers; ++l) {
    const EncLayerW& w = e.enc[l];
    const EncLayerW* nx = l + 1 < c.enc_layers ? &e.enc[l + 1] : nullptr;
    if (e.mma_enc) {
      TRY(ln_linear(e, T_ENC_LN, T_ENC_QKV, false, e.X, D, w.ln1w, w.ln1b, rows, Lin{e.Nrm, D, e.Nrm16, w.qkvw, w.qkvw16, w.qkvb, nullptr, 0, nullptr, 3 * D, e.QKV16, rows, 3 * D, D, 0, nullptr}, st));
      PE(T_ENC_ATTN, launch_enc_attention_mma(e.QKV16, nullptr, e.Att16, e.d_blk, n_blk, c.enc_heads, D, st));
    } else {
      TRY(ln_linear(e, T_ENC_LN, T_ENC_QKV, false, e.X, D, w.ln1w, w.ln1b, rows, Lin{e.Nrm, D, e.Nrm16, w.qkvw, w.qkvw16, w.qkvb, nullptr, 0, e.QKV, 3 * D, nullptr, rows, 3 * D, D, 0, nullptr}, st));
      PE(T_ENC_ATTN, launch_enc_attention(e.QKV, e.Att, tc ? e.Att16 : nullptr, e.d_blk, n_blk, c.enc_heads, D, st));
    }
    {
      Lin o{e.Att, D, e.Att16, w.ow, w.ow16, w.ob, e.X, D, e.X, D, nullptr, rows, D, D, 0, nullptr};
      if (fl) o = with_ln(o, w.ln2w, w.ln2b, e.Nrm16);
      PE(T_ENC_O, linear(e, o, st));
    }
    // algorithmic FLOPs of the roofline: 41 useful rows per block; 42 are executed
    if (e.prof_tag == T_ENC_FFN1 || e.prof_tag == T_ENC_FFN2) e.prof_flops += 2.0 * n_blk * 41 * (double)F * D;
    if (tc && e.fused_ffn && !fl && !e.ln_prologue && rows >= e.fused_ffn_min_rows) {
      // LayerNorm -> one fused FFN kernel (hidden activation stays in TMEM / shared memory); timed under the FFN2 tag
      e.launches += 2;
      PE(T_ENC_LN, launch_layernorm_bf16(e.X, D, w.ln2w, w.ln2b, e.Nrm16, D, rows, D, nullptr, st));
      if (e.prof_tag == T_ENC_FFN2) e.prof_flops += 2.0 * n_blk * 41 * (double)F * D;
      PE(T_ENC_FFN2, launch_ffn_fused_bf16(e.Nrm16, D, w.f1w16, w.f1b, w.f2w16, w.f2b, e.X, D, 1, rows, F,
                                           ffn_auto_splits(e, rows, F), nullptr, st));
    } else {
      TRY(ln_linear(e, T_ENC_LN, T_ENC_FFN1, false, e.X, D, w.ln2w, w.ln2b, rows, Lin{e.Nrm, D, e.Nrm16, w.f1w, w.f1w16, w.f1b, nullptr, 0, tc ? nullptr : e.FF, F, tc ? e.FF16 : nullptr, rows, F, D, 1, nullptr}, st));
      {
        Lin f2{e.FF, F, tc ? e.FF16 : nullptr, w.f2w, w.f2w16, w.f2b, e.X, D, e.X, D, nullptr, rows, D, F, 0, nullptr};
        if (fl && nx) f2 = with_ln(f2, nx->ln1w, nx->ln1b, e.Nrm16);
        PE(T_ENC_FFN2, linear(e, f2, st));
      }
    }
    PE(T_ENC_HANDOVER, launch_ctx_handover(e.X, e.enc_ctx, l, c.enc_layers, e.d_blk, n_blk, D, (fl && nx) ? nx->ln1w : nullptr,
                                           (fl && nx) ? nx->ln1b : nullptr, e.Nrm16, st));
    e.launches += fl ? 2 : 4;
  }
  return 0;
}

// ---------------------------------------------------------------- step trace (direct K6/K7/K8 parity tests)
// Record of one search iteration, taken after score combination and before pruning (all sizes in bytes, 256-aligned):
//   0 n_rows int[1] | 1 row_sh int[R] | 2 logp float[R][V] | 3 pre_ids int[R][40] | 4 psi float[R][40] | 5 psi_eos float[R]
//   | 6 ctc_s float[2][S][B] | 7 ctl int[S][16]
static void trace_layout(const Engine& e, int64_t off[8], int64_t* rec_bytes) {
  const int64_t R = e.cap.R, V = e.cfg.vocab, S = e.cfg.n_streams, B = e.cfg.beam;
  const int64_t sz[8] = {4, 4 * R, 4 * R * V, 4 * R * kPreBeam, 4 * R * kPreBeam, 4 * R, 4 * 2 * S * B, (int64_t)sizeof(StreamCtl) * S};
  int64_t o = 0;
  for (int i = 0; i < 8; ++i) { off[i] = o; o += (int64_t)align_up((size_t)sz[i]); }
  *rec_bytes = o;
}

static int trace_step(Engine& e, cudaStream_t st) {
  int64_t off[8], rec; trace_layout(e, off, &rec);
  unsigned char* d = e.trace_buf + (size_t)rec * e.trace_used;
  const SearchBuffers& sb = e.sb;
  const size_t R = e.cap.R, V = e.cfg.vocab, S = e.cfg.n_streams, B = e.cfg.beam;
  auto cp = [&](int i, const void* src, size_t bytes) { return cudaMemcpyAsync(d + off[i], src, bytes, cudaMemcpyDeviceToDevice, st); };
  SCB_CUDA_CHECK(cp(0, sb.n_rows, 4));
  SCB_CUDA_CHECK(cp(1, sb.row_sh, 4 * R));
  SCB_CUDA_CHECK(cp(2, e.dlogp, 4 * R * V));
  SCB_CUDA_CHECK(cp(3, sb.pre_ids, 4 * R * kPreBeam));
  SCB_CUDA_CHECK(cp(4, sb.psi, 4 * R * kPreBeam));
  SCB_CUDA_CHECK(cp(5, sb.psi_eos, 4 * R));
  SCB_CUDA_CHECK(cp(6, sb.ctc_s, 4 * 2 * S * B));
  SCB_CUDA_CHECK(cp(7, sb.ctl, sizeof(StreamCtl) * S));
  e.trace_used++;
  return 0;
}

// ---------------------------------------------------------------- one search iteration for all active streams

// ---------------------------------------------------------------- decode-step chains (kernels_chain_x3.cu)
// Built once per engine: every operand of the decode step lives at a fixed address (activation buffers of the
// workspace, weight planes), so the tensor maps and the stage lists never change.
static int build_decode_chains(Engine& e) {
  const ScConfig& c = e.cfg;
  const int D = c.d_model, F = c.ffn, V = c.vocab, R = e.cap.R, Ld = c.dec_layers;
  e.chain_ready = false;
  if (c.precision != 2 || D != 256 || F % 512 != 0 || V % 128 != 0 || !e.d_chain_maps) return 0;
  std::vector<CUtensorMap> maps;
  auto add_load = [&](const void* planes, size_t plane_elems, int rows, int cols, int ld) -> int {       // hi, lo: 64 x 128 boxes
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(planes);
    CUtensorMap m0, m1;
    if (tc_get_map(h, rows, cols, ld, 128, &m0) || tc_get_map(h + plane_elems, rows, cols, ld, 128, &m1)) return -1;
    maps.push_back(m0); maps.push_back(m1);
    return (int)maps.size() - 2;
  };
  auto add_out_planes = [&](void* planes, size_t plane_elems, int rows, int cols, int ld) -> int {       // hi, lo: 64 x 32 boxes
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(planes);
    CUtensorMap m0, m1;
    if (tc_get_map(h, rows, cols, ld, 32, &m0) || tc_get_map(h + plane_elems, rows, cols, ld, 32, &m1)) return -1;
    maps.push_back(m0); maps.push_back(m1);
    return (int)maps.size() - 2;
  };
  auto add_out_f32 = [&](float* ptr, int rows, int cols, int ld) -> int {
    CUtensorMap m0;
    if (tc_get_map_f32(ptr, rows, cols, ld, 32, &m0)) return -1;
    maps.push_back(m0);
    return (int)maps.size() - 1;
  };
  auto wplanes = [&](const float* w) -> const void* { auto it = e.x3map.find(w); return it == e.x3map.end() ? nullptr : it->second; };
  // activations
  const int a_dn = add_load(e.dn, (size_t)R * D, R, D, D), a_da = add_load(e.dattn, (size_t)R * D, R, D, D);
  const int a_df = add_load(e.dffn, (size_t)R * F, R, F, F);
  const int o_df = add_out_planes(e.dffn, (size_t)R * F, R, F, F);
  const int o_dx = add_out_f32(e.dx, R, D, D), o_qkv = add_out_f32(e.dqkv, R, 3 * D, 3 * D), o_dq = add_out_f32(e.dq, R, D, D);
  const int o_logp = add_out_f32(e.dlogp, R, V, V), o_part = add_out_f32(e.d_chain_part, 4 * e.chain_rp, D, D);
  if (a_dn < 0 || a_da < 0 || a_df < 0 || o_df < 0 || o_dx < 0 || o_qkv < 0 || o_dq < 0 || o_logp < 0 || o_part < 0) return -1;
  struct LW { int qkv, so, cq, co, f1, f2; };
  std::vector<LW> lw(Ld);
  for (int l = 0; l < Ld; ++l) {
    const DecLayerW& w = e.dec[l];
    const void *pq = wplanes(w.sqkvw), *ps = wplanes(w.sow), *pc = wplanes(w.cqw), *po = wplanes(w.cow), *p1 = wplanes(w.f1w), *p2 = wplanes(w.f2w);
    if (!pq || !ps || !pc || !po || !p1 || !p2) return 0;
    lw[l].qkv = add_load(pq, (size_t)3 * D * D, 3 * D, D, D); lw[l].so = add_load(ps, (size_t)D * D, D, D, D);
    lw[l].cq = add_load(pc, (size_t)D * D, D, D, D); lw[l].co = add_load(po, (size_t)D * D, D, D, D);
    lw[l].f1 = add_load(p1, (size_t)F * D, F, D, D); lw[l].f2 = add_load(p2, (size_t)D * F, D, F, F);
    if (lw[l].qkv < 0 || lw[l].so < 0 || lw[l].cq < 0 || lw[l].co < 0 || lw[l].f1 < 0 || lw[l].f2 < 0) return -1;
  }
  const void* pout = wplanes(e.doutw);
  if (!pout) return 0;
  const int w_out = add_load(pout, (size_t)V * D, V, D, D);
  if (w_out < 0) return -1;
  if (maps.size() > 32 + 16 * (size_t)Ld) { set_last_error("decode chains: %zu tensor maps exceed the carved table", maps.size()); return -1; }
  SCB_CUDA_CHECK(cudaMemcpy(e.d_chain_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));

  const int RT = (R + 127) / 128, region = 8 * RT * 4;                 // counters per chain launch: 8 groups of RT * 4
  auto gemm = [&](int map_a, int map_w, int map_o, int n_ct, int n_ks, int kb_item, int out_mode, int relu, const float* bias) {
    ChainStage g{};
    g.type = 0; g.n_ct = n_ct; g.n_ks = n_ks; g.kb_item = kb_item; g.map_a = map_a; g.map_w = map_w; g.map_o = map_o;
    g.out_mode = out_mode; g.relu = relu; g.bias = bias; g.wait_base = -1; g.sig_base = -1; g.sig_div = 1 << 20; g.sig_stride = 1; g.wait_stride = 1;
    return g;
  };
  auto lnorm = [&](const float* w, const float* b, int n_part, const float* pbias) {
    ChainStage g{};
    g.type = 1; g.ln_w = w; g.ln_b = b; g.x = e.dx; g.n_part = n_part; g.part = e.d_chain_part; g.pbias = pbias;
    g.out_hi = reinterpret_cast<__half*>(e.dn); g.out_plane = (size_t)R * D; g.wait_base = -1; g.sig_base = -1; g.sig_div = 1 << 20;
    g.sig_stride = 1; g.wait_stride = 1;
    return g;
  };
  // stage i signals counter group i (stride 1 per row tile unless stated), stage i + 1 waits for it
  auto link = [&](ChainStage& prod, ChainStage& cons, int group, int target_items) {
    prod.sig_base = group * RT * 4; prod.sig_stride = 1;
    cons.wait_base = group * RT * 4; cons.wait_stride = 1; cons.wait_ks = 0; cons.wait_target = 8 * target_items;
  };
  e.chains.assign(1 + 2 * Ld, ChainParams{});
  e.chain_items.assign(1 + 2 * Ld, 0);
  auto finish = [&](int idx, std::vector<ChainStage>& st) {
    ChainParams& cp = e.chains[idx];
    cp.n_stages = (int)st.size();
    int items = 0;
    for (size_t i = 0; i < st.size(); ++i) { cp.st[i] = st[i]; items += RT * (st[i].type == 0 ? st[i].n_ct * st[i].n_ks : 1); }
    cp.maps = e.d_chain_maps; cp.ctr = e.d_chain_ctr + (size_t)idx * region; cp.n_rows_dev = e.sb.n_rows; cp.M = R;
    cp.part_stride_rows = e.chain_rp;
    e.chain_items[idx] = items;
  };
  {                                                                    // chain 0: LayerNorm1 -> QKV of layer 0
    std::vector<ChainStage> st{lnorm(e.dec[0].ln1w, e.dec[0].ln1b, 0, nullptr),
                               gemm(a_dn, lw[0].qkv, o_qkv, 3 * D / 128, 1, D / 64, 0, 0, e.dec[0].sqkvb)};
    link(st[0], st[1], 0, 1);
    finish(0, st);
  }
  for (int l = 0; l < Ld; ++l) {
    const DecLayerW& w = e.dec[l];
    {                                                                  // after self-attention: O (+x) -> LayerNorm2 -> cross-Q
      std::vector<ChainStage> st{gemm(a_da, lw[l].so, o_dx, D / 128, 1, D / 64, 1, 0, w.sob), lnorm(w.ln2w, w.ln2b, 0, nullptr),
                                 gemm(a_dn, lw[l].cq, o_dq, D / 128, 1, D / 64, 0, 0, w.cqb)};
      link(st[0], st[1], 0, D / 128);
      link(st[1], st[2], 1, 1);
      finish(1 + 2 * l, st);
    }
    {   // after cross-attention: O (+x) -> LayerNorm3 -> FFN1 -> FFN2 (4 K splits) -> fold + next LayerNorm -> next projection
      const bool last = l + 1 == Ld;
      std::vector<ChainStage> st{gemm(a_da, lw[l].co, o_dx, D / 128, 1, D / 64, 1, 0, w.cob), lnorm(w.ln3w, w.ln3b, 0, nullptr),
                                 gemm(a_dn, lw[l].f1, o_df, F / 128, 1, D / 64, 2, 1, w.f1b),
                                 gemm(a_df, lw[l].f2, o_part, D / 128, 4, F / 4 / 64, 3, 0, nullptr),
                                 last ? lnorm(e.daw, e.dab, 4, w.f2b) : lnorm(e.dec[l + 1].ln1w, e.dec[l + 1].ln1b, 4, w.f2b),
                                 last ? gemm(a_dn, w_out, o_logp, V / 128, 1, D / 64, 0, 0, e.doutb)
                                      : gemm(a_dn, lw[l + 1].qkv, o_qkv, 3 * D / 128, 1, D / 64, 0, 0, e.dec[l + 1].sqkvb)};
      link(st[0], st[1], 0, D / 128);
      link(st[1], st[2], 1, 1);
      // FFN1 column tiles of 128 -> the K split of FFN2 that consumes them (F / 4 hidden units = F / 512 column tiles each)
      st[2].sig_base = 2 * RT * 4; st[2].sig_stride = 4; st[2].sig_div = F / 4 / 128;
      st[3].wait_base = 2 * RT * 4; st[3].wait_stride = 4; st[3].wait_ks = 1; st[3].wait_target = 8 * (F / 4 / 128);
      link(st[3], st[4], 3, (D / 128) * 4);
      link(st[4], st[5], 4, 1);
      finish(2 + 2 * l, st);
    }
  }
  e.chain_ready = true;
  return 0;
}

static int decode_tail(Engine& e, cudaStream_t st);

// one search iteration with the row-local chains: embed, [LN1 -> QKV], then per layer self-attention, chain, cross-attention, chain
static int run_decode_step_x3_chain(Engine& e, cudaStream_t st) {
  const ScConfig& c = e.cfg; const int D = c.d_model, R = e.cap.R;
  const SearchBuffers& sb = e.sb;
  const Planes da{e.dattn, (size_t)R * D, R};
  SCB_CUDA_CHECK(cudaMemsetAsync(e.d_chain_ctr, 0, (size_t)e.chain_ctr_ints * sizeof(int), st));
  PD(T_DEC_EMBED, launch_dec_embed(sb, e.demb, e.pe, e.dx, nullptr, nullptr, nullptr, st));
  PD(T_DEC_EMBED, launch_build_self_keys(sb, st));
  PD(T_DEC_QKV, launch_chain_x3(e.chains[0], e.chain_items[0], st));
  e.launches += 3;
  for (int l = 0; l < c.dec_layers; ++l) {
    PD(T_DEC_SELF_ATTN, launch_dec_attention_x3(sb, 0, l, e.dqkv, 3 * D, nullptr, split_out(da, D), st));
    PD(T_DEC_SO, launch_chain_x3(e.chains[1 + 2 * l], e.chain_items[1 + 2 * l], st));
    PD(T_DEC_CROSS_ATTN, launch_dec_attention_x3(sb, 1, l, e.dq, D, nullptr, split_out(da, D), st));
    PD(T_DEC_FFN1, launch_chain_x3(e.chains[2 + 2 * l], e.chain_items[2 + 2 * l], st));
    e.launches += 4;
  }
  return decode_tail(e, st);
}

static int run_decode_step_x3(Engine& e, cudaStream_t st) {
  const ScConfig& c = e.cfg; const int D = c.d_model, F = c.ffn, V = c.vocab, R = e.cap.R;
  const SearchBuffers& sb = e.sb;
  if (e.x3_chain && e.chain_ready && sb.kv_split) return run_decode_step_x3_chain(e, st);
  const int* nr = sb.n_rows;
  const Planes dn{e.dn, (size_t)R * D, R}, da{e.dattn, (size_t)R * D, R}, df{e.dffn, (size_t)R * F, R};
  PD(T_DEC_EMBED, launch_dec_embed(sb, e.demb, e.pe, e.dx, nullptr, nullptr, nullptr, st));
  if (e.attn_f32_rows || sb.kv_split) { PD(T_DEC_EMBED, launch_build_self_keys(sb, st)); e.launches++; }
  for (int l = 0; l < c.dec_layers; ++l) {
    const DecLayerW& w = e.dec[l];
    TRY(x3_ln_linear(e, T_DEC_LN, T_DEC_QKV, true, e.dx, D, w.ln1w, w.ln1b, dn, w.sqkvw, w.sqkvb, e.dqkv, 3 * D, nullptr, R, 3 * D, 0, nr, st));
    if (sb.kv_split) PD(T_DEC_SELF_ATTN, launch_dec_attention_x3(sb, 0, l, e.dqkv, 3 * D, nullptr, split_out(da, D), st));
    else if (e.attn_f32_rows) PD(T_DEC_SELF_ATTN, launch_dec_attention_f32(sb, 0, l, e.dqkv, 3 * D, nullptr, split_out(da, D), st));
    else PD(T_DEC_SELF_ATTN, launch_dec_self_attention(sb, l, e.dqkv, 3 * D, nullptr, nullptr, st, split_out(da, D)));
    PD(T_DEC_SO, x3_linear(e, da, D, w.sow, w.sob, e.dx, e.dx, D, nullptr, R, D, 0, nr, st));
    TRY(x3_ln_linear(e, T_DEC_LN, T_DEC_CQ, true, e.dx, D, w.ln2w, w.ln2b, dn, w.cqw, w.cqb, e.dq, D, nullptr, R, D, 0, nr, st));
    if (sb.kv_split) PD(T_DEC_CROSS_ATTN, launch_dec_attention_x3(sb, 1, l, e.dq, D, nullptr, split_out(da, D), st));
    else if (e.attn_f32_rows) PD(T_DEC_CROSS_ATTN, launch_dec_attention_f32(sb, 1, l, e.dq, D, nullptr, split_out(da, D), st));
    else PD(T_DEC_CROSS_ATTN, launch_dec_cross_attention(sb, l, e.dq, D, nullptr, nullptr, st, split_out(da, D)));
    PD(T_DEC_CO, x3_linear(e, da, D, w.cow, w.cob, e.dx, e.dx, D, nullptr, R, D, 0, nr, st));
    TRY(x3_ln_linear(e, T_DEC_LN, T_DEC_FFN1, true, e.dx, D, w.ln3w, w.ln3b, dn, w.f1w, w.f1b, nullptr, F, &df, R, F, 1, nr, st));
    PD(T_DEC_FFN2, x3_linear(e, df, F, w.f2w, w.f2b, e.dx, e.dx, D, nullptr, R, D, 0, nr, st));
    e.launches += 2;
  }
  TRY(x3_ln_linear(e, T_DEC_LN, T_DEC_OUT, true, e.dx, D, e.daw, e.dab, dn, e.doutw, e.doutb, e.dlogp, V, nullptr, R, V, 0, nr, st));
  return decode_tail(e, st);
}

static int run_decode_step(Engine& e, cudaStream_t st) {
  const ScConfig& c = e.cfg; const int D = c.d_model, F = c.ffn, V = c.vocab, R = e.cap.R;
  const SearchBuffers& sb = e.sb;
  const int* nr = sb.n_rows;
  const bool tc = c.precision == 1;
  e.prof_sample = (e.step_seq++ % e.prof_stride) == 0;
  const bool ptot = prof_on(e, T_DEC_STEP_TOTAL, true);
  if (ptot) prof_mark(e, T_DEC_STEP_TOTAL, st, true);
  if (c.precision == 2) {
    TRY(run_decode_step_x3(e, st));
    if (ptot) prof_mark(e, T_DEC_STEP_TOTAL, st, false);
    return 0;
  }
  const bool fl = tc && e.fuse_ln_dec;
  PD(T_DEC_EMBED, launch_dec_embed(sb, e.demb, e.pe, e.dx, fl ? e.dec[0].ln1w : nullptr, fl ? e.dec[0].ln1b : nullptr, e.dn16, st));
  const bool rows32 = !tc && e.attn_f32_rows;      // fp32 mode: warp-per-head attention over the per-step key list
  if (e.mma_attn || rows32) PD(T_DEC_EMBED, launch_build_self_keys(sb, st));
  for (int l = 0; l < c.dec_layers; ++l) {
    const DecLayerW& w = e.dec[l];
    // bf16 mode: norm1 comes fused from dec_embed / the previous layer's FFN2, norm2 from self-O, norm3 from cross-O
    TRY(ln_linear(e, T_DEC_LN, T_DEC_QKV, true, e.dx, D, w.ln1w, w.ln1b, R, Lin{e.dn, D, e.dn16, w.sqkvw, w.sqkvw16, w.sqkvb, nullptr, 0, e.dqkv, 3 * D, nullptr, R, 3 * D, D, 0, nr}, st));
    if (e.mma_attn) PD(T_DEC_SELF_ATTN, launch_dec_attention_mma(sb, 0, l, e.dqkv, 3 * D, e.dattn, e.dattn16, st));
    else if (rows32) PD(T_DEC_SELF_ATTN, launch_dec_attention_f32(sb, 0, l, e.dqkv, 3 * D, e.dattn, SplitOut(), st));
    else PD(T_DEC_SELF_ATTN, launch_dec_self_attention(sb, l, e.dqkv, 3 * D, e.dattn, tc ? e.dattn16 : nullptr, st));
    {
      Lin o{e.dattn, D, e.dattn16, w.sow, w.sow16, w.sob, e.dx, D, e.dx, D, nullptr, R, D, D, 0, nr};
      if (fl) o = with_ln(o, w.ln2w, w.ln2b, e.dn16);
      PD(T_DEC_SO, linear(e, o, st));
    }
    TRY(ln_linear(e, T_DEC_LN, T_DEC_CQ, true, e.dx, D, w.ln2w, w.ln2b, R, Lin{e.dn, D, e.dn16, w.cqw, w.cqw16, w.cqb, nullptr, 0, e.dq, D, nullptr, R, D, D, 0, nr}, st));
    if (e.mma_attn) PD(T_DEC_CROSS_ATTN, launch_dec_attention_mma(sb, 1, l, e.dq, D, e.dattn, e.dattn16, st));
    else if (rows32) PD(T_DEC_CROSS_ATTN, launch_dec_attention_f32(sb, 1, l, e.dq, D, e.dattn, SplitOut(), st));
    else PD(T_DEC_CROSS_ATTN, launch_dec_cross_attention(sb, l, e.dq, D, e.dattn, tc ? e.dattn16 : nullptr, st));
    {
      Lin o{e.dattn, D, e.dattn16, w.cow, w.cow16, w.cob, e.dx, D, e.dx, D, nullptr, R, D, D, 0, nr};
      if (fl) o = with_ln(o, w.ln3w, w.ln3b, e.dn16);
      PD(T_DEC_CO, linear(e, o, st));
    }
    if (tc && e.fused_ffn_dec && !fl && !e.ln_prologue) {
      e.launches += 2;
      PD(T_DEC_LN, launch_layernorm_bf16(e.dx, D, w.ln3w, w.ln3b, e.dn16, D, R, D, nr, st));
      PD(T_DEC_FFN2, launch_ffn_fused_bf16(e.dn16, D, w.f1w16, w.f1b, w.f2w16, w.f2b, e.dx, D, 1, R, F,
                                           ffn_auto_splits(e, R, F), nr, st));
    } else {
      TRY(ln_linear(e, T_DEC_LN, T_DEC_FFN1, true, e.dx, D, w.ln3w, w.ln3b, R, Lin{e.dn, D, e.dn16, w.f1w, w.f1w16, w.f1b, nullptr, 0, tc ? nullptr : e.dffn, F, tc ? e.dffn16 : nullptr, R, F, D, 1, nr}, st));
      {
        Lin f2{e.dffn, F, tc ? e.dffn16 : nullptr, w.f2w, w.f2w16, w.f2b, e.dx, D, e.dx, D, nullptr, R, D, F, 0, nr};
        if (fl) f2 = (l + 1 < c.dec_layers) ? with_ln(f2, e.dec[l + 1].ln1w, e.dec[l + 1].ln1b, e.dn16) : with_ln(f2, e.daw, e.dab, e.dn16);
        PD(T_DEC_FFN2, linear(e, f2, st));
      }
    }
    e.launches += fl ? 2 : 5;
  }
  TRY(ln_linear(e, T_DEC_LN, T_DEC_OUT, true, e.dx, D, e.daw, e.dab, R, Lin{e.dn, D, e.dn16, e.doutw, e.doutw16, e.doutb, nullptr, 0, e.dlogp, V, nullptr, R, V, D, 0, nr}, st));
  TRY(decode_tail(e, st));
  if (ptot) prof_mark(e, T_DEC_STEP_TOTAL, st, false);
  return 0;
}

// log-softmax + pre-beam, CTC prefix scores, score combination, pruning, state update, commit (shared by all modes)
static int decode_tail(Engine& e, cudaStream_t st) {
  const SearchBuffers& sb = e.sb;
  PD(T_PREBEAM, launch_logsoftmax_prebeam(sb, e.dlogp, st));
  PD(T_CTC_PREFIX, launch_ctc_prefix(sb, st));
  PD(T_COMBINE, launch_combine_topk(sb, e.dlogp, st));
  if (e.trace_buf && e.trace_used < e.trace_max) TRY(trace_step(e, st));
  PD(T_PRUNE, launch_beam_prune(sb, st));
  PD(T_CTC_UPDATE, launch_ctc_state_update(sb, st));
  PD(T_STEP_FINISH, launch_step_finish(sb, st));
  e.launches += 9;
  return 0;
}

// ---------------------------------------------------------------- CUDA-graph replay of the two launch-heavy parts
static void drop_graphs(Engine& e) {
  if (e.step_graph) { cudaGraphExecDestroy(e.step_graph); e.step_graph = nullptr; }
  for (auto& kv : e.enc_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e.enc_graphs.clear();
  e.graph_failed = false; e.graph_warm = 0;
}

// Captures fn(st) into an executable graph.  Returns 0 and *exec == nullptr when capture is not possible here (the
// caller then launches plainly); a non-zero return is a real launch error.
template <typename Fn>
static int capture_graph(Engine& e, cudaStream_t st, Fn fn, cudaGraphExec_t* exec, int* n_launches) {
  *exec = nullptr;
  const int before = e.launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError(); e.graph_failed = true; return 0;
  }
  const int rc = fn(st);
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &g);
  *n_launches = e.launches - before;
  e.launches = before;                                    // captured, not launched
  if (rc != 0 || ce != cudaSuccess || !g) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    e.graph_failed = true;
    return 0;
  }
  cudaGraphExec_t x = nullptr;
  if (cudaGraphInstantiate(&x, g, 0) != cudaSuccess || !x) { cudaGetLastError(); e.graph_failed = true; x = nullptr; }
  cudaGraphDestroy(g);
  *exec = x;
  return 0;
}

static int decode_step(Engine& e, cudaStream_t st) {
  if (!e.graph_decode || e.graph_failed || e.prof_tag != 0 || (e.trace_buf && e.trace_used < e.trace_max)) return run_decode_step(e, st);
  if (!e.step_graph) {
    if (e.graph_warm < 2) { e.graph_warm++; return run_decode_step(e, st); }
    TRY(capture_graph(e, st, [&](cudaStream_t s) { return run_decode_step(e, s); }, &e.step_graph, &e.step_graph_launches));
    if (!e.step_graph) return run_decode_step(e, st);
  }
  SCB_CUDA_CHECK(cudaGraphLaunch(e.step_graph, st));
  e.graphs_replayed++;
  e.launches += e.step_graph_launches;
  e.step_seq++;
  return 0;
}

static int encoder_layers(Engine& e, int n_blk, cudaStream_t st) {
  if (!e.graph_encoder || e.graph_failed || e.prof_tag != 0) return run_encoder_layers(e, n_blk, st);
  auto it = e.enc_graphs.find(n_blk);
  if (it == e.enc_graphs.end()) {
    // a block count is captured the second time it is seen (steady-state pushes repeat a handful of counts)
    e.enc_graphs[n_blk] = Engine::EncGraph{nullptr, 0, 1};
    return run_encoder_layers(e, n_blk, st);
  }
  Engine::EncGraph& g = it->second;
  if (!g.exec) {
    if (e.enc_graphs.size() > 64) return run_encoder_layers(e, n_blk, st);      // bounded cache
    TRY(capture_graph(e, st, [&](cudaStream_t s) { return run_encoder_layers(e, n_blk, s); }, &g.exec, &g.launches));
    if (!g.exec) return run_encoder_layers(e, n_blk, st);
  }
  SCB_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
  e.graphs_replayed++;
  e.launches += g.launches;
  g.uses++;
  return 0;
}

}  // namespace scb

using namespace scb;

// ================================================================== C ABI
extern "C" {

const char* sc_version(void) { return "speechcatcher_b200 0.1 (sm_100a)"; }
const char* sc_last_error(void) { return get_last_error(); }

static int check_cfg(const ScConfig* c) {
  if (!c) { set_last_error("null config"); return SC_ERR_ARG; }
  if (c->d_model != 256 || c->vocab != 1024 || c->ffn % 16 != 0) {
    set_last_error("unsupported architecture: d_model=%d vocab=%d ffn=%d (kernels are specialised for 256/1024)",
                   c->d_model, c->vocab, c->ffn);
    return SC_ERR_ARG;
  }
  if (c->precision < 0 || c->precision > 2) { set_last_error("precision %d out of range [0, 2]", c->precision); return SC_ERR_ARG; }
  if (c->beam < 1 || c->beam > 20) { set_last_error("beam %d out of range [1, 20]", c->beam); return SC_ERR_ARG; }
  if (c->n_streams < 1 || c->max_chunk < 1 || c->max_frames < 32) { set_last_error("bad capacities"); return SC_ERR_ARG; }
  // the positional table holds 5000 positions like the reference's (positional_encoding.py:31,72): an un-reset
  // stream ends there (200 s); the encoder indexes up to max_frames + one block of look-ahead frames
  if (c->max_frames + 64 > 5000) {
    set_last_error("max_frames %d exceeds the 5000-position table of the model (at most 4936 frames = 197 s per stream)", c->max_frames);
    return SC_ERR_ARG;
  }
  int dk_e = c->d_model / c->enc_heads, dk_d = c->d_model / c->dec_heads;
  if ((dk_e != 32 && dk_e != 64) || (dk_d != 32 && dk_d != 64)) { set_last_error("head dim must be 32 or 64"); return SC_ERR_ARG; }
  return SC_OK;
}

int sc_engine_workspace_bytes(const ScConfig* cfg, size_t* bytes) {
  int r = check_cfg(cfg); if (r) return r;
  Engine e(*cfg);
  Carver cv(nullptr, true);
  carve(e, cv);
  *bytes = align_up(cv.off) + 256;
  return SC_OK;
}

int sc_engine_create(const ScConfig* cfg, void* workspace, size_t bytes, void** handle) {
  int r = check_cfg(cfg); if (r) return r;
  size_t need = 0; sc_engine_workspace_bytes(cfg, &need);
  if (!workspace || bytes < need || ((uintptr_t)workspace & 255)) {
    set_last_error("workspace too small or misaligned: have %zu need %zu", bytes, need);
    return SC_ERR_ARG;
  }
  Engine* e = new Engine(*cfg);
  Carver cv(workspace, false);
  carve(*e, cv);
  const Caps& k = e->cap;
  const size_t S = cfg->n_streams;
  e->h_stage_bytes = sizeof(FrontendDesc) * S + sizeof(SubDesc) * S + sizeof(BlockDesc) * k.nb_max + sizeof(int) * (6 * S) +
                     (sizeof(int64_t) * 3 + sizeof(int)) * S * k.sub_cap + sizeof(int) * (2 * S + 2 * S * k.qpush) + sizeof(int) * S + 4096;
  if (cudaMallocHost(&e->h_stage, e->h_stage_bytes) != cudaSuccess || cudaMallocHost(&e->h_flag, 4 * sizeof(int)) != cudaSuccess) {
    set_last_error("pinned host allocation failed"); delete e; return SC_ERR_CUDA;
  }
  cudaEventCreateWithFlags(&e->ev[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev[1], cudaEventDisableTiming);
  {
    const char* es = getenv("SCB_ENC_SMS");
    make_encoder_stream(*e, es ? atoi(es) : 0);
  }
  cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_wave, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_enc, cudaEventDisableTiming);
  e->enc.resize(cfg->enc_layers); e->dec.resize(cfg->dec_layers);
  {
    const char* a = getenv("SCB_ATTN");     // "simt" forces the CUDA-core attention kernels in the bf16 mode (A/B tests)
    // beam 17..32 takes the tiled (WIDE) tensor-core kernel only on request ("mma_wide"): not yet run on a device
    const int mma_beam = (a && strcmp(a, "mma_wide") == 0) ? 32 : 16;
    e->mma_attn = cfg->precision == 1 && cfg->beam <= mma_beam && !(a && strcmp(a, "simt") == 0);
    e->mma_enc = cfg->precision == 1 && !(a && strcmp(a, "simt") == 0);
    e->attn_f32_rows = !(a && strcmp(a, "cta") == 0);
    e->enc_attn_x3 = !(a && (strcmp(a, "rows") == 0 || strcmp(a, "cta") == 0));
    if (const char* xd = getenv("SCB_X3_DEC_PERSIST")) e->x3_dec_persist = atoi(xd);
    if (const char* xc = getenv("SCB_X3_CHAIN")) e->x3_chain = atoi(xc) != 0;
    if (const char* xl = getenv("SCB_X3_LN_FUSED")) {
      e->x3_ln_fused_enc = strcmp(xl, "1") == 0 || strcmp(xl, "enc") == 0;
      e->x3_ln_fused_dec = strcmp(xl, "1") == 0 || strcmp(xl, "dec") == 0;
    }
    const char* lp = getenv("SCB_LN_PROLOGUE");
    // measured slower than a separate LayerNorm kernel (every N-tile CTA re-normalises its 128 rows): opt-in only
    e->ln_prologue = cfg->precision == 1 && cfg->d_model == 256 && lp && strcmp(lp, "1") == 0;
    const char* fln = getenv("SCB_FUSE_LN");   // "enc" / "dec" / "1": LayerNorm in the epilogue of the producing GEMM
    if (fln && cfg->precision == 1) {
      e->fuse_ln = strcmp(fln, "1") == 0 || strcmp(fln, "enc") == 0;
      e->fuse_ln_dec = strcmp(fln, "1") == 0 || strcmp(fln, "dec") == 0;
    }
    const char* ff = getenv("SCB_FUSED_FFN");   // "0" keeps the two-GEMM encoder FFN (A/B tests)
    e->fused_ffn = cfg->precision == 1 && cfg->d_model == 256 && cfg->ffn % 128 == 0 && !(ff && strcmp(ff, "0") == 0);
    const char* ffd = getenv("SCB_FUSED_FFN_DEC");   // "0" keeps LayerNorm -> FFN1 -> FFN2 in the decode step
    e->fused_ffn_dec = e->fused_ffn && !(ffd && strcmp(ffd, "0") == 0);
    const char* fsp = getenv("SCB_FFN_SPLITS");
    if (fsp) e->ffn_splits = atoi(fsp);
    const char* gr = getenv("SCB_GRAPH");    // bit 0: replay the search iteration as a CUDA graph, bit 1: the encoder stack
    if (gr) { e->graph_decode = (atoi(gr) & 1) != 0; e->graph_encoder = (atoi(gr) & 2) != 0; }
    const char* lpd = getenv("SCB_LN_PROLOGUE_DEC");
    e->ln_prologue_dec = cfg->precision == 1 && cfg->d_model == 256 && lpd && strcmp(lpd, "1") == 0;
    const char* pdl = getenv("SCB_PDL");     // programmatic dependent launch of the decode-step kernel chain (default on)
    g_use_pdl = !(pdl && strcmp(pdl, "0") == 0);
  }
  *handle = e;
  return SC_OK;
}

int sc_engine_destroy(void* handle) {
  Engine* e = (Engine*)handle;
  if (!e) return SC_OK;
  drop_graphs(*e);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  if (e->h_flag) cudaFreeHost(e->h_flag);
  cudaEventDestroy(e->ev[0]); cudaEventDestroy(e->ev[1]);
  if (e->st_enc) { cudaStreamSynchronize(e->st_enc); cudaStreamDestroy(e->st_enc); }
  destroy_green_ctx(*e);
  if (e->ev_in) { cudaEventDestroy(e->ev_in); cudaEventDestroy(e->ev_wave); cudaEventDestroy(e->ev_enc); }
  delete e;
  return SC_OK;
}

int sc_engine_set_weight(void* handle, const char* name, const void* dev_ptr, size_t /*n_elem*/) {
  Engine* e = (Engine*)handle;
  if (!e || !name || !dev_ptr) { set_last_error("set_weight: null argument"); return SC_ERR_ARG; }
  e->wmap[name] = dev_ptr;
  return SC_OK;
}

int sc_engine_set_frontend(void* handle, const float* window400, const float* mel_fb, const double* mean, const double* std_) {
  Engine* e = (Engine*)handle;
  if (!e || !window400 || !mel_fb) { set_last_error("set_frontend: null argument"); return SC_ERR_ARG; }
  if (frontend_upload_tables(e->d_fe_tab, window400, mel_fb)) return SC_ERR_CUDA;
  e->fe_tab_ready = true;
  e->has_stats = mean && std_;
  if (e->has_stats) {
    SCB_CUDA_CHECK(cudaMemcpy(e->d_mean, mean, 80 * sizeof(double), cudaMemcpyHostToDevice));
    SCB_CUDA_CHECK(cudaMemcpy(e->d_std, std_, 80 * sizeof(double), cudaMemcpyHostToDevice));
  }
  return SC_OK;
}

int sc_engine_finalize(void* handle) {
  Engine* e = (Engine*)handle;
  if (!e) return SC_ERR_ARG;
  std::string missing;
  const bool tc = e->cfg.precision == 1;
  auto getf = [&](const std::string& n) -> const float* {
    auto it = e->wmap.find(n);
    if (it == e->wmap.end()) { missing += n + " "; return nullptr; }
    return (const float*)it->second;
  };
  auto geth = [&](const std::string& n) -> const __nv_bfloat16* {
    if (!tc) return nullptr;
    auto it = e->wmap.find(n + ".bf16");
    if (it == e->wmap.end()) { missing += n + ".bf16 "; return nullptr; }
    return (const __nv_bfloat16*)it->second;
  };
  e->pe = getf("pe");
  e->c1w = getf("enc.conv1.w"); e->c1b = getf("enc.conv1.b");
  e->c2w = getf("enc.conv2.w"); e->c2b = getf("enc.conv2.b");
  e->eow = getf("enc.out.w"); e->eob = getf("enc.out.b");
  e->eaw = getf("enc.after.w"); e->eab = getf("enc.after.b");
  e->ctcw = getf("ctc.w"); e->ctcb = getf("ctc.b");
  e->demb = getf("dec.emb"); e->daw = getf("dec.after.w"); e->dab = getf("dec.after.b");
  e->doutw = getf("dec.out.w"); e->doutb = getf("dec.out.b");
  e->eow16 = geth("enc.out.w"); e->ctcw16 = geth("ctc.w"); e->doutw16 = geth("dec.out.w"); e->c2w16 = geth("enc.conv2.w");
  for (int l = 0; l < e->cfg.enc_layers; ++l) {
    std::string p = "enc." + std::to_string(l) + ".";
    EncLayerW& w = e->enc[l];
    w.ln1w = getf(p + "ln1.w"); w.ln1b = getf(p + "ln1.b"); w.qkvw = getf(p + "qkv.w"); w.qkvb = getf(p + "qkv.b");
    w.ow = getf(p + "o.w"); w.ob = getf(p + "o.b"); w.ln2w = getf(p + "ln2.w"); w.ln2b = getf(p + "ln2.b");
    w.f1w = getf(p + "ff1.w"); w.f1b = getf(p + "ff1.b"); w.f2w = getf(p + "ff2.w"); w.f2b = getf(p + "ff2.b");
    w.qkvw16 = geth(p + "qkv.w"); w.ow16 = geth(p + "o.w"); w.f1w16 = geth(p + "ff1.w"); w.f2w16 = geth(p + "ff2.w");
  }
  for (int l = 0; l < e->cfg.dec_layers; ++l) {
    std::string p = "dec." + std::to_string(l) + ".";
    DecLayerW& w = e->dec[l];
    w.ln1w = getf(p + "ln1.w"); w.ln1b = getf(p + "ln1.b"); w.sqkvw = getf(p + "self_qkv.w"); w.sqkvb = getf(p + "self_qkv.b");
    w.sow = getf(p + "self_o.w"); w.sob = getf(p + "self_o.b"); w.ln2w = getf(p + "ln2.w"); w.ln2b = getf(p + "ln2.b");
    w.cqw = getf(p + "src_q.w"); w.cqb = getf(p + "src_q.b"); w.ckvw = getf(p + "src_kv.w"); w.ckvb = getf(p + "src_kv.b");
    w.cow = getf(p + "src_o.w"); w.cob = getf(p + "src_o.b"); w.ln3w = getf(p + "ln3.w"); w.ln3b = getf(p + "ln3.b");
    w.f1w = getf(p + "ff1.w"); w.f1b = getf(p + "ff1.b"); w.f2w = getf(p + "ff2.w"); w.f2b = getf(p + "ff2.b");
    w.sqkvw16 = geth(p + "self_qkv.w"); w.sow16 = geth(p + "self_o.w"); w.cqw16 = geth(p + "src_q.w");
    w.ckvw16 = geth(p + "src_kv.w"); w.cow16 = geth(p + "src_o.w"); w.f1w16 = geth(p + "ff1.w"); w.f2w16 = geth(p + "ff2.w");
  }
  if (e->cfg.precision == 2) {
    // every GEMM weight needs its fp16 hi / lo planes ("<name>.x3", weights.split_f16)
    e->x3map.clear();
    for (auto& kv : e->wmap) {
      const std::string& n = kv.first;
      if (n.size() > 3 && n.compare(n.size() - 3, 3, ".x3") == 0) {
        auto base = e->wmap.find(n.substr(0, n.size() - 3));
        if (base != e->wmap.end()) e->x3map[(const float*)base->second] = kv.second;
      }
    }
    auto need = [&](const float* w, const std::string& n) { if (w && !e->x3map.count(w)) missing += n + ".x3 "; };
    need(e->c2w, "enc.conv2.w"); need(e->eow, "enc.out.w"); need(e->ctcw, "ctc.w"); need(e->doutw, "dec.out.w");
    for (int l = 0; l < e->cfg.enc_layers; ++l) {
      const EncLayerW& w = e->enc[l]; const std::string p = "enc." + std::to_string(l) + ".";
      need(w.qkvw, p + "qkv.w"); need(w.ow, p + "o.w"); need(w.f1w, p + "ff1.w"); need(w.f2w, p + "ff2.w");
    }
    for (int l = 0; l < e->cfg.dec_layers; ++l) {
      const DecLayerW& w = e->dec[l]; const std::string p = "dec." + std::to_string(l) + ".";
      need(w.sqkvw, p + "self_qkv.w"); need(w.sow, p + "self_o.w"); need(w.cqw, p + "src_q.w"); need(w.ckvw, p + "src_kv.w");
      need(w.cow, p + "src_o.w"); need(w.f1w, p + "ff1.w"); need(w.f2w, p + "ff2.w");
    }
  }
  if (!missing.empty()) { set_last_error("missing weights: %.400s", missing.c_str()); return SC_ERR_STATE; }
  // conv2 implicit-GEMM segment offsets: K = (kt, kf, c) -> ((kt * 39) + kf) * D
  int seg[9];
  for (int kt = 0; kt < 3; ++kt) for (int kf = 0; kf < 3; ++kf) seg[kt * 3 + kf] = (kt * 39 + kf) * e->cfg.d_model;
  SCB_CUDA_CHECK(cudaMemcpy(e->d_c2_seg, seg, sizeof(seg), cudaMemcpyHostToDevice));
  // all streams start reset
  std::vector<int> all(e->cfg.n_streams);
  for (int i = 0; i < e->cfg.n_streams; ++i) all[i] = i;
  SCB_CUDA_CHECK(cudaMemcpy(e->d_reset, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (launch_search_reset(e->sb, e->d_reset, e->cfg.n_streams, 0)) return SC_ERR_CUDA;
  if (build_decode_chains(*e)) return SC_ERR_CUDA;
  SCB_CUDA_CHECK(cudaDeviceSynchronize());
  e->finalized = true;
  return SC_OK;
}

int sc_engine_reset(void* handle, const int32_t* streams, int32_t n, void* stream) {
  Engine* e = (Engine*)handle;
  if (!e || !e->finalized) { set_last_error("engine not finalized"); return SC_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (e->enc_pending) { SCB_CUDA_CHECK(cudaEventSynchronize(e->ev_enc)); e->enc_pending = false; }
  for (int i = 0; i < n; ++i) {
    if (streams[i] < 0 || streams[i] >= e->cfg.n_streams) { set_last_error("bad stream id %d", streams[i]); return SC_ERR_ARG; }
    e->planner.reset(streams[i]);
    e->pending_bound[streams[i]] = 0;
  }
  int* hs = (int*)e->h_stage;
  memcpy(hs, streams, n * sizeof(int));
  SCB_CUDA_CHECK(cudaMemcpyAsync(e->d_reset, hs, n * sizeof(int), cudaMemcpyHostToDevice, st));
  if (launch_search_reset(e->sb, e->d_reset, n, st)) return SC_ERR_CUDA;
  SCB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SC_OK;
}

// One push of n streams.  feats_dev == nullptr: waveform chunks (n_samples = samples per stream).  feats_dev != nullptr:
// pre-computed, already normalised feature frames [i][n_samples[i]][80] with row pitch ld_wave floats per stream
// (speech2text_streaming.py:438-450): the frontend is skipped and the frames are copied behind each stream's carried ones.
static int push_impl(void* handle, const float* wave_dev, int32_t ld_wave, const float* feats_dev, const int32_t* streams,
                     const int32_t* n_samples, const int32_t* is_final, int32_t n, void* stream, ScPushStats* stats) {
  Engine* e = (Engine*)handle;
  if (!e || !e->finalized) { set_last_error("engine not finalized"); return SC_ERR_STATE; }
  const ScConfig& c = e->cfg; const Caps& k = e->cap;
  cudaStream_t sd = (cudaStream_t)stream;          // caller's stream: the search runs here
  const bool use_enc_stream = e->overlap && e->lazy_threshold > 0;
  cudaStream_t st = use_enc_stream ? e->st_enc : sd;  // frontend + encoder part of this push
  const int S = c.n_streams, D = c.d_model, V = c.vocab;
  e->launches = 0;
  Engine& eref = *e;
  // the pinned descriptor staging is reused: the previous push's uploads (on the encoder stream) must be done
  if (e->enc_pending) { SCB_CUDA_CHECK(cudaEventSynchronize(e->ev_enc)); e->enc_pending = false; }
  if (use_enc_stream) {                             // the encoder stream starts after the caller's H2D of the waveforms
    SCB_CUDA_CHECK(cudaEventRecord(e->ev_in, sd));
    SCB_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_in, 0));
  }
  if (n < 0 || n > S) { set_last_error("push: n=%d out of range", n); return SC_ERR_ARG; }
  // ---------------- plan on the host
  // The planner advances the host-side stream state; nothing reaches the device before every stream of the push has
  // been validated, so any error up to the descriptor upload rolls the planner (and last_plan) back: a failed push
  // leaves every stream exactly where it was.
  std::vector<StreamPush> plans(n);
  struct Rollback {
    Engine* e; std::vector<std::pair<int, StreamHost>> saved; std::vector<ScStreamPlan> plans_before; bool armed = true;
    ~Rollback() {
      if (!armed) return;
      for (size_t i = saved.size(); i-- > 0;) { e->planner.restore(saved[i].first, saved[i].second); e->last_plan[saved[i].first] = plans_before[i]; }
    }
  } rollback{e, {}, {}};
  rollback.saved.reserve(n); rollback.plans_before.reserve(n);
  {
    std::vector<char> seen(S, 0);
    for (int i = 0; i < n; ++i) {
      const int s = streams[i];
      if (s < 0 || s >= S) { set_last_error("bad stream id %d", s); return SC_ERR_ARG; }
      if (seen[s]) { set_last_error("stream %d listed twice in one push", s); return SC_ERR_ARG; }
      seen[s] = 1;
    }
  }
  for (int i = 0; i < n; ++i) {
    const int s = streams[i];
    if (feats_dev) {
      // every capacity downstream (conv rows, sub-sampled frames, blocks) is sized for 12 carried + fmax new frames,
      // fmax = STFT frames of the largest waveform slab = feat_cap - 16 (make_caps)
      if (n_samples[i] < 0 || n_samples[i] > k.feat_cap - 16) {
        set_last_error("stream %d: %d feature frames exceed the per-push capacity %d (raise max_chunk)", s, n_samples[i],
                       k.feat_cap - 16);
        return SC_ERR_CAPACITY;
      }
      if ((int64_t)n_samples[i] * 80 > ld_wave) { set_last_error("feature row pitch too small"); return SC_ERR_ARG; }
    } else {
      if (n_samples[i] < 0 || n_samples[i] > c.max_chunk) {
        set_last_error("stream %d: chunk of %d samples exceeds max_chunk %d", s, n_samples[i], c.max_chunk);
        return SC_ERR_CAPACITY;
      }
      if (n_samples[i] > ld_wave) { set_last_error("ld_wave too small"); return SC_ERR_ARG; }
    }
    rollback.saved.emplace_back(s, e->planner.state(s));
    rollback.plans_before.push_back(e->last_plan[s]);
    plans[i] = feats_dev ? e->planner.push_features(s, n_samples[i], is_final[i] != 0)
                         : e->planner.push(s, n_samples[i], is_final[i] != 0);
    StreamPush& p = plans[i];
    if (p.error) {
      set_last_error("stream %d: input shape the reference cannot process (code %d)", s, p.error);
      return SC_ERR_STATE;
    }
    if (e->planner.state(s).enc_len > k.Tcap) {
      set_last_error("stream %d: encoder buffer overflow (%d > max_frames %d)", s, e->planner.state(s).enc_len, k.Tcap);
      return SC_ERR_CAPACITY;
    }
    ScStreamPlan& lp = e->last_plan[s];
    lp.called = p.called; lp.n_feat = p.n_feat; lp.n_sub = p.run_sub ? p.sd.t2 : 0; lp.n_blocks = (int)p.blocks.size();
    lp.n_enc_out = p.n_enc_out; lp.enc_len = e->planner.state(s).enc_len; lp.n_decode_blocks = (int)p.dq_T.size();
    lp.last_T = p.dq_T.empty() ? 0 : p.dq_T.back();
  }
  // ---------------- flatten into pinned staging
  unsigned char* hp = e->h_stage;
  auto stage = [&](size_t bytes) { unsigned char* r = hp; hp += align_up(bytes, 16); return r; };
  FrontendDesc* h_fd = (FrontendDesc*)stage(sizeof(FrontendDesc) * S);
  SubDesc* h_sd = (SubDesc*)stage(sizeof(SubDesc) * S);
  BlockDesc* h_blk = (BlockDesc*)stage(sizeof(BlockDesc) * k.nb_max);
  int* h_cf = (int*)stage(sizeof(int) * 3 * S);
  int* h_cs = (int*)stage(sizeof(int) * 3 * S);
  int64_t* h_en_a = (int64_t*)stage(sizeof(int64_t) * S * k.sub_cap);
  int64_t* h_en_ctc = (int64_t*)stage(sizeof(int64_t) * S * k.sub_cap);
  int64_t* h_en_kv = (int64_t*)stage(sizeof(int64_t) * S * k.sub_cap);
  int* h_en_flag = (int*)stage(sizeof(int) * S * k.sub_cap);
  int* h_q = (int*)stage(sizeof(int) * (2 * S + 2 * S * k.qpush));
  int n_fd = 0, n_fd_all = 0, n_sd = 0, n_blk = 0, n_cf = 0, n_cs = 0, n_en = 0, n_q = 0, frame_base = 0, sub_rows = 0;
  int n_feat_total = 0;
  bool need_drain = false, any_final = false;
  // frontend descriptors: emitting ones first (frame_base indexing), buffer-only ones after
  std::vector<FrontendDesc> buf_only;
  for (int i = 0; i < n; ++i) {
    StreamPush& p = plans[i];
    if (!p.has_fd) continue;
    if (p.fd.emit1 > p.fd.emit0) { p.fd.frame_base = frame_base; frame_base += p.fd.n_frames; h_fd[n_fd++] = p.fd; n_feat_total += p.n_feat; }
    else buf_only.push_back(p.fd);
  }
  n_fd_all = n_fd;
  for (auto& f : buf_only) { f.frame_base = frame_base; h_fd[n_fd_all++] = f; }
  for (int i = 0; i < n; ++i) {
    StreamPush& p = plans[i];
    const int s = streams[i];
    if (is_final[i]) any_final = true;
    if (p.run_sub) {
      p.sd.row0 = sub_rows; sub_rows += p.sd.t2;
      h_sd[n_sd++] = p.sd;
    }
    if (p.feat_carry_move) { h_cf[n_cf] = s; h_cf[S + n_cf] = p.sd.carry_src; h_cf[2 * S + n_cf] = p.sd.carry_n; n_cf++; }
    if (p.sub_carry_move) { h_cs[n_cs] = s; h_cs[S + n_cs] = p.sub_carry_src; h_cs[2 * S + n_cs] = p.sub_carry_n; n_cs++; }
    const int blk0 = n_blk;
    for (auto b : p.blocks) {
      if ((int)n_blk >= k.nb_max) { set_last_error("block capacity exceeded"); return SC_ERR_CAPACITY; }
      if (b.prev_blk >= 0) b.prev_blk += blk0;
      b.out_row0 = n_en + (b.out_t0 - p.enc_t0);
      h_blk[n_blk++] = b;
    }
    for (int t = 0; t < p.n_enc_out; ++t) {
      const int tt = p.enc_t0 + t;
      h_en_a[n_en] = ((int64_t)s * k.Tcap + tt) * D;
      h_en_ctc[n_en] = ((int64_t)s * k.Tcap + tt) * V;
      h_en_kv[n_en] = ((int64_t)s * k.Tcap + tt) * 2 * D;
      h_en_flag[n_en] = tt < (kBlock - kLook) ? 1 : 0;      // rows of the first decode block are log-softmaxed (Q1)
      n_en++;
    }
    if (!p.dq_T.empty()) {
      if ((int)p.dq_T.size() > k.qpush) { set_last_error("decode queue capacity exceeded"); return SC_ERR_CAPACITY; }
      h_q[n_q] = s; h_q[S + n_q] = (int)p.dq_T.size();
      for (size_t j = 0; j < p.dq_T.size(); ++j) {
        h_q[2 * S + n_q * k.qpush + j] = p.dq_T[j];
        h_q[2 * S + S * k.qpush + n_q * k.qpush + j] = p.dq_final[j];
      }
      if (e->pending_bound[s] + (int)p.dq_T.size() > k.qcap) need_drain = true;
      n_q++;
    }
  }
  rollback.armed = false;      // validated: from here on device work is enqueued and the host state stands
  // ---------------- upload descriptors
  auto up = [&](void* dst, const void* src, size_t bytes) -> int {
    if (bytes == 0) return 0;
    SCB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  };
  TRY(up(e->d_fd, h_fd, sizeof(FrontendDesc) * n_fd_all));
  TRY(up(e->d_sd, h_sd, sizeof(SubDesc) * n_sd));
  TRY(up(e->d_blk, h_blk, sizeof(BlockDesc) * n_blk));
  if (n_cf) TRY(up(e->d_carry_f, h_cf, sizeof(int) * 3 * S));
  if (n_cs) TRY(up(e->d_carry_s, h_cs, sizeof(int) * 3 * S));
  TRY(up(e->d_en_a, h_en_a, sizeof(int64_t) * n_en));
  TRY(up(e->d_en_ctc, h_en_ctc, sizeof(int64_t) * n_en));
  TRY(up(e->d_en_kv, h_en_kv, sizeof(int64_t) * n_en));
  TRY(up(e->d_en_flag, h_en_flag, sizeof(int) * n_en));
  if (n_q) TRY(up(e->d_q, h_q, sizeof(int) * (2 * S + 2 * S * k.qpush)));
  // ---------------- frontend
#define e eref
  const bool petot = prof_on(e, T_ENC_TOTAL, false);
  if (petot) prof_mark(e, T_ENC_TOTAL, st, true);
  if (!feats_dev) {
    PE(T_FRONTEND, launch_frontend(e.fe_tab_ready ? e.d_fe_tab : nullptr, wave_dev, ld_wave, e.wbuf, 512, e.d_fd, n_fd, frame_base, e.has_stats ? e.d_mean : nullptr,
                        e.has_stats ? e.d_std : nullptr, e.featbuf, k.feat_cap, st));
  }
#undef e
  if (!feats_dev) {
    TRY(launch_wavebuf_update(wave_dev, ld_wave, e->wbuf, 512, e->d_fd, n_fd_all, st));
    e->launches += 2;
  } else {
    for (int i = 0; i < n; ++i) {                       // feature frames go behind the stream's carried frames
      const StreamPush& p = plans[i];
      if (p.n_feat <= 0) continue;
      n_feat_total += p.n_feat;
      SCB_CUDA_CHECK(cudaMemcpyAsync(e->featbuf + ((size_t)streams[i] * k.feat_cap + p.fd.feat_off) * 80,
                                     feats_dev + (size_t)i * ld_wave, sizeof(float) * 80 * (size_t)p.n_feat,
                                     cudaMemcpyDeviceToDevice, st));
    }
  }
  if (use_enc_stream) {                             // later work on the caller's stream (e.g. the next H2D into wave_dev)
    SCB_CUDA_CHECK(cudaEventRecord(e->ev_wave, st));   // is ordered after the waveforms have been consumed
    SCB_CUDA_CHECK(cudaStreamWaitEvent(sd, e->ev_wave, 0));
  }
  // ---------------- conv2d sub-sampling
  const bool tc = c.precision == 1;
  if (n_sd > 0) {
#define e eref
    PE(T_SUBOUT, launch_conv2_rows(e.d_sd, n_sd, k.t1_cap, k.sub_cap, D, e.d_c2_a, e.d_c2_c, st));
    if (tc) {
      // conv1 (+ReLU) fused with im2col -> bf16 A operand; conv2 and the output projection on tensor cores
      PE(T_CONV1, launch_conv1_im2col_bf16(e.featbuf, k.feat_cap, e.c1w, e.c1b, e.im2col16, k.t2_cap, e.d_sd, n_sd, D, st));
      if (e.prof_tag == T_CONV2) e.prof_flops += 2.0 * sub_rows * 19 * (double)D * 9 * D;
      Lin g{nullptr, 9 * D, e.im2col16, e.c2w, e.c2w16, e.c2b, nullptr, 0, nullptr, D, e.h2_16, sub_rows * 19, D, 9 * D, 1, nullptr};
      PE(T_CONV2, linear(e, g, st));
    } else {
      PE(T_CONV1, launch_conv1(e.featbuf, k.feat_cap, e.c1w, e.c1b, e.h1, k.t1_cap, e.d_sd, n_sd, D, st));
      GemmArgs g;
      g.A = e.h1; g.a_row_off = e.d_c2_a; g.a_seg_off = e.d_c2_seg; g.seg_len = D; g.W = e.c2w; g.bias = e.c2b;
      g.C = e.h2; g.ldc = D; g.M = sub_rows * 19; g.N = D; g.K = 9 * D; g.relu = 1;
      if (e.prof_tag == T_CONV2) e.prof_flops += 2.0 * g.M * (double)g.N * g.K;
      X3Extra x;
      if (c.precision == 2) {      // the conv2 output only feeds embed.out: written as split planes [sub_rows * 19][D]
        g.C = nullptr; x.C2 = e.h2; x.c2_plane = (size_t)k.sub_rows_max * 19 * D; x.ldc2 = D;
      }
      PE(T_CONV2, gemm_fp32(e, g, st, x));
    }
    if (c.precision == 2) {
      const Planes h2p{e.h2, (size_t)k.sub_rows_max * 19 * D, k.sub_rows_max};      // the same planes viewed as [sub_rows][19 D]
      PE(T_SUBOUT, x3_linear(e, h2p, 19 * D, e.eow, e.eob, nullptr, e.subbuf, 0, nullptr, sub_rows, D, 0, nullptr, st, e.d_c2_c));
      e.launches--;
    } else {
      Lin o{e.h2, 19 * D, e.h2_16, e.eow, e.eow16, e.eob, nullptr, 0, e.subbuf, 0, nullptr, sub_rows, D, 19 * D, 0, nullptr};
      o.c_row_off = e.d_c2_c;
      PE(T_SUBOUT, linear(e, o, st));
    }
#undef e
    e->launches += 4;
  }
  if (n_cf) { TRY(launch_carry_rows(e->featbuf, k.feat_cap, 80, e->d_carry_f, e->d_carry_f + S, e->d_carry_f + 2 * S, n_cf, st)); e->launches++; }
  // ---------------- encoder blocks
  if (n_blk > 0) {
#define e eref
    PE(T_BLOCK_ASM, launch_block_assemble(e.subbuf, k.sub_cap, e.pe, e.d_blk, n_blk, e.addin, e.prev_addin, e.X, D, st));
    TRY(encoder_layers(e, n_blk, st));
    PE(T_STITCH, launch_stitch_norm(e.X, e.d_blk, n_blk, e.eaw, e.eab, e.encbuf, k.Tcap, D, tc ? e.encnew16 : nullptr, st));
#undef e
    e->launches += 4;
  }
  if (n_cs) { TRY(launch_carry_rows(e->subbuf, k.sub_cap, D, e->d_carry_s, e->d_carry_s + S, e->d_carry_s + 2 * S, n_cs, st)); e->launches++; }
  // ---------------- CTC head + cross-attention K|V for the new encoder frames
  if (n_en > 0) {
    if (tc) {
      Lin g{nullptr, D, e->encnew16, e->ctcw, e->ctcw16, e->ctcb, nullptr, 0, e->ctcx, 0, nullptr, n_en, V, D, 0, nullptr};
      g.c_row_off = e->d_en_ctc;
#define e eref
      PE(T_CTC_HEAD, linear(e, g, st));
#undef e
    } else {
      GemmArgs g;
      g.A = e->encbuf; g.a_row_off = e->d_en_a; g.W = e->ctcw; g.bias = e->ctcb; g.C = e->ctcx; g.c_row_off = e->d_en_ctc;
      g.M = n_en; g.N = V; g.K = D;
#define e eref
      PE(T_CTC_HEAD, gemm_fp32(e, g, st));
#undef e
    }
    TRY(launch_logsoftmax_rows(e->ctcx, e->d_en_ctc, e->d_en_flag, n_en, V, st));
    for (int l = 0; l < c.dec_layers; ++l) {
      float* dst = e->sb.xkv + (size_t)l * S * k.Tcap * 2 * D;
      if (tc) {
        __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(e->sb.xkv) + (size_t)l * S * k.Tcap * 2 * D;
        Lin kv{nullptr, D, e->encnew16, e->dec[l].ckvw, e->dec[l].ckvw16, e->dec[l].ckvb, nullptr, 0, nullptr, 0, dst16, n_en, 2 * D, D, 0, nullptr};
        kv.c_row_off = e->d_en_kv;
#define e eref
        PE(T_XKV, linear(e, kv, st));
#undef e
      } else {
        GemmArgs kv;
        kv.A = e->encbuf; kv.a_row_off = e->d_en_a; kv.W = e->dec[l].ckvw; kv.bias = e->dec[l].ckvb;
        kv.C = dst; kv.c_row_off = e->d_en_kv; kv.M = n_en; kv.N = 2 * D; kv.K = D;
        X3Extra xs;
        if (e->sb.kv_split) {        // cache rows as split planes [hi K|V][lo K|V] (same row pitch as fp32)
          kv.C = nullptr; xs.C2 = dst; xs.c2_plane = 2 * (size_t)D; xs.ldc2 = 2 * D;
        }
#define e eref
        PE(T_XKV, gemm_fp32(e, kv, st, xs));
#undef e
      }
    }
    e->launches += 2 + c.dec_layers;
  }
  if (petot) prof_mark(eref, T_ENC_TOTAL, st, false);
  if (use_enc_stream) { SCB_CUDA_CHECK(cudaEventRecord(e->ev_enc, st)); e->enc_pending = true; }
  // ---------------- block-synchronous beam search (caller's stream)
  int steps = 0;
  // the loop stops when fewer than `stop_below` streams are active: 1 = drain (strict mode, final calls)
  const int stop_below = (e->lazy_threshold > 0 && !any_final) ? e->lazy_threshold : 1;
  auto decode_loop = [&](int below) -> int {
    SCB_CUDA_CHECK(cudaMemcpyAsync(&e->h_flag[0], e->sb.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, sd));
    SCB_CUDA_CHECK(cudaEventRecord(e->ev[0], sd));
    // one step is always in flight ahead of the host's view of n_active (kernels of an empty step exit at once)
    for (int i = 0;; ++i) {
      TRY(decode_step(*e, sd));
      steps++;
      SCB_CUDA_CHECK(cudaMemcpyAsync(&e->h_flag[2 * ((i + 1) & 1)], e->sb.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, sd));
      SCB_CUDA_CHECK(cudaEventRecord(e->ev[(i + 1) & 1], sd));
      SCB_CUDA_CHECK(cudaEventSynchronize(e->ev[i & 1]));
      if (e->h_flag[2 * (i & 1)] < below) break;
      if (steps > 64 * kMaxLength) { set_last_error("decode loop did not terminate"); return SC_ERR_STATE; }
    }
    SCB_CUDA_CHECK(cudaStreamSynchronize(sd));
    if (e->h_flag[1] || e->h_flag[3]) {
      set_last_error("decode capacity exceeded (flag %d): token capacity %d / queue capacity %d", e->h_flag[1] | e->h_flag[3], k.Lcap, k.qcap);
      return SC_ERR_CAPACITY;
    }
    const int last = e->h_flag[0] < e->h_flag[2] ? e->h_flag[0] : e->h_flag[2];
    if (below == 1 || last == 0) std::fill(e->pending_bound.begin(), e->pending_bound.end(), 0);
    return 0;
  };
  auto any_pending = [&]() { for (int v : e->pending_bound) if (v > 0) return true; return false; };
  auto enqueue = [&]() -> int {                      // this push's decode blocks -> device queues (needs the encoder output)
    if (use_enc_stream) SCB_CUDA_CHECK(cudaStreamWaitEvent(sd, e->ev_enc, 0));
    TRY(launch_search_begin(e->sb, e->d_q, e->d_q + S, e->d_q + 2 * S, e->d_q + 2 * S + S * k.qpush, k.qpush, n_q, sd));
    e->launches += 2;
    for (int i = 0; i < n; ++i) e->pending_bound[streams[i]] += (int)plans[i].dq_T.size();
    return 0;
  };
  if (use_enc_stream && !any_final && !need_drain) {
    // overlapped: keep iterating blocks queued by earlier pushes while this push's encoder runs on its own stream,
    // then queue the new blocks (they are decoded during the next push, or by the final drain)
    if (any_pending()) TRY(decode_loop(stop_below));
    TRY(enqueue());
  } else {
    if (need_drain && any_pending()) TRY(decode_loop(1));   // deferred blocks would overflow a device queue: drain first
    TRY(enqueue());
    if (!any_pending()) { SCB_CUDA_CHECK(cudaStreamSynchronize(sd)); }
    else TRY(decode_loop(stop_below));
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_feature_frames = n_feat_total; stats->n_encoder_blocks = n_blk; stats->n_encoder_frames = n_en;
    stats->n_decode_steps = steps; stats->n_kernel_launches = e->launches;
  }
  return SC_OK;
}

int sc_engine_push(void* handle, const float* wave_dev, int32_t ld_wave, const int32_t* streams,
                   const int32_t* n_samples, const int32_t* is_final, int32_t n, void* stream, ScPushStats* stats) {
  return push_impl(handle, wave_dev, ld_wave, nullptr, streams, n_samples, is_final, n, stream, stats);
}

int sc_engine_push_features(void* handle, const float* feats_dev, int32_t ld_feats, const int32_t* streams,
                            const int32_t* n_frames, const int32_t* is_final, int32_t n, void* stream, ScPushStats* stats) {
  if (!feats_dev) { set_last_error("push_features: null feature pointer"); return SC_ERR_ARG; }
  return push_impl(handle, nullptr, ld_feats, feats_dev, streams, n_frames, is_final, n, stream, stats);
}

int sc_engine_read_beam(void* handle, int32_t s, int32_t max_len, int32_t* n_hyp, int32_t* len, int32_t* process_idx,
                        int32_t* yseq, int32_t* xpos, double* score, void* stream) {
  Engine* e = (Engine*)handle;
  if (!e || !e->finalized) { set_last_error("engine not finalized"); return SC_ERR_STATE; }
  if (s < 0 || s >= e->cfg.n_streams) { set_last_error("bad stream id %d", s); return SC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const SearchBuffers& sb = e->sb;
  StreamCtl ctl;
  SCB_CUDA_CHECK(cudaMemcpyAsync(&ctl, sb.ctl + s, sizeof(ctl), cudaMemcpyDeviceToHost, st));
  SCB_CUDA_CHECK(cudaStreamSynchronize(st));
  *n_hyp = ctl.n_hyp; *len = ctl.len; *process_idx = ctl.process_idx;
  if (ctl.len > max_len || ctl.len > sb.Lcap) { set_last_error("read_beam: max_len %d < len %d", max_len, ctl.len); return SC_ERR_ARG; }
  for (int h = 0; h < ctl.n_hyp; ++h) {
    const size_t o = (((size_t)ctl.cur * sb.S + s) * sb.B + h);
    SCB_CUDA_CHECK(cudaMemcpyAsync(yseq + (size_t)h * max_len, sb.yseq + o * sb.Lcap, ctl.len * sizeof(int), cudaMemcpyDeviceToHost, st));
    SCB_CUDA_CHECK(cudaMemcpyAsync(xpos + (size_t)h * max_len, sb.xpos + o * sb.Lcap, ctl.len * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  SCB_CUDA_CHECK(cudaMemcpyAsync(score, sb.score + (((size_t)ctl.cur * sb.S + s) * sb.B), ctl.n_hyp * sizeof(double), cudaMemcpyDeviceToHost, st));
  SCB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SC_OK;
}

int sc_engine_read_all(void* handle, int32_t* ctl16, int32_t* yseq, int32_t* xpos, double* score, void* stream) {
  Engine* e = (Engine*)handle;
  if (!e || !e->finalized) { set_last_error("engine not finalized"); return SC_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  const SearchBuffers& sb = e->sb;
  const size_t S = sb.S, B = sb.B, L = sb.Lcap;
  SCB_CUDA_CHECK(cudaMemcpyAsync(ctl16, sb.ctl, S * sizeof(StreamCtl), cudaMemcpyDeviceToHost, st));
  SCB_CUDA_CHECK(cudaMemcpyAsync(yseq, sb.yseq, 2 * S * B * L * sizeof(int), cudaMemcpyDeviceToHost, st));
  SCB_CUDA_CHECK(cudaMemcpyAsync(xpos, sb.xpos, 2 * S * B * L * sizeof(int), cudaMemcpyDeviceToHost, st));
  SCB_CUDA_CHECK(cudaMemcpyAsync(score, sb.score, 2 * S * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  SCB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SC_OK;
}

int sc_engine_token_capacity(void* handle, int32_t* lcap) {
  Engine* e = (Engine*)handle;
  if (!e || !lcap) { set_last_error("token_capacity: bad argument"); return SC_ERR_ARG; }
  *lcap = e->cap.Lcap;
  return SC_OK;
}

int sc_engine_last_plan(void* handle, int32_t s, ScStreamPlan* plan) {
  Engine* e = (Engine*)handle;
  if (!e || s < 0 || s >= e->cfg.n_streams || !plan) { set_last_error("last_plan: bad argument"); return SC_ERR_ARG; }
  *plan = e->last_plan[s];
  return SC_OK;
}

int sc_engine_buffer(void* handle, const char* name, void** ptr, size_t* n_elem) {
  Engine* e = (Engine*)handle;
  if (!e || !name) return SC_ERR_ARG;
  auto it = e->named.find(name);
  if (it == e->named.end()) { set_last_error("unknown buffer %s", name); return SC_ERR_ARG; }
  *ptr = it->second.first; *n_elem = it->second.second;
  return SC_OK;
}

int sc_engine_trace_layout(void* handle, int64_t* offsets8, int64_t* record_bytes) {
  Engine* e = (Engine*)handle;
  if (!e || !offsets8 || !record_bytes) { set_last_error("trace_layout: null argument"); return SC_ERR_ARG; }
  trace_layout(*e, offsets8, record_bytes);
  return SC_OK;
}

int sc_engine_set_trace(void* handle, void* dev_buf, size_t bytes, int32_t max_steps) {
  Engine* e = (Engine*)handle;
  if (!e) { set_last_error("set_trace: null handle"); return SC_ERR_ARG; }
  int64_t off[8], rec; trace_layout(*e, off, &rec);
  if (dev_buf && (max_steps < 0 || (size_t)rec * (size_t)max_steps > bytes)) { set_last_error("set_trace: buffer too small (%zu < %lld x %d)", bytes, (long long)rec, max_steps); return SC_ERR_ARG; }
  e->trace_buf = (unsigned char*)dev_buf; e->trace_max = dev_buf ? max_steps : 0; e->trace_used = 0;
  return SC_OK;
}

int sc_engine_counter(void* handle, const char* name, int64_t* value) {
  Engine* e = (Engine*)handle;
  if (!e || !name || !value) { set_last_error("counter: null argument"); return SC_ERR_ARG; }
  if (strcmp(name, "graphs_replayed") == 0) { *value = e->graphs_replayed; return SC_OK; }
  if (strcmp(name, "trace_steps") == 0) { *value = e->trace_used; return SC_OK; }
  if (strcmp(name, "graph_failed") == 0) { *value = e->graph_failed ? 1 : 0; return SC_OK; }
  if (strcmp(name, "encoder_sms") == 0) { *value = e->enc_sms; return SC_OK; }
  set_last_error("unknown counter %s", name);
  return SC_ERR_ARG;
}

int sc_engine_set_option(void* handle, const char* name, int32_t value) {
  Engine* e = (Engine*)handle;
  if (!e || !name) { set_last_error("set_option: null argument"); return SC_ERR_ARG; }
  if (strcmp(name, "lazy_threshold") == 0) { e->lazy_threshold = value < 0 ? 0 : value; return SC_OK; }
  if (strcmp(name, "overlap") == 0) { e->overlap = value != 0; return SC_OK; }
  if (strcmp(name, "encoder_sms") == 0) {
    if (e->enc_pending) { cudaEventSynchronize(e->ev_enc); e->enc_pending = false; }
    drop_graphs(*e);
    make_encoder_stream(*e, value);
    return SC_OK;
  }
  drop_graphs(*e);     // everything below changes which kernels a step launches: captured graphs are stale
  if (strcmp(name, "graph_decode") == 0) { e->graph_decode = value != 0; return SC_OK; }
  if (strcmp(name, "graph_encoder") == 0) { e->graph_encoder = value != 0; return SC_OK; }
  if (strcmp(name, "pdl") == 0) { g_use_pdl = value != 0; return SC_OK; }
  if (strcmp(name, "x3_chain") == 0) { e->x3_chain = value != 0; drop_graphs(*e); return SC_OK; }
  if (strcmp(name, "x3_ln_fused_encoder") == 0) { e->x3_ln_fused_enc = value != 0; drop_graphs(*e); return SC_OK; }
  if (strcmp(name, "x3_ln_fused_decoder") == 0) { e->x3_ln_fused_dec = value != 0; drop_graphs(*e); return SC_OK; }
  if (strcmp(name, "ln_prologue") == 0) { e->ln_prologue = value != 0 && e->cfg.precision == 1 && e->cfg.d_model == 256; return SC_OK; }
  if (strcmp(name, "ln_prologue_decoder") == 0) { e->ln_prologue_dec = value != 0 && e->cfg.precision == 1 && e->cfg.d_model == 256; return SC_OK; }
  if (strcmp(name, "fused_ffn") == 0) {
    e->fused_ffn = value != 0 && e->cfg.precision == 1 && e->cfg.d_model == 256 && e->cfg.ffn % 128 == 0;
    return SC_OK;
  }
  if (strcmp(name, "fused_ffn_min_rows") == 0) { e->fused_ffn_min_rows = value; return SC_OK; }
  if (strcmp(name, "fused_ffn_decoder") == 0) { e->fused_ffn_dec = value != 0 && e->fused_ffn; return SC_OK; }
  if (strcmp(name, "ffn_splits") == 0) { e->ffn_splits = value < 0 ? 0 : value; return SC_OK; }
  if (strcmp(name, "fuse_layernorm") == 0) { e->fuse_ln = e->fuse_ln_dec = value != 0 && e->cfg.precision == 1; return SC_OK; }
  if (strcmp(name, "fuse_layernorm_decoder") == 0) { e->fuse_ln_dec = value != 0 && e->cfg.precision == 1; return SC_OK; }
  if (strcmp(name, "attn_f32_rows") == 0) { e->attn_f32_rows = value != 0; return SC_OK; }
  if (strcmp(name, "mma_attention") == 0) {
    e->mma_attn = value != 0 && e->cfg.precision == 1 && e->cfg.beam <= (value >= 2 ? 32 : 16);   // 2: allow the tiled kernel
    e->mma_enc = value != 0 && e->cfg.precision == 1;
    return SC_OK;
  }
  set_last_error("unknown option %s", name);
  return SC_ERR_ARG;
}

// ---------------- live kernel timing
int sc_engine_profile_begin(void* handle, int32_t tag, int32_t max_launches, int32_t stride) {
  Engine* e = (Engine*)handle;
  if (!e || !e->finalized) { set_last_error("engine not finalized"); return SC_ERR_STATE; }
  while ((int)e->prof_ev.size() < 2 * max_launches) {
    cudaEvent_t ev;
    SCB_CUDA_CHECK(cudaEventCreate(&ev));
    e->prof_ev.push_back(ev);
  }
  e->prof_ev_tag.assign(e->prof_ev.size(), 0);
  e->prof_tag = tag; e->prof_used = 0; e->prof_flops = 0.0; e->prof_stride = stride < 1 ? 1 : stride; e->step_seq = 0;
  SCB_CUDA_CHECK(cudaMemset(e->sb.prof, 0, 8 * sizeof(unsigned long long)));
  return SC_OK;
}

int sc_engine_profile_end(void* handle, int32_t n_tags, int32_t* launches_per_tag, double* ms_per_tag, double* host_flops,
                          uint64_t* counters8) {
  Engine* e = (Engine*)handle;
  if (!e || !e->finalized) { set_last_error("engine not finalized"); return SC_ERR_STATE; }
  SCB_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < n_tags; ++i) { launches_per_tag[i] = 0; ms_per_tag[i] = 0.0; }
  // events come in (begin, end) pairs per tag; totals (T_DEC_STEP_TOTAL / T_ENC_TOTAL) enclose other pairs
  std::vector<int> open_idx(T_COUNT, -1);
  for (int i = 0; i < e->prof_used; ++i) {
    const int t = e->prof_ev_tag[i];
    if (t > 0) { if (t < T_COUNT) open_idx[t] = i; continue; }
    const int tag = -t;
    if (tag <= 0 || tag >= T_COUNT || open_idx[tag] < 0) continue;
    float ms = 0.f;
    SCB_CUDA_CHECK(cudaEventElapsedTime(&ms, e->prof_ev[open_idx[tag]], e->prof_ev[i]));
    open_idx[tag] = -1;
    if (tag < n_tags) { ms_per_tag[tag] += ms; launches_per_tag[tag]++; }
  }
  *host_flops = e->prof_flops;
  SCB_CUDA_CHECK(cudaMemcpy(counters8, e->sb.prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  e->prof_tag = 0; e->prof_stride = 1;
  return SC_OK;
}

// ---------------- host-only planner
int sc_planner_create(int32_t n_streams, void** planner) {
  if (n_streams < 1 || !planner) return SC_ERR_ARG;
  *planner = new Planner(n_streams);
  return SC_OK;
}
int sc_planner_destroy(void* planner) { delete (Planner*)planner; return SC_OK; }
int sc_planner_reset(void* planner, int32_t s) { ((Planner*)planner)->reset(s); return SC_OK; }
static int planner_push_any(void* planner, int32_t s, int32_t count, int32_t is_final, int features, ScStreamPlan* plan);
int sc_planner_push(void* planner, int32_t s, int32_t n_samples, int32_t is_final, ScStreamPlan* plan) {
  return planner_push_any(planner, s, n_samples, is_final, 0, plan);
}
int sc_planner_push_features(void* planner, int32_t s, int32_t n_frames, int32_t is_final, ScStreamPlan* plan) {
  return planner_push_any(planner, s, n_frames, is_final, 1, plan);
}
static int planner_push_any(void* planner, int32_t s, int32_t count, int32_t is_final, int features, ScStreamPlan* plan) {
  Planner* p = (Planner*)planner;
  StreamPush r = features ? p->push_features(s, count, is_final != 0) : p->push(s, count, is_final != 0);
  if (r.error) { set_last_error("planner: unsupported shape (code %d)", r.error); return SC_ERR_STATE; }
  plan->called = r.called; plan->n_feat = r.n_feat; plan->n_sub = r.run_sub ? r.sd.t2 : 0; plan->n_blocks = (int)r.blocks.size();
  plan->n_enc_out = r.n_enc_out; plan->enc_len = p->state(s).enc_len; plan->n_decode_blocks = (int)r.dq_T.size();
  plan->last_T = r.dq_T.empty() ? 0 : r.dq_T.back();
  return SC_OK;
}

// ---------------- single operators
size_t sc_frontend_workspace_bytes(void) { return align_up(sizeof(FrontendTables)) + 256; }

int sc_frontend_init(void* workspace_dev, const float* window400, const float* mel_fb) {
  if (!workspace_dev || !window400 || !mel_fb) { set_last_error("frontend_init: null argument"); return SC_ERR_ARG; }
  return frontend_upload_tables((FrontendTables*)workspace_dev, window400, mel_fb) ? SC_ERR_CUDA : SC_OK;
}

int sc_frontend_fbank_mvn(void* workspace_dev, const float* wave_dev, int32_t n_samples, const double* mean_dev,
                          const double* std_dev, float* feats_dev, int32_t* n_frames, void* stream) {
  if (!workspace_dev || !wave_dev || !feats_dev || n_samples < 1) { set_last_error("frontend_fbank_mvn: bad argument"); return SC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  FrontendDesc d{};
  d.stream = 0; d.n_prev = 0; d.n_new = n_samples; d.slab = n_samples; d.n_frames = 1 + n_samples / 160;
  d.emit0 = 0; d.emit1 = d.n_frames; d.feat_off = 0; d.new_buf = 0; d.frame_base = 0;
  FrontendDesc* d_dev = (FrontendDesc*)((unsigned char*)workspace_dev + align_up(sizeof(FrontendTables)));
  SCB_CUDA_CHECK(cudaMemcpyAsync(d_dev, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  SCB_CUDA_CHECK(cudaStreamSynchronize(st));                   // `d` lives on this stack frame
  if (launch_frontend((const FrontendTables*)workspace_dev, wave_dev, n_samples, wave_dev, 0, d_dev, 1, d.n_frames,
                      mean_dev, mean_dev ? std_dev : nullptr, feats_dev, d.n_frames, st)) return SC_ERR_CUDA;
  if (n_frames) *n_frames = d.n_frames;
  return SC_OK;
}

int sc_ctc_prefix_step(const float* x_dev, int32_t t, int32_t v, const float* r_prev_dev, const int32_t* last_tok_dev,
                       int32_t prefix_len, const int32_t* cand_ids_dev, int32_t n_hyp, float* psi_dev, float* psi_eos_dev,
                       float* r_new_dev, void* stream) {
  if (!x_dev || !r_prev_dev || !last_tok_dev || !cand_ids_dev || !psi_dev || !psi_eos_dev) { set_last_error("ctc_prefix_step: null argument"); return SC_ERR_ARG; }
  return launch_ctc_prefix_op(x_dev, t, v, r_prev_dev, last_tok_dev, prefix_len, cand_ids_dev, n_hyp, psi_dev, psi_eos_dev,
                              r_new_dev, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}

int sc_layernorm_f32(const float* x, const float* w, const float* b, float* y, int32_t rows, int32_t d, void* stream) {
  return launch_layernorm(x, d, w, b, y, d, rows, d, nullptr, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}
int sc_linear_f32(const float* x, const float* w, const float* bias, const float* residual, float* y, int32_t m, int32_t n,
                  int32_t k, int32_t relu, void* stream) {
  GemmArgs g;
  g.A = x; g.lda = k; g.W = w; g.bias = bias; g.R = residual; g.ldr = n; g.C = y; g.ldc = n; g.M = m; g.N = n; g.K = k; g.relu = relu;
  return launch_gemm_f32(g, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}
int sc_linear_x3(const float* x, const void* w_planes_f16, const float* bias, const float* residual, float* y, int32_t m,
                 int32_t n, int32_t k, int32_t relu, void* stream) {
  GemmArgs g;
  g.A = x; g.lda = k; g.bias = bias; g.R = residual; g.ldr = n; g.C = y; g.ldc = n; g.M = m; g.N = n; g.K = k; g.relu = relu;
  return launch_gemm_x3(g, X3Extra(), w_planes_f16, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}
int sc_linear_x3_planes(const void* x_planes_f16, int64_t x_plane_elems, int32_t x_rows, const void* w_planes_f16,
                        const float* bias, const float* residual, float* y, void* y_planes_f16, int64_t y_plane_elems,
                        int32_t m, int32_t n, int32_t k, int32_t relu, int32_t kernel, void* stream) {
  GemmArgs g;
  g.lda = k; g.bias = bias; g.R = residual; g.ldr = n; g.C = y; g.ldc = n; g.M = m; g.N = n; g.K = k; g.relu = relu;
  X3Extra x;
  x.A2 = x_planes_f16; x.a2_plane = (size_t)x_plane_elems; x.a2_rows = x_rows;
  x.C2 = y_planes_f16; x.c2_plane = (size_t)y_plane_elems; x.ldc2 = n; x.kernel = kernel;
  return launch_gemm_x3(g, x, w_planes_f16, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}
int sc_linear_x3_ln(const float* x, const float* ln_w, const float* ln_b, const void* w_planes_f16, const float* bias, float* y,
                    void* y_planes_f16, int64_t y_plane_elems, int32_t m, int32_t n, int32_t relu, const int32_t* n_rows_dev,
                    void* stream) {
  GemmArgs g;
  g.lda = 256; g.bias = bias; g.C = y; g.ldc = n; g.ldr = n; g.M = m; g.N = n; g.K = 256; g.relu = relu; g.n_rows_dev = n_rows_dev;
  X3Extra ex;
  ex.lnX = x; ex.ldx = 256; ex.ln_w = ln_w; ex.ln_b = ln_b;
  ex.C2 = y_planes_f16; ex.c2_plane = (size_t)y_plane_elems; ex.ldc2 = n;
  return launch_gemm_x3(g, ex, w_planes_f16, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}
int sc_layernorm_split(const float* x, const float* w, const float* b, void* y_planes_f16, int64_t y_plane_elems,
                       int32_t rows, int32_t d, void* stream) {
  return launch_layernorm_split(x, d, w, b, y_planes_f16, (size_t)y_plane_elems, d, rows, d, nullptr, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}
int sc_linear_bf16(const void* x, const void* w, const float* bias, const float* residual, float* y, void* y16, int32_t m,
                   int32_t n, int32_t k, int32_t relu, void* stream) {
  return launch_gemm_bf16((const __nv_bfloat16*)x, k, (const __nv_bfloat16*)w, bias, residual, n, y, n, (__nv_bfloat16*)y16, n,
                          m, n, k, relu, nullptr, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}

int sc_linear_bf16_ln(const void* x, const void* w, const float* bias, const float* residual, float* y, const float* ln_w,
                      const float* ln_b, void* ln_out_bf16, int32_t m, int32_t k, void* stream) {
  return launch_gemm_bf16_ln((const __nv_bfloat16*)x, k, (const __nv_bfloat16*)w, bias, residual, 256, y, 256, nullptr, 0, nullptr,
                             m, 256, k, 0, nullptr, ln_w, ln_b, (__nv_bfloat16*)ln_out_bf16, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}

int sc_ffn_bf16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, float* y, int32_t accumulate,
                int32_t m, int32_t f, int32_t splits, void* stream) {
  return launch_ffn_fused_bf16((const __nv_bfloat16*)x, 256, (const __nv_bfloat16*)w1, b1, (const __nv_bfloat16*)w2, b2, y, 256,
                               accumulate, m, f, splits, nullptr, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}

int sc_ffn_bf16_timeline(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, float* y,
                         int32_t accumulate, int32_t m, int32_t f, int64_t* stamps, void* stream) {
  return launch_ffn_fused_bf16((const __nv_bfloat16*)x, 256, (const __nv_bfloat16*)w1, b1, (const __nv_bfloat16*)w2, b2, y, 256,
                               accumulate, m, f, 1, nullptr, (cudaStream_t)stream, (long long*)stamps) ? SC_ERR_CUDA : SC_OK;
}

int sc_linear_bf16_lnA(const float* x_f32, const float* ln_w, const float* ln_b, const void* w, const float* bias, float* y,
                       void* y16, int32_t m, int32_t n, int32_t relu, void* stream) {
  return launch_gemm_bf16_lnA(x_f32, 256, ln_w, ln_b, (const __nv_bfloat16*)w, bias, y, n, (__nv_bfloat16*)y16, n, m, n, relu,
                              nullptr, (cudaStream_t)stream) ? SC_ERR_CUDA : SC_OK;
}

}  // extern "C"
