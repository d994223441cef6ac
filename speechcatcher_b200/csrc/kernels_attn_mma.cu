// Decoder attention on tensor cores for the bf16 mode (bf16 K|V caches): flash-style, split-KV across the
// four warps of a CTA, mma.sync.m16n8k16 (bf16 x bf16 -> fp32) for Q*K^T and P*V, K|V tiles streamed with
// cp.async double buffering.  One CTA per (active stream, head); the <=16 hypotheses of the stream are the
// 16 rows of the MMA tile, so every K|V byte is read once per stream and the kernel is HBM-bound instead
// of shared-memory-issue bound like the fp32 SIMT variant in kernels_search.cu.
//
// Keys are an abstract list so that one kernel serves both attentions:
//   cross: key u = encoder frame u of the stream (visible to every hypothesis);
//   self : keys [0, Lc) = the beam's common ancestor chain (visible to all), followed by the divergent tail
//          as (hypothesis, position) pairs, each visible to its own hypothesis only (owner mask).
//
// Replaces the attention part of speechcatcher/model/decoder/decoder_layer.py:80-113.
#include "kernels.h"

namespace scb {

constexpr int A_KPW = 32;          // keys per warp per step (each warp runs its own cp.async pipeline)
constexpr int A_STEP = 4 * A_KPW;  // keys per CTA step
constexpr int A_NT = A_KPW / 8;    // score n-tiles per warp
constexpr int A_KK = A_KPW / 16;   // k-steps of the P*V product per warp
constexpr int MMA_MAXB = 32;       // hypotheses per stream: one m16 tile, or two tiles over blockIdx.z (WIDE)
constexpr int MMA_KEYS_SMEM = 768;  // self-attention key list entries staged in shared memory (longer lists spill to global)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One CTA per active stream, once per search iteration.
__global__ void __launch_bounds__(128) build_self_keys_kernel(SearchBuffers sb) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  const StreamCtl& c = sb.ctl[s];
  const int nb = c.n_hyp, len = c.len, B = sb.B, tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char ancs[];       // [nb][Lcap]
  __shared__ int s_lc;
  if (tid == 0) s_lc = len;
  const int words = sb.Lcap / 4;
  for (int i = tid; i < nb * words; i += 128) {               // rows are Lcap (multiple of 32) bytes: 4-byte loads
    const int b = i / words, w = i % words;
    const unsigned int* src = reinterpret_cast<const unsigned int*>(sb.anc + (((size_t)c.cur * sb.S + s) * B + b) * sb.Lcap);
    reinterpret_cast<unsigned int*>(ancs + (size_t)b * sb.Lcap)[w] = (4 * w < len) ? src[w] : 0u;
  }
  __syncthreads();
  if (tid < nb) ancs[(size_t)tid * sb.Lcap + len - 1] = (unsigned char)tid;     // the scored token lives at the own slot
  __syncthreads();
  for (int j = tid; j < len; j += 128) {
    const unsigned char a0 = ancs[j];
    bool same = true;
    for (int b = 1; b < nb; ++b) same &= (ancs[(size_t)b * sb.Lcap + j] == a0);
    if (!same) atomicMin(&s_lc, j);
  }
  __syncthreads();
  const int Lc = s_lc, n_div = len - Lc, n_keys = Lc + nb * n_div;
  int* keys = sb.self_keys + (size_t)s * sb.key_cap;
  for (int u = tid; u < n_keys && u < sb.key_cap; u += 128) {
    int j, slot, owner;
    if (u < Lc) { j = u; slot = ancs[u]; owner = -1; }
    else { const int p = u - Lc; owner = p / n_div; j = Lc + p % n_div; slot = ancs[(size_t)owner * sb.Lcap + j]; }
    keys[u] = j | (slot << 16) | ((owner + 1) << 24);
  }
  if (tid == 0) sb.self_nkeys[s] = min(n_keys, sb.key_cap);
}

int launch_build_self_keys(const SearchBuffers& sb, cudaStream_t st) {
  const size_t smem = (size_t)sb.B * sb.Lcap;
  static PerDeviceMark mk;
  size_t& attr = mk.cur();
  if (smem > 48 * 1024 && attr < smem) {
    cudaFuncSetAttribute(build_self_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  launch_k(build_self_keys_kernel, dim3(sb.S), dim3(128), smem, st, sb);
  SCB_LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ float ex2_approx(float x) {      // 2^x, ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ncu (mid-utterance, T ~ 400): the first version of this kernel was ISSUE-bound, not memory-bound -- 1900 warp
// instructions per 64-key tile for 32 HMMAs (CTA-wide double buffering with two __syncthreads per tile, per-element
// branches around expf, per-element mask loads, 64-bit address arithmetic per cp.async).  This version gives every
// warp a private two-stage cp.async pipeline over its own 32 keys per step (no CTA barrier in the main loop), works
// on raw scores with one ex2 per element (scale * log2(e) folded into an FMA), masks only where a mask can exist
// (self: the divergent tail, found with one warp vote; cross: the last step), and fetches B fragments with x4
// ldmatrix.
// WIDE = 1 (beam 17..32, opt-in, not yet run on a device): the hypotheses of a stream are split into m16 tiles over
// blockIdx.z; a CTA loads / appends / writes only the rows h0..h0+15 of its tile and compares key owners against the
// stream-wide hypothesis index.  A hypothesis' new token is visible to itself only, so the per-tile appends of the
// self-attention K|V do not race.  WIDE = 0 compiles to exactly the single-tile kernel (h0 = 0 folds away).
template <int DK, int MODE, int WIDE>
__global__ void __launch_bounds__(128) dec_attn_mma_kernel(SearchBuffers sb, __nv_bfloat16* kv_layer,
                                                           const float* __restrict__ q, int ldq, int q_off,
                                                           float* __restrict__ out, __nv_bfloat16* __restrict__ out16) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  const int head = blockIdx.y;
  const StreamCtl& c = sb.ctl[s];
  const int h0 = WIDE ? 16 * (int)blockIdx.z : 0;            // first hypothesis of this CTA's tile
  if (WIDE && h0 >= c.n_hyp) return;
  const int nb = WIDE ? min(16, c.n_hyp - h0) : c.n_hyp;     // hypotheses in the tile
  const int row0 = sb.row_base[s] + h0, D = sb.D, B = sb.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int RS = DK + 8;                 // padded smem row (bf16 elements): conflict-free ldmatrix
  constexpr int CPR = DK / 8;                // 16-byte chunks per K (or V) row
  constexpr int KSTEPS = DK / 16;
  constexpr int NDT = DK / 8;                // n-tiles of the output
  constexpr int RPP = 32 / CPR;              // rows one pass of the warp's 32 lanes covers
  constexpr int PASSES = A_KPW / RPP;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);               // [16][RS]
  __nv_bfloat16* KV = Qs + 16 * RS;                                             // [4 warps][2 stages][K|V][A_KPW][RS]
  signed char* own = reinterpret_cast<signed char*>(KV + 4 * 2 * 2 * A_KPW * RS);   // [4][2][A_KPW]
  int* keys_s = reinterpret_cast<int*>(own + 4 * 2 * A_KPW);                     // self only: [MMA_KEYS_SMEM]
  // merge scratch aliases the K|V stages after the main loop
  float* mrg_m = reinterpret_cast<float*>(KV);                                   // [4][16]
  float* mrg_l = mrg_m + 64;                                                     // [4][16]
  float* mrg_o = mrg_l + 64;                                                     // [4][16][DK]

  const size_t row_stride = 2 * (size_t)D;
  const int len = c.len;
  __nv_bfloat16* base;
  if (MODE == 1) base = kv_layer + (size_t)s * sb.Tcap * row_stride + head * DK;
  else base = kv_layer + (size_t)s * sb.Lcap * B * row_stride + head * DK;

  // ---- Q tile (rows >= nb are zero: they produce finite values that are never written), self: append K|V
  for (int i = tid; i < 16 * DK; i += 128) {
    const int r = i / DK, d = i % DK;
    const float v = r < nb ? q[(size_t)(row0 + r) * ldq + q_off + head * DK + d] : 0.f;
    Qs[r * RS + d] = __float2bfloat16(v);
  }
  int n_keys;
  const int* keys = nullptr;
  if (MODE == 0) {
    if (tid == 0) atomicAdd(&sb.prof[3], (unsigned long long)((long long)nb * 2ll * len * DK * 2));   // per tile: sums to n_hyp
    for (int i = tid; i < nb * 2 * DK; i += 128) {        // append K|V of the scored token at [len-1][b]
      const int b = i / (2 * DK), rem = i % (2 * DK), which = rem / DK, cc = rem % DK;
      const float v = q[(size_t)(row0 + b) * ldq + D + which * D + head * DK + cc];
      base[((size_t)(len - 1) * B + (h0 + b)) * row_stride + which * D + cc] = __float2bfloat16(v);
    }
    n_keys = sb.self_nkeys[s];
    keys = sb.self_keys + (size_t)s * sb.key_cap;
    // the key list is read once, coalesced, instead of one dependent global load in front of every step
    for (int i = tid; i < n_keys && i < MMA_KEYS_SMEM; i += 128) keys_s[i] = keys[i];
  } else {
    n_keys = c.Tb;
    if (tid == 0 && h0 == 0) atomicAdd(&sb.prof[2], (unsigned long long)(2ll * n_keys * DK * 2));
  }
  __syncthreads();                      // Qs, key list staged; appended rows visible to the loads below
  const int n_steps = (n_keys + A_STEP - 1) / A_STEP;

  __nv_bfloat16* kw = KV + (size_t)warp * (2 * 2 * A_KPW * RS);      // this warp's stages: [2][K|V][A_KPW][RS]
  signed char* ownw = own + warp * 2 * A_KPW;
  const int lr = lane / CPR, ch8 = (lane % CPR) * 8;

  auto issue = [&](int t, int buf) {
    const int u0 = t * A_STEP + A_KPW * warp;
    __nv_bfloat16* kd = kw + (size_t)buf * (2 * A_KPW * RS);
    __nv_bfloat16* vd = kd + A_KPW * RS;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = lr + RPP * ps, u = u0 + r;
      if (u < n_keys) {
        const __nv_bfloat16* src;
        if (MODE == 1) src = base + (size_t)u * row_stride + ch8;
        else {
          const int kd_ = u < MMA_KEYS_SMEM ? keys_s[u] : keys[u];
          src = base + ((size_t)(kd_ & 0xffff) * B + ((kd_ >> 16) & 0xff)) * row_stride + ch8;
          if (ch8 == 0) ownw[buf * A_KPW + r] = (signed char)((kd_ >> 24) - 1);
        }
        cp_async16(kd + r * RS + ch8, src);
        cp_async16(vd + r * RS + ch8, src + D);
      } else {
        *reinterpret_cast<uint4*>(kd + r * RS + ch8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(vd + r * RS + ch8) = make_uint4(0, 0, 0, 0);
        if (MODE == 0 && ch8 == 0) ownw[buf * A_KPW + r] = (signed char)-2;      // invisible to everyone
      }
    }
    cp_async_commit();
  };

  if (n_steps > 0) issue(0, 0);
  uint32_t qa[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    ldsm_x4(qa[ks], Qs + ((lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 16 * ks + 8 * (lane >> 4));

  const float cs = 1.4426950408889634f / sqrtf((float)DK);      // softmax scale * log2(e): p = 2^(cs * (s - m))
  const int r0 = lane >> 2, r1 = r0 + 8, qd = lane & 3;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};   // running max of the RAW scores
  float o[NDT][4];
#pragma unroll
  for (int i = 0; i < NDT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

  for (int t = 0; t < n_steps; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_steps) { issue(t + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();                       // the other lanes' copies / zero fills of this stage are visible
    const int u0 = t * A_STEP + A_KPW * warp;
    if (u0 < n_keys) {
      const __nv_bfloat16* kb = kw + (size_t)buf * (2 * A_KPW * RS);
      const __nv_bfloat16* vb = kb + A_KPW * RS;
      // ---- S = Q K^T for this warp's A_KPW keys (raw scores)
      float sacc[A_NT][4];
#pragma unroll
      for (int nt = 0; nt < A_NT; ++nt) {
        sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < KSTEPS / 2; ++k2) {
          uint32_t bfr[4];              // B fragments of two k-steps with one ldmatrix
          ldsm_x4(bfr, kb + (8 * nt + (lane & 7)) * RS + 32 * k2 + 8 * (lane >> 3));
          mma_bf16(sacc[nt], qa[2 * k2], bfr);
          mma_bf16(sacc[nt], qa[2 * k2 + 1], bfr + 2);
        }
      }
      // ---- mask: only where one can exist
      if (MODE == 0) {
        const signed char* ob = ownw + buf * A_KPW;
        if (!__all_sync(0xffffffffu, ob[lane] == -1)) {       // step touches the divergent tail (or padding)
#pragma unroll
          for (int nt = 0; nt < A_NT; ++nt) {
            const int o0 = ob[8 * nt + 2 * qd], o1 = ob[8 * nt + 2 * qd + 1];
            if (!(o0 == -1 || o0 == h0 + r0)) sacc[nt][0] = -INFINITY;
            if (!(o1 == -1 || o1 == h0 + r0)) sacc[nt][1] = -INFINITY;
            if (!(o0 == -1 || o0 == h0 + r1)) sacc[nt][2] = -INFINITY;
            if (!(o1 == -1 || o1 == h0 + r1)) sacc[nt][3] = -INFINITY;
          }
        }
      } else if (u0 + A_KPW > n_keys) {                        // cross: zero-filled rows past the last frame
#pragma unroll
        for (int nt = 0; nt < A_NT; ++nt) {
          const int k0 = u0 + 8 * nt + 2 * qd;
          if (k0 >= n_keys) { sacc[nt][0] = -INFINITY; sacc[nt][2] = -INFINITY; }
          if (k0 + 1 >= n_keys) { sacc[nt][1] = -INFINITY; sacc[nt][3] = -INFINITY; }
        }
      }
      // ---- online softmax (rows r0, r1 of this thread)
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < A_NT; ++nt) {
        mx[0] = fmaxf(mx[0], fmaxf(sacc[nt][0], sacc[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(sacc[nt][2], sacc[nt][3]));
      }
      float scale[2], nm[2], psum[2] = {0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        const float m_new = fmaxf(m_run[h], mx[h]);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;     // row has seen no visible key yet
        scale[h] = ex2_approx((m_run[h] - m_use) * cs);             // m_run = -inf -> 0 (o, l are still 0)
        m_run[h] = m_new;
        nm[h] = -m_use * cs;
      }
      uint32_t pa[A_KK][4];
#pragma unroll
      for (int nt = 0; nt < A_NT; ++nt) {
        const float p0 = ex2_approx(fmaf(sacc[nt][0], cs, nm[0])), p1 = ex2_approx(fmaf(sacc[nt][1], cs, nm[0]));
        const float p2 = ex2_approx(fmaf(sacc[nt][2], cs, nm[1])), p3 = ex2_approx(fmaf(sacc[nt][3], cs, nm[1]));
        psum[0] += p0 + p1; psum[1] += p2 + p3;
        // C fragment of S -> A fragment of P: n-tile 2kk -> a0/a1, n-tile 2kk+1 -> a2/a3
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        psum[h] += __shfl_xor_sync(0xffffffffu, psum[h], 1);
        psum[h] += __shfl_xor_sync(0xffffffffu, psum[h], 2);
        l_run[h] = l_run[h] * scale[h] + psum[h];
      }
#pragma unroll
      for (int nd = 0; nd < NDT; ++nd) { o[nd][0] *= scale[0]; o[nd][1] *= scale[0]; o[nd][2] *= scale[1]; o[nd][3] *= scale[1]; }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < A_KK; ++kk) {
#pragma unroll
        for (int n2 = 0; n2 < NDT / 2; ++n2) {
          uint32_t bfr[4];              // V fragments of two output n-tiles with one transposing ldmatrix
          ldsm_x4_trans(bfr, vb + (16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 16 * n2 + 8 * (lane >> 4));
          mma_bf16(o[2 * n2], pa[kk], bfr);
          mma_bf16(o[2 * n2 + 1], pa[kk], bfr + 2);
        }
      }
    }
    __syncwarp();                       // stage fully consumed before this warp refills it
  }
  __syncthreads();                      // every warp is done with its stages: the merge scratch aliases them
  // ---- merge the four warps' partial (m, l, O)
  if (qd == 0) {
    mrg_m[warp * 16 + r0] = m_run[0]; mrg_m[warp * 16 + r1] = m_run[1];
    mrg_l[warp * 16 + r0] = l_run[0]; mrg_l[warp * 16 + r1] = l_run[1];
  }
#pragma unroll
  for (int nd = 0; nd < NDT; ++nd) {
    float* dst = mrg_o + (size_t)warp * 16 * DK;
    dst[r0 * DK + 8 * nd + 2 * qd] = o[nd][0]; dst[r0 * DK + 8 * nd + 2 * qd + 1] = o[nd][1];
    dst[r1 * DK + 8 * nd + 2 * qd] = o[nd][2]; dst[r1 * DK + 8 * nd + 2 * qd + 1] = o[nd][3];
  }
  __syncthreads();
  for (int i = tid; i < nb * DK; i += 128) {
    const int r = i / DK, d = i % DK;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, mrg_m[w * 16 + r]);
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = mrg_m[w * 16 + r];
      const float f = (mw == -INFINITY) ? 0.f : ex2_approx((mw - M) * cs);
      L += mrg_l[w * 16 + r] * f;
      acc += mrg_o[((size_t)w * 16 + r) * DK + d] * f;
    }
    const float res = acc / L;
    out[(size_t)(row0 + r) * D + head * DK + d] = res;
    if (out16) out16[(size_t)(row0 + r) * D + head * DK + d] = __float2bfloat16(res);
  }
}

// ---------------------------------------------------------------- encoder block attention on tensor cores (bf16 mode)
// One CTA (3 warps) per (block, head): the 42 block rows padded to 48 = three m16 tiles, one per warp; each warp
// holds all 48 keys of its 16 query rows, so the masked softmax needs no cross-warp exchange.
// Mask as in the fp32 kernel: rows 1..41 attend keys 0..40 (short path: rows/keys < n_rows, no mask).
template <int DK>
__global__ void __launch_bounds__(96) enc_attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv16, float* __restrict__ out,
                                                          __nv_bfloat16* __restrict__ out16,
                                                          const BlockDesc* __restrict__ blk, int D) {
  const BlockDesc b = blk[blockIdx.x];
  const int head = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int RS = DK + 8, CPR = DK / 8, KSTEPS = DK / 16, NDT = DK / 8, ROWS = 48;
  __shared__ __align__(16) __nv_bfloat16 Qs[ROWS * RS];
  __shared__ __align__(16) __nv_bfloat16 Ks[ROWS * RS];
  __shared__ __align__(16) __nv_bfloat16 Vs[ROWS * RS];
  const __nv_bfloat16* base = qkv16 + (size_t)blockIdx.x * kSlots * 3 * D + head * DK;
  for (int idx = tid; idx < ROWS * CPR; idx += 96) {
    const int r = idx / CPR, ch = idx % CPR;
    if (r < kSlots) {
      const __nv_bfloat16* src = base + (size_t)r * 3 * D + ch * 8;
      cp_async16(Qs + r * RS + ch * 8, src);
      cp_async16(Ks + r * RS + ch * 8, src + D);
      cp_async16(Vs + r * RS + ch * 8, src + 2 * D);
    } else {
      *reinterpret_cast<uint4*>(Qs + r * RS + ch * 8) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Ks + r * RS + ch * 8) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Vs + r * RS + ch * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int q_lo = b.short_path ? 0 : 1, q_hi = b.short_path ? b.n_rows : kSlots;
  const int k_hi = b.short_path ? b.n_rows : kBlock + 1;
  uint32_t qa[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    ldsm_x4(qa[ks], Qs + (16 * warp + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 16 * ks + 8 * (lane >> 4));
  float sacc[6][4];
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      uint32_t bfr[2];
      ldsm_x2(bfr, Ks + (8 * nt + (lane & 7)) * RS + 16 * ks + 8 * ((lane >> 3) & 1));
      mma_bf16(sacc[nt], qa[ks], bfr);
    }
  }
  const float inv_sqrt = 1.0f / sqrtf((float)DK);
  const int r0 = 16 * warp + (lane >> 2), r1 = r0 + 8, qd = lane & 3;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 6; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = 8 * nt + 2 * qd + (e & 1);
      const int row = (e & 2) ? r1 : r0;
      const bool vis = row >= q_lo && row < q_hi && key < k_hi;
      const float v = vis ? sacc[nt][e] * inv_sqrt : -INFINITY;
      sacc[nt][e] = v;
      mx[e >> 1] = fmaxf(mx[e >> 1], v);
    }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
  }
#pragma unroll
  for (int nt = 0; nt < 6; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float m = mx[e >> 1];
      const float p = (m == -INFINITY) ? 0.f : expf(sacc[nt][e] - m);
      sacc[nt][e] = p;
      sum[e >> 1] += p;
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
    sum[h] = sum[h] > 0.f ? 1.0f / sum[h] : 0.f;
  }
  uint32_t pa[3][4];
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(sacc[nt][0] * sum[0], sacc[nt][1] * sum[0]);
    pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(sacc[nt][2] * sum[1], sacc[nt][3] * sum[1]);
  }
  float o[NDT][4];
#pragma unroll
  for (int nd = 0; nd < NDT; ++nd) {
    o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      uint32_t bfr[2];
      ldsm_x2_trans(bfr, Vs + (16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 8 * nd);
      mma_bf16(o[nd], pa[kk], bfr);
    }
  }
#pragma unroll
  for (int nd = 0; nd < NDT; ++nd) {
    const int col = head * DK + 8 * nd + 2 * qd;
    if (r0 < kSlots) {
      const size_t off = ((size_t)blockIdx.x * kSlots + r0) * D + col;
      *reinterpret_cast<uint32_t*>(out16 + off) = pack_bf16(o[nd][0], o[nd][1]);
      if (out) { out[off] = o[nd][0]; out[off + 1] = o[nd][1]; }
    }
    if (r1 < kSlots) {
      const size_t off = ((size_t)blockIdx.x * kSlots + r1) * D + col;
      *reinterpret_cast<uint32_t*>(out16 + off) = pack_bf16(o[nd][2], o[nd][3]);
      if (out) { out[off] = o[nd][2]; out[off + 1] = o[nd][3]; }
    }
  }
}

int launch_enc_attention_mma(const __nv_bfloat16* qkv16, float* out, __nv_bfloat16* out16, const BlockDesc* blk, int n_blk,
                             int n_head, int d_model, cudaStream_t st) {
  if (n_blk <= 0) return 0;
  dim3 grid(n_blk, n_head);
  const int dk = d_model / n_head;
  if (dk == 32) enc_attn_mma_kernel<32><<<grid, 96, 0, st>>>(qkv16, out, out16, blk, d_model);
  else if (dk == 64) enc_attn_mma_kernel<64><<<grid, 96, 0, st>>>(qkv16, out, out16, blk, d_model);
  else { set_last_error("enc_attn_mma: unsupported head dim %d", dk); return -1; }
  SCB_LAUNCH_CHECK();
  return 0;
}

template <int DK, int MODE>
static int launch_mma_t(const SearchBuffers& sb, __nv_bfloat16* kv_layer, const float* q, int ldq, int q_off, float* out,
                        __nv_bfloat16* out16, cudaStream_t st) {
  constexpr int RS = DK + 8;
  const size_t stages = sizeof(__nv_bfloat16) * 4 * 2 * 2 * (size_t)A_KPW * RS;     // [4 warps][2 stages][K|V][A_KPW][RS]
  size_t smem = sizeof(__nv_bfloat16) * (size_t)16 * RS + stages + 4 * 2 * A_KPW + 16;
  if (MODE == 0) smem += sizeof(int) * MMA_KEYS_SMEM;
  const size_t merge = sizeof(float) * (128 + 4 * 16 * DK);
  if (stages < merge) { set_last_error("attn_mma: merge scratch does not fit"); return -1; }
  if (sb.B > 16) {                                  // beam 17..32: one CTA per (stream, head, m16 tile of hypotheses)
    static PerDeviceMark mk_w;
    size_t& attr_w = mk_w.cur();
    if (attr_w < smem) {
      if (cudaFuncSetAttribute(dec_attn_mma_kernel<DK, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_last_error("attn_mma: cudaFuncSetAttribute(%zu) failed", smem);
        return -1;
      }
      attr_w = smem;
    }
    launch_k(dec_attn_mma_kernel<DK, MODE, 1>, dim3(sb.S, sb.H, (sb.B + 15) / 16), dim3(128), smem, st, sb, kv_layer, q, ldq,
             q_off, out, out16);
    SCB_LAUNCH_CHECK();
    return 0;
  }
  static PerDeviceMark mk;
  size_t& attr = mk.cur();
  if (attr < smem) {
    if (cudaFuncSetAttribute(dec_attn_mma_kernel<DK, MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error("attn_mma: cudaFuncSetAttribute(%zu) failed", smem);
      return -1;
    }
    attr = smem;
  }
  dim3 grid(sb.S, sb.H);
  launch_k(dec_attn_mma_kernel<DK, MODE, 0>, grid, dim3(128), smem, st, sb, kv_layer, q, ldq, q_off, out, out16);
  SCB_LAUNCH_CHECK();
  return 0;
}

// mode 0: self attention (q = fused QKV GEMM output, row stride ldq = 3D; K|V of the new token at +D);
// mode 1: cross attention.  Requires bf16 KV caches and beam <= 32 (beam > 16 takes the tiled WIDE variant).
int launch_dec_attention_mma(const SearchBuffers& sb, int mode, int layer, const float* q, int ldq, float* out,
                             __nv_bfloat16* out16, cudaStream_t st) {
  if (!sb.kv_bf16 || sb.B > MMA_MAXB) { set_last_error("attn_mma: needs bf16 KV and beam <= %d", MMA_MAXB); return -1; }
  const int dk = sb.D / sb.H;
  __nv_bfloat16* kv = mode == 0
      ? reinterpret_cast<__nv_bfloat16*>(sb.skv) + (size_t)layer * sb.S * sb.Lcap * sb.B * 2 * sb.D
      : reinterpret_cast<__nv_bfloat16*>(sb.xkv) + (size_t)layer * sb.S * sb.Tcap * 2 * sb.D;
  if (dk == 32) return mode == 0 ? launch_mma_t<32, 0>(sb, kv, q, ldq, 0, out, out16, st) : launch_mma_t<32, 1>(sb, kv, q, ldq, 0, out, out16, st);
  if (dk == 64) return mode == 0 ? launch_mma_t<64, 0>(sb, kv, q, ldq, 0, out, out16, st) : launch_mma_t<64, 1>(sb, kv, q, ldq, 0, out, out16, st);
  set_last_error("attn_mma: unsupported head dim %d", dk);
  return -1;
}

}  // namespace scb
