// fp32 decoder attention for the precise modes (fp32 K|V caches): one CTA per active stream, one warp per head, no
// CTA-wide barrier in the main loop.
//
// The fp32 decode step spends most of its time in attention (self-attention over the KV tree, cross-attention over the
// stream's encoder memory), and the first SIMT kernels (kernels_search.cu: one CTA per (stream, head), thread-per-key
// scoring out of shared-memory tiles with four __syncthreads per 128 keys) ran at a small fraction of both the HBM and
// the FMA roofline.  Here every warp runs a private pipeline over its head's slice of the K|V rows:
//   * lane j owns key 32 t + j of tile t: its K row slice (DK floats) is loaded straight into registers (next tile's
//     prefetched while this one is scored), the V slice of the same row goes to a warp-private shared-memory stage by
//     cp.async (two stages);
//   * scores: lane j computes q_b . k_j for the hypotheses b of the stream (q rows broadcast from shared memory);
//   * online softmax per hypothesis (running max via warp shuffles; the running sum stays per lane until the end);
//   * P V: the tile's probabilities are transposed through shared memory, lane c then accumulates output dimension c
//     for every hypothesis (V[j][c] conflict-free, p[b][j] broadcast).
// Keys are an abstract list like in kernels_attn_mma.cu: cross = encoder frames 0..Tb-1 (visible to every hypothesis),
// self = the per-step key list of the KV tree (build_self_keys_kernel): common ancestor chain once, then the divergent
// tail as (hypothesis, position) pairs visible to their owner only.
//
// Arithmetic per element is that of the reference's attention (scores = q.k / sqrt(dk), softmax, weighted sum of V;
// speechcatcher/model/attention/multi_head_attention.py:92-133) in fp32 with expf; only the summation order differs.
//
// Replaces the attention part of speechcatcher/model/decoder/decoder_layer.py:80-113.
#include "kernels.h"

namespace scb {

__device__ __forceinline__ void cpa16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int DK, int MAXB, int MODE>
__global__ void __launch_bounds__(256, (MAXB <= 10 && DK == 32) ? 2 : 1) dec_attn_f32_kernel(SearchBuffers sb, float* kv_layer, const float* __restrict__ q,
                                                           int ldq, float* __restrict__ out, SplitOut so) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  const StreamCtl& c = sb.ctl[s];
  const int nb = c.n_hyp, row0 = sb.row_base[s], D = sb.D, B = sb.B, len = c.len;
  const int tid = threadIdx.x, head = tid >> 5, lane = tid & 31, H = blockDim.x >> 5;
  constexpr int C4 = DK / 4;                  // 16-byte chunks per K (or V) row slice
  constexpr int NO = DK / 32;                 // output dimensions per lane
  constexpr bool PF = DK <= 32;               // register prefetch of the next tile's K rows

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);                       // [MAXB][D]
  float* vs_all = qs + MAXB * D;                                        // [H][2][32][DK]
  float* ps_all = vs_all + (size_t)H * 2 * 32 * DK;                     // [H][MAXB][32]
  float* vs = vs_all + (size_t)head * 2 * 32 * DK;
  float* ps = ps_all + (size_t)head * MAXB * 32;

  const size_t row_stride = 2 * (size_t)D;
  float* base;                                                          // this head's column slice of the K|V rows
  if (MODE == 1) base = kv_layer + (size_t)s * sb.Tcap * row_stride + head * DK;
  else base = kv_layer + (size_t)s * sb.Lcap * B * row_stride + head * DK;

  for (int i = tid; i < nb * D; i += blockDim.x) qs[i] = q[(size_t)(row0 + i / D) * ldq + i % D];
  int n_keys;
  const int* keys = nullptr;
  if (MODE == 0) {
    if (tid == 0) atomicAdd(&sb.prof[3], (unsigned long long)((long long)nb * 2ll * len * D * 4));
    // append K|V of the scored token at [len-1][b] (columns D.. and 2D.. of the fused QKV row)
    for (int b = 0; b < nb; ++b) {
      const float* src = q + (size_t)(row0 + b) * ldq + D + head * DK;
      float* dst = base + ((size_t)(len - 1) * B + b) * row_stride;
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        dst[lane + 32 * i] = src[lane + 32 * i];
        dst[D + lane + 32 * i] = src[D + lane + 32 * i];
      }
    }
    n_keys = sb.self_nkeys[s];
    keys = sb.self_keys + (size_t)s * sb.key_cap;
  } else {
    n_keys = c.Tb;
    if (tid == 0) atomicAdd(&sb.prof[2], (unsigned long long)(2ll * n_keys * D * 4));
  }
  __syncthreads();                      // q rows staged; appended rows visible to the loads below
  const int n_tiles = (n_keys + 31) >> 5;
  const float sqrt_dk = sqrtf((float)DK);

  // per-lane description of "my key" of a tile: row pointer (nullptr past the end) and owner (-1 = visible to all)
  auto key_of = [&](int t, const float*& row, int& owner) {
    const int u = t * 32 + lane;
    row = nullptr; owner = -1;
    if (u < n_keys) {
      if (MODE == 1) row = base + (size_t)u * row_stride;
      else {
        const int kw = keys[u];
        row = base + ((size_t)(kw & 0xffff) * B + ((kw >> 16) & 0xff)) * row_stride;
        owner = (kw >> 24) - 1;
      }
    }
  };
  auto issue_v = [&](const float* row, int buf) {                       // lane j stages the V slice of its own key
    float* dst = vs + ((size_t)buf * 32 + lane) * DK;
    if (row) {
#pragma unroll
      for (int i = 0; i < C4; ++i) cpa16(dst + 4 * i, row + D + 4 * i);
    } else {
#pragma unroll
      for (int i = 0; i < C4; ++i) *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cpa_commit();
  };
  auto load_k = [&](const float* row, float* k) {
    if (row) {
#pragma unroll
      for (int i = 0; i < C4; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(row) + i);
        k[4 * i] = v.x; k[4 * i + 1] = v.y; k[4 * i + 2] = v.z; k[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < DK; ++i) k[i] = 0.f;
    }
  };

  float m_run[MAXB], l_lane[MAXB], o[MAXB][NO];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    m_run[b] = -INFINITY; l_lane[b] = 0.f;
#pragma unroll
    for (int i = 0; i < NO; ++i) o[b][i] = 0.f;
  }
  float kn[DK];
  const float* row_n; int own_n;
  if (n_tiles > 0) {
    key_of(0, row_n, own_n);
    issue_v(row_n, 0);
    if (PF) load_k(row_n, kn);
  }
#pragma unroll 1
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    const float* row_c = row_n; const int own_c = own_n;
    float kc[DK];
    if (PF) {
#pragma unroll
      for (int i = 0; i < DK; ++i) kc[i] = kn[i];
    } else {
      load_k(row_c, kc);
    }
    if (t + 1 < n_tiles) {
      key_of(t + 1, row_n, own_n);
      issue_v(row_n, buf ^ 1);
      if (PF) load_k(row_n, kn);
      cpa_wait<1>();
    } else {
      cpa_wait<0>();
    }
    __syncwarp();                       // the other lanes' V copies / zero fills of this stage are visible
    const bool valid = row_c != nullptr;
    // ---- scores of my key for every hypothesis that can see it, online softmax, probabilities -> ps[b][lane]
    unsigned vis_mask = 0;              // hypotheses with at least one visible key in this tile (warp-uniform)
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < nb) {
        const bool vis = valid && (own_c < 0 || own_c == b);
        if (__any_sync(0xffffffffu, vis)) {
          vis_mask |= 1u << b;
          float sc = -INFINITY;
          if (vis) {
            const float4* qb = reinterpret_cast<const float4*>(qs + b * D + head * DK);
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < C4; ++i) {
              const float4 qv = qb[i];
              d = fmaf(qv.x, kc[4 * i], d); d = fmaf(qv.y, kc[4 * i + 1], d);
              d = fmaf(qv.z, kc[4 * i + 2], d); d = fmaf(qv.w, kc[4 * i + 3], d);
            }
            sc = d / sqrt_dk;
          }
          const float m_new = fmaxf(m_run[b], warp_max(sc));
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          const float f = expf(m_run[b] - m_use);             // m_run = -inf -> 0 (o, l are still 0)
          const float p = vis ? expf(sc - m_use) : 0.f;
          m_run[b] = m_new;
          l_lane[b] = l_lane[b] * f + p;
#pragma unroll
          for (int i = 0; i < NO; ++i) o[b][i] *= f;
          ps[b * 32 + lane] = p;
        }
      }
    }
    __syncwarp();                       // probabilities of all 32 keys visible
    // ---- O += P V: lane owns output dimension(s) lane + 32 i
    const float* vt = vs + (size_t)buf * 32 * DK;
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
      float v[4][NO];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < NO; ++i) v[jj][i] = vt[(j4 + jj) * DK + lane + 32 * i];
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        if (b < nb && (vis_mask >> b & 1u)) {
          const float4 p4 = *reinterpret_cast<const float4*>(ps + b * 32 + j4);
#pragma unroll
          for (int i = 0; i < NO; ++i) {
            float a = o[b][i];
            a = fmaf(p4.x, v[0][i], a); a = fmaf(p4.y, v[1][i], a); a = fmaf(p4.z, v[2][i], a); a = fmaf(p4.w, v[3][i], a);
            o[b][i] = a;
          }
        }
      }
    }
    __syncwarp();                       // stage and probabilities fully consumed before the next tile refills them
  }
  // ---- normalise and write: row (row0 + b), columns head * DK + lane + 32 i
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    if (b < nb) {
      const float l = warp_sum(l_lane[b]);
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        const float r = o[b][i] / l;
        const int col = head * DK + lane + 32 * i;
        if (so.base) so.put((size_t)(row0 + b), col, r);
        else out[(size_t)(row0 + b) * D + col] = r;
      }
    }
  }
}

template <int DK, int MAXB, int MODE>
static int attn_f32_launch(const SearchBuffers& sb, float* kv_layer, const float* q, int ldq, float* out, SplitOut so,
                           cudaStream_t st) {
  const int H = sb.D / DK;
  const size_t smem = sizeof(float) * ((size_t)MAXB * sb.D + (size_t)H * 2 * 32 * DK + (size_t)H * MAXB * 32);
  static PerDeviceMark mk;
  size_t& attr = mk.cur();
  if (attr < smem) {
    if (cudaFuncSetAttribute(dec_attn_f32_kernel<DK, MAXB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(dec_attn_f32, smem=%zu) failed", smem);
      return -1;
    }
    attr = smem;
  }
  launch_k(dec_attn_f32_kernel<DK, MAXB, MODE>, dim3(sb.S), dim3(H * 32), smem, st, sb, kv_layer, q, ldq, out, so);
  SCB_LAUNCH_CHECK();
  return 0;
}

// mode 0: self-attention (q = fused QKV rows, needs launch_build_self_keys earlier in the step), 1: cross-attention
int launch_dec_attention_f32(const SearchBuffers& sb, int mode, int layer, const float* q, int ldq, float* out,
                             SplitOut so, cudaStream_t st) {
  if (sb.kv_bf16) { set_last_error("dec_attention_f32 needs fp32 K|V caches"); return -1; }
  const int dk = sb.D / sb.H;
  if ((dk != 32 && dk != 64) || sb.B > 20 || sb.D % dk != 0 || sb.D / dk > 8) {
    set_last_error("dec_attention_f32: unsupported head dim %d / beam %d", dk, sb.B);
    return -1;
  }
  float* kvl = mode == 1 ? sb.xkv + (size_t)layer * sb.S * sb.Tcap * 2 * sb.D
                         : sb.skv + (size_t)layer * sb.S * sb.Lcap * sb.B * 2 * sb.D;
#define SCB_ATTN_F32(DKV, MB)                                                                         \
  return mode == 1 ? attn_f32_launch<DKV, MB, 1>(sb, kvl, q, ldq, out, so, st)                        \
                   : attn_f32_launch<DKV, MB, 0>(sb, kvl, q, ldq, out, so, st)
  if (dk == 32) { if (sb.B <= 10) { SCB_ATTN_F32(32, 10); } else { SCB_ATTN_F32(32, 20); } }
  if (sb.B <= 10) { SCB_ATTN_F32(64, 10); }
  SCB_ATTN_F32(64, 20);
#undef SCB_ATTN_F32
}

}  // namespace scb
