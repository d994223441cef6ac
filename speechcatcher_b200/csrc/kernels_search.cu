// Block-synchronous beam search on the device: decoder-step attention over a tree-structured self-KV
// store and a per-stream cross-KV cache, log-softmax + pre-beam top-40, the batched CTC prefix scorer,
// score combination + per-hypothesis top-k, global pruning and all hypothesis / endpoint bookkeeping
// (EOS stop, BBD rollback, rewind, global step cap, per-push block queue).
//
// Replaces: speechcatcher/beam_search/beam_search.py:71-185, 403-505, 655-838
//           speechcatcher/beam_search/ctc_prefix_score_full.py:88-368
//           speechcatcher/beam_search/scorers.py:238-431
//           speechcatcher/beam_search/hypothesis.py:132-168
//           speechcatcher/model/decoder/transformer_decoder.py:210-251 (attention part)
#include "kernels.h"

namespace scb {

enum { OUT_CONTINUE = 0, OUT_BREAK_NEW = 1, OUT_BREAK_OLD = 2 };

// ---------------------------------------------------------------- indexing helpers
__device__ __forceinline__ size_t beam_off(const SearchBuffers& sb, int buf, int s, int h) {
  return (((size_t)buf * sb.S + s) * sb.B + h);
}

// ---------------------------------------------------------------- reset
__global__ void search_reset_kernel(SearchBuffers sb, const int* __restrict__ streams, int n) {
  int i = blockIdx.x;
  if (i >= n) return;
  int s = streams[i];
  if (threadIdx.x == 0) {
    StreamCtl c;
    memset(&c, 0, sizeof(c));
    c.cur = 0; c.n_hyp = 1; c.len = 1;
    sb.ctl[s] = c;
    size_t o = beam_off(sb, 0, s, 0);
    sb.yseq[o * sb.Lcap] = sb.V - 1;          // [sos]
    sb.xpos[o * sb.Lcap] = 0;
    sb.score[o] = 0.0; sb.sc_dec[o] = 0.0; sb.sc_ctc[o] = 0.0;
    sb.ctc_s[o] = 0.f;
    if (i == 0) sb.n_active[1] = 0;            // clear the capacity-error flag
  }
}

int launch_search_reset(const SearchBuffers& sb, const int* streams, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  search_reset_kernel<<<n, 32, 0, st>>>(sb, streams, n);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- block start / end (device functions)
// extend_state (ctc_prefix_score_full.py:326-368) for every hypothesis of the current beam, or the
// initial state r^b = cumsum(blank) for the start hypothesis (ctc_prefix_score_full.py:121-134).
__device__ void start_block_ctc(const SearchBuffers& sb, int s, StreamCtl& c, int tid, int nthreads) {
  const float* x = sb.ctcx + (size_t)s * sb.Tcap * sb.V;      // blank id 0 -> column 0
  for (int h = tid; h < c.n_hyp; h += nthreads) {
    float* r = sb.ctc_r + beam_off(sb, c.cur, s, h) * sb.Tcap * 2;
    if (!c.has_ctc) {
      float acc = 0.f;
      for (int t = 0; t < c.Tb; ++t) {
        acc = (t == 0) ? x[0] : acc + x[(size_t)t * sb.V];
        r[2 * t] = kLogZero;
        r[2 * t + 1] = acc;
      }
      sb.ctc_s[beam_off(sb, c.cur, s, h)] = 0.f;
    } else {
      int start = min(max(c.ctc_T, 1), c.Tb);
      for (int t = start; t < c.Tb; ++t) {
        r[2 * t] = kLogZero;
        r[2 * t + 1] = r[2 * (t - 1) + 1] + x[(size_t)t * sb.V];
      }
    }
  }
}

// Runs with the whole CTA; thread 0 owns the control word in shared memory (`c`).
__device__ void advance_blocks(const SearchBuffers& sb, int s, StreamCtl& c) {
  // pops queue entries until one can iterate (process_idx < max_length) or the queue is empty
  while (true) {
    __syncthreads();
    if (c.blk_next >= c.blk_count) {
      if (threadIdx.x == 0) c.active = 0;
      __syncthreads();
      return;
    }
    if (threadIdx.x == 0) {
      c.Tb = sb.blkq_T[s * sb.qcap + c.blk_next];
      c.is_final = sb.blkq_final[s * sb.qcap + c.blk_next];
      c.blk_next++;
      c.iters_done = 0;
    }
    __syncthreads();
    start_block_ctc(sb, s, c, threadIdx.x, blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) { c.has_ctc = 1; c.ctc_T = c.Tb; }
    __syncthreads();
    if (c.process_idx < kMaxLength) {
      if (threadIdx.x == 0) c.active = 1;
      __syncthreads();
      return;
    }
    // loop exhausted before the first iteration: nothing to rewind (iters_done == 0)
  }
}

// Queue the decode blocks of this push (appended behind blocks still pending from earlier pushes when the
// engine defers decoding) and start the first one if the stream is idle.  One CTA per queued stream.
__global__ void search_begin_kernel(SearchBuffers sb, const int* __restrict__ q_stream, const int* __restrict__ q_n,
                                    const int* __restrict__ q_T, const int* __restrict__ q_final, int q_stride) {
  __shared__ StreamCtl c;
  const int s = q_stream[blockIdx.x];
  if (threadIdx.x == 0) {
    c = sb.ctl[s];
    const int rem = c.blk_count - c.blk_next;
    int* qT = sb.blkq_T + (size_t)s * sb.qcap;
    int* qF = sb.blkq_final + (size_t)s * sb.qcap;
    if (c.blk_next > 0)
      for (int i = 0; i < rem; ++i) { qT[i] = qT[c.blk_next + i]; qF[i] = qF[c.blk_next + i]; }
    const int n_new = q_n[blockIdx.x];
    for (int i = 0; i < n_new && rem + i < sb.qcap; ++i) {
      qT[rem + i] = q_T[blockIdx.x * q_stride + i];
      qF[rem + i] = q_final[blockIdx.x * q_stride + i];
    }
    if (rem + n_new > sb.qcap) sb.n_active[1] = 2;          // queue overflow (host keeps a bound; must not happen)
    c.blk_count = min(rem + n_new, sb.qcap);
    c.blk_next = 0;
  }
  __syncthreads();
  if (!c.active) advance_blocks(sb, s, c);
  __syncthreads();
  if (threadIdx.x == 0) sb.ctl[s] = c;
}

// ---------------------------------------------------------------- compaction of active rows
__global__ void compact_rows_kernel(SearchBuffers sb) {
  pdl_sync();
  __shared__ int s_scan[1024];
  __shared__ int s_base, s_nact;
  if (threadIdx.x == 0) { s_base = 0; s_nact = 0; }
  __syncthreads();
  for (int s0 = 0; s0 < sb.S; s0 += blockDim.x) {
    int s = s0 + threadIdx.x;
    int n = 0;
    if (s < sb.S && sb.ctl[s].active) n = sb.ctl[s].n_hyp;
    s_scan[threadIdx.x] = n;
    __syncthreads();
    for (int off = 1; off < blockDim.x; off <<= 1) {          // Hillis-Steele inclusive scan
      int v = threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    int incl = s_scan[threadIdx.x];
    int base = s_base + incl - n;
    if (s < sb.S) {
      sb.row_base[s] = n > 0 ? base : -1;
      for (int h = 0; h < n; ++h) sb.row_sh[base + h] = s * sb.B + h;
    }
    // active stream list (order of stream ids)
    int flag = n > 0 ? 1 : 0;
    __syncthreads();
    s_scan[threadIdx.x] = flag;
    __syncthreads();
    for (int off = 1; off < blockDim.x; off <<= 1) {
      int v = threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    if (flag) sb.act_streams[s_nact + s_scan[threadIdx.x] - 1] = s;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) { s_base += incl; s_nact += s_scan[threadIdx.x]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *sb.n_rows = s_base; *sb.n_active = s_nact;
    if (s_base > 0) { atomicAdd(&sb.prof[1], (unsigned long long)s_base); atomicAdd(&sb.prof[4], 1ull); }
  }
}

int launch_search_begin(const SearchBuffers& sb, const int* q_stream, const int* q_n, const int* q_T,
                        const int* q_final, int q_stride, int n_q, cudaStream_t st) {
  if (n_q > 0) {
    search_begin_kernel<<<n_q, 64, 0, st>>>(sb, q_stream, q_n, q_T, q_final, q_stride);
    SCB_LAUNCH_CHECK();
  }
  compact_rows_kernel<<<1, 1024, 0, st>>>(sb);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- decoder input: sqrt(D)*Emb[y_last] + pe[len-1]
__global__ void __launch_bounds__(128) dec_embed_kernel(SearchBuffers sb, const float* __restrict__ emb, const float* __restrict__ pe,
                                                        float* __restrict__ x, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out16) {
  pdl_sync();
  const int r = blockIdx.x;
  if (r >= *sb.n_rows) return;
  const int sh = sb.row_sh[r], s = sh / sb.B, h = sh % sb.B;
  const StreamCtl& c = sb.ctl[s];
  const int tok = sb.yseq[beam_off(sb, c.cur, s, h) * sb.Lcap + c.len - 1];
  const float scale = sqrtf((float)sb.D);
  float v[4];                                   // D <= 512 with 128 threads
  const int nv = sb.D / 128;
  float sum = 0.f;
  for (int i = 0; i < nv; ++i) {
    const int d = threadIdx.x + 128 * i;
    v[i] = emb[(size_t)tok * sb.D + d] * scale + pe[(size_t)(c.len - 1) * sb.D + d];
    x[(size_t)r * sb.D + d] = v[i];
    sum += v[i];
  }
  if (!ln_w) return;
  // fused LayerNorm (first decoder layer's norm1) -> bf16 operand of the QKV GEMM
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float mean = (red[0] + red[1] + red[2] + red[3]) / (float)sb.D;
  float q = 0.f;
  for (int i = 0; i < nv; ++i) { const float dd = v[i] - mean; q += dd * dd; }
  q = warp_sum(q);
  if (lane == 0) red[4 + warp] = q;
  __syncthreads();
  const float rstd = 1.0f / sqrtf((red[4] + red[5] + red[6] + red[7]) / (float)sb.D + 1e-12f);
  for (int i = 0; i < nv; ++i) {
    const int d = threadIdx.x + 128 * i;
    out16[(size_t)r * sb.D + d] = __float2bfloat16((v[i] - mean) * rstd * ln_w[d] + ln_b[d]);
  }
}

int launch_dec_embed(const SearchBuffers& sb, const float* emb, const float* pe, float* x, const float* ln_w,
                     const float* ln_b, __nv_bfloat16* out16, cudaStream_t st) {
  if (sb.D > 512 || sb.D % 128 != 0) { set_last_error("dec_embed: D=%d unsupported", sb.D); return -1; }
  launch_k(dec_embed_kernel, dim3(sb.S * sb.B), dim3(128), 0, st, sb, emb, pe, x, ln_w, ln_b, out16);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- decoder attention (fp32 mode / beam > 16)
constexpr int AMAXB = 20;     // beam capacity of the attention and pruning kernels

// ---------------------------------------------------------------- cross attention, shared-memory staged
// One CTA per (active stream, head).  All hypotheses of a stream see the same memory, so each K|V tile is
// loaded once (coalesced, converted to fp32), staged in shared memory and reused by every hypothesis.
// K rows are padded to DK+4 floats so that the per-position float4 row reads are bank-conflict free.
template <int DK, typename KVT>
__global__ void __launch_bounds__(128) dec_cross_attn_kernel(SearchBuffers sb, const KVT* __restrict__ xkv_layer,
                                                             const float* __restrict__ q, int ldq,
                                                             float* __restrict__ out, __nv_bfloat16* __restrict__ out16,
                                                             SplitOut so) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  const int head = blockIdx.y;
  const StreamCtl& c = sb.ctl[s];
  const int nb = c.n_hyp, row0 = sb.row_base[s], D = sb.D, npos = c.Tb;
  const int tid = threadIdx.x;
  constexpr int TILE = 4096 / DK;          // 128 positions (DK = 32) or 64 (DK = 64)
  constexpr int KS = DK + 4;
  constexpr int CH = DK / 4;               // 4-element chunks per row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Ks = reinterpret_cast<float*>(smem_raw);       // [TILE][KS]
  float* Vs = Ks + TILE * KS;                            // [TILE][DK]
  float* sc = Vs + TILE * DK;                            // [AMAXB][TILE]
  float* qs = sc + AMAXB * TILE;                         // [AMAXB][DK]
  float* sm_m = qs + AMAXB * DK;
  float* sm_l = sm_m + AMAXB;
  float* sm_f = sm_l + AMAXB;
  if (tid == 0) atomicAdd(&sb.prof[2], (unsigned long long)(2ll * npos * DK * sizeof(KVT)));
  for (int i = tid; i < nb * DK; i += 128) qs[i] = q[(size_t)(row0 + i / DK) * ldq + head * DK + i % DK];
  if (tid < AMAXB) { sm_m[tid] = -INFINITY; sm_l[tid] = 0.f; }
  const size_t row_stride = 2 * (size_t)D;
  const KVT* base = xkv_layer + (size_t)s * sb.Tcap * row_stride + head * DK;
  const float sqrt_dk = sqrtf((float)DK);
  constexpr int G = 128 / DK;
  constexpr int NACC = (AMAXB + G - 1) / G;
  const int g = tid / DK, cdim = tid % DK;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;

  for (int t0 = 0; t0 < npos; t0 += TILE) {
    const int jn = min(TILE, npos - t0);
    __syncthreads();                       // previous tile fully consumed (also orders the q / m / l init)
    for (int idx = tid; idx < jn * CH; idx += 128) {
      const int r = idx / CH, ch = idx % CH;
      const KVT* kp = base + (size_t)(t0 + r) * row_stride + ch * 4;
      float4 kf, vf;
      if constexpr (sizeof(KVT) == 4) {
        kf = *reinterpret_cast<const float4*>(kp);
        vf = *reinterpret_cast<const float4*>(kp + D);
      } else {
        const uint2 ku = *reinterpret_cast<const uint2*>(kp);
        const uint2 vu = *reinterpret_cast<const uint2*>(kp + D);
        const float2 k0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ku.x));
        const float2 k1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ku.y));
        const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vu.x));
        const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vu.y));
        kf = make_float4(k0.x, k0.y, k1.x, k1.y);
        vf = make_float4(v0.x, v0.y, v1.x, v1.y);
      }
      *reinterpret_cast<float4*>(Ks + r * KS + ch * 4) = kf;
      *reinterpret_cast<float4*>(Vs + r * DK + ch * 4) = vf;
    }
    __syncthreads();
    if (tid < TILE) {
      if (tid < jn) {
        float kreg[DK];
#pragma unroll
        for (int i = 0; i < DK; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Ks + tid * KS + i);
          kreg[i] = v.x; kreg[i + 1] = v.y; kreg[i + 2] = v.z; kreg[i + 3] = v.w;
        }
        for (int b = 0; b < nb; ++b) {
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < DK; i += 4) {        // one broadcast LDS.128 per four FMAs, same summation order
            const float4 qv = *reinterpret_cast<const float4*>(qs + b * DK + i);
            d = fmaf(qv.x, kreg[i], d); d = fmaf(qv.y, kreg[i + 1], d);
            d = fmaf(qv.z, kreg[i + 2], d); d = fmaf(qv.w, kreg[i + 3], d);
          }
          sc[b * TILE + tid] = d / sqrt_dk;
        }
      } else {
        for (int b = 0; b < nb; ++b) sc[b * TILE + tid] = -INFINITY;
      }
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int b = warp; b < nb; b += 4) {
        float m = -INFINITY;
        for (int i = lane; i < TILE; i += 32) m = fmaxf(m, sc[b * TILE + i]);
        m = warp_max(m);
        const float m_old = sm_m[b];
        const float m_new = fmaxf(m_old, m);
        float ssum = 0.f;
        for (int i = lane; i < TILE; i += 32) {
          const float e = expf(sc[b * TILE + i] - m_new);
          sc[b * TILE + i] = e;
          ssum += e;
        }
        ssum = warp_sum(ssum);
        if (lane == 0) {
          const float f = expf(m_old - m_new);
          sm_f[b] = f;
          sm_l[b] = sm_l[b] * f + ssum;
          sm_m[b] = m_new;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int b = g + i * G;
      if (b < nb) acc[i] *= sm_f[b];
    }
    {
      int jj = 0;
      for (; jj + 4 <= jn; jj += 4) {           // same accumulation order as the scalar loop
        const float v0 = Vs[jj * DK + cdim], v1 = Vs[(jj + 1) * DK + cdim];
        const float v2 = Vs[(jj + 2) * DK + cdim], v3 = Vs[(jj + 3) * DK + cdim];
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
          const int b = g + i * G;
          if (b < nb) {
            const float4 p4 = *reinterpret_cast<const float4*>(sc + b * TILE + jj);
            float a = acc[i];
            a = fmaf(p4.x, v0, a); a = fmaf(p4.y, v1, a); a = fmaf(p4.z, v2, a); a = fmaf(p4.w, v3, a);
            acc[i] = a;
          }
        }
      }
      for (; jj < jn; ++jj) {
        const float v = Vs[jj * DK + cdim];
#pragma unroll
        for (int i = 0; i < NACC; ++i) { const int b = g + i * G; if (b < nb) acc[i] = fmaf(sc[b * TILE + jj], v, acc[i]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    const int b = g + i * G;
    if (b < nb) {
      const float o = acc[i] / sm_l[b];
      if (so.base) so.put((size_t)(row0 + b), head * DK + cdim, o);
      else out[(size_t)(row0 + b) * D + head * DK + cdim] = o;
      if (out16) out16[(size_t)(row0 + b) * D + head * DK + cdim] = __float2bfloat16(o);
    }
  }
}

template <typename KVT>
static int launch_cross_t(const SearchBuffers& sb, int layer, const float* q, int ldq, float* out,
                          __nv_bfloat16* out16, cudaStream_t st, SplitOut so) {
  const int dk = sb.D / sb.H;
  const KVT* base = reinterpret_cast<const KVT*>(sb.xkv) + (size_t)layer * sb.S * sb.Tcap * 2 * sb.D;
  dim3 grid(sb.S, sb.H);
  const int tile = 4096 / dk;
  const size_t smem = sizeof(float) * ((size_t)tile * (dk + 4) + (size_t)tile * dk + (size_t)AMAXB * tile + AMAXB * dk + 3 * AMAXB);
  if (dk == 32) {
    static PerDeviceMark mk;
    if (mk.cur() < smem) { cudaFuncSetAttribute(dec_cross_attn_kernel<32, KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mk.cur() = smem; }
    launch_k(dec_cross_attn_kernel<32, KVT>, grid, dim3(128), smem, st, sb, base, q, ldq, out, out16, so);
  } else if (dk == 64) {
    static PerDeviceMark mk;
    if (mk.cur() < smem) { cudaFuncSetAttribute(dec_cross_attn_kernel<64, KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mk.cur() = smem; }
    launch_k(dec_cross_attn_kernel<64, KVT>, grid, dim3(128), smem, st, sb, base, q, ldq, out, out16, so);
  } else { set_last_error("cross attention: unsupported head dim %d", dk); return -1; }
  SCB_LAUNCH_CHECK();
  return 0;
}

int launch_dec_cross_attention(const SearchBuffers& sb, int layer, const float* q, int ldq, float* out,
                               __nv_bfloat16* out16, cudaStream_t st, SplitOut so) {
  if (sb.B > AMAXB) { set_last_error("cross attention: beam %d > %d", sb.B, AMAXB); return -1; }
  return sb.kv_bf16 ? launch_cross_t<__nv_bfloat16>(sb, layer, q, ldq, out, out16, st, so)
                    : launch_cross_t<float>(sb, layer, q, ldq, out, out16, st, so);
}

// ---------------------------------------------------------------- self attention over the KV tree, shared-memory staged
// One CTA per (active stream, head).  The hypotheses of a beam share a common ancestor chain for all
// but the last few positions (paths in the tree never re-merge), so positions [0, Lc) are loaded once
// and scored for every hypothesis like in the cross-attention kernel; the divergent tail [Lc, len) is
// processed as a flat list of (hypothesis, position) pairs, each pair contributing to its own
// hypothesis only.  This step's K|V (columns D..3D of the fused QKV GEMM) is appended first.
template <int DK, typename KVT>
__global__ void __launch_bounds__(128) dec_self_attn_kernel(SearchBuffers sb, KVT* skv_layer,
                                                            const float* __restrict__ qkv, int ldq,
                                                            float* __restrict__ out, __nv_bfloat16* __restrict__ out16,
                                                            SplitOut so) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  const int head = blockIdx.y;
  const StreamCtl& c = sb.ctl[s];
  const int nb = c.n_hyp, row0 = sb.row_base[s], D = sb.D, len = c.len, B = sb.B;
  const int tid = threadIdx.x;
  constexpr int TILE = 4096 / DK;
  constexpr int KS = DK + 4;
  constexpr int CH = DK / 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Ks = reinterpret_cast<float*>(smem_raw);
  float* Vs = Ks + TILE * KS;
  float* sc = Vs + TILE * DK;                            // [AMAXB][TILE]; row 0 doubles as the flat pair scores
  float* qs = sc + AMAXB * TILE;
  float* sm_m = qs + AMAXB * DK;
  float* sm_l = sm_m + AMAXB;
  float* sm_f = sm_l + AMAXB;
  int* s_lc = reinterpret_cast<int*>(sm_f + AMAXB);
  unsigned char* ancs = reinterpret_cast<unsigned char*>(s_lc + 4);   // [AMAXB][Lcap]
  const size_t row_stride = 2 * (size_t)D;
  KVT* store = skv_layer + (size_t)s * sb.Lcap * B * row_stride;
  if (tid == 0) {
    atomicAdd(&sb.prof[3], (unsigned long long)((long long)nb * 2ll * len * DK * sizeof(KVT)));
    *s_lc = len;
  }
  for (int i = tid; i < nb * DK; i += 128) qs[i] = qkv[(size_t)(row0 + i / DK) * ldq + head * DK + i % DK];
  if (tid < AMAXB) { sm_m[tid] = -INFINITY; sm_l[tid] = 0.f; }
  for (int i = tid; i < nb * 2 * DK; i += 128) {        // append K|V of the scored token at [len-1][b]
    const int b = i / (2 * DK), rem = i % (2 * DK), which = rem / DK, cc = rem % DK;
    const float v = qkv[(size_t)(row0 + b) * ldq + D + which * D + head * DK + cc];
    store[((size_t)(len - 1) * B + b) * row_stride + which * D + head * DK + cc] = (KVT)v;
  }
  for (int i = tid; i < nb * len; i += 128) {
    const int b = i / len, j = i % len;
    ancs[b * sb.Lcap + j] = (j == len - 1) ? (unsigned char)b : sb.anc[beam_off(sb, c.cur, s, b) * sb.Lcap + j];
  }
  __syncthreads();
  for (int j = tid; j < len; j += 128) {
    const unsigned char a0 = ancs[j];
    bool same = true;
    for (int b = 1; b < nb; ++b) same &= (ancs[b * sb.Lcap + j] == a0);
    if (!same) atomicMin(s_lc, j);
  }
  __syncthreads();
  const int Lc = *s_lc;
  const float sqrt_dk = sqrtf((float)DK);
  constexpr int G = 128 / DK;
  constexpr int NACC = (AMAXB + G - 1) / G;
  const int g = tid / DK, cdim = tid % DK;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;

  auto load_row = [&](int r, int ch, const KVT* kp) {
    float4 kf, vf;
    if constexpr (sizeof(KVT) == 4) {
      kf = *reinterpret_cast<const float4*>(kp);
      vf = *reinterpret_cast<const float4*>(kp + D);
    } else {
      const uint2 ku = *reinterpret_cast<const uint2*>(kp);
      const uint2 vu = *reinterpret_cast<const uint2*>(kp + D);
      const float2 k0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ku.x));
      const float2 k1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ku.y));
      const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vu.x));
      const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vu.y));
      kf = make_float4(k0.x, k0.y, k1.x, k1.y);
      vf = make_float4(v0.x, v0.y, v1.x, v1.y);
    }
    *reinterpret_cast<float4*>(Ks + r * KS + ch * 4) = kf;
    *reinterpret_cast<float4*>(Vs + r * DK + ch * 4) = vf;
  };

  // ---------------- common ancestor chain: one row per position, shared by every hypothesis
  for (int t0 = 0; t0 < Lc; t0 += TILE) {
    const int jn = min(TILE, Lc - t0);
    __syncthreads();
    for (int idx = tid; idx < jn * CH; idx += 128) {
      const int r = idx / CH, ch = idx % CH, j = t0 + r;
      load_row(r, ch, store + ((size_t)j * B + ancs[j]) * row_stride + head * DK + ch * 4);
    }
    __syncthreads();
    if (tid < TILE) {
      if (tid < jn) {
        float kreg[DK];
#pragma unroll
        for (int i = 0; i < DK; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Ks + tid * KS + i);
          kreg[i] = v.x; kreg[i + 1] = v.y; kreg[i + 2] = v.z; kreg[i + 3] = v.w;
        }
        for (int b = 0; b < nb; ++b) {
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < DK; i += 4) {        // one broadcast LDS.128 per four FMAs, same summation order
            const float4 qv = *reinterpret_cast<const float4*>(qs + b * DK + i);
            d = fmaf(qv.x, kreg[i], d); d = fmaf(qv.y, kreg[i + 1], d);
            d = fmaf(qv.z, kreg[i + 2], d); d = fmaf(qv.w, kreg[i + 3], d);
          }
          sc[b * TILE + tid] = d / sqrt_dk;
        }
      } else {
        for (int b = 0; b < nb; ++b) sc[b * TILE + tid] = -INFINITY;
      }
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int b = warp; b < nb; b += 4) {
        float m = -INFINITY;
        for (int i = lane; i < TILE; i += 32) m = fmaxf(m, sc[b * TILE + i]);
        m = warp_max(m);
        const float m_old = sm_m[b], m_new = fmaxf(m_old, m);
        float ssum = 0.f;
        for (int i = lane; i < TILE; i += 32) {
          const float e = expf(sc[b * TILE + i] - m_new);
          sc[b * TILE + i] = e;
          ssum += e;
        }
        ssum = warp_sum(ssum);
        if (lane == 0) {
          const float f = expf(m_old - m_new);
          sm_f[b] = f; sm_l[b] = sm_l[b] * f + ssum; sm_m[b] = m_new;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NACC; ++i) { const int b = g + i * G; if (b < nb) acc[i] *= sm_f[b]; }
    {
      int jj = 0;
      for (; jj + 4 <= jn; jj += 4) {           // same accumulation order as the scalar loop
        const float v0 = Vs[jj * DK + cdim], v1 = Vs[(jj + 1) * DK + cdim];
        const float v2 = Vs[(jj + 2) * DK + cdim], v3 = Vs[(jj + 3) * DK + cdim];
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
          const int b = g + i * G;
          if (b < nb) {
            const float4 p4 = *reinterpret_cast<const float4*>(sc + b * TILE + jj);
            float a = acc[i];
            a = fmaf(p4.x, v0, a); a = fmaf(p4.y, v1, a); a = fmaf(p4.z, v2, a); a = fmaf(p4.w, v3, a);
            acc[i] = a;
          }
        }
      }
      for (; jj < jn; ++jj) {
        const float v = Vs[jj * DK + cdim];
#pragma unroll
        for (int i = 0; i < NACC; ++i) { const int b = g + i * G; if (b < nb) acc[i] = fmaf(sc[b * TILE + jj], v, acc[i]); }
      }
    }
  }
  // ---------------- divergent tail: flat list of (hypothesis, position) pairs, ordered by hypothesis
  const int n_div = len - Lc;
  const int n_pairs = nb * n_div;
  for (int p0 = 0; p0 < n_pairs; p0 += TILE) {
    const int pn = min(TILE, n_pairs - p0);
    __syncthreads();
    for (int idx = tid; idx < pn * CH; idx += 128) {
      const int r = idx / CH, ch = idx % CH, u = p0 + r, b = u / n_div, j = Lc + u % n_div;
      load_row(r, ch, store + ((size_t)j * B + ancs[b * sb.Lcap + j]) * row_stride + head * DK + ch * 4);
    }
    __syncthreads();
    if (tid < pn) {
      const int b = (p0 + tid) / n_div;
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < DK; i += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(Ks + tid * KS + i);
        const float4 qv = *reinterpret_cast<const float4*>(qs + b * DK + i);
        d = fmaf(qv.x, kv.x, d); d = fmaf(qv.y, kv.y, d); d = fmaf(qv.z, kv.z, d); d = fmaf(qv.w, kv.w, d);
      }
      sc[tid] = d / sqrt_dk;
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int b = warp; b < nb; b += 4) {
        const int lo = max(b * n_div, p0) - p0, hi = min((b + 1) * n_div, p0 + pn) - p0;
        if (lo >= hi) { if (lane == 0) sm_f[b] = 1.f; continue; }
        float m = -INFINITY;
        for (int i = lo + lane; i < hi; i += 32) m = fmaxf(m, sc[i]);
        m = warp_max(m);
        const float m_old = sm_m[b], m_new = fmaxf(m_old, m);
        float ssum = 0.f;
        for (int i = lo + lane; i < hi; i += 32) {
          const float e = expf(sc[i] - m_new);
          sc[i] = e;
          ssum += e;
        }
        ssum = warp_sum(ssum);
        if (lane == 0) {
          const float f = expf(m_old - m_new);
          sm_f[b] = f; sm_l[b] = sm_l[b] * f + ssum; sm_m[b] = m_new;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int b = g + i * G;
      if (b < nb) {
        const int lo = max(b * n_div, p0) - p0, hi = min((b + 1) * n_div, p0 + pn) - p0;
        float a = acc[i] * sm_f[b];
        for (int jj = lo; jj < hi; ++jj) a = fmaf(sc[jj], Vs[jj * DK + cdim], a);
        acc[i] = a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    const int b = g + i * G;
    if (b < nb) {
      const float o = acc[i] / sm_l[b];
      if (so.base) so.put((size_t)(row0 + b), head * DK + cdim, o);
      else out[(size_t)(row0 + b) * D + head * DK + cdim] = o;
      if (out16) out16[(size_t)(row0 + b) * D + head * DK + cdim] = __float2bfloat16(o);
    }
  }
}

template <typename KVT>
static int launch_self_t(const SearchBuffers& sb, int layer, const float* qkv, int ldq, float* out,
                         __nv_bfloat16* out16, cudaStream_t st, SplitOut so) {
  const int dk = sb.D / sb.H;
  KVT* base = reinterpret_cast<KVT*>(sb.skv) + (size_t)layer * sb.S * sb.Lcap * sb.B * 2 * sb.D;
  dim3 grid(sb.S, sb.H);
  const int tile = 4096 / dk;
  const size_t smem = sizeof(float) * ((size_t)tile * (dk + 4) + (size_t)tile * dk + (size_t)AMAXB * tile + AMAXB * dk + 3 * AMAXB) +
                      16 + (size_t)AMAXB * sb.Lcap;
  if (dk == 32) {
    static PerDeviceMark mk;
    if (mk.cur() < smem) { cudaFuncSetAttribute(dec_self_attn_kernel<32, KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mk.cur() = smem; }
    launch_k(dec_self_attn_kernel<32, KVT>, grid, dim3(128), smem, st, sb, base, qkv, ldq, out, out16, so);
  } else if (dk == 64) {
    static PerDeviceMark mk;
    if (mk.cur() < smem) { cudaFuncSetAttribute(dec_self_attn_kernel<64, KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mk.cur() = smem; }
    launch_k(dec_self_attn_kernel<64, KVT>, grid, dim3(128), smem, st, sb, base, qkv, ldq, out, out16, so);
  } else { set_last_error("self attention: unsupported head dim %d", dk); return -1; }
  SCB_LAUNCH_CHECK();
  return 0;
}

int launch_dec_self_attention(const SearchBuffers& sb, int layer, const float* qkv, int ldq, float* out,
                              __nv_bfloat16* out16, cudaStream_t st, SplitOut so) {
  if (sb.B > AMAXB) { set_last_error("self attention: beam %d > %d", sb.B, AMAXB); return -1; }
  return sb.kv_bf16 ? launch_self_t<__nv_bfloat16>(sb, layer, qkv, ldq, out, out16, st, so)
                    : launch_self_t<float>(sb, layer, qkv, ldq, out, out16, st, so);
}

// ---------------------------------------------------------------- log-softmax + pre-beam top-40
// One warp per row (V = 1024 -> 32 values per lane).  Writes log-probs in place and the 40 best ids of
// w_dec * logp in descending order (lowest index first on ties).   (beam_search.py:121-154)
constexpr int VMAX_PER_LANE = 32;

__global__ void __launch_bounds__(128) logsoftmax_prebeam_kernel(SearchBuffers sb, float* __restrict__ logits) {
  pdl_sync();
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= *sb.n_rows) return;
  const int lane = threadIdx.x & 31, V = sb.V;
  float* row = logits + (size_t)r * V;
  float v[VMAX_PER_LANE];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < VMAX_PER_LANE; ++i) { v[i] = row[lane + 32 * i]; m = fmaxf(m, v[i]); }
  m = warp_max(m);
  float ssum = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX_PER_LANE; ++i) ssum += expf(v[i] - m);
  ssum = warp_sum(ssum);
  const float lse = logf(ssum);
#pragma unroll
  for (int i = 0; i < VMAX_PER_LANE; ++i) {
    v[i] = (v[i] - m) - lse;
    row[lane + 32 * i] = v[i];
    v[i] = __fmul_rn(sb.w_dec, v[i]);             // full-scorer score used for the pre-beam
  }
  // 40 rounds of warp arg-max
  float lbest = -INFINITY; int lidx = 0;
  auto rescan = [&]() {
    lbest = -INFINITY; lidx = 0;
#pragma unroll
    for (int i = 0; i < VMAX_PER_LANE; ++i) if (v[i] > lbest) { lbest = v[i]; lidx = i; }
  };
  rescan();
  for (int k = 0; k < kPreBeam; ++k) {
    float bv = lbest; int bi = lane + 32 * lidx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) sb.pre_ids[(size_t)r * kPreBeam + k] = bi;
    if ((bi & 31) == lane) {
      const int slot = bi >> 5;
#pragma unroll
      for (int i = 0; i < VMAX_PER_LANE; ++i) if (i == slot) v[i] = -INFINITY;
      rescan();
    }
  }
}

int launch_logsoftmax_prebeam(const SearchBuffers& sb, float* logits, cudaStream_t st) {
  if (sb.V != 1024) { set_last_error("logsoftmax_prebeam: V=%d unsupported (1024 only)", sb.V); return -1; }
  launch_k(logsoftmax_prebeam_kernel, dim3(cdiv(sb.S * sb.B, 4)), dim3(128), 0, st, sb, logits);
  SCB_LAUNCH_CHECK();
  return 0;
}

// rows with row_flag != 0 get an in-place log-softmax (CTC store rows t < 24)
__global__ void __launch_bounds__(128) logsoftmax_rows_kernel(float* __restrict__ x, const int64_t* __restrict__ row_off,
                                                              const int* __restrict__ row_flag, int rows, int V) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows || !row_flag[r]) return;
  const int lane = threadIdx.x & 31;
  float* row = x + row_off[r];
  float m = -INFINITY;
  for (int i = lane; i < V; i += 32) m = fmaxf(m, row[i]);
  m = warp_max(m);
  float ssum = 0.f;
  for (int i = lane; i < V; i += 32) ssum += expf(row[i] - m);
  const float lse = logf(warp_sum(ssum));
  for (int i = lane; i < V; i += 32) row[i] = (row[i] - m) - lse;
}

int launch_logsoftmax_rows(float* x, const int64_t* row_off, const int* row_flag, int rows, int V, cudaStream_t st) {
  if (rows <= 0) return 0;
  logsoftmax_rows_kernel<<<cdiv(rows, 4), 128, 0, st>>>(x, row_off, row_flag, rows, V);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- CTC prefix scorer (ctc_prefix_score_full.py:88-291)
// 64 threads per row, thread k < 40 owns candidate ids[row][k].  The forward variables r are NOT
// materialised for all 40 candidates (the reference writes a (T,2,n_bh,40) tensor every step); only the
// column each surviving hypothesis inherits is rebuilt after pruning (ctc_state_update_kernel).
struct CtcRec { float rn, rb; };

__device__ __forceinline__ void ctc_forward_column(const float* __restrict__ x, int V, const float* __restrict__ rprev,
                                                   int T, int L, int c, bool same_as_last, float* __restrict__ r_out,
                                                   float& psi) {
  // r_out (if non-null): [T][2].  Returns psi = logsumexp_t(phi[t-1] + x[t,c]) (+) r[start-1, n].
  const int start = min(max(L, 1), T);
  float rn = kLogZero, rb = kLogZero;
  if (L == 0) rn = x[c];                                    // r[0, n] = x[0, c]
  if (r_out) {
    for (int t = 0; t < start; ++t) { r_out[2 * t] = (t == 0 && L == 0) ? rn : kLogZero; r_out[2 * t + 1] = kLogZero; }
  }
  // online logsumexp over {r[start-1, n]} U {phi[t-1] + x[t, c]}
  float mx = rn, sum = 1.f;
  float p_n = rprev[2 * (start - 1)], p_b = rprev[2 * (start - 1) + 1];
  // The recursion over t is sequential, but its inputs (x[t, c], x[t, blank], r_prev[t]) are not: they are
  // fetched eight time steps ahead into registers (two alternating batches) so that the dependent chain only
  // contains the log-sum-exp arithmetic.  Same operations in the same order as the plain loop (bit-identical).
  constexpr int CB = 8;
  struct Batch { float xc[CB], xb[CB]; float2 rp[CB]; };
  auto load = [&](Batch& b, int t0) {
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      const int t = t0 + i;
      if (t < T) {
        b.xc[i] = x[(size_t)t * V + c];
        b.xb[i] = x[(size_t)t * V];
        b.rp[i] = *reinterpret_cast<const float2*>(rprev + 2 * t);
      }
    }
  };
  auto process = [&](const Batch& b, int t0) {
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      const int t = t0 + i;
      if (t < T) {
        const float phi = same_as_last ? p_b : lse2(p_n, p_b);
        const float nn = lse2(rn, phi) + b.xc[i];
        const float nb = lse2(rn, rb) + b.xb[i];
        rn = nn; rb = nb;
        if (r_out) *reinterpret_cast<float2*>(r_out + 2 * t) = make_float2(rn, rb);
        const float e = phi + b.xc[i];
        if (e > mx) { sum = sum * expf(mx - e) + 1.f; mx = e; } else { sum += expf(e - mx); }
        p_n = b.rp[i].x; p_b = b.rp[i].y;
      }
    }
  };
  Batch A, B;
  load(A, start);
  for (int t0 = start; t0 < T; t0 += 2 * CB) {
    load(B, t0 + CB);
    process(A, t0);
    load(A, t0 + 2 * CB);
    process(B, t0 + CB);
  }
  psi = logf(sum) + mx;
}

__global__ void __launch_bounds__(64) ctc_prefix_kernel(SearchBuffers sb) {
  pdl_sync();
  const int r = blockIdx.x;
  if (r >= *sb.n_rows) return;
  const int sh = sb.row_sh[r], s = sh / sb.B, h = sh % sb.B;
  const StreamCtl& c = sb.ctl[s];
  const float* x = sb.ctcx + (size_t)s * sb.Tcap * sb.V;
  const float* rprev = sb.ctc_r + beam_off(sb, c.cur, s, h) * sb.Tcap * 2;
  const int L = c.len - 1, T = c.Tb;
  const int last = sb.yseq[beam_off(sb, c.cur, s, h) * sb.Lcap + c.len - 1];
  const int k = threadIdx.x;
  if (k == kPreBeam + 1)   // algorithmic bytes of one row = 4*T*(3K+2) + 4*V
    atomicAdd(&sb.prof[0], (unsigned long long)(4ll * T * (3 * kPreBeam + 2) + 4ll * sb.V));
  if (k < kPreBeam) {
    const int tok = sb.pre_ids[(size_t)r * kPreBeam + k];
    float psi;
    ctc_forward_column(x, sb.V, rprev, T, L, tok, tok == last, nullptr, psi);
    sb.psi[(size_t)r * kPreBeam + k] = psi;
  } else if (k == kPreBeam) {
    sb.psi_eos[r] = lse2(rprev[2 * (T - 1)], rprev[2 * (T - 1) + 1]);   // r_sum[T-1]
  }
}

int launch_ctc_prefix(const SearchBuffers& sb, cudaStream_t st) {
  launch_k(ctc_prefix_kernel, dim3(sb.S * sb.B), dim3(64), 0, st, sb);
  SCB_LAUNCH_CHECK();
  return 0;
}

// The same recursion as a stand-alone operator over explicit tensors (C ABI sc_ctc_prefix_step; direct parity tests of
// SURVEY.md row K7 against oracle/ctc_prefix.py): one CTA per hypothesis, thread k < 40 = candidate k.
__global__ void __launch_bounds__(64) ctc_prefix_op_kernel(const float* __restrict__ x, int T, int V,
                                                           const float* __restrict__ r_prev, const int* __restrict__ last_tok,
                                                           int L, const int* __restrict__ ids, float* __restrict__ psi,
                                                           float* __restrict__ psi_eos, float* __restrict__ r_new) {
  const int h = blockIdx.x, k = threadIdx.x;
  const float* rp = r_prev + (size_t)h * T * 2;
  if (k < kPreBeam) {
    const int tok = ids[(size_t)h * kPreBeam + k];
    float* ro = r_new ? r_new + ((size_t)h * kPreBeam + k) * T * 2 : nullptr;
    float ps;
    ctc_forward_column(x, V, rp, T, L, tok, tok == last_tok[h], ro, ps);
    psi[(size_t)h * kPreBeam + k] = ps;
  } else if (k == kPreBeam) {
    psi_eos[h] = lse2(rp[2 * (T - 1)], rp[2 * (T - 1) + 1]);
  }
}

int launch_ctc_prefix_op(const float* x, int T, int V, const float* r_prev, const int* last_tok, int L, const int* ids,
                         int n_hyp, float* psi, float* psi_eos, float* r_new, cudaStream_t st) {
  if (n_hyp <= 0) return 0;
  if (T < 1 || V < 2 || L < 0) { set_last_error("ctc_prefix_op: bad shape T=%d V=%d L=%d", T, V, L); return -1; }
  ctc_prefix_op_kernel<<<n_hyp, 64, 0, st>>>(x, T, V, r_prev, last_tok, L, ids, psi, psi_eos, r_new);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- combine + per-hypothesis top-B (beam_search.py:145-183, 723)
// One warp per row.  Only the 40 candidates and <eos> can carry a real CTC score; every other token
// scores w_ctc * (logzero - s_prev) and can never reach the top-B (B <= 20 < 39).
__global__ void __launch_bounds__(128) combine_topk_kernel(SearchBuffers sb, const float* __restrict__ logp) {
  pdl_sync();
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= *sb.n_rows) return;
  const int lane = threadIdx.x & 31;
  const int sh = sb.row_sh[r], s = sh / sb.B, h = sh % sb.B;
  const StreamCtl& c = sb.ctl[s];
  const float s_prev = sb.ctc_s[beam_off(sb, c.cur, s, h)];
  const int eos = sb.V - 1;
  const int* ids = sb.pre_ids + (size_t)r * kPreBeam;
  // lane owns candidates lane and lane+32 (slot 40 = <eos> when it is not among the 40)
  float val[2], dec[2], ctc[2], psi[2]; int tok[2];
  bool eos_in = false;
  for (int k = 0; k < kPreBeam; ++k) eos_in |= (ids[k] == eos);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int k = lane + 32 * q;
    val[q] = -INFINITY; tok[q] = -1; dec[q] = 0.f; ctc[q] = 0.f; psi[q] = kLogZero;
    int t = -1;
    if (k < kPreBeam) t = ids[k];
    else if (k == kPreBeam && !eos_in) t = eos;
    if (t >= 0) {
      float ps;
      if (t == 0) ps = kLogZero;                        // blank is forced to logzero (:288)
      else if (t == eos) ps = sb.psi_eos[r];            // eos = r_sum[T-1] (:284-285)
      else ps = sb.psi[(size_t)r * kPreBeam + k];
      const float d = logp[(size_t)r * sb.V + t];
      const float cs = ps - s_prev;
      tok[q] = t; dec[q] = d; ctc[q] = cs; psi[q] = ps;
      val[q] = __fadd_rn(__fmul_rn(sb.w_dec, d), __fmul_rn(sb.w_ctc, cs));
    }
  }
  // inherit-column token for select_state (scorers.py:418-425): the token itself if it was scored,
  // otherwise candidate 0 of this hypothesis
  const int col0 = ids[0];
  for (int i = 0; i < sb.B; ++i) {
    float bv = val[0]; int bq = 0;
    if (val[1] > bv) { bv = val[1]; bq = 1; }
    int bk = lane + 32 * bq;                             // candidate slot index (ties -> lowest token id)
    int bt = tok[bq];
    float cv = bv; int ck = bk, ct = bt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, cv, o);
      int ok = __shfl_xor_sync(0xffffffffu, ck, o);
      int ot = __shfl_xor_sync(0xffffffffu, ct, o);
      if (ov > cv || (ov == cv && ot >= 0 && (ct < 0 || ot < ct))) { cv = ov; ck = ok; ct = ot; }
    }
    if ((ck & 31) == lane) {
      const int q = ck >> 5;
      const size_t o = (size_t)r * sb.B + i;
      sb.cand_val[o] = val[q];
      sb.cand_tok[o] = tok[q];
      sb.cand_dec[o] = dec[q];
      sb.cand_ctc[o] = ctc[q];
      sb.cand_psi[o] = psi[q];
      sb.cand_col[o] = (ck < kPreBeam) ? tok[q] : col0;
      val[q] = -INFINITY;
    }
    __syncwarp();
  }
}

int launch_combine_topk(const SearchBuffers& sb, const float* logp, cudaStream_t st) {
  launch_k(combine_topk_kernel, dim3(cdiv(sb.S * sb.B, 4)), dim3(128), 0, st, sb, logp);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- global prune + new beam + stop decision
// One CTA (128 threads) per active stream.  (beam_search.py:721-809, hypothesis.py:132-168)
constexpr int PMAXC = 13;   // ceil(20*20 / 32)

__global__ void __launch_bounds__(128) beam_prune_kernel(SearchBuffers sb) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  StreamCtl& c = sb.ctl[s];
  const int B = sb.B, nb = c.n_hyp, row0 = sb.row_base[s], cur = c.cur, nxt = cur ^ 1, len = c.len;
  const int tid = threadIdx.x, eos = sb.V - 1;
  __shared__ int sel[AMAXB];        // chosen candidate index (parent * B + rank)
  __shared__ int s_tok[AMAXB];
  __shared__ int s_flag;
  if (tid < 32) {
    const int ncand = nb * B;
    double v[PMAXC];
#pragma unroll
    for (int i = 0; i < PMAXC; ++i) {
      int ci = tid + 32 * i;
      v[i] = -INFINITY;
      if (ci < ncand) {
        int p = ci / B, j = ci % B;
        v[i] = sb.score[beam_off(sb, cur, s, p)] + (double)sb.cand_val[(size_t)(row0 + p) * B + j];
      }
    }
    for (int k = 0; k < B; ++k) {
      double bv = -INFINITY; int bi = 1 << 30;
#pragma unroll
      for (int i = 0; i < PMAXC; ++i) {
        int ci = tid + 32 * i;
        if (v[i] > bv || (v[i] == bv && ci < bi && v[i] > -INFINITY)) { bv = v[i]; bi = ci; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (tid == 0) sel[k] = bi;
      if ((bi & 31) == tid) {
        const int slot = bi >> 5;
#pragma unroll
        for (int i = 0; i < PMAXC; ++i) if (i == slot) v[i] = -INFINITY;
      }
    }
  }
  if (tid == 0) s_flag = 0;
  __syncthreads();
  // scalar fields of the new hypotheses
  if (tid < B) {
    const int ci = sel[tid], p = ci / B, j = ci % B;
    const size_t co = (size_t)(row0 + p) * B + j, po = beam_off(sb, cur, s, p), no = beam_off(sb, nxt, s, tid);
    const int tok = sb.cand_tok[co];
    s_tok[tid] = tok;
    sb.score[no] = sb.score[po] + (double)sb.cand_val[co];
    sb.sc_dec[no] = sb.sc_dec[po] + (double)sb.cand_dec[co];
    sb.sc_ctc[no] = sb.sc_ctc[po] + (double)sb.cand_ctc[co];
    sb.ctc_s[no] = sb.cand_psi[co];
    sb.new_parent[s * B + tid] = p;
    sb.new_col[s * B + tid] = sb.cand_col[co];
    sb.yseq[no * sb.Lcap + len] = tok;
    sb.xpos[no * sb.Lcap + len] = c.Tb - 1;
    sb.anc[no * sb.Lcap + len - 1] = (unsigned char)p;
  }
  __syncthreads();
  // copy the parents' histories
  for (int i = tid; i < B * len; i += blockDim.x) {
    const int hn = i / len, j = i % len;
    const int p = sel[hn] / B;
    const size_t po = beam_off(sb, cur, s, p) * sb.Lcap, no = beam_off(sb, nxt, s, hn) * sb.Lcap;
    sb.yseq[no + j] = sb.yseq[po + j];
    sb.xpos[no + j] = sb.xpos[po + j];
    if (j < len - 1) sb.anc[no + j] = sb.anc[po + j];
  }
  // BBD repetition test on the new beam (beam_search.py:466-505): last token occurs in yseq[1:-1]
  if (sb.use_bbd && !c.is_final) {
    for (int i = tid; i < B * len; i += blockDim.x) {
      const int hn = i / len, j = i % len;
      const int last = s_tok[hn];
      if (j >= 1 && last != eos) {
        const int p = sel[hn] / B;
        if (sb.yseq[beam_off(sb, cur, s, p) * sb.Lcap + j] == last) s_flag = 1;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    bool any_eos = false, all_eos = true;
    for (int i = 0; i < B; ++i) { bool e = s_tok[i] == eos; any_eos |= e; all_eos &= e; }
    const bool best_eos = s_tok[0] == eos;
    int outcome = OUT_CONTINUE;
    if (any_eos && (!c.is_final || best_eos)) outcome = OUT_BREAK_NEW;
    if (outcome == OUT_CONTINUE && sb.use_bbd && !c.is_final && s_flag) outcome = OUT_BREAK_OLD;
    if (outcome == OUT_CONTINUE && c.is_final && all_eos) outcome = OUT_BREAK_NEW;
    c.outcome = outcome;
    c.steps_total++;
  }
}

int launch_beam_prune(const SearchBuffers& sb, cudaStream_t st) {
  if (sb.B > AMAXB) { set_last_error("beam_prune: beam %d > %d", sb.B, AMAXB); return -1; }
  launch_k(beam_prune_kernel, dim3(sb.S), dim3(128), 0, st, sb);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- CTC state of the new beam (select_state, scorers.py:382-431)
// One thread per (active stream, new hypothesis): rebuild the inherited forward column into the other
// beam buffer.  Wasted (but harmless) when the step is later discarded by a rewind.
__global__ void __launch_bounds__(32) ctc_state_update_kernel(SearchBuffers sb) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  const StreamCtl& c = sb.ctl[s];
  const int hn = threadIdx.x;
  if (hn >= sb.B) return;
  const int p = sb.new_parent[s * sb.B + hn], col = sb.new_col[s * sb.B + hn];
  const float* x = sb.ctcx + (size_t)s * sb.Tcap * sb.V;
  const float* rprev = sb.ctc_r + beam_off(sb, c.cur, s, p) * sb.Tcap * 2;
  float* rout = sb.ctc_r + beam_off(sb, c.cur ^ 1, s, hn) * sb.Tcap * 2;
  const int last = sb.yseq[beam_off(sb, c.cur, s, p) * sb.Lcap + c.len - 1];
  float psi;
  ctc_forward_column(x, sb.V, rprev, c.Tb, c.len - 1, col, col == last, rout, psi);
}

int launch_ctc_state_update(const SearchBuffers& sb, cudaStream_t st) {
  launch_k(ctc_state_update_kernel, dim3(sb.S), dim3(32), 0, st, sb);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- commit the step: block end / rewind / next block
__global__ void __launch_bounds__(64) step_commit_kernel(SearchBuffers sb) {
  pdl_sync();
  if ((int)blockIdx.x >= *sb.n_active) return;
  const int s = sb.act_streams[blockIdx.x];
  __shared__ StreamCtl c;
  __shared__ int ended;
  if (threadIdx.x == 0) {
    c = sb.ctl[s];
    ended = 0;
    const int outcome = c.outcome;
    if (outcome == OUT_CONTINUE) {
      c.cur ^= 1; c.len += 1; c.n_hyp = sb.B; c.ctc_T = c.Tb; c.has_ctc = 1;
      c.iters_done += 1; c.process_idx += 1;
      if (c.process_idx >= kMaxLength) {                  // while-loop exhausted
        ended = 1;
        if (c.process_idx > 1 && c.iters_done >= 1) c.process_idx -= 1;   // rewind to an identical snapshot
      }
    } else {
      ended = 1;
      const bool rewind = c.process_idx > 1 && c.iters_done >= 1;         // beam_search.py:827-836
      if (rewind) c.process_idx -= 1;                                     // beam = snapshot = current buffer
      else if (outcome == OUT_BREAK_NEW) { c.cur ^= 1; c.len += 1; c.n_hyp = sb.B; c.ctc_T = c.Tb; c.has_ctc = 1; }
    }
    if (c.len >= sb.Lcap - 1) {       // token capacity reached: stop this stream and raise the error flag
      c.blk_next = c.blk_count; ended = 1; sb.n_active[1] = 1;
    }
  }
  __syncthreads();
  if (ended) advance_blocks(sb, s, c);
  __syncthreads();
  if (threadIdx.x == 0) sb.ctl[s] = c;
}

int launch_step_finish(const SearchBuffers& sb, cudaStream_t st) {
  launch_k(step_commit_kernel, dim3(sb.S), dim3(64), 0, st, sb);
  SCB_LAUNCH_CHECK();
  launch_k(compact_rows_kernel, dim3(1), dim3(1024), 0, st, sb);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
