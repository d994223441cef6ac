// Contextual-block streaming encoder: conv2d sub-sampling helpers, block assembly with positional
// encoding and context slots, block-wise masked self-attention, context hand-over between blocks
// and layers, and the output stitch + final LayerNorm.  (GEMMs and row LayerNorms live in
// kernels_gemm*.cu.)
//
// Replaces: speechcatcher/model/encoder/subsampling.py:71-106
//           speechcatcher/model/encoder/contextual_block_transformer_encoder.py:278-419, 500-528
//           speechcatcher/model/encoder/contextual_block_encoder_layer.py:178-271
//           speechcatcher/model/attention/multi_head_attention.py:92-133 (masked vanilla attention)
#include <stdlib.h>
#include "kernels.h"

namespace scb {

constexpr int NMEL = 80, F1 = 39, F2 = 19;

// ---------------------------------------------------------------- conv1 (1 -> D channels, 3x3, stride 2) + ReLU
// Output is channels-last h1[stream][t1][f1][c] so that conv2 becomes a GEMM whose A rows are nine
// contiguous D-vectors (implicit GEMM, see launch_conv2_rows).
__global__ void conv1_kernel(const float* __restrict__ featbuf, int feat_cap, const float* __restrict__ w1,
                             const float* __restrict__ b1, float* __restrict__ h1, int t1_cap,
                             const SubDesc* __restrict__ desc, int D) {
  const SubDesc d = desc[blockIdx.y];
  const int t1 = blockIdx.x;
  if (t1 >= d.t1) return;
  __shared__ float xin[3][NMEL];
  const float* src = featbuf + ((size_t)d.stream * feat_cap + 2 * t1) * NMEL;
  for (int i = threadIdx.x; i < 3 * NMEL; i += blockDim.x) xin[i / NMEL][i % NMEL] = src[i];
  __syncthreads();
  float* dst = h1 + ((size_t)d.stream * t1_cap + t1) * F1 * D;
  for (int o = threadIdx.x; o < F1 * D; o += blockDim.x) {
    int c = o % D, f1 = o / D;
    const float* w = w1 + c * 9;
    float acc = b1[c];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt)
#pragma unroll
      for (int kf = 0; kf < 3; ++kf) acc = fmaf(w[kt * 3 + kf], xin[kt][2 * f1 + kf], acc);
    dst[o] = fmaxf(acc, 0.f);
  }
}

int launch_conv1(const float* featbuf, int feat_cap, const float* w1, const float* b1, float* h1, int t1_cap,
                 const SubDesc* desc, int n_desc, int D, cudaStream_t st) {
  if (n_desc <= 0) return 0;
  dim3 grid(t1_cap, n_desc);
  conv1_kernel<<<grid, 256, 0, st>>>(featbuf, feat_cap, w1, b1, h1, t1_cap, desc, D);
  SCB_LAUNCH_CHECK();
  return 0;
}

// bf16 mode: conv1 + ReLU fused with the im2col of conv2.  One CTA per (stream, t2), thread = channel c.
// A16[(row0 + t2) * 19 + f2][(kt * 3 + kf) * D + c] = bf16(relu(conv1(t1 = 2 t2 + kt, f1 = 2 f2 + kf, c))),
// i.e. the dense K-major operand of the conv2 tensor-core GEMM (K = 9 D).  conv1 is only 9 MACs per
// value, so recomputing it for the (at most four) windows a value belongs to is cheaper than a round trip.
__global__ void __launch_bounds__(256) conv1_im2col_bf16_kernel(const float* __restrict__ featbuf, int feat_cap,
                                                                const float* __restrict__ w1, const float* __restrict__ b1,
                                                                __nv_bfloat16* __restrict__ A16,
                                                                const SubDesc* __restrict__ desc, int D) {
  const SubDesc d = desc[blockIdx.y];
  const int t2 = blockIdx.x;
  if (t2 >= d.t2) return;
  __shared__ float xin[7][NMEL];
  const float* src = featbuf + ((size_t)d.stream * feat_cap + 4 * t2) * NMEL;
  for (int i = threadIdx.x; i < 7 * NMEL; i += blockDim.x) xin[i / NMEL][i % NMEL] = src[i];
  __syncthreads();
  const int c = threadIdx.x;
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = w1[c * 9 + i];
  const float bias = b1[c];
  __nv_bfloat16* dst = A16 + ((size_t)(d.row0 + t2) * F2) * 9 * D;
  for (int f2 = 0; f2 < F2; ++f2) {
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
#pragma unroll
      for (int kf = 0; kf < 3; ++kf) {
        const int f1 = 2 * f2 + kf;            // conv1 output column; its inputs are feature bins 2 f1 .. 2 f1 + 2
        float acc = bias;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) acc = fmaf(w[a * 3 + b], xin[2 * kt + a][2 * f1 + b], acc);
        dst[((size_t)f2 * 9 + kt * 3 + kf) * D + c] = __float2bfloat16(fmaxf(acc, 0.f));
      }
    }
  }
}

int launch_conv1_im2col_bf16(const float* featbuf, int feat_cap, const float* w1, const float* b1, __nv_bfloat16* A16,
                             int t2_cap, const SubDesc* desc, int n_desc, int D, cudaStream_t st) {
  if (n_desc <= 0) return 0;
  if (D != 256) { set_last_error("conv1_im2col: D=%d unsupported", D); return -1; }
  dim3 grid(t2_cap, n_desc);
  conv1_im2col_bf16_kernel<<<grid, 256, 0, st>>>(featbuf, feat_cap, w1, b1, A16, desc, D);
  SCB_LAUNCH_CHECK();
  return 0;
}

// Row tables for the conv2 implicit GEMM (rows = (stream, t2, f2)) and for the output projection
// (rows = (stream, t2) -> subbuf[stream][sub_off + t2]).
__global__ void conv2_rows_kernel(const SubDesc* __restrict__ desc, int t1_cap, int sub_cap, int D,
                                  int64_t* __restrict__ a_row_off, int64_t* __restrict__ c_row_off) {
  const SubDesc d = desc[blockIdx.x];
  for (int i = threadIdx.x; i < d.t2 * F2; i += blockDim.x) {
    int t2 = i / F2, f2 = i % F2;
    a_row_off[(size_t)d.row0 * F2 + i] = (((int64_t)d.stream * t1_cap + 2 * t2) * F1 + 2 * f2) * D;
  }
  for (int t2 = threadIdx.x; t2 < d.t2; t2 += blockDim.x)
    c_row_off[d.row0 + t2] = ((int64_t)d.stream * sub_cap + d.sub_off + t2) * D;
}

int launch_conv2_rows(const SubDesc* desc, int n_desc, int t1_cap, int sub_cap, int D, int64_t* a_row_off,
                      int64_t* c_row_off, cudaStream_t st) {
  if (n_desc <= 0) return 0;
  conv2_rows_kernel<<<n_desc, 256, 0, st>>>(desc, t1_cap, sub_cap, D, a_row_off, c_row_off);
  SCB_LAUNCH_CHECK();
  return 0;
}

// Move the carried frames (rows [src, src+n)) of a per-stream buffer to its front.  Source and
// destination may overlap, so each CTA stages its rows in registers first.
__global__ void carry_rows_kernel(float* __restrict__ buf, int cap, int width, const int* __restrict__ stream,
                                  const int* __restrict__ src, const int* __restrict__ n) {
  const int s = stream[blockIdx.x], r0 = src[blockIdx.x], cnt = n[blockIdx.x];
  float* base = buf + (size_t)s * cap * width;
  const int total = cnt * width;
  constexpr int MAXR = 48;               // 256 threads * 48 = 12288 floats (>= 40 frames * 256)
  float v[MAXR];
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    int idx = threadIdx.x + i * blockDim.x;
    v[i] = idx < total ? base[(size_t)r0 * width + idx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    int idx = threadIdx.x + i * blockDim.x;
    if (idx < total) base[idx] = v[i];
  }
}

int launch_carry_rows(float* buf, int cap, int width, const int* stream, const int* src, const int* n,
                      int n_desc, cudaStream_t st) {
  if (n_desc <= 0) return 0;
  carry_rows_kernel<<<n_desc, 256, 0, st>>>(buf, cap, width, stream, src, n);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- block assembly
// Pass 1: addin[blk] = sqrt(D) * mean(block frames) + pe[ctx offset]        (encoder.py:362-368)
__global__ void block_addin_kernel(const float* __restrict__ subbuf, int sub_cap, const float* __restrict__ pe,
                                   const BlockDesc* __restrict__ blk, float* __restrict__ addin, int D,
                                   float scale) {
  const BlockDesc b = blk[blockIdx.x];
  if (b.short_path) return;
  const int d = threadIdx.x;
  const float* src = subbuf + ((size_t)b.stream * sub_cap + b.sub_start) * D;
  float acc = 0.f;
  for (int j = 0; j < b.clen; ++j) acc += src[(size_t)j * D + d];
  float m = acc / (float)b.clen;
  addin[(size_t)blockIdx.x * D + d] = m * scale + pe[(size_t)b.pe_ctx_off * D + d];
}

// Pass 2: X[blk] = [prev ctx | sqrt(D)*frames + pe | zero padding | own ctx]   (encoder.py:354-380)
__global__ void block_fill_kernel(const float* __restrict__ subbuf, int sub_cap, const float* __restrict__ pe,
                                  const BlockDesc* __restrict__ blk, const float* __restrict__ addin,
                                  const float* __restrict__ prev_addin, float* __restrict__ X, int D, float scale) {
  const BlockDesc b = blk[blockIdx.x];
  const int d = threadIdx.x;
  float* xb = X + (size_t)blockIdx.x * kSlots * D;
  const float* src = subbuf + ((size_t)b.stream * sub_cap + b.sub_start) * D;
  if (b.short_path) {
    for (int j = 0; j < kSlots; ++j) {
      float v = 0.f;
      if (j < b.clen) v = src[(size_t)j * D + d] * scale + pe[(size_t)(b.pe_frame_off + j) * D + d];
      xb[(size_t)j * D + d] = v;
    }
    return;
  }
  float ctx_in;
  if (b.prev_blk >= 0) ctx_in = addin[(size_t)b.prev_blk * D + d];
  else if (b.has_prev_addin) ctx_in = prev_addin[(size_t)b.stream * D + d];
  else ctx_in = addin[(size_t)blockIdx.x * D + d];
  xb[d] = ctx_in;
  for (int j = 0; j < kBlock; ++j) {
    float v = 0.f;
    if (j < b.clen) v = src[(size_t)j * D + d] * scale + pe[(size_t)(b.pe_frame_off + j) * D + d];
    xb[(size_t)(j + 1) * D + d] = v;
  }
  xb[(size_t)(kBlock + 1) * D + d] = addin[(size_t)blockIdx.x * D + d];
}

// Pass 3 (after pass 2 of all blocks): remember the last block's addin for the next call.
__global__ void block_prev_addin_kernel(const BlockDesc* __restrict__ blk, const float* __restrict__ addin,
                                        float* __restrict__ prev_addin, int D) {
  const BlockDesc b = blk[blockIdx.x];
  if (!b.is_last || b.short_path) return;
  prev_addin[(size_t)b.stream * D + threadIdx.x] = addin[(size_t)blockIdx.x * D + threadIdx.x];
}

int launch_block_assemble(const float* subbuf, int sub_cap, const float* pe, const BlockDesc* blk, int n_blk,
                          float* addin, float* prev_addin, float* X, int D, cudaStream_t st) {
  if (n_blk <= 0) return 0;
  float scale = sqrtf((float)D);
  block_addin_kernel<<<n_blk, D, 0, st>>>(subbuf, sub_cap, pe, blk, addin, D, scale);
  SCB_LAUNCH_CHECK();
  block_fill_kernel<<<n_blk, D, 0, st>>>(subbuf, sub_cap, pe, blk, addin, prev_addin, X, D, scale);
  SCB_LAUNCH_CHECK();
  block_prev_addin_kernel<<<n_blk, D, 0, st>>>(blk, addin, prev_addin, D);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- block self-attention
// One CTA per (block, head).  Normal blocks: query rows 1..41 attend key rows 0..40 (row 0 is fully
// masked: softmax over finfo.min then zeroed -> its attention output is 0 and the row is later
// overwritten by the context hand-over; key 41 is never visible).  Short path: rows 0..T-1, no mask.
template <int DK>
__global__ void __launch_bounds__(128) enc_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                            __nv_bfloat16* __restrict__ out16,
                                                            const BlockDesc* __restrict__ blk, int D, SplitOut so) {
  const BlockDesc b = blk[blockIdx.x];
  const int head = blockIdx.y;
  __shared__ float Q[kSlots][DK + 1], Kt[kSlots][DK + 1], Vv[kSlots][DK + 1];
  __shared__ float P[kSlots][kSlots + 2];
  const float* base = qkv + (size_t)blockIdx.x * kSlots * 3 * D + head * DK;
  for (int i = threadIdx.x; i < kSlots * DK; i += blockDim.x) {
    int r = i / DK, c = i % DK;
    const float* row = base + (size_t)r * 3 * D;
    Q[r][c] = row[c];
    Kt[r][c] = row[D + c];
    Vv[r][c] = row[2 * D + c];
  }
  __syncthreads();
  const int q_lo = b.short_path ? 0 : 1, q_hi = b.short_path ? b.n_rows : kSlots;       // [q_lo, q_hi)
  const int k_lo = 0, k_hi = b.short_path ? b.n_rows : kBlock + 1;                       // [k_lo, k_hi)
  const float sqrt_dk = sqrtf((float)DK);
  const int nq = q_hi - q_lo, nk = k_hi - k_lo;
  for (int i = threadIdx.x; i < nq * nk; i += blockDim.x) {
    int qi = q_lo + i / nk, ki = k_lo + i % nk;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DK; ++c) acc = fmaf(Q[qi][c], Kt[ki][c], acc);
    P[qi][ki] = acc / sqrt_dk;
  }
  __syncthreads();
  // softmax per query row: one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int qi = q_lo + warp; qi < q_hi; qi += nwarp) {
    float m = -INFINITY;
    for (int ki = k_lo + lane; ki < k_hi; ki += 32) m = fmaxf(m, P[qi][ki]);
    m = warp_max(m);
    float s = 0.f;
    for (int ki = k_lo + lane; ki < k_hi; ki += 32) { float e = expf(P[qi][ki] - m); P[qi][ki] = e; s += e; }
    s = warp_sum(s);
    for (int ki = k_lo + lane; ki < k_hi; ki += 32) P[qi][ki] /= s;
  }
  __syncthreads();
  float* obase = out + (size_t)blockIdx.x * kSlots * D + head * DK;
  for (int i = threadIdx.x; i < kSlots * DK; i += blockDim.x) {
    int qi = i / DK, c = i % DK;
    float acc = 0.f;
    if (qi >= q_lo && qi < q_hi)
      for (int ki = k_lo; ki < k_hi; ++ki) acc = fmaf(P[qi][ki], Vv[ki][c], acc);
    if (so.base) so.put((size_t)blockIdx.x * kSlots + qi, head * DK + c, acc);
    else obase[(size_t)qi * D + c] = acc;     // rows outside [q_lo, q_hi) get 0 (fully masked rows)
    if (out16) out16[((size_t)blockIdx.x * kSlots + qi) * D + head * DK + c] = __float2bfloat16(acc);
  }
}


// Register-tiled version (fp32 modes): one CTA per block, one thread per (head, query row) task.  K and V of all heads
// are staged once in shared memory ([42][D] each); a task keeps its query row, its 42 scores and its output row in
// registers, so the inner loops are FMAs fed by warp-broadcast LDS.128 of K / V rows (4 FMAs per shared-memory
// instruction) and the softmax needs no cross-lane traffic.  Same arithmetic per element as enc_attention_kernel:
// score = (FMA chain over c) / sqrt(dk), p = exp(s - max) / sum, out = FMA chain over keys.
template <int DK>
__global__ void __launch_bounds__(192) enc_attention_rows_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                 const BlockDesc* __restrict__ blk, int D, int H, SplitOut so) {
  const BlockDesc b = blk[blockIdx.x];
  extern __shared__ __align__(16) float sm_kv[];          // K [42][D] then V [42][D]
  float* Ks = sm_kv;
  float* Vs = sm_kv + kSlots * D;
  const float* base = qkv + (size_t)blockIdx.x * kSlots * 3 * D;
  const int k_hi = b.short_path ? b.n_rows : kBlock + 1;                        // keys [0, k_hi)
  const int q_lo = b.short_path ? 0 : 1, q_hi = b.short_path ? b.n_rows : kSlots;   // queries [q_lo, q_hi)
  const int d4 = D / 4;
  for (int i = threadIdx.x; i < k_hi * d4; i += blockDim.x) {
    const int r = i / d4, c4 = i % d4;
    const float4* row = reinterpret_cast<const float4*>(base + (size_t)r * 3 * D);
    reinterpret_cast<float4*>(Ks)[r * d4 + c4] = row[d4 + c4];
    reinterpret_cast<float4*>(Vs)[r * d4 + c4] = row[2 * d4 + c4];
  }
  __syncthreads();
  const float sqrt_dk = sqrtf((float)DK);
  for (int task = threadIdx.x; task < H * kSlots; task += blockDim.x) {
    const int head = task / kSlots, qi = task % kSlots;
    float o[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) o[c] = 0.f;
    if (qi >= q_lo && qi < q_hi) {
      float q[DK];
      const float4* qrow = reinterpret_cast<const float4*>(base + (size_t)qi * 3 * D + head * DK);
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) { const float4 v = qrow[c]; q[4 * c] = v.x; q[4 * c + 1] = v.y; q[4 * c + 2] = v.z; q[4 * c + 3] = v.w; }
      float sc[kSlots];
      float m = -INFINITY;
#pragma unroll
      for (int ki = 0; ki < kSlots; ++ki) {
        sc[ki] = -INFINITY;
        if (ki < k_hi) {
          const float4* kr = reinterpret_cast<const float4*>(Ks + ki * D + head * DK);
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < DK / 4; ++c) {
            const float4 kv = kr[c];
            acc = fmaf(q[4 * c], kv.x, acc); acc = fmaf(q[4 * c + 1], kv.y, acc);
            acc = fmaf(q[4 * c + 2], kv.z, acc); acc = fmaf(q[4 * c + 3], kv.w, acc);
          }
          sc[ki] = acc / sqrt_dk;
          m = fmaxf(m, sc[ki]);
        }
      }
      float ssum = 0.f;
#pragma unroll
      for (int ki = 0; ki < kSlots; ++ki) if (ki < k_hi) { sc[ki] = expf(sc[ki] - m); ssum += sc[ki]; }
#pragma unroll
      for (int ki = 0; ki < kSlots; ++ki) {
        if (ki < k_hi) {
          const float pk = sc[ki] / ssum;
          const float4* vr = reinterpret_cast<const float4*>(Vs + ki * D + head * DK);
#pragma unroll
          for (int c = 0; c < DK / 4; ++c) {
            const float4 vv = vr[c];
            o[4 * c] = fmaf(pk, vv.x, o[4 * c]); o[4 * c + 1] = fmaf(pk, vv.y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(pk, vv.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pk, vv.w, o[4 * c + 3]);
          }
        }
      }
    }
    // rows outside [q_lo, q_hi) get 0 (fully masked rows)
    const size_t row = (size_t)blockIdx.x * kSlots + qi;
    if (so.base) {
#pragma unroll
      for (int c = 0; c < DK; c += 8) {
        uint4 uh, ul;
        x3_split8(o + c, uh, ul);
        *reinterpret_cast<uint4*>(so.base + row * so.ld + head * DK + c) = uh;
        *reinterpret_cast<uint4*>(so.base + so.plane + row * so.ld + head * DK + c) = ul;
      }
    } else {
      float4* orow = reinterpret_cast<float4*>(out + row * D + head * DK);
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) orow[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    }
  }
}

int launch_enc_attention(const float* qkv, float* out, __nv_bfloat16* out16, const BlockDesc* blk, int n_blk,
                         int n_head, int d_model, cudaStream_t st, SplitOut so) {
  if (n_blk <= 0) return 0;
  dim3 grid(n_blk, n_head);
  int dk = d_model / n_head;
  static const bool rows_kernel = [] { const char* v = getenv("SCB_ENC_ATTN"); return !(v && v[0] == 'c'); }();   // "cta": old kernel
  if (!out16 && rows_kernel && (dk == 32 || dk == 64) && d_model % 8 == 0) {
    const size_t smem = sizeof(float) * 2 * kSlots * d_model;
    if (dk == 32) {
      static PerDeviceMark mk;
      if (mk.cur() < smem) { cudaFuncSetAttribute(enc_attention_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mk.cur() = smem; }
      enc_attention_rows_kernel<32><<<n_blk, 192, smem, st>>>(qkv, out, blk, d_model, n_head, so);
    } else {
      static PerDeviceMark mk;
      if (mk.cur() < smem) { cudaFuncSetAttribute(enc_attention_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mk.cur() = smem; }
      enc_attention_rows_kernel<64><<<n_blk, 192, smem, st>>>(qkv, out, blk, d_model, n_head, so);
    }
    SCB_LAUNCH_CHECK();
    return 0;
  }
  if (dk == 32) enc_attention_kernel<32><<<grid, 128, 0, st>>>(qkv, out, out16, blk, d_model, so);
  else if (dk == 64) enc_attention_kernel<64><<<grid, 128, 0, st>>>(qkv, out, out16, blk, d_model, so);
  else { set_last_error("enc_attention: unsupported head dim %d", dk); return -1; }
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- context hand-over (encoder_layer.py:253-267)
// The CTA of a stream's first block walks that stream's chain of blocks (they are consecutive in the
// descriptor array): slot 0 of block i <- slot 41 of block i-1 (or the carried context / own slot 41
// for the first block), and the last slot 41 becomes the context carried to the next call.
__global__ void ctx_handover_kernel(float* __restrict__ X, float* __restrict__ enc_ctx, int layer, int n_layers,
                                    const BlockDesc* __restrict__ blk, int D, const float* __restrict__ ln_w,
                                    const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ nrm16) {
  const BlockDesc b = blk[blockIdx.x];
  if (b.short_path || b.prev_blk >= 0) return;
  const int d = threadIdx.x;
  __shared__ float red[32];
  const int warp = d >> 5, lane = d & 31, nwarp = blockDim.x >> 5;
  float* ctx = enc_ctx + ((size_t)b.stream * n_layers + layer) * D + d;
  float carry = b.has_past_ctx ? *ctx : X[((size_t)blockIdx.x * kSlots + kBlock + 1) * D + d];
  for (int i = blockIdx.x;; ++i) {
    float* xb = X + (size_t)i * kSlots * D;
    const float own = xb[(size_t)(kBlock + 1) * D + d];
    xb[d] = carry;
    if (ln_w) {
      // bf16 mode: the next layer's norm1 of this slot-0 row (the GEMM epilogue normalised the value it replaced)
      float sm = warp_sum(carry);
      if (lane == 0) red[warp] = sm;
      __syncthreads();
      float tot = 0.f;
      for (int w = 0; w < nwarp; ++w) tot += red[w];
      const float mean = tot / (float)D;
      const float dd = carry - mean;
      float q = warp_sum(dd * dd);
      __syncthreads();
      if (lane == 0) red[warp] = q;
      __syncthreads();
      float qt = 0.f;
      for (int w = 0; w < nwarp; ++w) qt += red[w];
      const float rstd = 1.0f / sqrtf(qt / (float)D + 1e-12f);
      nrm16[(size_t)i * kSlots * D + d] = __float2bfloat16(dd * rstd * ln_w[d] + ln_b[d]);
      __syncthreads();
    }
    carry = own;
    if (blk[i].is_last) break;
  }
  *ctx = carry;
}

int launch_ctx_handover(float* X, float* enc_ctx, int layer, int n_layers, const BlockDesc* blk, int n_blk,
                        int D, const float* ln_w, const float* ln_b, __nv_bfloat16* nrm16, cudaStream_t st) {
  if (n_blk <= 0) return 0;
  ctx_handover_kernel<<<n_blk, D, 0, st>>>(X, enc_ctx, layer, n_layers, blk, D, ln_w, ln_b, nrm16);
  SCB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- stitch + after_norm (encoder.py:391-405, 500-522)
// One warp per emitted frame: LayerNorm of X[blk][slot] -> encbuf[stream][t].
__global__ void stitch_norm_kernel(const float* __restrict__ X, const BlockDesc* __restrict__ blk,
                                   const float* __restrict__ w, const float* __restrict__ bb,
                                   float* __restrict__ encbuf, int t_cap, int D,
                                   __nv_bfloat16* __restrict__ dense16) {
  const BlockDesc b = blk[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int r = warp; r < b.out_count; r += nwarp) {
    const float* xr = X + ((size_t)blockIdx.x * kSlots + b.out_slot0 + r) * D;
    float v[16];
    const int nv = D >> 5;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (i < nv) { v[i] = xr[lane + 32 * i]; s += v[i]; }
    float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (i < nv) { float dd = v[i] - mean; q += dd * dd; }
    float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + 1e-12f);
    float* yr = encbuf + ((size_t)b.stream * t_cap + b.out_t0 + r) * D;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (i < nv) {
      int c = lane + 32 * i;
      float o = (v[i] - mean) * rstd * w[c] + bb[c];
      yr[c] = o;
      if (dense16) dense16[(size_t)(b.out_row0 + r) * D + c] = __float2bfloat16(o);
    }
  }
}

int launch_stitch_norm(const float* X, const BlockDesc* blk, int n_blk, const float* w, const float* b,
                       float* encbuf, int t_cap, int D, __nv_bfloat16* dense16, cudaStream_t st) {
  if (n_blk <= 0) return 0;
  stitch_norm_kernel<<<n_blk, 256, 0, st>>>(X, blk, w, b, encbuf, t_cap, D, dense16);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
