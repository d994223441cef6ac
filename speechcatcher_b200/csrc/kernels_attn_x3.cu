// Decoder attention on tensor cores with fp32-class accuracy (precise mode, precision 2): the K|V caches are kept as
// split fp16 planes (x3_split.cuh: hi + lo 2^-11, the same 4 bytes per element as fp32) and both products run as three
// mma.sync.m16n8k16 (fp16 x fp16 -> fp32) each, like the split-precision GEMM (kernels_gemm_x3.cu):
//     S = Q_hi K_hi^T + 2^-11 (Q_hi K_lo^T + Q_lo K_hi^T)            O = P_hi V_hi + 2^-11 (P_hi V_lo + P_lo V_hi)
// The CUDA-core fp32 attention kernels are instruction-bound (~1400 warp instructions per 32 keys and head); here the
// same tile costs 48 HMMAs plus the softmax.  Structure as in kernels_attn_mma.cu (bf16 mode): one CTA per (active
// stream, head), the <= 16 hypotheses of the stream are the rows of the MMA tile, flash-style split-KV over the four
// warps, every warp with a private two-stage cp.async pipeline over its own 32 keys per step, abstract key list
// (cross: encoder frames; self: common ancestor chain + owner-masked divergent tail, build_self_keys_kernel).
//
// Row layout of both caches: [row][hi: K(D) | V(D)][lo: K(D) | V(D)] fp16, i.e. 2 KB per row for D = 256 -- the row
// pitch of the fp32 caches, so the buffers are the same.  Cross rows are written by the K|V projection GEMM (split-plane
// output with row scatter), self rows by this kernel when it appends the scored token.
//
// Accumulation: a tensor-core instruction adds its products to the fp32 accumulator with truncation (measured,
// profiles/r2_tc_accum_probe.json), so P V of every 32-key tile is formed in fresh accumulators and added to the
// running output with round-to-nearest FMAs (the rescale o * f + tile does it), and the low-order terms have their own
// accumulators.  exp2f (not the approximate ex2) is used throughout.
//
// Replaces the attention part of speechcatcher/model/decoder/decoder_layer.py:80-113
// (speechcatcher/model/attention/multi_head_attention.py:92-133).
#include <cuda_fp16.h>
#include <stdlib.h>
#include "kernels.h"

namespace scb {

// Keys per warp per step.  ncu (profiles/r2_ncu_dec_attn_x3.txt) showed the first version (32 keys, 85 KB of stages per
// CTA, two CTAs = 8 warps per SM) at 12 % warp occupancy and ~1.2 TB/s: latency-bound.  16 keys per warp-step halve
// the stages (44 KB per CTA, five CTAs = 20 warps per SM).
#ifndef SCB_X_KPW
#define SCB_X_KPW 16
#endif
#ifndef SCB_X_NS
#define SCB_X_NS 3
#endif
// compile-time knobs for tuning runs (SCB_NVCC_EXTRA="-DSCB_X_KPW=32 -DSCB_X_NS=2").  Measured on the bench workload, strict mode
// (keys, stages) -> audio-s/s: (16, 2 padded) 3 893, (16, 3) 4 046, (16, 4) 3 972, (32, 2) 3 995, (32, 3) 3 805
constexpr int X_KPW = SCB_X_KPW;
constexpr int X_NSTAGES = SCB_X_NS;
constexpr int X_MINB = (X_KPW * X_NSTAGES <= 48) ? 4 : (X_KPW * X_NSTAGES <= 64 ? 3 : 2);   // CTAs per SM the stages leave room for
constexpr int X_STEP = 4 * X_KPW;  // keys per CTA step
constexpr int X_NT = X_KPW / 8;    // score n-tiles per warp
constexpr int X_KK = X_KPW / 16;   // k-steps of the P*V product per warp
constexpr int X_KEYS_SMEM = 768;   // self-attention key list entries staged in shared memory

__device__ __forceinline__ void xcp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void xcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void xcp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void xldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void xldsm_x4_trans(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_f16(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// two fp32 values -> packed fp16 hi pair and packed fp16 lo pair
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __half h0, l0, h1, l1;
  x3_split(x0, h0, l0);
  x3_split(x1, h1, l1);
  hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
  lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

// WH = false: one CTA (4 warps) per (stream, head), the four warps split the keys of a 64-key step and merge at the end.
// WH = true ("warp = head", opt-in experiment): one CTA per stream with one warp per head; a warp walks ALL keys of its head in
// 16-key steps, so there is no merge, the fixed cost of a CTA (Q split, append, key list, first round trip) is paid once
// per stream instead of once per (stream, head) in ~3 waves, and the eight head slices of a cache row are fetched together.
template <int DK, int MODE, bool WH>
__global__ void __launch_bounds__(WH ? 256 : 128, WH ? 1 : X_MINB) dec_attn_x3_kernel(SearchBuffers sb, __half* kv_layer, const float* __restrict__ q,
                                                          int ldq, float* __restrict__ out, SplitOut so) {
  pdl_sync();
  // head-major grid (SCB_ATTN_HEAD_MAJOR, default): the CTAs of the eight heads of a stream are neighbours in launch order,
  // so the 64-byte head slices of one 2 KB cache row are fetched at about the same time (DRAM page / L2 locality)
  // instead of in eight separate passes over the cache
  const int bs = WH ? blockIdx.x : (sb.attn_head_major ? blockIdx.y : blockIdx.x);
  if (bs >= *sb.n_active) return;
  const int s = sb.act_streams[bs];
  const int head = WH ? (int)(threadIdx.x >> 5) : (sb.attn_head_major ? blockIdx.x : blockIdx.y);
  const StreamCtl& c = sb.ctl[s];
  // beam 17..32: two m16 tiles of hypotheses per (stream, head), one CTA each (blockIdx.z); a tile reads every key but
  // appends, scores and writes only its own rows h0 .. h0 + nb - 1 (a hypothesis' new token is visible to itself only,
  // so the other tile's concurrent appends are always masked; the caches are zero-initialised, so masked values are finite)
  const int h0 = blockIdx.z * 16;
  const int nb = min(16, c.n_hyp - h0);
  if (nb <= 0) return;
  const int row0 = sb.row_base[s] + h0, D = sb.D, B = sb.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int RS = DK + 8;                 // padded smem row (fp16 elements): conflict-free ldmatrix
  constexpr int CPR = DK / 8;                // 16-byte chunks per row slice
  constexpr int KSTEPS = DK / 16;
  constexpr int NDT = DK / 8;                // n-tiles of the output
  constexpr int RPP = 32 / CPR;              // rows one pass of the warp's 32 lanes covers
  constexpr int PASSES = X_KPW / RPP;
  // K|V stages: unpadded rows with the 16-byte chunks XOR-swizzled by the row (conflict-free ldmatrix like the padding, but
  // 4 KB instead of 5 KB per 16-key stage), three stages per warp: two loads stay in flight while a tile is being
  // multiplied.  (A load-only probe with this kernel's grid and two stages reaches 4.4-6.2 TB/s, scripts/probes/
  // kv_pattern_probe.cu; the arithmetic between a warp's loads is what the third stage hides.)
  constexpr int RSK = DK;
  constexpr int NS = X_NSTAGES;
  constexpr int ST_HALFS = 4 * X_KPW * RSK;  // one stage: K hi, K lo, V hi, V lo
  auto swz = [](int r) { return (r / (8 / CPR)) & (CPR - 1); };

  const int NWARP = WH ? sb.H : 4;           // warps per CTA
  const int nthr = NWARP * 32;
  const int qsets = WH ? NWARP : 1;          // WH: every warp has its own Q tile (its head)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* Qh = reinterpret_cast<__half*>(smem_raw) + (WH ? warp * 2 * 16 * RS : 0);   // [16][RS]
  __half* Ql = Qh + 16 * RS;                                                     // [16][RS]
  __half* KV = reinterpret_cast<__half*>(smem_raw) + qsets * 2 * 16 * RS;        // [warps][2 stages][4][X_KPW][RS]
  signed char* own = reinterpret_cast<signed char*>(KV + NWARP * NS * ST_HALFS); // [warps][NS][X_KPW]
  int* keys_s = reinterpret_cast<int*>(own + ((NWARP * NS * X_KPW + 15) & ~15));   // self only: [X_KEYS_SMEM]
  // merge scratch aliases the K|V stages after the main loop
  float* mrg_m = reinterpret_cast<float*>(KV);                                   // [4][16]
  float* mrg_l = mrg_m + 64;                                                     // [4][16]
  float* mrg_o = mrg_l + 64;                                                     // [4][16][DK]

  const size_t row_stride = 4 * (size_t)D;   // fp16 elements per cache row: hi K|V, lo K|V
  const int len = c.len;
  __half* base;
  if (MODE == 1) base = kv_layer + (size_t)s * sb.Tcap * row_stride + head * DK;
  else base = kv_layer + (size_t)s * sb.Lcap * B * row_stride + head * DK;

  // ---- Q tile as hi / lo planes (rows >= nb are zero), self: append K|V of the scored token as split rows
  const int t0i = WH ? lane : tid, tstep = WH ? 32 : 128;      // WH: each warp prepares its own head
  for (int i = t0i; i < 16 * DK; i += tstep) {
    const int r = i / DK, d = i % DK;
    const float v = r < nb ? q[(size_t)(row0 + r) * ldq + head * DK + d] : 0.f;
    __half h, l;
    x3_split(v, h, l);
    Qh[r * RS + d] = h;
    Ql[r * RS + d] = l;
  }
  int n_keys;
  const int* keys = nullptr;
  if (MODE == 0) {
    if (t0i == 0) atomicAdd(&sb.prof[3], (unsigned long long)((long long)nb * 2ll * len * DK * 4));
    for (int i = t0i; i < nb * 2 * DK; i += tstep) {      // [len-1][b]: K at +0 / V at +D of the hi half, lo half at +2D
      const int b = i / (2 * DK), rem = i % (2 * DK), which = rem / DK, cc = rem % DK;
      const float v = q[(size_t)(row0 + b) * ldq + D + which * D + head * DK + cc];
      __half h, l;
      x3_split(v, h, l);
      __half* dst = base + ((size_t)(len - 1) * B + h0 + b) * row_stride + which * D + cc;
      dst[0] = h;
      dst[2 * D] = l;
    }
    n_keys = sb.self_nkeys[s];
    keys = sb.self_keys + (size_t)s * sb.key_cap;
    // bytes this kernel really reads: every key row of the list once per (stream, head, tile), 4 planes x DK fp16
    if (t0i == 0) atomicAdd(&sb.prof[5], (unsigned long long)((long long)n_keys * 4 * DK * 2));
    for (int i = tid; i < n_keys && i < X_KEYS_SMEM; i += nthr) keys_s[i] = keys[i];
  } else {
    n_keys = c.Tb;
    if (t0i == 0 && h0 == 0) atomicAdd(&sb.prof[2], (unsigned long long)(2ll * n_keys * DK * 4));
  }
  __syncthreads();                      // Q planes, key list staged; appended rows visible to the loads below
  constexpr int STEP = WH ? X_KPW : X_STEP;                     // keys all warps of a CTA advance per step
  const int wk = WH ? 0 : X_KPW * warp;                         // this warp's first key inside a step
  const int n_steps = (n_keys + STEP - 1) / STEP;

  __half* kw = KV + (size_t)warp * (NS * ST_HALFS);
  signed char* ownw = own + warp * NS * X_KPW;
  const int lr = lane / CPR, ch8 = (lane % CPR) * 8;

  auto issue = [&](int t, int buf) {
    const int u0 = t * STEP + wk;
    __half* st = kw + (size_t)buf * ST_HALFS;           // K hi | K lo | V hi | V lo, [X_KPW][RSK] each
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = lr + RPP * ps, u = u0 + r;
      __half* d0 = st + r * RSK + (((lane % CPR) ^ swz(r)) << 3);
      if (u < n_keys) {
        const __half* src;
        if (MODE == 1) src = base + (size_t)u * row_stride + ch8;
        else {
          const int kd_ = u < X_KEYS_SMEM ? keys_s[u] : keys[u];
          src = base + ((size_t)(kd_ & 0xffff) * B + ((kd_ >> 16) & 0xff)) * row_stride + ch8;
          if (ch8 == 0) ownw[buf * X_KPW + r] = (signed char)((kd_ >> 24) - 1);
        }
        xcp16(d0, src);                                   // K hi
        xcp16(d0 + X_KPW * RSK, src + 2 * D);             // K lo
        xcp16(d0 + 2 * X_KPW * RSK, src + D);             // V hi
        xcp16(d0 + 3 * X_KPW * RSK, src + 3 * D);         // V lo
      } else {
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(d0) = z;
        *reinterpret_cast<uint4*>(d0 + X_KPW * RSK) = z;
        *reinterpret_cast<uint4*>(d0 + 2 * X_KPW * RSK) = z;
        *reinterpret_cast<uint4*>(d0 + 3 * X_KPW * RSK) = z;
        if (MODE == 0 && ch8 == 0) ownw[buf * X_KPW + r] = (signed char)-2;      // invisible to everyone
      }
    }
    xcp_commit();
  };

  if (n_steps > 0) issue(0, 0);
  if (NS > 2 && n_steps > 1) issue(1, 1);
  uint32_t qah[KSTEPS][4], qal[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const int off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 16 * ks + 8 * (lane >> 4);
    xldsm_x4(qah[ks], Qh + off);
    xldsm_x4(qal[ks], Ql + off);
  }

  const float cs = 1.4426950408889634f / sqrtf((float)DK);      // softmax scale * log2(e): p = 2^(cs * (s - m))
  const int r0 = lane >> 2, r1 = r0 + 8, qd = lane & 3;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};   // running max of the RAW scores
  float o[NDT][4];
#pragma unroll
  for (int i = 0; i < NDT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

  for (int t = 0; t < n_steps; ++t) {
    const int buf = t % NS;
    if (NS > 2) {
      if (t + 2 < n_steps) { issue(t + 2, (t + 2) % NS); xcp_wait<2>(); }
      else if (t + 1 < n_steps) { xcp_wait<1>(); }
      else { xcp_wait<0>(); }
    } else {
      if (t + 1 < n_steps) { issue(t + 1, (t + 1) % NS); xcp_wait<1>(); } else { xcp_wait<0>(); }
    }
    __syncwarp();                       // the other lanes' copies / zero fills of this stage are visible
    const int u0 = t * STEP + wk;
    if (u0 < n_keys) {
      const __half* kh = kw + (size_t)buf * ST_HALFS;
      const __half* kl = kh + X_KPW * RSK;
      const __half* vh = kl + X_KPW * RSK;
      const __half* vl = vh + X_KPW * RSK;
      // ---- S = Q K^T for this warp's keys (raw scores): main term and low-order terms (x 2^11) separately
      float sacc[X_NT][4];
#pragma unroll
      for (int nt = 0; nt < X_NT; ++nt) {
        float sm[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k2 = 0; k2 < KSTEPS / 2; ++k2) {
          uint32_t bh[4], bl[4];          // B fragments of two k-steps with one ldmatrix per plane
          const int krow = 8 * nt + (lane & 7);
          const int off = krow * RSK + (((4 * k2 + (lane >> 3)) ^ swz(krow)) << 3);
          xldsm_x4(bh, kh + off);
          xldsm_x4(bl, kl + off);
          mma_f16(sm, qah[2 * k2], bh);
          mma_f16(sm, qah[2 * k2 + 1], bh + 2);
          mma_f16(sc, qah[2 * k2], bl);
          mma_f16(sc, qah[2 * k2 + 1], bl + 2);
          mma_f16(sc, qal[2 * k2], bh);
          mma_f16(sc, qal[2 * k2 + 1], bh + 2);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[nt][e] = fmaf(sc[e], X3_INV_SCALE, sm[e]);
      }
      // ---- mask: only where one can exist
      if (MODE == 0) {
        const signed char* ob = ownw + buf * X_KPW;
        if (!__all_sync(0xffffffffu, ob[lane & (X_KPW - 1)] == -1)) {       // step touches the divergent tail (or padding)
#pragma unroll
          for (int nt = 0; nt < X_NT; ++nt) {
            const int o0 = ob[8 * nt + 2 * qd], o1 = ob[8 * nt + 2 * qd + 1];
            if (!(o0 == -1 || o0 == h0 + r0)) sacc[nt][0] = -INFINITY;
            if (!(o1 == -1 || o1 == h0 + r0)) sacc[nt][1] = -INFINITY;
            if (!(o0 == -1 || o0 == h0 + r1)) sacc[nt][2] = -INFINITY;
            if (!(o1 == -1 || o1 == h0 + r1)) sacc[nt][3] = -INFINITY;
          }
        }
      } else if (u0 + X_KPW > n_keys) {                        // cross: zero-filled rows past the last frame
#pragma unroll
        for (int nt = 0; nt < X_NT; ++nt) {
          const int k0 = u0 + 8 * nt + 2 * qd;
          if (k0 >= n_keys) { sacc[nt][0] = -INFINITY; sacc[nt][2] = -INFINITY; }
          if (k0 + 1 >= n_keys) { sacc[nt][1] = -INFINITY; sacc[nt][3] = -INFINITY; }
        }
      }
      // ---- online softmax (rows r0, r1 of this thread)
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < X_NT; ++nt) {
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
        scale[h] = exp2f((m_run[h] - m_use) * cs);                  // m_run = -inf -> 0 (o, l are still 0)
        m_run[h] = m_new;
        nm[h] = -m_use * cs;
      }
      uint32_t pah[X_KK][4], pal[X_KK][4];
#pragma unroll
      for (int nt = 0; nt < X_NT; ++nt) {
        const float p0 = exp2f(fmaf(sacc[nt][0], cs, nm[0])), p1 = exp2f(fmaf(sacc[nt][1], cs, nm[0]));
        const float p2 = exp2f(fmaf(sacc[nt][2], cs, nm[1])), p3 = exp2f(fmaf(sacc[nt][3], cs, nm[1]));
        psum[0] += p0 + p1; psum[1] += p2 + p3;
        // C fragment of S -> A fragment of P: n-tile 2kk -> a0/a1, n-tile 2kk+1 -> a2/a3
        split_pair(p0, p1, pah[nt >> 1][(nt & 1) * 2 + 0], pal[nt >> 1][(nt & 1) * 2 + 0]);
        split_pair(p2, p3, pah[nt >> 1][(nt & 1) * 2 + 1], pal[nt >> 1][(nt & 1) * 2 + 1]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        psum[h] += __shfl_xor_sync(0xffffffffu, psum[h], 1);
        psum[h] += __shfl_xor_sync(0xffffffffu, psum[h], 2);
        l_run[h] = l_run[h] * scale[h] + psum[h];
      }
      // ---- O = O * scale + P V, the tile's product in fresh accumulators
#pragma unroll
      for (int n2 = 0; n2 < NDT / 2; ++n2) {
        float om[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, oc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kk = 0; kk < X_KK; ++kk) {
          uint32_t bh[4], bl[4];        // V fragments of two output n-tiles with one transposing ldmatrix per plane
          const int vrow = 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1);
          const int off = vrow * RSK + (((2 * n2 + (lane >> 4)) ^ swz(vrow)) << 3);
          xldsm_x4_trans(bh, vh + off);
          xldsm_x4_trans(bl, vl + off);
          mma_f16(om[0], pah[kk], bh);
          mma_f16(om[1], pah[kk], bh + 2);
          mma_f16(oc[0], pah[kk], bl);
          mma_f16(oc[1], pah[kk], bl + 2);
          mma_f16(oc[0], pal[kk], bh);
          mma_f16(oc[1], pal[kk], bh + 2);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float* od = o[2 * n2 + j];
          od[0] = fmaf(od[0], scale[0], fmaf(oc[j][0], X3_INV_SCALE, om[j][0]));
          od[1] = fmaf(od[1], scale[0], fmaf(oc[j][1], X3_INV_SCALE, om[j][1]));
          od[2] = fmaf(od[2], scale[1], fmaf(oc[j][2], X3_INV_SCALE, om[j][2]));
          od[3] = fmaf(od[3], scale[1], fmaf(oc[j][3], X3_INV_SCALE, om[j][3]));
        }
      }
    }
    __syncwarp();                       // stage fully consumed before this warp refills it
  }
  if (WH) {
    // ---- one warp owns the head: normalise and write rows r0, r1 straight from the accumulators
#pragma unroll
    for (int nd = 0; nd < NDT; ++nd) {
      const int col = head * DK + 8 * nd + 2 * qd;
      if (r0 < nb) {
        const float a0 = o[nd][0] / l_run[0], a1 = o[nd][1] / l_run[0];
        const size_t row = (size_t)(row0 + r0);
        if (so.base) {
          uint32_t h2, l2;
          split_pair(a0, a1, h2, l2);
          *reinterpret_cast<uint32_t*>(so.base + row * so.ld + col) = h2;
          *reinterpret_cast<uint32_t*>(so.base + so.plane + row * so.ld + col) = l2;
        } else { out[row * D + col] = a0; out[row * D + col + 1] = a1; }
      }
      if (r1 < nb) {
        const float a2 = o[nd][2] / l_run[1], a3 = o[nd][3] / l_run[1];
        const size_t row = (size_t)(row0 + r1);
        if (so.base) {
          uint32_t h2, l2;
          split_pair(a2, a3, h2, l2);
          *reinterpret_cast<uint32_t*>(so.base + row * so.ld + col) = h2;
          *reinterpret_cast<uint32_t*>(so.base + so.plane + row * so.ld + col) = l2;
        } else { out[row * D + col] = a2; out[row * D + col + 1] = a3; }
      }
    }
    return;
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
      const float f = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * cs);
      L += mrg_l[w * 16 + r] * f;
      acc += mrg_o[((size_t)w * 16 + r) * DK + d] * f;
    }
    const float res = acc / L;
    if (so.base) so.put((size_t)(row0 + r), head * DK + d, res);
    else out[(size_t)(row0 + r) * D + head * DK + d] = res;
  }
}


// ---------------------------------------------------------------- encoder block attention, same arithmetic
// One CTA (3 warps) per (block, head): the 42 block rows padded to 48 = three m16 query tiles, one per warp; Q, K, V of
// the head are read as fp32 from the fused QKV rows, split once into fp16 hi / lo planes in shared memory, and both
// products run as three HMMAs each.  Mask as in the CUDA-core kernels: rows 1..41 attend keys 0..40 (short path:
// rows / keys < n_rows, no mask); rows outside the query range get 0.  (structure: enc_attn_mma_kernel of the bf16 mode)
template <int DK>
__global__ void __launch_bounds__(96) enc_attn_x3_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                         const BlockDesc* __restrict__ blk, int D, SplitOut so, int head_major) {
  const int blk_i = head_major ? blockIdx.y : blockIdx.x;
  const BlockDesc b = blk[blk_i];
  const int head = head_major ? blockIdx.x : blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int RS = DK + 8, CPR = DK / 8, KSTEPS = DK / 16, NDT = DK / 8, ROWS = 48;
  __shared__ __align__(16) __half sm_all[6 * ROWS * RS];      // Q hi, Q lo, K hi, K lo, V hi, V lo
  __half* Qh = sm_all;
  __half* Ql = Qh + ROWS * RS;
  __half* Kh = Ql + ROWS * RS;
  __half* Kl = Kh + ROWS * RS;
  __half* Vh = Kl + ROWS * RS;
  __half* Vl = Vh + ROWS * RS;
  const float* base = qkv + (size_t)blk_i * kSlots * 3 * D + head * DK;
  for (int idx = tid; idx < 3 * ROWS * CPR; idx += 96) {
    const int which = idx / (ROWS * CPR), rem = idx % (ROWS * CPR), r = rem / CPR, ch = rem % CPR;
    uint4 uh = make_uint4(0, 0, 0, 0), ul = uh;
    if (r < kSlots) {
      const float4* src = reinterpret_cast<const float4*>(base + (size_t)r * 3 * D + which * D + ch * 8);
      const float4 a = src[0], c = src[1];
      const float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      x3_split8(x, uh, ul);
    }
    __half* dh = sm_all + (size_t)(2 * which) * ROWS * RS + r * RS + ch * 8;
    *reinterpret_cast<uint4*>(dh) = uh;
    *reinterpret_cast<uint4*>(dh + ROWS * RS) = ul;
  }
  __syncthreads();
  const int q_lo = b.short_path ? 0 : 1, q_hi = b.short_path ? b.n_rows : kSlots;
  const int k_hi = b.short_path ? b.n_rows : kBlock + 1;
  uint32_t qah[KSTEPS][4], qal[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const int off = (16 * warp + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 16 * ks + 8 * (lane >> 4);
    xldsm_x4(qah[ks], Qh + off);
    xldsm_x4(qal[ks], Ql + off);
  }
  float sacc[6][4];
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    float smn[4] = {0.f, 0.f, 0.f, 0.f}, scr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k2 = 0; k2 < KSTEPS / 2; ++k2) {
      uint32_t bh[4], bl[4];
      const int off = (8 * nt + (lane & 7)) * RS + 32 * k2 + 8 * (lane >> 3);
      xldsm_x4(bh, Kh + off);
      xldsm_x4(bl, Kl + off);
      mma_f16(smn, qah[2 * k2], bh);
      mma_f16(smn, qah[2 * k2 + 1], bh + 2);
      mma_f16(scr, qah[2 * k2], bl);
      mma_f16(scr, qah[2 * k2 + 1], bl + 2);
      mma_f16(scr, qal[2 * k2], bh);
      mma_f16(scr, qal[2 * k2 + 1], bh + 2);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) sacc[nt][e] = fmaf(scr[e], X3_INV_SCALE, smn[e]);
  }
  const float sqrt_dk = sqrtf((float)DK);
  const int r0 = 16 * warp + (lane >> 2), r1 = r0 + 8, qd = lane & 3;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 6; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = 8 * nt + 2 * qd + (e & 1);
      const int row = (e & 2) ? r1 : r0;
      const bool vis = row >= q_lo && row < q_hi && key < k_hi;
      const float v = vis ? sacc[nt][e] / sqrt_dk : -INFINITY;
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
  }
  uint32_t pah[3][4], pal[3][4];
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    const float p0 = sum[0] > 0.f ? sacc[nt][0] / sum[0] : 0.f, p1 = sum[0] > 0.f ? sacc[nt][1] / sum[0] : 0.f;
    const float p2 = sum[1] > 0.f ? sacc[nt][2] / sum[1] : 0.f, p3 = sum[1] > 0.f ? sacc[nt][3] / sum[1] : 0.f;
    split_pair(p0, p1, pah[nt >> 1][(nt & 1) * 2 + 0], pal[nt >> 1][(nt & 1) * 2 + 0]);
    split_pair(p2, p3, pah[nt >> 1][(nt & 1) * 2 + 1], pal[nt >> 1][(nt & 1) * 2 + 1]);
  }
#pragma unroll
  for (int n2 = 0; n2 < NDT / 2; ++n2) {
    float om[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, oc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      uint32_t bh[4], bl[4];
      const int off = (16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * RS + 16 * n2 + 8 * (lane >> 4);
      xldsm_x4_trans(bh, Vh + off);
      xldsm_x4_trans(bl, Vl + off);
      mma_f16(om[0], pah[kk], bh);
      mma_f16(om[1], pah[kk], bh + 2);
      mma_f16(oc[0], pah[kk], bl);
      mma_f16(oc[1], pah[kk], bl + 2);
      mma_f16(oc[0], pal[kk], bh);
      mma_f16(oc[1], pal[kk], bh + 2);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = head * DK + 8 * (2 * n2 + j) + 2 * qd;
      const float o0 = fmaf(oc[j][0], X3_INV_SCALE, om[j][0]), o1 = fmaf(oc[j][1], X3_INV_SCALE, om[j][1]);
      const float o2 = fmaf(oc[j][2], X3_INV_SCALE, om[j][2]), o3 = fmaf(oc[j][3], X3_INV_SCALE, om[j][3]);
      if (r0 < kSlots) {
        const size_t row = (size_t)blk_i * kSlots + r0;
        if (so.base) {
          uint32_t h2, l2;
          split_pair(o0, o1, h2, l2);
          *reinterpret_cast<uint32_t*>(so.base + row * so.ld + col) = h2;
          *reinterpret_cast<uint32_t*>(so.base + so.plane + row * so.ld + col) = l2;
        } else { out[row * D + col] = o0; out[row * D + col + 1] = o1; }
      }
      if (r1 < kSlots) {
        const size_t row = (size_t)blk_i * kSlots + r1;
        if (so.base) {
          uint32_t h2, l2;
          split_pair(o2, o3, h2, l2);
          *reinterpret_cast<uint32_t*>(so.base + row * so.ld + col) = h2;
          *reinterpret_cast<uint32_t*>(so.base + so.plane + row * so.ld + col) = l2;
        } else { out[row * D + col] = o2; out[row * D + col + 1] = o3; }
      }
    }
  }
}

int launch_enc_attention_x3(const float* qkv, float* out, const BlockDesc* blk, int n_blk, int n_head, int d_model,
                            SplitOut so, cudaStream_t st) {
  if (n_blk <= 0) return 0;
  static const int head_major = [] { const char* v = getenv("SCB_ATTN_HEAD_MAJOR"); return v ? atoi(v) : 1; }();
  dim3 grid = head_major ? dim3(n_head, n_blk) : dim3(n_blk, n_head);
  const int dk = d_model / n_head;
  if (dk == 32) enc_attn_x3_kernel<32><<<grid, 96, 0, st>>>(qkv, out, blk, d_model, so, head_major);
  else if (dk == 64) enc_attn_x3_kernel<64><<<grid, 96, 0, st>>>(qkv, out, blk, d_model, so, head_major);
  else { set_last_error("enc_attn_x3: unsupported head dim %d", dk); return -1; }
  SCB_LAUNCH_CHECK();
  return 0;
}

template <int DK, int MODE, bool WH>
static int launch_x3_t(const SearchBuffers& sb, __half* kv_layer, const float* q, int ldq, float* out, SplitOut so,
                       cudaStream_t st) {
  constexpr int RS = DK + 8;
  const int nw = WH ? sb.H : 4;
  const size_t stages = sizeof(__half) * 4 * X_NSTAGES * (size_t)nw * X_KPW * DK;   // [warps][stages][4 planes][X_KPW][DK]
  size_t smem = sizeof(__half) * (size_t)(WH ? nw : 1) * 2 * 16 * RS + stages + (((size_t)nw * X_NSTAGES * X_KPW + 15) & ~(size_t)15) + 16;
  if (MODE == 0) smem += sizeof(int) * X_KEYS_SMEM;
  static PerDeviceMark mk;
  size_t& attr = mk.cur();
  if (attr < smem) {
    if (cudaFuncSetAttribute(dec_attn_x3_kernel<DK, MODE, WH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error("attn_x3: cudaFuncSetAttribute(%zu) failed", smem);
      return -1;
    }
    attr = smem;
  }
  const int tiles = (sb.B + 15) / 16;
  const dim3 grid = WH ? dim3(sb.S, 1, tiles) : (sb.attn_head_major ? dim3(sb.H, sb.S, tiles) : dim3(sb.S, sb.H, tiles));
  launch_k(dec_attn_x3_kernel<DK, MODE, WH>, grid, dim3(WH ? 32 * sb.H : 128), smem, st, sb, kv_layer, q, ldq, out, so);
  SCB_LAUNCH_CHECK();
  return 0;
}

// mode 0: self attention (q = fused QKV GEMM output, row stride ldq = 3D; K|V of the new token at +D / +2D), needs
// launch_build_self_keys earlier in the step; mode 1: cross attention.  Requires split-plane K|V caches and beam <= 32.
int launch_dec_attention_x3(const SearchBuffers& sb, int mode, int layer, const float* q, int ldq, float* out,
                            SplitOut so, cudaStream_t st) {
  if (!sb.kv_split || sb.B > 32) { set_last_error("attn_x3: needs split-plane KV caches and beam <= 32"); return -1; }
  const int dk = sb.D / sb.H;
  __half* kv = mode == 0
      ? reinterpret_cast<__half*>(sb.skv) + (size_t)layer * sb.S * sb.Lcap * sb.B * 4 * sb.D
      : reinterpret_cast<__half*>(sb.xkv) + (size_t)layer * sb.S * sb.Tcap * 4 * sb.D;
  // warp = head form: measured slower (a warp then walks four times as many 16-key steps, and the step latency is what
  // bounds the kernel: self 49.6 -> 68.5 us, cross 44.3 -> 55.1 us per launch); opt-in, SCB_ATTN_WARP_HEAD=1; <= 8 heads
  static const bool wh_env = [] { const char* v = getenv("SCB_ATTN_WARP_HEAD"); return v && v[0] == '1'; }();
  const bool wh = wh_env && sb.H <= 8;
  if (dk == 32) {
    if (wh) return mode == 0 ? launch_x3_t<32, 0, true>(sb, kv, q, ldq, out, so, st) : launch_x3_t<32, 1, true>(sb, kv, q, ldq, out, so, st);
    return mode == 0 ? launch_x3_t<32, 0, false>(sb, kv, q, ldq, out, so, st) : launch_x3_t<32, 1, false>(sb, kv, q, ldq, out, so, st);
  }
  if (dk == 64) {
    if (wh) return mode == 0 ? launch_x3_t<64, 0, true>(sb, kv, q, ldq, out, so, st) : launch_x3_t<64, 1, true>(sb, kv, q, ldq, out, so, st);
    return mode == 0 ? launch_x3_t<64, 0, false>(sb, kv, q, ldq, out, so, st) : launch_x3_t<64, 1, false>(sb, kv, q, ldq, out, so, st);
  }
  set_last_error("attn_x3: unsupported head dim %d", dk);
  return -1;
}

}  // namespace scb
