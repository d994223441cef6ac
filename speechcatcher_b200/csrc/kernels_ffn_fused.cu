// Fused position-wise feed-forward for sm_100a:  Y = X_res + W2 * ReLU(W1 * A + b1) + b2  (A = LayerNorm(X), bf16),
// the reference's PositionwiseFeedForward inside the encoder layer (speechcatcher/model/layers/feed_forward.py:41-50,
// encoder/contextual_block_encoder_layer.py:243-251), as ONE kernel per layer instead of two GEMMs.
//
// Why: with d_model 256 the two un-fused GEMMs sit at the ridge of the machine and are bound by L2 -> SM operand
// traffic (a 128 x 128 x 256 tile re-reads A for every N tile) plus an HBM/L2 round trip of the [M][2048] hidden
// activation.  Here a CTA owns 128 rows: the A tile stays resident in shared memory, the hidden activation never
// leaves the SM (TMEM -> registers -> bf16 -> swizzled shared memory -> operand of the second MMA), and only the
// weights stream through a TMA ring.
//
// Per CTA (one 128-row tile), hidden processed in 16 chunks of 128:
//   G1_j : acc1[j&1] (TMEM, 128 cols)  = A[128 x 256] * W1[j*128 .. +128, :]^T      2 stages x 2 K-blocks, 16 UMMAs N=128
//   E1_j : epilogue warps: acc1 -> +b1 -> ReLU -> bf16 -> H[j&1] (128B-swizzled K-major smem, 2 K-blocks)
//   G2_j : acc2 (TMEM, 256 cols)      += H[j&1][128 x 128] * W2[:, j*128 .. +128]^T  2 stages (K-blocks), 8 UMMAs N=256
//   final: acc2 + b2 -> 128B-swizzled fp32 staging tile in the (now free) A/H shared memory -> TMA reduce-add into
//   the output rows: the in-place residual connection  X += FFN(LN(X))  is performed at the L2, the kernel never
//   reads the residual and never issues a row-per-thread (uncoalesced) global access.
// The MMA warp issues  G1_0, G1_1, { G2_j, G1_{j+2} }  so the tensor pipe runs G1 of the next chunks while the
// epilogue warps convert chunk j; acc1 and H are double-buffered.  TMEM: 2 x 128 + 256 = 512 columns.
// Shared memory: A 64 KB + H 2 x 32 KB + weight ring 3 x 32 KB = 224 KB -> one CTA per SM.
//
// Measured design points (scripts/ffn_timeline.py, clock64 stamps inside the kernel):
//   * the MMA-issue and TMA-issue loops are executed by whole warps with one elected lane issuing; under
//     `if (lane == 0)` ptxas wraps every UTCHMMA in a lane-serialising R2UR loop (~100 cycles per MMA);
//   * UTCHMMA issue is paced by the tensor pipe (shallow queue), so every barrier wait of the issuing warp is pipe idle
//     time: a ring stage is 32 KB = 8 (N=128) or 4 (N=256) UMMAs = 512 pipe cycles per wait;
//   * with one epilogue warp per scheduler all latencies are exposed: eight epilogue warps (two per TMEM lane quarter),
//     bias vectors requested before the TMEM read.
//   * a row-per-thread fp32 residual read / output write of the 128 x 256 tile costs ~8k cycles each (32 L1 tags per
//     instruction); the TMA reduce-add epilogue replaces both.
// Numerics: those of the two-GEMM path (fp32 accumulation, hidden rounded to bf16 after bias + ReLU, then
// (acc + b2) + residual in that order).
#include <stdlib.h>
#include "kernels.h"
#include "tc_ptx.cuh"

namespace scb {

constexpr int FF_D = 256;              // d_model (K of GEMM 1, N of GEMM 2)
constexpr int FF_CH = 128;             // hidden columns per chunk
constexpr int FF_BOX = 128 * 64 * 2;   // one 128-row x 64-col bf16 TMA box = 16 KB
constexpr int FF_STAGE = 2 * FF_BOX;   // ring stage: two boxes
constexpr int FF_RING = 3;
constexpr int FF_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

struct FfnParams {
  const float* b1; const float* b2;
  int M, F, accumulate;   // accumulate: C += tile (in-place residual) instead of C = tile
  const int* n_rows_dev;  // optional device-side row bound (decode step: rows of the active hypotheses)
  long long* dbg;   // optional timeline of CTA 0 (clock64 stamps), see sc_ffn_bf16_timeline
};

__global__ void __launch_bounds__(FF_THREADS, 1) ffn_fused_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                       const __grid_constant__ CUtensorMap map_w1,
                                                                       const __grid_constant__ CUtensorMap map_w2,
                                                                       const __grid_constant__ CUtensorMap map_c,
                                                                       FfnParams p) {
  const int m0 = blockIdx.x * TC_BM;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                              // 4 K-blocks x 16 KB, resident
  unsigned char* sH = sA + 4 * FF_BOX;                   // 2 buffers x (2 K-blocks x 16 KB)
  unsigned char* sW = sH + 4 * FF_BOX;                   // weight ring, FF_RING stages of 32 KB
  uint64_t* full_bar = (uint64_t*)(sW + FF_RING * FF_STAGE);
  uint64_t* empty_bar = full_bar + FF_RING;
  uint64_t* a_full = empty_bar + FF_RING;
  uint64_t* acc1_full = a_full + 1;                      // [2]
  uint64_t* h_ready = acc1_full + 2;                     // [2]  256 arrivals (epilogue threads)
  uint64_t* h_free = h_ready + 2;                        // [2]  G2_j has finished reading H[j&1]
  uint64_t* acc2_full = h_free + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc2_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // gridDim.y > 1: the hidden dimension is split across CTAs; each adds its partial tile to C at the L2 (TMA
  // reduce-add), so a small-M launch (decode step) still spreads over many SMs.  Split 0 carries b2.
  const int n_chunks = p.F / FF_CH / (int)gridDim.y;
  const int jb = blockIdx.y * n_chunks;          // first hidden chunk of this CTA

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int i = 0; i < FF_RING; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&h_ready[i], 256); mbar_init(&h_free[i], 1); }
    mbar_init(acc2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc2 = tmem_base + 256;

  pdl_sync();
  int M = p.M;
  if (p.n_rows_dev) M = min(M, *p.n_rows_dev);
  const bool cta_active = m0 < M;
  const bool stamp = p.dbg && blockIdx.x == 0;

  if (!cta_active) {
    // nothing to compute
  } else if (warp == 0) {
    // ===================== TMA producer: A once, then the weight stages in the order the MMA warp consumes them.
    // The whole warp runs the loop (uniform control flow); one elected lane issues.
    if (elect_one_sync()) {
      mbar_expect_tx(a_full, 4 * FF_BOX);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(&map_a, a_full, sA + kb * FF_BOX, kb * TC_BK, m0);
    }
    __syncwarp();
    int c = 0;
    // one stage = two boxes: (col0,row0) and (col1,row1) of the same tensor map
    auto put = [&](const CUtensorMap* map, int col0, int row0, int col1, int row1) {
      const int s = c % FF_RING, ph = (c / FF_RING) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one_sync()) {
        if (stamp && c < 64) p.dbg[256 + c] = clock64();      // stage free, loads about to be issued
        mbar_expect_tx(&full_bar[s], FF_STAGE);
        tma_load_2d(map, &full_bar[s], sW + s * FF_STAGE, col0, row0);
        tma_load_2d(map, &full_bar[s], sW + s * FF_STAGE + FF_BOX, col1, row1);
      }
      __syncwarp();
      ++c;
    };
    // W1 chunk j: stage t holds K-blocks 2t, 2t+1 of rows [j*128, j*128+128)
    auto put_w1 = [&](int j) { for (int t = 0; t < 2; ++t) put(&map_w1, (2 * t) * TC_BK, (jb + j) * FF_CH, (2 * t + 1) * TC_BK, (jb + j) * FF_CH); };
    // W2 chunk j: stage kb holds all 256 output rows of K-block kb (columns j*128 + kb*64 ..)
    auto put_w2 = [&](int j) { for (int kb = 0; kb < 2; ++kb) put(&map_w2, (jb + j) * FF_CH + kb * TC_BK, 0, (jb + j) * FF_CH + kb * TC_BK, 128); };
    put_w1(0);
    if (n_chunks > 1) put_w1(1);
    for (int j = 0; j < n_chunks; ++j) {
      put_w2(j);
      if (j + 2 < n_chunks) put_w1(j + 2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp waits on the barriers, one elected lane issues
    // instruction descriptors: D fp32, A/B bf16, K-major both, M = 128; N = 128 (GEMM 1) / 256 (GEMM 2)
    const uint32_t idesc_hi = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc1 = idesc_hi | ((uint32_t)(128 >> 3) << 17);
    const uint32_t idesc2 = idesc_hi | ((uint32_t)(256 >> 3) << 17);
    int c = 0;
    auto g1 = [&](int j) {
      const uint32_t d = tmem_base + (uint32_t)((j & 1) * 128);
      for (int t = 0; t < 2; ++t) {
        const int s = c % FF_RING, ph = (c / FF_RING) & 1;
        if (stamp && c < 64 && lane == 0) p.dbg[384 + c] = clock64();
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (stamp && c < 64 && lane == 0) p.dbg[512 + c] = clock64();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t da = make_smem_desc(smem_u32(sA + (2 * t + kk) * FF_BOX));
            const uint64_t db = make_smem_desc(smem_u32(sW + s * FF_STAGE + kk * FF_BOX));
#pragma unroll
            for (int k = 0; k < TC_BK / UMMA_K; ++k) umma_bf16(d, da + 2 * k, db + 2 * k, idesc1, (t | kk | k) != 0);
          }
          umma_commit(&empty_bar[s]);
          if (t == 1) umma_commit(&acc1_full[j & 1]);
        }
        __syncwarp();
        ++c;
      }
    };
    auto g2 = [&](int j) {
      const int b = j & 1;
      mbar_wait(&h_ready[b], (j >> 1) & 1);
      tc_fence_after();
      if (stamp && lane == 0) p.dbg[8 + j * 4 + 3] = clock64();   // H_j ready
      for (int kb = 0; kb < 2; ++kb) {
        const int s = c % FF_RING, ph = (c / FF_RING) & 1;
        if (stamp && c < 64 && lane == 0) p.dbg[384 + c] = clock64();
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (stamp && c < 64 && lane == 0) p.dbg[512 + c] = clock64();
        if (elect_one_sync()) {
          const uint64_t da = make_smem_desc(smem_u32(sH + (b * 2 + kb) * FF_BOX));
          const uint64_t db = make_smem_desc(smem_u32(sW + s * FF_STAGE));     // 256 rows: the two boxes are contiguous
#pragma unroll
          for (int k = 0; k < TC_BK / UMMA_K; ++k) umma_bf16(tm_acc2, da + 2 * k, db + 2 * k, idesc2, (j | kb | k) != 0);
          umma_commit(&empty_bar[s]);
          if (kb == 1) {
            umma_commit(&h_free[b]);
            if (j == n_chunks - 1) umma_commit(acc2_full);
          }
        }
        __syncwarp();
        ++c;
      }
    };
    if (stamp && lane == 0) p.dbg[0] = clock64();
    mbar_wait(a_full, 0);
    tc_fence_after();
    if (stamp && lane == 0) p.dbg[1] = clock64();
    g1(0);
    if (n_chunks > 1) g1(1);
    for (int j = 0; j < n_chunks; ++j) {
      if (stamp && lane == 0) p.dbg[8 + j * 4 + 0] = clock64();      // before waiting for H_j
      g2(j);
      if (stamp && lane == 0) p.dbg[8 + j * 4 + 1] = clock64();      // G2_j issued
      if (j + 2 < n_chunks) g1(j + 2);
      if (stamp && lane == 0) p.dbg[8 + j * 4 + 2] = clock64();      // G1_{j+2} issued
    }
  } else {
    // ===================== epilogue warps 2..9: thread <-> accumulator row; the two warps of a TMEM lane quarter
    // split the columns (with one warp per scheduler every bias load and TMEM read latency would be exposed)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int R = q * 32 + lane;                     // row inside the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool estamp = stamp && warp == 2 && lane == 0;
    for (int j = 0; j < n_chunks; ++j) {
      const int b = j & 1;
      if (estamp) p.dbg[128 + j * 4 + 0] = clock64();
      mbar_wait(&acc1_full[b], (j >> 1) & 1);
      if (estamp) p.dbg[128 + j * 4 + 1] = clock64();   // acc1_j complete
      if (j >= 2) mbar_wait(&h_free[b], ((j - 2) >> 1) & 1);
      tc_fence_after();
      if (estamp) p.dbg[128 + j * 4 + 2] = clock64();   // H buffer free
      // this warp converts hidden columns [half*64, half*64+64) of the chunk = K-block `half` of H[b]
      const float4* b1v = reinterpret_cast<const float4*>(p.b1 + (jb + j) * FF_CH + half * 64);
      unsigned char* tile = sH + (b * 2 + half) * FF_BOX + R * 128;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float4 bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = __ldg(b1v + g * 8 + i);      // in flight while the TMEM read completes
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)(b * 128 + half * 64 + g * 32), v);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 ba = bb[2 * t], bc = bb[2 * t + 1];
          const float o0 = fmaxf(__uint_as_float(v[t * 8 + 0]) + ba.x, 0.f), o1 = fmaxf(__uint_as_float(v[t * 8 + 1]) + ba.y, 0.f);
          const float o2 = fmaxf(__uint_as_float(v[t * 8 + 2]) + ba.z, 0.f), o3 = fmaxf(__uint_as_float(v[t * 8 + 3]) + ba.w, 0.f);
          const float o4 = fmaxf(__uint_as_float(v[t * 8 + 4]) + bc.x, 0.f), o5 = fmaxf(__uint_as_float(v[t * 8 + 5]) + bc.y, 0.f);
          const float o6 = fmaxf(__uint_as_float(v[t * 8 + 6]) + bc.z, 0.f), o7 = fmaxf(__uint_as_float(v[t * 8 + 7]) + bc.w, 0.f);
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(o4, o5), h3 = __floats2bfloat162_rn(o6, o7);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(tile + (((g * 4 + t) ^ (R & 7)) << 4)) = u;
        }
      }
      tc_fence_before();                                             // TMEM accesses done before the MMA warp proceeds
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_ready[b])) : "memory");
      if (estamp) p.dbg[128 + j * 4 + 3] = clock64();   // E1_j done
    }
    // ---- final epilogue.  All MMAs are complete, so sA and sH (128 KB, contiguous) are free: they become eight
    // 128B-swizzled fp32 staging tiles [128 rows][32 cols].  Each warp converts its 32 rows x 128 columns, then one
    // elected lane hands its four 32 x 32 boxes to the TMA engine (store, or fp32 add at the L2 for the residual).
    mbar_wait(acc2_full, 0);
    tc_fence_after();
    if (estamp) p.dbg[2] = clock64();
#pragma unroll 1
    for (int g = half * 4; g < half * 4 + 4; ++g) {
      float4 bb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        bb[i] = blockIdx.y == 0 ? __ldg(reinterpret_cast<const float4*>(p.b2 + g * 32) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t v[32];
      tmem_ld32(tm_acc2 + lane_base + (uint32_t)(g * 32), v);
      unsigned char* row = sA + g * FF_BOX + R * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 o = make_float4(__uint_as_float(v[4 * i]) + bb[i].x, __uint_as_float(v[4 * i + 1]) + bb[i].y,
                                     __uint_as_float(v[4 * i + 2]) + bb[i].z, __uint_as_float(v[4 * i + 3]) + bb[i].w);
        *reinterpret_cast<float4*>(row + ((i ^ (R & 7)) << 4)) = o;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one_sync()) {
      for (int g = half * 4; g < half * 4 + 4; ++g) {
        const void* src = sA + g * FF_BOX + q * 32 * 128;
        if (p.accumulate || gridDim.y > 1) tma_reduce_add_2d(&map_c, src, g * 32, m0 + q * 32);
        else tma_store_2d(&map_c, src, g * 32, m0 + q * 32);
      }
      tma_store_commit();
      tma_store_wait_all();      // the staging tile must outlive the copy; completes before the CTA exits
    }
    __syncwarp();
    tc_fence_before();
    if (estamp) p.dbg[3] = clock64();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// C[M][256] (+)= W2 * ReLU(W1 * A + b1) + b2.  A [M][256] bf16 (lda), W1 [F][256] bf16, W2 [256][F] bf16; F % 128 == 0.
// accumulate != 0: C already holds the residual and the tile is added to it (TMA reduce-add), else C is overwritten.
// splits > 1 (needs accumulate): the hidden dimension is divided over `splits` CTAs per row tile; the partial tiles
// meet in C through fp32 adds at the L2, so the result depends on their arrival order at rounding level.
int launch_ffn_fused_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W1, const float* b1,
                          const __nv_bfloat16* W2, const float* b2, float* C, int ldc, int accumulate, int M, int F,
                          int splits, const int* n_rows_dev, cudaStream_t st, long long* dbg) {
  if (M <= 0) return 0;
  if (splits < 1) splits = 1;
  if (F % (FF_CH * splits) != 0 || F < FF_CH || lda % 8 != 0 || ldc % 4 != 0 || !C || (splits > 1 && !accumulate)) {
    set_last_error("ffn_fused: unsupported shape M=%d F=%d lda=%d ldc=%d splits=%d accumulate=%d", M, F, lda, ldc, splits, accumulate);
    return -1;
  }
  CUtensorMap ma, mw1, mw2, mc;
  if (tc_get_map(A, M, FF_D, lda, TC_BM, &ma)) return -1;
  if (tc_get_map(W1, F, FF_D, FF_D, 128, &mw1)) return -1;
  if (tc_get_map(W2, FF_D, F, F, 128, &mw2)) return -1;
  if (tc_get_map_f32(C, M, FF_D, ldc, 32, &mc)) return -1;
  constexpr size_t smem = 1024 + (size_t)(4 + 4) * FF_BOX + (size_t)FF_RING * FF_STAGE + 256;
  static PerDeviceMark attr_mk;
  if (!attr_mk.cur()) {
    if (cudaFuncSetAttribute(ffn_fused_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(ffn_fused, smem=%zu) failed", smem);
      return -1;
    }
    attr_mk.cur() = 1;
  }
  FfnParams p{b1, b2, M, F, accumulate, n_rows_dev, dbg};
  launch_k(ffn_fused_bf16_kernel, dim3(cdiv(M, TC_BM), splits), dim3(FF_THREADS), smem, st, ma, mw1, mw2, mc, p);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
