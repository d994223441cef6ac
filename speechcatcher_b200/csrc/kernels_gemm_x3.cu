// Split-precision tensor-core GEMM for sm_100a ("x3"): C = act(A * W^T + bias) + R with fp32 operands and
// fp32-class accuracy on the tcgen05 tensor cores.
//
// Why: the reference runs every torch.nn.functional.linear of the path in fp32 (MKL on the CPU), and the n-best of
// the block-synchronous beam search flips on score differences of ~1e-4, which plain bf16
// operands (error ~1e-2 on the log-probs) cannot resolve.  Each fp32 operand x is therefore split into two fp16 values
//     x_hi = fp16(x),   x_lo = fp16((x - x_hi) * 2^11)          (x - x_hi is exact; |x - (x_hi + x_lo 2^-11)| <= 2^-22 |x|)
// and the product is evaluated as three tensor-core GEMMs into two TMEM accumulators
//     D0 += A_hi W_hi^T          D1 += A_hi W_lo^T + A_lo W_hi^T          C = D0 + 2^-11 D1
// (the dropped A_lo W_lo^T term is 2^-22 relative).  fp16 x fp16 products are exact in fp32 and the accumulators are
// fp32, so the result carries fp32-class error at three times the bf16 tensor-core cost instead of the CUDA-core fp32
// FMA rate.  Keeping the correction terms in their own accumulator (scaled by 2^11) avoids both fp16 underflow of the
// low parts and absorption of the small terms into the large sum.
//
// Accumulator rounding (measured on the B200, scripts/tc_accum_probe.py -> profiles/r2_tc_accum_probe.json): every
// tcgen05.mma adds its 16 products to the fp32 accumulator with TRUNCATION (all-positive data: mean signed error
// -3.1e-7 relative at K = 256, -4.3e-6 at K = 2048, i.e. ~1 ulp(2^-24) lost per instruction, where a round-to-nearest
// FMA chain shows 1.2e-7 / 5.7e-7 rms and no bias).  The main term is therefore spread round-robin over J accumulators
// (K step g -> D0[g % J]) that are summed with round-to-nearest adds in the epilogue: each accumulator sees 1/J of the
// instructions and 1/J of the magnitude, which brings the truncation error back to the level of an fp32 FMA chain.
//
//   * W is split once at load time into two K-major fp16 planes [2][N][K] and staged by TMA (128-byte swizzle);
//   * A stays fp32 in global memory (optionally gathered through a row-offset table and per-K-segment offsets: the
//     implicit-GEMM conv2): four converter warps load fp32 rows, split them in registers and write the hi / lo tiles
//     straight into the 128-byte-swizzled K-major shared-memory layout the MMA reads (generic-proxy stores + proxy fence);
//   * warp 0 = TMA producer (W planes), warp 1 = TMEM allocator + MMA issuer (tcgen05.mma.kind::f16, fp16 operands,
//     M = 128, N = BN, K = 16; twelve UMMAs per 64-wide K block), warps 2..5 = A converters, then the epilogue
//     (tcgen05.ld, sum_j D0[j] + 2^-11 D1, bias / ReLU / residual, fp32 rows, optional row scatter).
//
// This is the precise-mode counterpart of every torch.nn.functional.linear on the hot path
// (speechcatcher/model/attention/multi_head_attention.py:79-83,133, layers/feed_forward.py:50,
//  decoder/transformer_decoder.py:249, ctc.py:40, encoder/subsampling.py:87-105).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "kernels.h"
#include "tc_ptx.cuh"
#include "x3_split.cuh"

namespace scb {

constexpr int X3_THREADS = 192;
constexpr int X3_CONV_THREADS = 128;         // warps 2..5

struct X3Params {
  const float* A; int lda; const int64_t* a_row_off; const int* a_seg_off; int seg_len;
  const float* bias; const float* R; int ldr; float* C; int ldc; const int64_t* c_row_off;
  int M, N, K, relu; const int* n_rows_dev;
  __half* C2; size_t c2_plane; int ldc2;      // optional split-plane output (hi plane, lo plane at + c2_plane elements)
};

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}

constexpr int x3_tmem_cols(int need) { return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }

// A_TMA: the A operand already exists as split planes in global memory (written by the producing kernel) and is staged
// by TMA like the weights; otherwise the converter warps build the planes from fp32 rows (gathered A: conv2).
template <int BN, int STAGES, int J, bool A_TMA, int MINB = 1>
__global__ void __launch_bounds__(X3_THREADS, MINB) gemm_x3_kernel(const __grid_constant__ CUtensorMap map_wh,
                                                                const __grid_constant__ CUtensorMap map_wl,
                                                                const __grid_constant__ CUtensorMap map_ah,
                                                                const __grid_constant__ CUtensorMap map_al,
                                                                X3Params p) {
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = TC_BM * TC_BK * 2;           // one fp16 plane of the A stage (16 KB)
  constexpr int B_BYTES = BN * TC_BK * 2;              // one fp16 plane of the W stage
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  uint64_t* w_full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* a_full = w_full + STAGES;
  uint64_t* empty_bar = a_full + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.K / TC_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
    if (A_TMA) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
    }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&a_full[i], X3_CONV_THREADS); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  static_assert((J + 1) * BN <= 512, "TMEM holds 512 fp32 columns");
  constexpr int TM_COLS = x3_tmem_cols((J + 1) * BN);
  if (warp == 1) {                           // J main accumulators + the correction accumulator, BN fp32 columns each
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_sync();                                // from here on the producer kernel's writes (A, n_rows) are visible
  int M = p.M;
  if (p.n_rows_dev) M = min(M, *p.n_rows_dev);
  const bool cta_active = m0 < M;            // uniform for the whole CTA; inactive CTAs only tear down

  if (!cta_active) {
    // nothing to compute
  } else if (warp == 0) {
    // ===================== TMA producer: the two weight planes of every K block =====================
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one_sync()) {
        unsigned char* sA = smem + s * STAGE_BYTES;
        unsigned char* sW = sA + 2 * A_BYTES;
        mbar_expect_tx(&w_full[s], A_TMA ? 2 * A_BYTES + 2 * B_BYTES : 2 * B_BYTES);
        if (A_TMA) {
          tma_load_2d(&map_ah, &w_full[s], sA, kb * TC_BK, m0);
          tma_load_2d(&map_al, &w_full[s], sA + A_BYTES, kb * TC_BK, m0);
        }
        tma_load_2d(&map_wh, &w_full[s], sW, kb * TC_BK, n0);
        tma_load_2d(&map_wl, &w_full[s], sW + B_BYTES, kb * TC_BK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t d1 = tmem_base + J * BN;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, ph = (kb / STAGES) & 1;
      mbar_wait(&w_full[s], ph);
      if (!A_TMA) mbar_wait(&a_full[s], ph);
      tc_fence_after();
      if (elect_one_sync()) {
        unsigned char* st = smem + s * STAGE_BYTES;
        const uint64_t ah = make_smem_desc(smem_u32(st)), al = make_smem_desc(smem_u32(st + A_BYTES));
        const uint64_t wh = make_smem_desc(smem_u32(st + 2 * A_BYTES)), wl = make_smem_desc(smem_u32(st + 2 * A_BYTES + B_BYTES));
#pragma unroll
        for (int k = 0; k < TC_BK / UMMA_K; ++k) {
          // 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in 16-byte units
          const int g = kb * (TC_BK / UMMA_K) + k;              // K step -> main accumulator g % J
          umma_f16(tmem_base + (uint32_t)((g % J) * BN), ah + 2 * k, wh + 2 * k, idesc, g >= J);
          umma_f16(d1, ah + 2 * k, wl + 2 * k, idesc, g != 0);
          umma_f16(d1, al + 2 * k, wh + 2 * k, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);           // frees the stage once the MMAs have read it
        if (kb == nkb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== A converters (warps 2..5), then the epilogue =====================
    if (!A_TMA) {
    const int ct = threadIdx.x - 64;          // 0..127
    const int ch = ct & 7;                    // 16-byte chunk of the fp16 row = 8 consecutive k
    const int rg = ct >> 3;                   // rows rg + 16 i
    const float* arow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + rg + 16 * i;
      arow[i] = nullptr;
      if (m < M) arow[i] = p.a_row_off ? p.A + p.a_row_off[m] : p.A + (size_t)m * p.lda;
    }
    float4 v[8][2];
    auto load_kb = [&](int kb) {
      int koff = kb * TC_BK + ch * 8;
      if (p.a_seg_off) {
        const int seg = (kb * TC_BK) / p.seg_len;
        koff = p.a_seg_off[seg] + (kb * TC_BK - seg * p.seg_len) + ch * 8;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (arow[i]) {
          v[i][0] = *reinterpret_cast<const float4*>(arow[i] + koff);
          v[i][1] = *reinterpret_cast<const float4*>(arow[i] + koff + 4);
        } else {
          v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
          v[i][1] = v[i][0];
        }
      }
    };
    load_kb(0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      unsigned char* sAh = smem + s * STAGE_BYTES;
      unsigned char* sAl = sAh + A_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int Rr = rg + 16 * i;
        const float x[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
        uint4 uh, ul;
        x3_split8(x, uh, ul);
        const int off = Rr * 128 + ((ch ^ (Rr & 7)) << 4);
        *reinterpret_cast<uint4*>(sAh + off) = uh;
        *reinterpret_cast<uint4*>(sAl + off) = ul;
      }
      if (kb + 1 < nkb) load_kb(kb + 1);       // next block's global loads are in flight while the MMAs of this one run
      fence_proxy_async_smem();                // generic-proxy stores -> visible to the tensor core (async proxy)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[s])) : "memory");
    }
    }

    // ---- epilogue: thread <-> accumulator row (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < M;
    float* crow = nullptr;
    const float* rrow = nullptr;
    __half* c2row = nullptr;
    if (row_ok) {
      if (p.C) crow = p.c_row_off ? p.C + p.c_row_off[m] : p.C + (size_t)m * p.ldc;
      if (p.R) rrow = p.R + (size_t)m * p.ldr;
      // split-plane rows; with a row-offset table (given in fp32 elements of a [hi | lo] row, e.g. the K|V cache rows) the
      // planes of a row are adjacent: hi at 2 * offset, lo c2_plane elements further
      if (p.C2) c2row = p.c_row_off ? p.C2 + 2 * p.c_row_off[m] : p.C2 + (size_t)m * p.ldc2;
    }
    float4 rn[4];
    if (rrow) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rn[i] = *reinterpret_cast<const float4*>(rrow + n0 + 4 * i);
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      const int n = n0 + c0;
      float4 bb[4], rr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) bb[i] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t v0[J][16], v1[16];
      // (K >= 64 = 4 K steps >= J: every main accumulator has received at least one instruction)
#pragma unroll
      for (int j = 0; j < J; ++j) tmem_ld16_nowait(lane_base + (uint32_t)(j * BN + c0), v0[j]);
      tmem_ld16_nowait(lane_base + (uint32_t)(J * BN + c0), v1);
#pragma unroll
      for (int i = 0; i < 4; ++i) rr[i] = rn[i];
      if (rrow && c0 + 16 < BN) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rn[i] = *reinterpret_cast<const float4*>(rrow + n + 16 + 4 * i);
      }
      tmem_ld_wait();
      if (row_ok) {
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(v0[0][j]);
          if (J == 2) { a += __uint_as_float(v0[1][j]); }
          if (J == 4) {
            a += __uint_as_float(v0[1][j]);
            a += __uint_as_float(v0[2][j]) + __uint_as_float(v0[3][j]);
          }
          if (J == 7) {
            a += __uint_as_float(v0[1][j]);
            a += __uint_as_float(v0[2][j]) + __uint_as_float(v0[3][j]);
            a += (__uint_as_float(v0[4][j]) + __uint_as_float(v0[5][j])) + __uint_as_float(v0[6][j]);
          }
          o[j] = fmaf(__uint_as_float(v1[j]), X3_INV_SCALE, a);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[4 * i] += bb[i].x; o[4 * i + 1] += bb[i].y; o[4 * i + 2] += bb[i].z; o[4 * i + 3] += bb[i].w; }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        if (rrow) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { o[4 * i] += rr[i].x; o[4 * i + 1] += rr[i].y; o[4 * i + 2] += rr[i].z; o[4 * i + 3] += rr[i].w; }
        }
        if (crow) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(crow + n + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
        if (c2row) {                         // the consumer is another Linear: hand the result over as split planes
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 uh, ul;
            x3_split8(o + j, uh, ul);
            *reinterpret_cast<uint4*>(c2row + n + j) = uh;
            *reinterpret_cast<uint4*>(c2row + p.c2_plane + n + j) = ul;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TM_COLS));
  }
}

template <int BN, int STAGES, int J, bool A_TMA, int MINB = 1>
static int x3_launch(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mah, const CUtensorMap& mal,
                     const X3Params& p, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)STAGES * (2 * TC_BM * TC_BK * 2 + 2 * BN * TC_BK * 2) + 256;
  static PerDeviceMark attr_mk;
  if (!attr_mk.cur()) {
    if (cudaFuncSetAttribute(gemm_x3_kernel<BN, STAGES, J, A_TMA, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm_x3, smem=%zu) failed", smem);
      return -1;
    }
    attr_mk.cur() = 1;
  }
  dim3 grid(p.N / BN, cdiv(p.M, TC_BM));
  launch_k(gemm_x3_kernel<BN, STAGES, J, A_TMA, MINB>, grid, dim3(X3_THREADS), smem, st, mh, ml, mah, mal, p);
  SCB_LAUNCH_CHECK();
  return 0;
}

// W2: the two fp16 planes [2][N][K] produced by weights.split_f16 (hi plane, then lo plane scaled by 2^11).
// A: fp32 rows (g.A, optionally gathered) or split planes (x.A2: hi plane [a2_rows][K] with leading dimension g.lda,
// lo plane a2_plane elements further; a2_rows = row capacity of the buffer, so that one tensor map serves every M).
// Output: fp32 rows (g.C, optional when x.C2 is given) and / or split planes (x.C2).
int launch_gemm_x3(const GemmArgs& g, const X3Extra& x, const void* W2, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  // encoder-sized products with plane operands: the persistent kernel (SCB_X3_PERSIST_MIN_M rows and up; 0 disables)
  static const int persist_min_m = [] { const char* v = getenv("SCB_X3_PERSIST_MIN_M"); return v ? atoi(v) : 1; }();
  if (x.kernel == 2 || x.kernel == 3 || x.lnX ||
      (x.kernel == 0 && !g.n_rows_dev && persist_min_m > 0 && g.M >= persist_min_m && gemm_x3p_eligible(g, x))) {
    // K = 256 without a residual can take the form with A in tensor memory (kernels_gemm_x3t.cu).  Verified bit-identical to
    // the shared-memory form but measured slower as it stands (QKV 20 -> 33 us, FFN1 41 -> 55 us: its thread-per-row A loader
    // costs ~12 us per row tile and the direct-store epilogue is slower than the TMA one, while the UMMA rate only improves
    // from one 128x128x16 per ~108 cycles to the equivalent of ~92): opt-in, SCB_X3T=1.
    static const bool use_ts = [] { const char* v = getenv("SCB_X3T"); return v && v[0] == '1'; }();
    if (x.kernel != 3 && use_ts && gemm_x3t_eligible(g, x)) return launch_gemm_x3t(g, x, W2, st);
    return launch_gemm_x3p(g, x, W2, st);
  }
  const bool a_tma = x.A2 != nullptr;
  if (g.K % TC_BK != 0 || g.N % 64 != 0 || (g.a_seg_off && g.seg_len % TC_BK != 0) || (!g.a_row_off && g.lda % 8 != 0) || !W2 ||
      (!g.C && !x.C2) || (a_tma && (g.a_row_off || x.a2_rows < g.M)) || (x.C2 && x.ldc2 % 8 != 0)) {
    set_last_error("gemm_x3: unsupported shape M=%d N=%d K=%d lda=%d seg=%d", g.M, g.N, g.K, g.lda, g.seg_len);
    return -1;
  }
  // tile shapes: 128-wide N tiles with 2 main accumulators when that still fills the GPU, else 64-wide with 4; long-K
  // products (FFN2, conv2, embed.out: 128+ tensor-core instructions per output) take 7 main accumulators so that
  // the truncating accumulation stays at the error level of a K = 256 product
  const bool long_k = g.K >= 1024;
  const bool small = long_k || (g.N % 128 != 0) || ((long)cdiv(g.M, TC_BM) * (g.N / 128) < kNumSMs);
  const int BN = small ? 64 : 128;
  const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(W2);        // 16-bit elements: the map only moves bytes
  const __nv_bfloat16* wl = wh + (size_t)g.N * g.K;
  CUtensorMap mh, ml, mah, mal;
  if (tc_get_map(wh, g.N, g.K, g.K, BN, &mh)) return -1;
  if (tc_get_map(wl, g.N, g.K, g.K, BN, &ml)) return -1;
  if (a_tma) {
    const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(x.A2);
    if (tc_get_map(ah, x.a2_rows, g.K, g.lda, TC_BM, &mah)) return -1;
    if (tc_get_map(ah + x.a2_plane, x.a2_rows, g.K, g.lda, TC_BM, &mal)) return -1;
  } else { mah = mh; mal = ml; }
  X3Params p{g.A, g.lda, g.a_row_off, g.a_seg_off, g.seg_len, g.bias, g.R, g.ldr, g.C, g.ldc, g.c_row_off,
             g.M, g.N, g.K, g.relu, g.n_rows_dev, (__half*)x.C2, x.c2_plane, x.ldc2};
  // stage = 32 KB (A hi + lo) + 2 * BN * 128 B (W hi + lo): BN 128 -> 64 KB (3 stages), BN 64 -> 48 KB (4 stages)
  // decode-step projections (device-side row count, K = 256: four K blocks): a two-stage ring and two main accumulators
  // (8 instructions each, like the 128-wide tiles) make the CTA small enough -- 97 KB of shared memory, 256 TMEM columns
  // -- for two per SM, so that one CTA's epilogue overlaps the other's loads and the next kernel of the chain can
  // become resident early (SCB_X3_DEC2=0 keeps the four-stage form)
  static const bool dec2 = [] { const char* v = getenv("SCB_X3_DEC2"); return v && v[0] == '1'; }();
  if (a_tma && dec2 && g.n_rows_dev && g.K == 256 && small) return x3_launch<64, 2, 2, true, 2>(mh, ml, mah, mal, p, st);
  if (a_tma) {
    if (long_k) return x3_launch<64, 4, 7, true>(mh, ml, mah, mal, p, st);
    return small ? x3_launch<64, 4, 4, true>(mh, ml, mah, mal, p, st) : x3_launch<128, 3, 2, true>(mh, ml, mah, mal, p, st);
  }
  if (long_k) return x3_launch<64, 4, 7, false>(mh, ml, mah, mal, p, st);
  return small ? x3_launch<64, 4, 4, false>(mh, ml, mah, mal, p, st) : x3_launch<128, 3, 2, false>(mh, ml, mah, mal, p, st);
}

}  // namespace scb
