// bf16 tensor-core GEMM for sm_100a: C = act(A * W^T + bias) + R, fp32 accumulation in TMEM.
//
//   * operands: A [M][K] and W [N][K], both K-major bf16, staged by TMA (cp.async.bulk.tensor.2d,
//     128-byte swizzle) into a 4-stage shared-memory ring;
//   * math: tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16, issued by one elected thread,
//     accumulator (128 lanes x BN fp32 columns) lives in tensor memory;
//   * epilogue: four warps read their TMEM lane quarter with tcgen05.ld (32x32b), fuse bias / ReLU /
//     residual and write fp32 and/or bf16 rows (optionally scattered through a row-offset table);
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
//
// This is the bf16 counterpart of every torch.nn.functional.linear on the hot path
// (speechcatcher/model/attention/multi_head_attention.py:79-83,133, layers/feed_forward.py:50,
//  decoder/transformer_decoder.py:249, ctc.py:40).
#include <cuda.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include "kernels.h"
#include "tc_ptx.cuh"

namespace scb {

struct TcParams {
  const float* bias; const float* R; int ldr; float* C; int ldc; __nv_bfloat16* Cb; int ldcb;
  const int64_t* c_row_off; int M, N, K, relu; const int* n_rows_dev;
  // optional fused LayerNorm of the output rows (needs BN == N): y = LN(C_row) * ln_w + ln_b -> bf16 ln_out[M][N]
  const float* ln_w; const float* ln_b; __nv_bfloat16* ln_out;
  // LNA variant: the A operand is LayerNorm(X) computed in the kernel from fp32 rows (K == 256 == all four K-blocks)
  const float* lna_x; int lna_ldx; const float* lna_w; const float* lna_b;
};

template <int BN, int TC_STAGES, bool LNA>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b,
                                                                     TcParams p) {
  // Everything up to pdl_sync() is independent of the preceding kernel (barrier init, TMEM allocation, tensor-map
  // prefetch), so under programmatic dependent launch it overlaps with that kernel's tail.
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = TC_BM * TC_BK * 2, B_BYTES = BN * TC_BK * 2;
  unsigned char* sA = smem;
  unsigned char* sB = smem + TC_STAGES * A_BYTES;
  uint64_t* full_bar = (uint64_t*)(sB + TC_STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tmem_full = empty_bar + TC_STAGES;
  uint64_t* a_ready = tmem_full + 1;         // LNA: the four epilogue warps have written the normalised A tile
  uint32_t* tmem_slot = (uint32_t*)(a_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.K / TC_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(a_ready, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                           // TMEM allocation: BN fp32 columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
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
  } else
  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop and waits on the barriers (uniform control flow); one elected lane issues.
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES, ph = (kb / TC_STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one_sync()) {
        if (LNA) {
          mbar_expect_tx(&full_bar[s], B_BYTES);
        } else {
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          tma_load_2d(&map_a, &full_bar[s], sA + s * A_BYTES, kb * TC_BK, m0);
        }
        tma_load_2d(&map_b, &full_bar[s], sB + s * B_BYTES, kb * TC_BK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Same shape: warp-uniform loop, elected lane issues.  Under `if (lane == 0)` ptxas cannot prove the descriptors
    // uniform and wraps every UTCHMMA in a lane-serialising R2UR loop (~100 cycles per MMA, measured).
    // instruction descriptor: D fp32, A/B bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    if (LNA) mbar_wait(a_ready, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES, ph = (kb / TC_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t da = make_smem_desc(smem_u32(sA + s * A_BYTES));
        const uint64_t db = make_smem_desc(smem_u32(sB + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < TC_BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in 16-byte units
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);           // frees the smem stage once the MMAs have read it
        if (kb == nkb - 1) umma_commit(tmem_full);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    if (LNA) {
      // A tile = bf16(LayerNorm(X rows)) written straight into the 128-byte-swizzled K-major layout TMA would
      // produce: K-block kb = lane / 8, 16-byte chunk c = lane % 8 of tile row R lands at chunk (c ^ (R & 7)).
      // Same arithmetic as layernorm_kernel (two-pass moments, eps 1e-12), so results are bit-identical to the
      // unfused LayerNorm -> GEMM pair.
      float w8[8], b8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { w8[i] = __ldg(p.lna_w + lane * 8 + i); b8[i] = __ldg(p.lna_b + lane * 8 + i); }
      const int kbl = lane >> 3, ch = lane & 7;
      constexpr int RB = 8;                   // rows per batch: all loads of a batch are in flight together
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += RB) {
        float v[RB][8];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int m = m0 + q * 32 + r0 + j;
          if (m < M) {
            const float* xr = p.lna_x + (size_t)m * p.lna_ldx + lane * 8;
            const float4 a = *reinterpret_cast<const float4*>(xr), b = *reinterpret_cast<const float4*>(xr + 4);
            v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w; v[j][4] = b.x; v[j][5] = b.y; v[j][6] = b.z; v[j][7] = b.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int R = q * 32 + r0 + j, m = m0 + R;
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) sm += v[j][i];
          const float mean = warp_sum(sm) / 256.f;
          float qq = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; qq += d * d; }
          const float rstd = 1.0f / sqrtf(warp_sum(qq) / 256.f + 1e-12f);
          uint4 u = make_uint4(0u, 0u, 0u, 0u);
          if (m < M) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = (v[j][i] - mean) * rstd * w8[i] + b8[i];
            __nv_bfloat162 h0 = __floats2bfloat162_rn(y[0], y[1]), h1 = __floats2bfloat162_rn(y[2], y[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(y[4], y[5]), h3 = __floats2bfloat162_rn(y[6], y[7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          }
          *reinterpret_cast<uint4*>(sA + kbl * A_BYTES + R * 128 + ((ch ^ (R & 7)) << 4)) = u;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(a_ready)) : "memory");
    }
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < M;
    float* crow = nullptr;
    __nv_bfloat16* cbrow = nullptr;
    const float* rrow = nullptr;
    if (row_ok) {
      if (p.C) crow = p.c_row_off ? p.C + p.c_row_off[m] : p.C + (size_t)m * p.ldc;
      if (p.Cb) cbrow = p.c_row_off ? p.Cb + p.c_row_off[m] : p.Cb + (size_t)m * p.ldcb;
      if (p.R) rrow = p.R + (size_t)m * p.ldr;
    }
    float s1 = 0.f, s2 = 0.f;
    // 16 columns per step.  With one epilogue warp per scheduler every load latency is exposed, so the loads are
    // hoisted: the residual of the first step is requested BEFORE waiting for the accumulator (it overlaps the
    // main loop), each later step's residual one step ahead, and the bias vector ahead of the TMEM read.
    // (R may alias C: a step only ever stores columns whose residual it has already consumed.)
    float4 rn[4];
    if (rrow) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rn[i] = *reinterpret_cast<const float4*>(rrow + n0 + 4 * i);
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      const int n = n0 + c0;
      float4 bb[4], rr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) bb[i] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t v[16];
      tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
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
        for (int i = 0; i < 4; ++i) {
          o[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb[i].x; o[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb[i].y;
          o[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb[i].z; o[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb[i].w;
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        if (rrow) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { o[4 * i] += rr[i].x; o[4 * i + 1] += rr[i].y; o[4 * i + 2] += rr[i].z; o[4 * i + 3] += rr[i].w; }
        }
        if (p.ln_w) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { s1 += o[j]; s2 = fmaf(o[j], o[j], s2); }
        }
        if (crow) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(crow + n + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
        if (cbrow) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[j], o[j + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(cbrow + n + j) = u;
          }
        }
      }
    }
    if (p.ln_w && row_ok) {
      // fused LayerNorm of the finished row (this thread wrote it above, so its own global stores are visible):
      // single-pass moments in fp32, eps 1e-12 as in the reference's LayerNorm
      const float mean = s1 / (float)BN;
      const float var = fmaxf(s2 / (float)BN - mean * mean, 0.f);
      const float rstd = 1.0f / sqrtf(var + 1e-12f);
      __nv_bfloat16* lrow = p.ln_out + (size_t)m * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 8) {
        const float4 a = *reinterpret_cast<const float4*>(crow + c0);
        const float4 b = *reinterpret_cast<const float4*>(crow + c0 + 4);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ln_w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(p.ln_w + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0 + 4));
        __nv_bfloat162 h0 = __floats2bfloat162_rn((a.x - mean) * rstd * w0.x + b0.x, (a.y - mean) * rstd * w0.y + b0.y);
        __nv_bfloat162 h1 = __floats2bfloat162_rn((a.z - mean) * rstd * w0.z + b0.z, (a.w - mean) * rstd * w0.w + b0.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn((b.x - mean) * rstd * w1.x + b1.x, (b.y - mean) * rstd * w1.y + b1.y);
        __nv_bfloat162 h3 = __floats2bfloat162_rn((b.z - mean) * rstd * w1.z + b1.z, (b.w - mean) * rstd * w1.w + b1.w);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(lrow + c0) = u;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

// ---------------------------------------------------------------- host: tensor maps + launch
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::mutex g_map_mu;
static std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> g_maps;

// caller holds g_map_mu
static int ensure_encode_locked() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    set_last_error("cuTensorMapEncodeTiled entry point not found");
    return -1;
  }
  g_encode = (PFN_encodeTiled)fn;
  return 0;
}

int tc_get_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (ensure_encode_locked()) return -1;
  auto key = std::make_tuple(ptr, rows, cols, ld, box_rows);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return -1; }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return 0;
}

// fp32 row-major [rows][cols] (ld in elements), box = 32 columns (128 bytes) x box_rows, 128-byte swizzle: the staging
// layout of the fused-FFN output tile (TMA store / reduce-add).
int tc_get_map_f32(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  static std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> maps;
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (ensure_encode_locked()) return -1;
  auto key = std::make_tuple(ptr, rows, cols, ld, box_rows);
  auto it = maps.find(key);
  if (it != maps.end()) { *out = it->second; return 0; }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(f32) failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return -1; }
  if (maps.size() > 4096) maps.clear();
  maps[key] = m;
  *out = m;
  return 0;
}

template <int BN, int TC_STAGES, bool LNA = false>
static int launch_bn(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = 1024 + TC_STAGES * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 256;
  static PerDeviceMark attr_mk;
  if (!attr_mk.cur()) {
    if (cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, TC_STAGES, LNA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(smem=%zu) failed", smem);
      return -1;
    }
    attr_mk.cur() = 1;
  }
  dim3 grid(p.N / BN, cdiv(p.M, TC_BM));
  launch_k(gemm_bf16_tc_kernel<BN, TC_STAGES, LNA>, grid, dim3(TC_THREADS), smem, st, ma, mb, p);
  SCB_LAUNCH_CHECK();
  return 0;
}

int launch_gemm_bf16_ln(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, const float* R,
                        int ldr, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, const int64_t* c_row_off, int M, int N,
                        int K, int relu, const int* n_rows_dev, const float* ln_w, const float* ln_b,
                        __nv_bfloat16* ln_out, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (K % TC_BK != 0 || N % 64 != 0 || lda % 8 != 0) {
    set_last_error("gemm_bf16: unsupported shape M=%d N=%d K=%d lda=%d", M, N, K, lda);
    return -1;
  }
  const bool ln = ln_w != nullptr;
  if (ln && (N != 256 || !C || c_row_off || !ln_b || !ln_out || ldc != N)) {
    set_last_error("gemm_bf16: fused LayerNorm needs N == 256 and a dense fp32 output (N=%d)", N);
    return -1;
  }
  const bool small = !ln && ((N % 128 != 0) || ((long)cdiv(M, TC_BM) * (N / 128) < kNumSMs));
  const int BN = ln ? 256 : (small ? 64 : 128);
  CUtensorMap ma, mb;
  if (tc_get_map(A, M, K, lda, TC_BM, &ma)) return -1;
  if (tc_get_map(W, N, K, K, BN, &mb)) return -1;
  TcParams p{bias, R, ldr, C, ldc, Cb, ldcb, c_row_off, M, N, K, relu, n_rows_dev, ln_w, ln_b, ln_out, nullptr, 0, nullptr, nullptr};
  // K <= 256 has only four K-blocks: two stages let several CTAs share an SM so that one CTA's epilogue
  // overlaps another's main loop; deeper K keeps the four-stage ring
  // ... unless the grid is so small (decoder GEMMs) that no two CTAs share an SM anyway: then the deeper ring
  // loads all four K-blocks at once and exposes the TMA latency once instead of twice
  const long n_cta = (long)cdiv(M, TC_BM) * (N / BN);
  static const bool deep_small = [] { const char* v = getenv("SCB_GEMM_DEEP_SMALL"); return !(v && v[0] == '0'); }();
  const bool shallow = K / TC_BK <= 4 && (n_cta > 2 * kNumSMs || !deep_small);
  if (ln) return shallow ? launch_bn<256, 2>(ma, mb, p, st) : launch_bn<256, 4>(ma, mb, p, st);
  if (small) return shallow ? launch_bn<64, 2>(ma, mb, p, st) : launch_bn<64, 4>(ma, mb, p, st);
  return shallow ? launch_bn<128, 2>(ma, mb, p, st) : launch_bn<128, 4>(ma, mb, p, st);
}

// C = act(LayerNorm(X) * W^T + bias): the LayerNorm of the fp32 rows X [M][256] is computed inside the GEMM (K = 256).
int launch_gemm_bf16_lnA(const float* X, int ldx, const float* lna_w, const float* lna_b, const __nv_bfloat16* W,
                         const float* bias, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, int M, int N, int relu,
                         const int* n_rows_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int K = 256;
  if (N % 64 != 0 || ldx % 4 != 0) { set_last_error("gemm_bf16_lnA: unsupported shape M=%d N=%d ldx=%d", M, N, ldx); return -1; }
  const bool small = (N % 128 != 0) || ((long)cdiv(M, TC_BM) * (N / 128) < kNumSMs);
  const int BN = small ? 64 : 128;
  CUtensorMap mb;
  if (tc_get_map(W, N, K, K, BN, &mb)) return -1;
  TcParams p{bias, nullptr, 0, C, ldc, Cb, ldcb, nullptr, M, N, K, relu, n_rows_dev, nullptr, nullptr, nullptr, X, ldx, lna_w, lna_b};
  return small ? launch_bn<64, 4, true>(mb, mb, p, st) : launch_bn<128, 4, true>(mb, mb, p, st);
}

int launch_gemm_bf16_ex(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, const float* R,
                        int ldr, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, const int64_t* c_row_off, int M, int N,
                        int K, int relu, const int* n_rows_dev, cudaStream_t st) {
  return launch_gemm_bf16_ln(A, lda, W, bias, R, ldr, C, ldc, Cb, ldcb, c_row_off, M, N, K, relu, n_rows_dev, nullptr,
                             nullptr, nullptr, st);
}

int launch_gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, const float* R,
                     int ldr, float* C, int ldc, __nv_bfloat16* Cb, int ldcb, int M, int N, int K, int relu,
                     const int* n_rows_dev, cudaStream_t st) {
  return launch_gemm_bf16_ex(A, lda, W, bias, R, ldr, C, ldc, Cb, ldcb, nullptr, M, N, K, relu, n_rows_dev, st);
}

}  // namespace scb
