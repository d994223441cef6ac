// Persistent split-precision tensor-core GEMM with the A operand in TENSOR MEMORY ("x3t"; K = 256, encoder QKV / FFN1).
//     C = act(A * W^T + bias)        A, W as split fp16 planes (x3_split.cuh), fp32-class result
//
// Why: the shared-memory form (kernels_gemm_x3p.cu) issues one 128 x 128 x 16 UMMA per ~108 cycles instead of 64
// (clock64 instrumentation on the B200, profiles/r2_gemm_x3p_issuer_waits.txt: the issuer spends 72 % of its time
// issuing, not waiting): with both operands in shared memory every UMMA reads 4 KB of A and 4 KB of B -- 128 B/cycle,
// the whole shared-memory bandwidth of the SM -- on top of the TMA writes of the W ring.  tcgen05.mma can take A from
// tensor memory instead.  Here
//   * the row tile's A planes (128 rows x 256 K, hi + lo) live in TMEM: 2 fp16 per 32-bit column, 128 columns per plane;
//     four loader warps (one per TMEM lane quarter, thread <-> row) read the rows from global memory and tcgen05.st them,
//     once per row tile;
//   * shared memory holds nothing but the W ring (13 stages x 16 KB: W hi | W lo of a 64-column tile x 64 k), so the ring
//     is deep enough to cover the TMA latency and the UMMAs only read B from it (64 B/cycle);
//   * tiles are 128 x 64: two accumulator stages of (main 64 + correction 64) columns next to A's 256 columns;
//     K chunks of 128 drained into registers with round-to-nearest adds as in kernels_gemm_x3p.cu;
//   * the eight epilogue warps (thread <-> row, 32 columns each) store straight to global memory (fp32 rows or split
//     planes): no staging traffic in shared memory.
// Status: bit-identical to the shared-memory form (tests/test_gpu_gemm_x3.py) but SLOWER as it stands (QKV 20 -> 33 us,
// FFN1 41 -> 55 us on one push of 256 streams): the thread-per-row loader takes ~24 k cycles per row tile and is not
// overlapped with the UMMAs of the previous row tile (no room for a second A in TMEM), the direct-store epilogue makes the
// issuer wait ~1.3 k cycles per tile, and the UMMA rate itself only improves from ~106 cycles per 128x128x16 to the
// equivalent of ~92 (profiles/r2_gemm_x3p_issuer_waits.txt).  Opt-in (SCB_X3T=1); the engine uses kernels_gemm_x3p.cu.
// Replaces torch.nn.functional.linear of the encoder layers where K = d_model
// (speechcatcher/model/attention/multi_head_attention.py:79-83, layers/feed_forward.py:50).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "kernels.h"
#include "tc_ptx.cuh"
#include "x3_split.cuh"

namespace scb {

constexpr int XT_THREADS = 448;                  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warps 10..13 A loaders
constexpr int XT_BN = 64;
constexpr int XT_WPLANE = XT_BN * TC_BK * 2;     // 8 KB: one fp16 plane of a 64-row x 64-k weight block
constexpr int XT_STAGE = 2 * XT_WPLANE;          // W hi | W lo
constexpr int XT_NST = 13;
constexpr size_t XT_SMEM = 1024 + (size_t)XT_NST * XT_STAGE + 512;
constexpr uint32_t XT_ACC_COL = 256;             // TMEM: A hi [0,128), A lo [128,256), accumulator stage s at 256 + 128 s

struct XtParams {
  const float* bias;
  const __half* A; size_t a_plane; int lda;      // A planes: hi rows at A, lo plane a_plane elements further
  float* C; int ldc;                             // fp32 output rows, or
  __half* C2; size_t c2_plane; int ldc2;         // split-plane output
  int M, N, relu, m_tiles, n_tiles;
};

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void xt_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void xt_ld32_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

__global__ void __launch_bounds__(XT_THREADS, 1) gemm_x3t_kernel(const __grid_constant__ CUtensorMap map_wh,
                                                                 const __grid_constant__ CUtensorMap map_wl, XtParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + XT_NST * XT_STAGE);
  uint64_t* empty_bar = full_bar + 16;
  uint64_t* a_full = empty_bar + 16;
  uint64_t* a_empty = a_full + 1;
  uint64_t* tmem_full = a_empty + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long T = (long)p.m_tiles * p.n_tiles;
  const int t0 = (int)((long)blockIdx.x * T / gridDim.x), t1 = (int)((long)(blockIdx.x + 1) * T / gridDim.x);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
    for (int i = 0; i < XT_NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_full, 4); mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
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

  pdl_sync();                                // the producing kernel's writes (A planes) are visible

  if (warp == 0) {
    // ===================== TMA producer: the W planes of every (tile, K block) =====================
    int it = 0;
    for (int t = t0; t < t1; ++t) {
      const int n = t % p.n_tiles;
      for (int kb = 0; kb < 4; ++kb, ++it) {
        const int s = it % XT_NST, ph = (it / XT_NST) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one_sync()) {
          unsigned char* st = smem + s * XT_STAGE;
          mbar_expect_tx(&full_bar[s], XT_STAGE);
          tma_load_2d(&map_wh, &full_bar[s], st, kb * TC_BK, n * XT_BN);
          tma_load_2d(&map_wl, &full_bar[s], st + XT_WPLANE, kb * TC_BK, n * XT_BN);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D fp32 (bit 4), A/B fp16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(XT_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t a_hi = tmem_base, a_lo = tmem_base + 128;
    int it = 0, a_gen = 0, prev_m = -1, chunk = 0;
    for (int t = t0; t < t1; ++t) {
      const int m = t / p.n_tiles;
      const bool last_of_m = (t + 1 == t1) || ((t + 1) / p.n_tiles != m);
      if (m != prev_m) {                     // the loaders have stored this row tile's A planes
        mbar_wait(a_full, a_gen & 1);
        tc_fence_after();
        ++a_gen; prev_m = m;
      }
      for (int c = 0; c < 2; ++c, ++chunk) {
        const int as = chunk & 1;
        mbar_wait(&tmem_empty[as], ((chunk >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + XT_ACC_COL + (uint32_t)(as * 128), d_corr = d_main + 64;
        for (int kb2 = 0; kb2 < 2; ++kb2, ++it) {
          const int kb = 2 * c + kb2;
          const int s = it % XT_NST, ph = (it / XT_NST) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one_sync()) {
            unsigned char* st = smem + s * XT_STAGE;
            const uint64_t wh = make_smem_desc(smem_u32(st)), wl = make_smem_desc(smem_u32(st + XT_WPLANE));
#pragma unroll
            for (int k = 0; k < TC_BK / UMMA_K; ++k) {
              const uint32_t ka = (uint32_t)((kb * 4 + k) * 8);            // 16 fp16 of K = 8 TMEM columns
              const uint32_t acc = (kb2 | k) != 0;
              xt_umma_ts(d_main, a_hi + ka, wh + 2 * k, idesc, acc);
              xt_umma_ts(d_corr, a_hi + ka, wl + 2 * k, idesc, acc);
              xt_umma_ts(d_corr, a_lo + ka, wh + 2 * k, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);
            if (kb2 == 1) umma_commit(&tmem_full[as]);
            if (last_of_m && c == 1 && kb2 == 1) umma_commit(a_empty);     // every UMMA that reads this A has been issued
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 10) {
    // ===================== A loaders (warps 10..13): thread <-> row of the TMEM lane quarter warp % 4 =====================
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int a_gen = 0, prev_m = -1;
    for (int t = t0; t < t1; ++t) {
      const int m = t / p.n_tiles;
      if (m == prev_m) continue;
      mbar_wait(a_empty, (a_gen & 1) ^ 1);   // the UMMAs of the previous row tile have read A
      tc_fence_after();
      const int row = m * TC_BM + q * 32 + lane;
#pragma unroll 1
      for (int pl = 0; pl < 2; ++pl) {
        const uint4* src = reinterpret_cast<const uint4*>(p.A + (size_t)pl * p.a_plane + (size_t)row * p.lda);
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {     // 64 K elements = 32 columns per store
          uint32_t v[32];
          if (row < p.M) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 u = src[cb * 8 + i];
              v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
          tmem_st32(lane_addr + (uint32_t)(pl * 128 + cb * 32), v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(a_full)) : "memory");
      ++a_gen; prev_m = m;
    }
  } else {
    // ===================== epilogue warps 2..9: thread <-> accumulator row, 32 of the 64 columns each =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + XT_ACC_COL + (uint32_t)(half * 32);
    int chunk = 0;
    for (int t = t0; t < t1; ++t) {
      const int m = t / p.n_tiles, n = t - m * p.n_tiles;
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      for (int c = 0; c < 2; ++c, ++chunk) {
        const int as = chunk & 1;
        mbar_wait(&tmem_full[as], (chunk >> 1) & 1);
        tc_fence_after();
        uint32_t vm[32], vc[32];
        xt_ld32_nowait(lane_base + (uint32_t)(as * 128), vm);
        xt_ld32_nowait(lane_base + (uint32_t)(as * 128 + 64), vc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[as])) : "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += fmaf(__uint_as_float(vc[j]), X3_INV_SCALE, __uint_as_float(vm[j]));
      }
      const int col0 = n * XT_BN + half * 32, row = m * TC_BM + q * 32 + lane;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
          acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
      if (row < p.M) {
        if (p.C2) {
          __half* dh = p.C2 + (size_t)row * p.ldc2 + col0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 uh, ul;
            x3_split8(acc + 8 * i, uh, ul);
            *reinterpret_cast<uint4*>(dh + 8 * i) = uh;
            *reinterpret_cast<uint4*>(dh + p.c2_plane + 8 * i) = ul;
          }
        } else {
          float* d = p.C + (size_t)row * p.ldc + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(d + 4 * i) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// K = 256, A as split planes (dense rows), N a multiple of 64, M known on the host, fp32 rows or split planes out, no residual
bool gemm_x3t_eligible(const GemmArgs& g, const X3Extra& x) {
  if (!x.A2 || x.lnX || g.a_row_off || g.a_seg_off || g.c_row_off || g.n_rows_dev || g.R) return false;
  if (g.M <= 0 || g.K != 256 || g.N % XT_BN != 0 || g.lda % 8 != 0) return false;
  if ((g.C != nullptr) == (x.C2 != nullptr)) return false;
  if (g.C && g.ldc % 4 != 0) return false;
  if (x.C2 && x.ldc2 % 8 != 0) return false;
  return true;
}

int launch_gemm_x3t(const GemmArgs& g, const X3Extra& x, const void* W2, cudaStream_t st) {
  if (!gemm_x3t_eligible(g, x) || !W2) { set_last_error("gemm_x3t: unsupported shape M=%d N=%d K=%d", g.M, g.N, g.K); return -1; }
  const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(W2);
  CUtensorMap mh, ml;
  if (tc_get_map(wh, g.N, g.K, g.K, XT_BN, &mh)) return -1;
  if (tc_get_map(wh + (size_t)g.N * g.K, g.N, g.K, g.K, XT_BN, &ml)) return -1;
  static PerDeviceMark attr_mk;
  if (!attr_mk.cur()) {
    if (cudaFuncSetAttribute(gemm_x3t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XT_SMEM) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm_x3t, smem=%zu) failed", XT_SMEM);
      return -1;
    }
    attr_mk.cur() = 1;
  }
  XtParams p{g.bias, reinterpret_cast<const __half*>(x.A2), x.a2_plane, g.lda, g.C, g.ldc, reinterpret_cast<__half*>(x.C2),
             x.c2_plane, x.ldc2, g.M, g.N, g.relu, cdiv(g.M, TC_BM), g.N / XT_BN};
  const long tiles = (long)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  launch_k(gemm_x3t_kernel, dim3(grid), dim3(XT_THREADS), XT_SMEM, st, mh, ml, p);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
