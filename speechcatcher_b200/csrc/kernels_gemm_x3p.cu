// Persistent split-precision tensor-core GEMM for sm_100a ("x3p"): the encoder-sized form of kernels_gemm_x3.cu.
//     C = act(A * W^T + bias) (+ C)        A, W as split fp16 planes (x3_split.cuh), fp32-class result
//
// Why a second kernel: the per-tile kernel (one 128 x BN tile per CTA) spends most of a CTA's life outside the tensor
// pipe -- launch, TMEM allocation, first TMA round trip, an epilogue that nothing overlaps, row-per-thread global
// stores -- and re-reads the A tile from L2 for every N tile.  Measured on the B200 (profiles/r2_gemm_x3_microbench.jsonl):
// encoder FFN1 (10752 x 2048 x 256) 97 us = 25 % of the tensor peak counted in executed UMMAs.  Here:
//   * persistent: grid = min(tiles, SMs); a CTA owns a contiguous range of (m, n) tiles, n fastest;
//   * A-resident (K = 256): the 128 x 256 A tile (hi + lo planes, 128 KB) is loaded once per m tile and stays in shared
//     memory while the W planes of successive n tiles stream through a TMA ring (the ring slots of A are released
//     K block by K block during the last n tile, so the next m tile's A streams in behind the MMAs);
//     long-K products (FFN2, K = 2048) stream A and W together through a three-stage ring;
//   * the accumulator lives in TMEM for one K chunk of 128 only (8 UMMA k-steps: main term in one 128-column
//     accumulator, the two correction terms in another), in two TMEM stages: eight epilogue warps drain chunk i into
//     fp32 registers with round-to-nearest adds while the tensor core works on chunk i + 1.  This is also what keeps
//     the truncating tensor-core accumulation (profiles/r2_tc_accum_probe.json) at the error level of an fp32 FMA
//     chain for any K -- the job of the J round-robin accumulators in the per-tile kernel;
//   * output through shared memory and TMA: fp32 rows (store, or fp32 add at the L2 = the in-place residual
//     connection), or split fp16 planes for a consuming Linear; rows beyond M are clipped by the tensor maps.
// Replaces torch.nn.functional.linear of the encoder layers
// (speechcatcher/model/attention/multi_head_attention.py:79-83,133, layers/feed_forward.py:50).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "kernels.h"
#include "tc_ptx.cuh"
#include "x3_split.cuh"

namespace scb {

constexpr int XP_THREADS = 320;                  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int XP_BN = 128;
constexpr int XP_PLANE = TC_BM * TC_BK * 2;      // one fp16 plane of a 128-row x 64-k operand block: 16 KB (A and W alike)
constexpr int XP_MAIN_BYTES = 192 * 1024;        // operand storage
constexpr int XP_OUT_BYTES = 8 * 4096;           // output staging: [32 rows][128 B] per epilogue warp
constexpr size_t XP_SMEM = 1024 + XP_MAIN_BYTES + XP_OUT_BYTES + 256;

struct XpParams {
  const float* bias;
  int M, N, K, relu, out_mode;                   // out_mode 0: fp32 store, 1: fp32 add into C, 2: split planes
  int m_tiles, n_tiles;
  const int* n_rows_dev;                         // optional device-side row count (decode step: active rows)
  const float* X; int ldx; const float* ln_w; const float* ln_b;   // LN form: A = LayerNorm(X rows), built in the kernel
  int dbg;                                       // SCB_XP_DBG (timing experiments only): 1 = no output, 2 = no TMEM drain
  // direct != 0 (SCB_XP_DIRECT=1, experiment): store / plane outputs go from registers straight to global memory (thread <->
  // row, 16-byte stores) instead of through the shared-memory staging + TMA store.  The kernel is bound by shared-memory
  // bandwidth -- every 128x128x16 UMMA with both operands in shared memory reads 8 KB, the W ring writes 2.7 KB per UMMA,
  // staging writes + reads another 2.7 KB: 13.4 KB / 128 B per cycle = 105 cycles per UMMA, measured 106
  // (profiles/r2_gemm_x3p_issuer_waits.txt) -- but the uncoalesced stores cost more than the staging saves (FFN1 40 -> 50 us)
  int direct; float* C; int ldc; __half* C2; size_t c2_plane; int ldc2;
};

__device__ __forceinline__ void xp_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void xp_ld32_nowait(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void xp_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <bool A_RES, bool LN>
__global__ void __launch_bounds__(XP_THREADS, 1) gemm_x3p_kernel(const __grid_constant__ CUtensorMap map_wh,
                                                                 const __grid_constant__ CUtensorMap map_wl,
                                                                 const __grid_constant__ CUtensorMap map_ah,
                                                                 const __grid_constant__ CUtensorMap map_al,
                                                                 const __grid_constant__ CUtensorMap map_c,
                                                                 const __grid_constant__ CUtensorMap map_ch,
                                                                 const __grid_constant__ CUtensorMap map_cl,
                                                                 XpParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // ring stages (SCB_XP_DBG & 8, timing experiment with & 1: the output staging becomes a third W stage of the A-resident form)
  const int NST = A_RES ? ((p.dbg & 9) == 9 ? 3 : 2) : 3;
  constexpr int STAGE_BYTES = A_RES ? 2 * XP_PLANE : 4 * XP_PLANE;     // W hi | W lo   or   A hi | A lo | W hi | W lo
  constexpr int A_KB_BYTES = 2 * XP_PLANE;                             // resident A: hi | lo per K block
  unsigned char* ring = smem + (A_RES ? 4 * A_KB_BYTES : 0);
  unsigned char* s_out = smem + XP_MAIN_BYTES;
  uint64_t* full_bar = (uint64_t*)(s_out + XP_OUT_BYTES);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* a_full = empty_bar + 4;
  uint64_t* a_empty = a_full + 4;
  uint64_t* tmem_full = a_empty + 4;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.K / TC_BK, n_chunks = nkb / 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
    for (int i = 0; i < 4; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&a_full[i], LN ? 8 : 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                           // two accumulator stages x (main 128 + correction 128) fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_sync();                                // the producing kernel's writes (A planes, C for the residual add) are visible
  int M = p.M;
  if (p.n_rows_dev) M = min(M, *p.n_rows_dev);
  const long T = (long)((M + TC_BM - 1) / TC_BM) * p.n_tiles;          // row tiles that hold active rows only
  const int t0 = (int)((long)blockIdx.x * T / gridDim.x), t1 = (int)((long)(blockIdx.x + 1) * T / gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    int it = 0, a_gen = 0, prev_m = -1;
    for (int t = t0; t < t1; ++t) {
      const int m = t / p.n_tiles, n = t - m * p.n_tiles;
      if (A_RES && !LN && m != prev_m) {     // new row tile: refill the resident A, K block by K block as they are released
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&a_empty[kb], (a_gen & 1) ^ 1);
          if (elect_one_sync()) {
            unsigned char* dst = smem + kb * A_KB_BYTES;
            mbar_expect_tx(&a_full[kb], 2 * XP_PLANE);
            tma_load_2d(&map_ah, &a_full[kb], dst, kb * TC_BK, m * TC_BM);
            tma_load_2d(&map_al, &a_full[kb], dst + XP_PLANE, kb * TC_BK, m * TC_BM);
          }
          __syncwarp();
        }
        ++a_gen; prev_m = m;
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % NST, ph = (it / NST) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one_sync()) {
          unsigned char* st = ring + s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          if (!A_RES) {
            tma_load_2d(&map_ah, &full_bar[s], st, kb * TC_BK, m * TC_BM);
            tma_load_2d(&map_al, &full_bar[s], st + XP_PLANE, kb * TC_BK, m * TC_BM);
            st += 2 * XP_PLANE;
          }
          tma_load_2d(&map_wh, &full_bar[s], st, kb * TC_BK, n * XP_BN);
          tma_load_2d(&map_wl, &full_bar[s], st + XP_PLANE, kb * TC_BK, n * XP_BN);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(XP_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int it = 0, a_gen = 0, prev_m = -1, chunk = 0;
    for (int t = t0; t < t1; ++t) {
      const int m = t / p.n_tiles;
      const bool first_of_m = m != prev_m;
      const bool last_of_m = (t + 1 == t1) || ((t + 1) / p.n_tiles != m);
      int a_ph = 0;
      if (first_of_m) { a_ph = a_gen & 1; ++a_gen; prev_m = m; }
      for (int c = 0; c < n_chunks; ++c, ++chunk) {
        const int as = chunk & 1;
        mbar_wait(&tmem_empty[as], ((chunk >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(as * 256), d_corr = d_main + 128;
        for (int kb2 = 0; kb2 < 2; ++kb2, ++it) {
          const int kb = 2 * c + kb2;
          const int s = it % NST, ph = (it / NST) & 1;
          mbar_wait(&full_bar[s], ph);
          if (A_RES && first_of_m) mbar_wait(&a_full[kb], a_ph);
          tc_fence_after();
          if (elect_one_sync()) {
            unsigned char* st = ring + s * STAGE_BYTES;
            unsigned char* sa = A_RES ? smem + kb * A_KB_BYTES : st;
            unsigned char* sw = A_RES ? st : st + 2 * XP_PLANE;
            const uint64_t ah = make_smem_desc(smem_u32(sa)), al = make_smem_desc(smem_u32(sa + XP_PLANE));
            const uint64_t wh = make_smem_desc(smem_u32(sw)), wl = make_smem_desc(smem_u32(sw + XP_PLANE));
#pragma unroll
            for (int k = 0; k < TC_BK / UMMA_K; ++k) {
              // 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in 16-byte units
              const uint32_t acc = (kb2 | k) != 0;
              xp_umma(d_main, ah + 2 * k, wh + 2 * k, idesc, acc);
              xp_umma(d_corr, ah + 2 * k, wl + 2 * k, idesc, acc);
              xp_umma(d_corr, al + 2 * k, wh + 2 * k, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);                        // frees the ring stage once the MMAs have read it
            if (A_RES && last_of_m) umma_commit(&a_empty[kb]);  // ... and this K block of the resident A
            if (kb2 == 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue warps 2..9: thread <-> accumulator row; the two warps of a TMEM lane quarter
    // split the 128 columns
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
    unsigned char* stg = s_out + (warp - 2) * 4096;             // [32 rows][128 B], 128-byte swizzle (1 KB aligned)
    unsigned char* my_row = stg + lane * 128;
    const int sw = lane & 7;
    int chunk = 0, a_gen = 0, prev_m = -1;
    for (int t = t0; t < t1; ++t) {
      const int m = t / p.n_tiles, n = t - m * p.n_tiles;
      if (LN && m != prev_m) {
        // ---- LayerNorm prologue: these eight warps build the resident A tile of a new row tile from the fp32 rows
        // (one warp per row; statistics exactly as layernorm_kernel, kernels_gemm.cu: element lane + 32 i per lane, the
        // row re-read in 16-byte chunk order for the normalise / split / swizzled store).  Rows beyond M are zero.
        mbar_wait(&a_empty[3], (a_gen & 1) ^ 1);                // the MMAs of the previous row tile have read all of A
        const int kb = lane >> 3, ch = lane & 7;                // columns 8 lane .. 8 lane + 7 = chunk ch of K block kb
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ln_w) + 2 * lane), w1 = __ldg(reinterpret_cast<const float4*>(p.ln_w) + 2 * lane + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 2 * lane), b1 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 2 * lane + 1);
        // 16 rows per warp in two batches of eight: all loads of a batch are in flight together and the eight rows'
        // shuffle reductions interleave (row by row the build is a chain of L2 and shuffle latencies, ~10 us per tile)
#pragma unroll 1
        for (int jb = 0; jb < 2; ++jb) {
          const int r_first = (warp - 2) + 64 * jb;             // rows r_first + 8 j, j = 0..7
          float v[8][8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = m * TC_BM + r_first + 8 * j;
            const float* xr = p.X + (size_t)row * p.ldx;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] = row < M ? xr[lane + 32 * i] : 0.f;
          }
          float mean[8], rstd[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += v[j][i];
            mean[j] = sum;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mean[j] += __shfl_xor_sync(0xffffffffu, mean[j], o);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            mean[j] = mean[j] / 256.0f;
            float qq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean[j]; qq += d * d; }
            rstd[j] = qq;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rstd[j] += __shfl_xor_sync(0xffffffffu, rstd[j], o);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            rstd[j] = 1.0f / sqrtf(rstd[j] / 256.0f + 1e-12f);
            const int r = r_first + 8 * j, row = m * TC_BM + r;
            uint4 uh = make_uint4(0, 0, 0, 0), ul = uh;
            if (row < M) {
              const float* xr = p.X + (size_t)row * p.ldx;
              const float4 x0 = *reinterpret_cast<const float4*>(xr + 8 * lane), x1 = *reinterpret_cast<const float4*>(xr + 8 * lane + 4);
              const float mu = mean[j], rs = rstd[j];
              const float o8[8] = {(x0.x - mu) * rs * w0.x + b0.x, (x0.y - mu) * rs * w0.y + b0.y,
                                   (x0.z - mu) * rs * w0.z + b0.z, (x0.w - mu) * rs * w0.w + b0.w,
                                   (x1.x - mu) * rs * w1.x + b1.x, (x1.y - mu) * rs * w1.y + b1.y,
                                   (x1.z - mu) * rs * w1.z + b1.z, (x1.w - mu) * rs * w1.w + b1.w};
              x3_split8(o8, uh, ul);
            }
            unsigned char* dst = smem + kb * A_KB_BYTES + r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = uh;
            *reinterpret_cast<uint4*>(dst + XP_PLANE) = ul;
          }
        }
        fence_proxy_async_smem();                // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[k4])) : "memory");
        }
        ++a_gen; prev_m = m;
      }
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      for (int c = 0; c < n_chunks; ++c, ++chunk) {
        const int as = chunk & 1;
        mbar_wait(&tmem_full[as], (chunk >> 1) & 1);
        tc_fence_after();
        const uint32_t tb = lane_base + (uint32_t)(as * 256);
        if (!(p.dbg & 2)) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t vm[32], vc[32];
          xp_ld32_nowait(tb + (uint32_t)(g * 32), vm);
          xp_ld32_nowait(tb + 128u + (uint32_t)(g * 32), vc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            acc[g * 32 + j] += fmaf(__uint_as_float(vc[j]), X3_INV_SCALE, __uint_as_float(vm[j]));
        }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[as])) : "memory");
      }
      // ---- this warp's 32 rows x 64 columns of the tile
      const int col0 = n * XP_BN + half * 64, row0 = m * TC_BM + q * 32;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
          acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
      if (p.direct && p.out_mode != 1) {
        const int row = row0 + lane;
        if (row < M && !(p.dbg & 1)) {
          if (p.out_mode == 2) {
            __half* dh = p.C2 + (size_t)row * p.ldc2 + col0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 uh, ul;
              x3_split8(acc + 8 * i, uh, ul);
              *reinterpret_cast<uint4*>(dh + 8 * i) = uh;
              *reinterpret_cast<uint4*>(dh + p.c2_plane + 8 * i) = ul;
            }
          } else {
            float* d = p.C + (size_t)row * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              *reinterpret_cast<float4*>(d + 4 * i) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
          }
        }
      } else if (row0 < M && !(p.dbg & 1)) {
        if (p.out_mode == 2) {
          uint4 ul[8];
          if (lane == 0) xp_store_wait_read();                 // the previous tile's copy has read the staging rows
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 uh;
            x3_split8(acc + 8 * i, uh, ul[i]);
            *reinterpret_cast<uint4*>(my_row + ((i ^ sw) << 4)) = uh;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&map_ch, stg, col0, row0); tma_store_commit(); xp_store_wait_read(); }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(my_row + ((i ^ sw) << 4)) = ul[i];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&map_cl, stg, col0, row0); tma_store_commit(); }
        } else {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (lane == 0) xp_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(my_row + ((i ^ sw) << 4)) =
                  make_float4(acc[32 * r + 4 * i], acc[32 * r + 4 * i + 1], acc[32 * r + 4 * i + 2], acc[32 * r + 4 * i + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.out_mode == 1) tma_reduce_add_2d(&map_c, stg, col0 + 32 * r, row0);
              else tma_store_2d(&map_c, stg, col0 + 32 * r, row0);
              tma_store_commit();
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();     // the staging rows must outlive the copies; completes before the CTA exits
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <bool A_RES, bool LN>
static int xp_launch(const CUtensorMap* maps, const XpParams& p, cudaStream_t st) {
  static PerDeviceMark attr_mk;
  if (!attr_mk.cur()) {
    if (cudaFuncSetAttribute(gemm_x3p_kernel<A_RES, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XP_SMEM) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm_x3p, smem=%zu) failed", XP_SMEM);
      return -1;
    }
    attr_mk.cur() = 1;
  }
  // SCB_XP_MAX_CTAS: leave some SMs to the concurrently running search chain (experiments)
  static const int max_ctas = [] { const char* v = getenv("SCB_XP_MAX_CTAS"); const int n = v ? atoi(v) : 0; return n > 0 && n < kNumSMs ? n : kNumSMs; }();
  const long tiles = (long)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  launch_k(gemm_x3p_kernel<A_RES, LN>, dim3(grid), dim3(XP_THREADS), XP_SMEM, st, maps[0], maps[1], maps[2], maps[3], maps[4],
           maps[5], maps[6], p);
  SCB_LAUNCH_CHECK();
  return 0;
}

// Shapes the persistent kernel takes: A as split planes (dense rows) or as LayerNorm(X) of fp32 rows (K = 256), dense
// output rows, N a multiple of 128, K = 256 (A-resident) or a multiple of 128 from 384 up (streamed), exactly one
// output form, and a residual only as the in-place add (R == C).  With a device-side row count (n_rows_dev) only the row
// tiles holding active rows are computed; rows of the last such tile beyond the count are written too (stale or zero
// operands -> values nobody reads), so the output buffers must hold M rows.
bool gemm_x3p_eligible(const GemmArgs& g, const X3Extra& x) {
  if ((!x.A2 && !x.lnX) || g.a_row_off || g.a_seg_off || g.c_row_off) return false;
  if (g.M <= 0 || g.N % XP_BN != 0 || g.K % 128 != 0 || g.K < 256 || (x.A2 && g.lda % 8 != 0)) return false;
  if (x.lnX && (g.K != 256 || x.ldx % 4 != 0 || !x.ln_w || !x.ln_b)) return false;
  if ((g.C != nullptr) == (x.C2 != nullptr)) return false;
  if (g.C && g.ldc % 4 != 0) return false;
  if (x.C2 && x.ldc2 % 8 != 0) return false;
  if (g.R && (g.R != g.C || g.ldr != g.ldc)) return false;
  return true;
}

int launch_gemm_x3p(const GemmArgs& g, const X3Extra& x, const void* W2, cudaStream_t st) {
  if (!gemm_x3p_eligible(g, x) || !W2) { set_last_error("gemm_x3p: unsupported shape M=%d N=%d K=%d", g.M, g.N, g.K); return -1; }
  const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(W2);        // 16-bit elements: the maps only move bytes
  const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(x.A2);
  CUtensorMap maps[7];
  if (tc_get_map(wh, g.N, g.K, g.K, XP_BN, &maps[0])) return -1;
  if (tc_get_map(wh + (size_t)g.N * g.K, g.N, g.K, g.K, XP_BN, &maps[1])) return -1;
  if (x.lnX) { maps[2] = maps[0]; maps[3] = maps[1]; }                         // A is built in the kernel
  else {
    if (tc_get_map(ah, g.M, g.K, g.lda, TC_BM, &maps[2])) return -1;           // rows = M: the tail tile is zero-filled
    if (tc_get_map(ah + x.a2_plane, g.M, g.K, g.lda, TC_BM, &maps[3])) return -1;
  }
  if (g.C) {
    if (tc_get_map_f32(g.C, g.M, g.N, g.ldc, 32, &maps[4])) return -1;
    maps[5] = maps[4]; maps[6] = maps[4];
  } else {
    const __nv_bfloat16* ch = reinterpret_cast<const __nv_bfloat16*>(x.C2);
    if (tc_get_map(ch, g.M, g.N, x.ldc2, 32, &maps[5])) return -1;
    if (tc_get_map(ch + x.c2_plane, g.M, g.N, x.ldc2, 32, &maps[6])) return -1;
    maps[4] = maps[5];
  }
  static const int dbg = [] { const char* v = getenv("SCB_XP_DBG"); return v ? atoi(v) : 0; }();
  static const int direct = [] { const char* v = getenv("SCB_XP_DIRECT"); return v ? atoi(v) : 0; }();
  XpParams p{g.bias, g.M, g.N, g.K, g.relu, g.C ? (g.R ? 1 : 0) : 2, cdiv(g.M, TC_BM), g.N / XP_BN, g.n_rows_dev,
             x.lnX, x.ldx, x.ln_w, x.ln_b, dbg, direct, g.C, g.ldc, reinterpret_cast<__half*>(x.C2), x.c2_plane, x.ldc2};
  if (x.lnX) return xp_launch<true, true>(maps, p, st);
  return g.K == 256 ? xp_launch<true, false>(maps, p, st) : xp_launch<false, false>(maps, p, st);
}

}  // namespace scb
