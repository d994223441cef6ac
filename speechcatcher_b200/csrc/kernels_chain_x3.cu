// Row-local chains of the decode step as ONE persistent kernel each (precise tensor-core mode, precision 2).
//
// Between two attention kernels of a decoder layer every operation is local to a row (= hypothesis):
//     self-attention -> [ O-proj (+x) -> LayerNorm2 -> cross-Q ] -> cross-attention
//                    -> [ O-proj (+x) -> LayerNorm3 -> FFN1 -> ReLU -> FFN2 (+x) -> LayerNorm1' -> QKV' ] -> self-attention'
// As separate kernels these are 3 + 6 dependent launches per layer, each of which pays its own launch / drain gap, its
// own TMEM allocation and its first TMA round trip on ~20 row tiles of work -- the decode step was bound by that chain
// latency, not by SM capacity.  Here a chain is a list of STAGES; a stage is a set of ITEMS
//   * GEMM item  (row tile rt of 128 rows, column tile ct of 128, K split ks): the persistent split-fp16 pipeline of
//     kernels_gemm_x3p.cu (TMA ring -> tcgen05.mma into two TMEM accumulator stages per 128-K chunk -> eight epilogue warps
//     accumulate the chunks in registers with round-to-nearest adds -> TMA store / fp32 add at the L2 / split planes);
//   * LayerNorm item (row tile rt): the eight epilogue warps normalise 128 rows (optionally first adding the split-K
//     partial sums of the preceding FFN2 and its bias into the residual stream, in a fixed order) and write split planes;
// all items of all stages are numbered in stage order and dealt round-robin to the persistent CTAs.  An item waits for
// the items that produce ITS row tile only, through global completion counters (release: results visible -> threadfence
// -> atomicAdd; acquire: poll -> fence.proxy.async -> TMA loads), so the stages pipeline across row tiles and a whole
// chain costs one launch.  Every dependency points to a lower item number and every CTA works in increasing order, so
// the schedule cannot deadlock as long as each CTA eventually runs.
//
// Replaces, per decoder layer: speechcatcher/model/decoder/decoder_layer.py:80-132 minus the two attention products
// (multi_head_attention.py:79-83,133; feed_forward.py:50; normalization.py:23).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "kernels.h"
#include "tc_ptx.cuh"
#include "x3_split.cuh"

namespace scb {

constexpr int CH_THREADS = 320;                  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue / LayerNorm
constexpr int CH_BN = 128;
constexpr int CH_PLANE = TC_BM * TC_BK * 2;      // 16 KB: one fp16 plane of a 128-row x 64-k operand block
constexpr int CH_STAGE_BYTES = 4 * CH_PLANE;     // A hi | A lo | W hi | W lo
constexpr int CH_NST = 3;
constexpr int CH_OUT_BYTES = 8 * 4096;
constexpr size_t CH_SMEM = 1024 + CH_NST * CH_STAGE_BYTES + CH_OUT_BYTES + 256;

__device__ __forceinline__ void ch_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void ch_ld32_nowait(uint32_t taddr, uint32_t* v) {
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

// Polls a completion counter until it reaches `target` (acquire), then orders the following async-proxy (TMA) reads
// after the producers' writes.  A wait of seconds can only be a protocol bug: trap instead of hanging the GPU.
__device__ __forceinline__ void ch_wait_counter(const int* ctr, int target) {
  int v;
  uint32_t spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    __nanosleep(64);
    if ((++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void ch_signal(int* ctr) {
  __threadfence();
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
}

struct ItemPos { int s, rt, ct, ks; };
// item number g -> (stage, row tile, column tile, K split); items of a stage are row-tile major
__device__ __forceinline__ bool ch_decode(const ChainParams& p, int RT, int g, ItemPos& it) {
  for (int s = 0; s < p.n_stages; ++s) {
    const int per_rt = p.st[s].type == 0 ? p.st[s].n_ct * p.st[s].n_ks : 1;
    const int cnt = RT * per_rt;
    if (g < cnt) {
      it.s = s; it.rt = g / per_rt;
      const int rem = g - it.rt * per_rt;
      it.ct = p.st[s].type == 0 ? rem / p.st[s].n_ks : 0;
      it.ks = p.st[s].type == 0 ? rem - it.ct * p.st[s].n_ks : 0;
      return true;
    }
    g -= cnt;
  }
  return false;
}

__global__ void __launch_bounds__(CH_THREADS, 1) chain_x3_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* s_out = smem + CH_NST * CH_STAGE_BYTES;
  uint64_t* full_bar = (uint64_t*)(s_out + CH_OUT_BYTES);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tmem_full = empty_bar + 4;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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

  pdl_sync();                                // the preceding kernel's writes (attention output planes, x, counters) are visible
  int M = p.M;
  if (p.n_rows_dev) M = min(M, *p.n_rows_dev);
  const int RT = (M + TC_BM - 1) / TC_BM;    // row tiles that hold active rows

  if (warp == 0) {
    // ===================== TMA producer =====================
    int itn = 0;
    ItemPos ip;
    for (int g = blockIdx.x; ch_decode(p, RT, g, ip); g += gridDim.x) {
      const ChainStage& S = p.st[ip.s];
      if (S.type != 0) continue;
      const CUtensorMap* ma = p.maps + S.map_a;
      const CUtensorMap* mw = p.maps + S.map_w;
      bool dep_ok = S.wait_base < 0;
      for (int kb = 0; kb < S.kb_item; ++kb, ++itn) {
        const int s = itn % CH_NST, ph = (itn / CH_NST) & 1;
        const int kc = (ip.ks * S.kb_item + kb) * TC_BK;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st = smem + s * CH_STAGE_BYTES;
        if (elect_one_sync()) {              // the weights do not depend on anything: in flight before the dependency is met
          mbar_expect_tx(&full_bar[s], CH_STAGE_BYTES);
          tma_load_2d(mw, &full_bar[s], st + 2 * CH_PLANE, kc, ip.ct * CH_BN);
          tma_load_2d(mw + 1, &full_bar[s], st + 3 * CH_PLANE, kc, ip.ct * CH_BN);
        }
        __syncwarp();
        if (!dep_ok) {
          ch_wait_counter(p.ctr + S.wait_base + ip.rt * S.wait_stride + (S.wait_ks ? ip.ks : 0), S.wait_target);
          dep_ok = true;
          __syncwarp();
        }
        if (elect_one_sync()) {
          tma_load_2d(ma, &full_bar[s], st, kc, ip.rt * TC_BM);
          tma_load_2d(ma + 1, &full_bar[s], st + CH_PLANE, kc, ip.rt * TC_BM);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(CH_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int itn = 0, chunk = 0;
    ItemPos ip;
    for (int g = blockIdx.x; ch_decode(p, RT, g, ip); g += gridDim.x) {
      const ChainStage& S = p.st[ip.s];
      if (S.type != 0) continue;
      const int n_chunks = S.kb_item >> 1;
      for (int c = 0; c < n_chunks; ++c, ++chunk) {
        const int as = chunk & 1;
        mbar_wait(&tmem_empty[as], ((chunk >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(as * 256), d_corr = d_main + 128;
        for (int kb2 = 0; kb2 < 2; ++kb2, ++itn) {
          const int s = itn % CH_NST, ph = (itn / CH_NST) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one_sync()) {
            unsigned char* st = smem + s * CH_STAGE_BYTES;
            const uint64_t ah = make_smem_desc(smem_u32(st)), al = make_smem_desc(smem_u32(st + CH_PLANE));
            const uint64_t wh = make_smem_desc(smem_u32(st + 2 * CH_PLANE)), wl = make_smem_desc(smem_u32(st + 3 * CH_PLANE));
#pragma unroll
            for (int k = 0; k < TC_BK / UMMA_K; ++k) {
              const uint32_t acc = (kb2 | k) != 0;
              ch_umma(d_main, ah + 2 * k, wh + 2 * k, idesc, acc);
              ch_umma(d_corr, ah + 2 * k, wl + 2 * k, idesc, acc);
              ch_umma(d_corr, al + 2 * k, wh + 2 * k, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);
            if (kb2 == 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue / LayerNorm warps 2..9 =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
    unsigned char* stg = s_out + (warp - 2) * 4096;             // [32 rows][128 B], 128-byte swizzle
    unsigned char* my_row = stg + lane * 128;
    const int sw = lane & 7;
    int chunk = 0;
    ItemPos ip;
    for (int g = blockIdx.x; ch_decode(p, RT, g, ip); g += gridDim.x) {
      const ChainStage& S = p.st[ip.s];
      int* sig = S.sig_base >= 0 ? p.ctr + S.sig_base + ip.rt * S.sig_stride + ip.ct / S.sig_div : nullptr;
      if (S.type == 1) {
        // ---------------- LayerNorm item: one warp per row, 16 rows per warp in two batches of eight
        if (S.wait_base >= 0) ch_wait_counter(p.ctr + S.wait_base + ip.rt * S.wait_stride, S.wait_target);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(S.ln_w) + 2 * lane), w1 = __ldg(reinterpret_cast<const float4*>(S.ln_w) + 2 * lane + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(S.ln_b) + 2 * lane), b1 = __ldg(reinterpret_cast<const float4*>(S.ln_b) + 2 * lane + 1);
#pragma unroll 1
        for (int jb = 0; jb < 2; ++jb) {
          const int r_first = ip.rt * TC_BM + (warp - 2) + 64 * jb;            // rows r_first + 8 j
          float v[8][8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = r_first + 8 * j;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
            if (row < M) {
              const float* xr = S.x + (size_t)row * 256 + 8 * lane;
              float4 a0 = __ldcg(reinterpret_cast<const float4*>(xr)), a1 = __ldcg(reinterpret_cast<const float4*>(xr + 4));
              if (S.n_part > 0) {
                // x += bias + sum of the split-K partial products, partials first and in K order
                const float* pr = S.part + (size_t)row * 256 + 8 * lane;
                float4 t0 = __ldcg(reinterpret_cast<const float4*>(pr)), t1 = __ldcg(reinterpret_cast<const float4*>(pr + 4));
                for (int k = 1; k < S.n_part; ++k) {
                  const float* pk = pr + (size_t)k * p.part_stride_rows * 256;
                  const float4 u0 = __ldcg(reinterpret_cast<const float4*>(pk)), u1 = __ldcg(reinterpret_cast<const float4*>(pk + 4));
                  t0.x += u0.x; t0.y += u0.y; t0.z += u0.z; t0.w += u0.w; t1.x += u1.x; t1.y += u1.y; t1.z += u1.z; t1.w += u1.w;
                }
                const float4 c0 = __ldg(reinterpret_cast<const float4*>(S.pbias) + 2 * lane), c1 = __ldg(reinterpret_cast<const float4*>(S.pbias) + 2 * lane + 1);
                t0.x += c0.x; t0.y += c0.y; t0.z += c0.z; t0.w += c0.w; t1.x += c1.x; t1.y += c1.y; t1.z += c1.z; t1.w += c1.w;
                a0.x += t0.x; a0.y += t0.y; a0.z += t0.z; a0.w += t0.w; a1.x += t1.x; a1.y += t1.y; a1.z += t1.z; a1.w += t1.w;
                float* xw = S.x + (size_t)row * 256 + 8 * lane;
                *reinterpret_cast<float4*>(xw) = a0;
                *reinterpret_cast<float4*>(xw + 4) = a1;
              }
              v[j][0] = a0.x; v[j][1] = a0.y; v[j][2] = a0.z; v[j][3] = a0.w; v[j][4] = a1.x; v[j][5] = a1.y; v[j][6] = a1.z; v[j][7] = a1.w;
            }
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
            const int row = r_first + 8 * j;
            if (row < M) {
              const float rs = 1.0f / sqrtf(rstd[j] / 256.0f + 1e-12f), mu = mean[j];
              const float o8[8] = {(v[j][0] - mu) * rs * w0.x + b0.x, (v[j][1] - mu) * rs * w0.y + b0.y,
                                   (v[j][2] - mu) * rs * w0.z + b0.z, (v[j][3] - mu) * rs * w0.w + b0.w,
                                   (v[j][4] - mu) * rs * w1.x + b1.x, (v[j][5] - mu) * rs * w1.y + b1.y,
                                   (v[j][6] - mu) * rs * w1.z + b1.z, (v[j][7] - mu) * rs * w1.w + b1.w};
              uint4 uh, ul;
              x3_split8(o8, uh, ul);
              __half* dst = S.out_hi + (size_t)row * 256 + 8 * lane;
              *reinterpret_cast<uint4*>(dst) = uh;
              *reinterpret_cast<uint4*>(dst + S.out_plane) = ul;
            }
          }
        }
        __syncwarp();
        if (sig && lane == 0) ch_signal(sig);
        continue;
      }
      // ---------------- GEMM item: accumulate the K chunks of the tile in registers
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      const int n_chunks = S.kb_item >> 1;
      for (int c = 0; c < n_chunks; ++c, ++chunk) {
        const int as = chunk & 1;
        mbar_wait(&tmem_full[as], (chunk >> 1) & 1);
        tc_fence_after();
        const uint32_t tb = lane_base + (uint32_t)(as * 256);
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          uint32_t vm[32], vc[32];
          ch_ld32_nowait(tb + (uint32_t)(gq * 32), vm);
          ch_ld32_nowait(tb + 128u + (uint32_t)(gq * 32), vc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            acc[gq * 32 + j] += fmaf(__uint_as_float(vc[j]), X3_INV_SCALE, __uint_as_float(vm[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[as])) : "memory");
      }
      const int col0 = ip.ct * CH_BN + half * 64, row0 = ip.rt * TC_BM + q * 32;
      if (S.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(S.bias + col0) + i);
          acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
        }
      }
      if (S.relu) {
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
      if (row0 < M) {
        const CUtensorMap* mo = p.maps + S.map_o;
        if (S.out_mode == 2) {
          uint4 ul[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 uh;
            x3_split8(acc + 8 * i, uh, ul[i]);
            *reinterpret_cast<uint4*>(my_row + ((i ^ sw) << 4)) = uh;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(mo, stg, col0, row0); tma_store_commit(); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(my_row + ((i ^ sw) << 4)) = ul[i];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(mo + 1, stg, col0, row0); tma_store_commit(); }
        } else {
          const int orow = S.out_mode == 3 ? ip.ks * p.part_stride_rows + row0 : row0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (r == 1) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(my_row + ((i ^ sw) << 4)) =
                  make_float4(acc[32 * r + 4 * i], acc[32 * r + 4 * i + 1], acc[32 * r + 4 * i + 2], acc[32 * r + 4 * i + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (S.out_mode == 1) tma_reduce_add_2d(mo, stg, col0 + 32 * r, row0);
              else tma_store_2d(mo, stg, col0 + 32 * r, orow);
              tma_store_commit();
            }
          }
        }
        // results complete in global memory (which also frees the staging rows), then visible, then counted
        if (lane == 0) tma_store_wait_all();
      }
      __syncwarp();
      if (sig && lane == 0) ch_signal(sig);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

int launch_chain_x3(const ChainParams& p, int max_items, cudaStream_t st) {
  static PerDeviceMark attr_mk;
  if (!attr_mk.cur()) {
    if (cudaFuncSetAttribute(chain_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM) != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(chain_x3, smem=%zu) failed", CH_SMEM);
      return -1;
    }
    attr_mk.cur() = 1;
  }
  const int grid = max_items < kNumSMs ? (max_items < 1 ? 1 : max_items) : kNumSMs;
  launch_k(chain_x3_kernel, dim3(grid), dim3(CH_THREADS), CH_SMEM, st, p);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
