// LayerNorm and the fp32 SIMT GEMM used by the fp32 (parity) mode.
//
// Every Linear of the path is C = A * W^T (+bias)(+ReLU)(+residual) with W in the torch layout
// [N][K].  The fp32 mode must be true-fp32 FMA with a fixed reduction order (SURVEY.md section 7
// "bit-exact n-best"), so this kernel runs on the CUDA cores; the bf16 mode uses the tcgen05 kernel
// in kernels_gemm_tc.cu instead.
#include <stdarg.h>
#include <mutex>
#include <type_traits>
#include "kernels.h"
#include "x3_split.cuh"

namespace scb {

static char g_err[512] = "";
static std::mutex g_err_mu;
void set_last_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }
bool g_use_pdl = false;

// ------------------------------------------------------------------ LayerNorm
// One warp per row; the row stays in registers between the mean and variance passes.
// OutT = float / __nv_bfloat16: plain rows.  OutT = __half: split fp16 planes (hi at y, lo at y + plane; x3_split.cuh),
// the operand format of the precise tensor-core GEMM.
template <typename OutT, int MAXV>
__global__ void layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                 const float* __restrict__ b, OutT* __restrict__ y, int ldy, int rows,
                                 int D, const int* __restrict__ n_rows_dev, size_t plane) {
  pdl_sync();
  if (n_rows_dev) rows = min(rows, *n_rows_dev);
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)row * ldx;
  float v[MAXV];
  int nv = D >> 5;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) { v[i] = xr[lane + 32 * i]; s += v[i]; }
  float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) { float d = v[i] - mean; q += d * d; }
  float var = warp_sum(q) / (float)D;
  float rstd = 1.0f / sqrtf(var + 1e-12f);
  OutT* yr = y + (size_t)row * ldy;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      int c = lane + 32 * i;
      float o = (v[i] - mean) * rstd * w[c] + b[c];
      if constexpr (sizeof(OutT) == 2 && !std::is_same<OutT, __nv_bfloat16>::value) {
        __half h, l;
        x3_split(o, h, l);
        yr[c] = h;
        yr[plane + c] = l;
      } else {
        yr[c] = (OutT)o;
      }
    }
}

template <typename OutT>
static int ln_launch(const float* x, int ldx, const float* w, const float* b, OutT* y, int ldy, int rows,
                     int D, const int* n_rows_dev, cudaStream_t st, size_t plane = 0) {
  if (rows <= 0) return 0;
  if (D % 32 != 0 || D > 512) { set_last_error("layernorm: unsupported D=%d", D); return -1; }
  const int warps = 8;
  launch_k(layernorm_kernel<OutT, 16>, dim3(cdiv(rows, warps)), dim3(warps * 32), 0, st, x, ldx, w, b, y, ldy, rows, D, n_rows_dev, plane);
  SCB_LAUNCH_CHECK();
  return 0;
}
int launch_layernorm(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows,
                     int D, const int* n_rows_dev, cudaStream_t st) {
  return ln_launch<float>(x, ldx, w, b, y, ldy, rows, D, n_rows_dev, st);
}
int launch_layernorm_bf16(const float* x, int ldx, const float* w, const float* b, __nv_bfloat16* y,
                          int ldy, int rows, int D, const int* n_rows_dev, cudaStream_t st) {
  return ln_launch<__nv_bfloat16>(x, ldx, w, b, y, ldy, rows, D, n_rows_dev, st);
}

int launch_layernorm_split(const float* x, int ldx, const float* w, const float* b, void* y2, size_t plane, int ldy,
                           int rows, int D, const int* n_rows_dev, cudaStream_t st) {
  return ln_launch<__half>(x, ldx, w, b, (__half*)y2, ldy, rows, D, n_rows_dev, st, plane);
}

// ------------------------------------------------------------------ fp32 GEMM
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread, register double buffering of the next
// K-slab.  K must be a multiple of 16; rows of A and W must be 16-byte aligned.
constexpr int GBM = 64, GBN = 64, GBK = 16;

struct GemmParams {
  const float* A; int lda; const int64_t* a_row_off; const int* a_seg_off; int seg_len;
  const float* W; const float* bias; const float* R; int ldr; float* C; int ldc;
  const int64_t* c_row_off; int M, N, K, relu; const int* n_rows_dev; __nv_bfloat16* Cb; int ldcb;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmParams p) {
  pdl_sync();
  int M = p.M;
  if (p.n_rows_dev) M = min(M, *p.n_rows_dev);
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  if (m0 >= M) return;
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int lrow = tid >> 2;         // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) * 4;      // 0,4,8,12 : k offset of its float4
  // A source row
  const int am = m0 + lrow;
  const bool a_ok = am < M;
  const float* a_ptr = nullptr;
  if (a_ok) a_ptr = p.a_row_off ? p.A + p.a_row_off[am] : p.A + (size_t)am * p.lda;
  const int wn = n0 + lrow;
  const bool w_ok = wn < p.N;
  const float* w_ptr = w_ok ? p.W + (size_t)wn * p.K : nullptr;

  const int tx = tid & 15, ty = tid >> 4;   // thread -> 4 cols (tx*4..), 4 rows (ty*4..)
  // Two-level accumulation: FMA chains of 128 products (`acc`) are added into `tot` with round-to-nearest adds, so the
  // rounding error of a K = 2048 dot product is that of 16 partial sums instead of one 2048-term chain (closer to the
  // blocked summation of the reference's CPU GEMM; the decisions of the beam search hinge on ~1e-6 differences).
  float acc[4][4], tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }

  auto load_a = [&](int k0) -> float4 {
    if (!a_ok) return make_float4(0.f, 0.f, 0.f, 0.f);
    int k = k0 + lk;
    if (p.a_seg_off) {
      int seg = k / p.seg_len;
      return *reinterpret_cast<const float4*>(a_ptr + p.a_seg_off[seg] + (k - seg * p.seg_len));
    }
    return *reinterpret_cast<const float4*>(a_ptr + k);
  };
  auto load_w = [&](int k0) -> float4 {
    if (!w_ok) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(w_ptr + k0 + lk);
  };

  float4 ra = load_a(0), rw = load_w(0);
  for (int k0 = 0; k0 < p.K; k0 += GBK) {
    As[lk + 0][lrow] = ra.x; As[lk + 1][lrow] = ra.y; As[lk + 2][lrow] = ra.z; As[lk + 3][lrow] = ra.w;
    Bs[lk + 0][lrow] = rw.x; Bs[lk + 1][lrow] = rw.y; Bs[lk + 2][lrow] = rw.z; Bs[lk + 3][lrow] = rw.w;
    __syncthreads();
    if (k0 + GBK < p.K) { ra = load_a(k0 + GBK); rw = load_w(k0 + GBK); }
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w};
      float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (((k0 / GBK) & 7) == 7) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += tot[i][j];
  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float* crow = p.c_row_off ? p.C + p.c_row_off[m] : p.C + (size_t)m * p.ldc;
    const float* rrow = p.R ? p.R + (size_t)m * p.ldr : nullptr;
    int n = n0 + tx * 4;
    if (n + 3 < p.N) {
      float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (p.bias) {
        float4 bb = *reinterpret_cast<const float4*>(p.bias + n);
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (rrow) {
        float4 rr = *reinterpret_cast<const float4*>(rrow + n);
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      *reinterpret_cast<float4*>(crow + n) = o;
      if (p.Cb) {
        __nv_bfloat16* cb = p.Cb + (size_t)m * p.ldcb + n;
        cb[0] = __float2bfloat16(o.x); cb[1] = __float2bfloat16(o.y); cb[2] = __float2bfloat16(o.z); cb[3] = __float2bfloat16(o.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j >= p.N) break;
        float o = acc[i][j];
        if (p.bias) o += p.bias[n + j];
        if (p.relu) o = fmaxf(o, 0.f);
        if (rrow) o += rrow[n + j];
        crow[n + j] = o;
      }
    }
  }
}

int launch_gemm_f32(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (g.K % GBK != 0 || (g.a_seg_off && g.seg_len % GBK != 0)) {
    set_last_error("gemm_f32: K=%d (seg %d) must be a multiple of %d", g.K, g.seg_len, GBK);
    return -1;
  }
  if (g.N % 4 != 0) { set_last_error("gemm_f32: N=%d must be a multiple of 4", g.N); return -1; }
  GemmParams p{g.A, g.lda, g.a_row_off, g.a_seg_off, g.seg_len, g.W, g.bias, g.R, g.ldr, g.C, g.ldc,
               g.c_row_off, g.M, g.N, g.K, g.relu, g.n_rows_dev, g.Cb, g.ldcb};
  dim3 grid(cdiv(g.N, GBN), cdiv(g.M, GBM));
  launch_k(gemm_f32_kernel, grid, dim3(256), 0, st, p);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
