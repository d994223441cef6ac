// Frontend: per-slab framing (reflect-padded, centred STFT), Hann window, 512-point FFT,
// power spectrum, 257x80 slaney mel matrix, clamp/log, global mean-variance normalisation in
// fp64, and the 2-frame trim at chunk boundaries -- one fused kernel, one warp per STFT frame.
//
// Replaces: speechcatcher/speech2text_streaming.py:278-400 (apply_frontend, normalize_features)
//           speechcatcher/model/frontend/stft_frontend.py:87-154 (STFTFrontend.forward)
#include <math.h>
#include "kernels.h"

namespace scb {

constexpr int NFFT = 512, HOPS = 160, WINL = 400, NBIN = 257, NMEL = 80;

// The constant tables live in the engine's caller-owned workspace (one copy per engine, hence per device): the library
// allocates no device memory and keeps no device-side global state.
//   window[512]  hann(400) zero-padded to 512 (offset 56)        twiddle[256]  exp(-2 pi i k / 512)
//   melrange[80] [k_lo, k_hi) of the non-zero taps of every (triangular) mel filter        melfb[257][80]
int frontend_upload_tables(FrontendTables* dev, const float* window400, const float* mel_fb) {
  static_assert(sizeof(FrontendTables) == sizeof(float) * NFFT + sizeof(float2) * (NFFT / 2) + sizeof(int2) * NMEL +
                                              sizeof(float) * NBIN * NMEL, "FrontendTables layout");
  FrontendTables* h = new FrontendTables;
  for (int i = 0; i < NFFT; ++i) h->window[i] = 0.f;
  for (int i = 0; i < WINL; ++i) h->window[(NFFT - WINL) / 2 + i] = window400[i];
  for (int k = 0; k < NFFT / 2; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)NFFT;
    h->twiddle[k] = make_float2((float)cos(a), (float)sin(a));
  }
  for (int m = 0; m < NMEL; ++m) {
    int lo = NBIN, hi = 0;
    for (int k = 0; k < NBIN; ++k)
      if (mel_fb[k * NMEL + m] != 0.f) { if (k < lo) lo = k; hi = k + 1; }
    if (lo >= hi) { lo = 0; hi = 0; }
    h->melrange[m] = make_int2(lo, hi);
  }
  for (int i = 0; i < NBIN * NMEL; ++i) h->melfb[i] = mel_fb[i];
  const cudaError_t e = cudaMemcpy(dev, h, sizeof(FrontendTables), cudaMemcpyHostToDevice);
  delete h;
  if (e != cudaSuccess) { set_last_error("frontend tables upload failed: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

__device__ __forceinline__ int bitrev9(int i) { return (int)(__brev((unsigned)i) >> 23); }

// blockDim = 128 (4 warps = 4 frames).  Dynamic smem: 4 * 512 float2.
__global__ void __launch_bounds__(128) frontend_kernel(
    const float* __restrict__ wave_in, int ld_wave, const float* __restrict__ wbuf, int ld_wbuf,
    const FrontendDesc* __restrict__ desc, int n_desc, int total_frames, const FrontendTables* __restrict__ tab,
    const double* __restrict__ mean, const double* __restrict__ std_, float* __restrict__ featbuf,
    int feat_cap) {
  __shared__ float2 sm[4][NFFT];
  __shared__ float pw[4][NBIN + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gf = blockIdx.x * 4 + warp;           // flat frame index over all descriptors
  if (gf >= total_frames) return;
  // locate the descriptor (n_desc is small; binary search over frame_base)
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (desc[mid].frame_base <= gf) lo = mid; else hi = mid - 1;
  }
  const FrontendDesc d = desc[lo];
  const int f = gf - d.frame_base;
  if (f < d.emit0 || f >= d.emit1) return;        // trimmed frames are never computed
  const float* wb = wbuf + (size_t)d.stream * ld_wbuf;
  const float* wi = wave_in + (size_t)d.stream * ld_wave;
  const int L = d.slab, have = d.n_prev + d.n_new;
  float2* x = sm[warp];
  const int c = f * HOPS - NFFT / 2;              // slab index of the frame's first sample
  for (int i = lane; i < NFFT; i += 32) {
    int idx = c + i;
    if (idx < 0) idx = -idx;                       // reflect (torch.stft center=True)
    if (idx >= L) idx = 2 * (L - 1) - idx;
    float v = 0.f;
    if (idx < d.n_prev) v = wb[idx];
    else if (idx < have) v = wi[idx - d.n_prev];   // beyond `have`: zero padding of a short final slab
    x[bitrev9(i)] = make_float2(v * __ldg(&tab->window[i]), 0.f);
  }
  __syncwarp();
  // radix-2 decimation-in-time FFT, 9 stages, 256 butterflies per stage
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = NFFT >> s;
    for (int bfly = lane; bfly < NFFT / 2; bfly += 32) {
      int j = bfly & (half - 1);
      int i0 = ((bfly >> (s - 1)) << s) + j;
      int i1 = i0 + half;
      float2 w = __ldg(&tab->twiddle[j * tstep]);
      float2 a = x[i0], b = x[i1];
      float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      x[i0] = make_float2(a.x + t.x, a.y + t.y);
      x[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncwarp();
  }
  float* p = pw[warp];
  for (int k = lane; k < NBIN; k += 32) p[k] = x[k].x * x[k].x + x[k].y * x[k].y;
  __syncwarp();
  float* out = featbuf + ((size_t)d.stream * feat_cap + d.feat_off + (f - d.emit0)) * NMEL;
  for (int m = lane; m < NMEL; m += 32) {
    // taps outside [k_lo, k_hi) are exactly zero, and fma(p, 0, acc) == acc, so skipping them is bit-identical
    float acc = 0.f;
    const int2 rg = __ldg(&tab->melrange[m]);
    for (int k = rg.x; k < rg.y; ++k) acc = fmaf(p[k], __ldg(&tab->melfb[k * NMEL + m]), acc);
    float lg = logf(fmaxf(acc, 1e-10f));
    if (mean) lg = (float)(((double)lg - mean[m]) / std_[m]);   // numpy fp64 round trip, :355-358
    out[m] = lg;
  }
}

int launch_frontend(const FrontendTables* tab, const float* wave_in, int ld_wave, const float* wbuf, int ld_wbuf,
                    const FrontendDesc* desc, int n_desc, int total_frames, const double* mean, const double* std_,
                    float* featbuf, int feat_cap, cudaStream_t st) {
  if (total_frames <= 0) return 0;
  if (!tab) { set_last_error("frontend tables not uploaded"); return -1; }
  frontend_kernel<<<cdiv(total_frames, 4), 128, 0, st>>>(wave_in, ld_wave, wbuf, ld_wbuf, desc, n_desc,
                                                        total_frames, tab, mean, std_, featbuf, feat_cap);
  SCB_LAUNCH_CHECK();
  return 0;
}

// New waveform buffer = tail of (old buffer ++ chunk).  One CTA per descriptor; the tail is staged in
// registers before it is written because source and destination overlap inside wbuf.
__global__ void __launch_bounds__(512) wavebuf_update_kernel(const float* __restrict__ wave_in, int ld_wave,
                                                             float* __restrict__ wbuf, int ld_wbuf,
                                                             const FrontendDesc* __restrict__ desc) {
  const FrontendDesc d = desc[blockIdx.x];
  float* wb = wbuf + (size_t)d.stream * ld_wbuf;
  const float* wi = wave_in + (size_t)d.stream * ld_wave;
  const int have = d.n_prev + d.n_new;
  const int src0 = have - d.new_buf;
  const int i = threadIdx.x;
  float v = 0.f;
  if (i < d.new_buf) {
    int idx = src0 + i;
    v = idx < d.n_prev ? wb[idx] : wi[idx - d.n_prev];
  }
  __syncthreads();
  if (i < d.new_buf) wb[i] = v;
}

int launch_wavebuf_update(const float* wave_in, int ld_wave, float* wbuf, int ld_wbuf,
                          const FrontendDesc* desc, int n_desc, cudaStream_t st) {
  if (n_desc <= 0) return 0;
  wavebuf_update_kernel<<<n_desc, 512, 0, st>>>(wave_in, ld_wave, wbuf, ld_wbuf, desc);
  SCB_LAUNCH_CHECK();
  return 0;
}

}  // namespace scb
