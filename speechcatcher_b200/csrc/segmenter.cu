// Offline segmentation of long files: the energy curve on the GPU, the
// cut-point search on the host.
//
// Replaces: speechcatcher/simple_endpointing.py:72-75  (logfbank energy + Gaussian smoothing, fp64 like numpy/scipy)
//           speechcatcher/simple_endpointing.py:22-69  (BeamSearch: cost function and cut-point search)
// The feature step restates python_speech_features 0.6 `logfbank` (pre-emphasis 0.97, 400-sample rectangular
// frames every 160 samples, zero-padded tail, 512-point power spectrum / 512, 26 HTK-mel triangles floored to FFT
// bins, zeros -> eps, log); that package is unpinned in the reference and absent from the build image.
//
// seg_energy_kernel : one warp per 25 ms frame; int16 PCM -> pre-emphasis -> radix-2 fp64 FFT in shared memory ->
//                     26 filter-bank sums -> sum(log) / 10.  Reads 2 B x 160 new samples per frame (HBM-trivial:
//                     115 MB for an hour of audio); bound by the fp64 butterflies (9 x 256 per frame).
// seg_smooth_kernel : scipy.ndimage.gaussian_filter1d (mode="reflect", radius int(4 sigma + .5)) in scipy's own
//                     summation order (centre tap, then symmetric pairs outwards-in), negated.
#include <math.h>
#include <algorithm>
#include <vector>
#include "../../include/speechcatcher_b200.h"
#include "common.cuh"

namespace scb {

constexpr int SEG_NFFT = 512, SEG_WIN = 400, SEG_HOP = 160, SEG_NBIN = 257, SEG_NFILT = 26;
constexpr int SEG_MAX_RADIUS = 200;

__constant__ double2 c_tw64[SEG_NFFT / 2];   // exp(-2 pi i k / 512) in fp64
static PerDeviceMark g_tw64_ready;     // the __constant__ twiddle table exists once per device

struct SegBins { double b[SEG_NFILT + 2]; };
struct SegGauss { int radius; double w[2 * SEG_MAX_RADIUS + 1]; };

// base.get_filterbanks: bin edges = floor((nfft + 1) * mel2hz(linspace(mel(0), mel(sr / 2), nfilt + 2)) / sr)
static void host_filterbank_bins(double* b) {
  const double lowmel = 2595.0 * log10(1.0 + 0.0 / 700.0);
  const double highmel = 2595.0 * log10(1.0 + 8000.0 / 700.0);
  const int n = SEG_NFILT + 2;
  const double step = (highmel - lowmel) / (double)(n - 1);          // numpy.linspace: start + i * step,
  for (int i = 0; i < n; ++i) {                                       // last point forced to `stop`
    double mel = (i == n - 1) ? highmel : lowmel + (double)i * step;
    double hz = 700.0 * (pow(10.0, mel / 2595.0) - 1.0);
    b[i] = floor((double)(SEG_NFFT + 1) * hz / 16000.0);
  }
}

__device__ __forceinline__ int seg_bitrev9(int i) { return (int)(__brev((unsigned)i) >> 23); }

// blockDim = 128: 4 warps = 4 frames.  32 KB of static shared memory.
__global__ void __launch_bounds__(128) seg_energy_kernel(const int16_t* __restrict__ pcm, long long n_samples,
                                                         long long n_frames, SegBins bins,
                                                         double* __restrict__ energy) {
  __shared__ double2 sm[4][SEG_NFFT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * 4 + warp;
  if (t >= n_frames) return;
  double2* x = sm[warp];
  const long long base = t * SEG_HOP;
  for (int i = lane; i < SEG_NFFT; i += 32) {
    const long long n = base + i;
    double v = 0.0;
    if (i < SEG_WIN && n < n_samples) {
      const double cur = (double)pcm[n];
      // signal[1:] - 0.97 * signal[:-1] in numpy: one rounded product, one rounded difference (no FMA)
      v = n > 0 ? __dsub_rn(cur, __dmul_rn(0.97, (double)pcm[n - 1])) : cur;
    }
    x[seg_bitrev9(i)] = make_double2(v, 0.0);
  }
  __syncwarp();
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1), tstep = SEG_NFFT >> s;
    for (int bfly = lane; bfly < SEG_NFFT / 2; bfly += 32) {
      const int j = bfly & (half - 1);
      const int i0 = ((bfly >> (s - 1)) << s) + j, i1 = i0 + half;
      const double2 w = c_tw64[j * tstep];
      const double2 a = x[i0], b = x[i1];
      const double2 tt = make_double2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      x[i0] = make_double2(a.x + tt.x, a.y + tt.y);
      x[i1] = make_double2(a.x - tt.x, a.y - tt.y);
    }
    __syncwarp();
  }
  // power spectrum / nfft in place (real part), then one filter per lane
  for (int k = lane; k < SEG_NBIN; k += 32) {
    const double2 c = x[k];
    x[k].x = (1.0 / SEG_NFFT) * (c.x * c.x + c.y * c.y);
  }
  __syncwarp();
  double lg = 0.0;
  if (lane < SEG_NFILT) {
    const double b0 = bins.b[lane], b1 = bins.b[lane + 1], b2 = bins.b[lane + 2];
    double acc = 0.0;
    for (int i = (int)b0; i < (int)b1; ++i) acc += x[i].x * (((double)i - b0) / (b1 - b0));
    for (int i = (int)b1; i < (int)b2; ++i) acc += x[i].x * ((b2 - (double)i) / (b2 - b1));
    if (acc == 0.0) acc = 2.220446049250313e-16;      // numpy.finfo(float).eps
    lg = log(acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
  if (lane == 0) energy[t] = lg / 10.0;
}

__device__ __forceinline__ long long seg_reflect(long long i, long long n) {
  // scipy "reflect": (d c b a | a b c d | d c b a), period 2n
  const long long p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

__global__ void __launch_bounds__(256) seg_smooth_kernel(const double* __restrict__ e, long long n, SegGauss g,
                                                         double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int r = g.radius;
  double tmp = __dmul_rn(e[t], g.w[r]);
  const bool inner = t - r >= 0 && t + r < n;
  for (int ii = -r; ii < 0; ++ii) {
    const double a = inner ? e[t + ii] : e[seg_reflect(t + ii, n)];
    const double b = inner ? e[t - ii] : e[seg_reflect(t - ii, n)];
    tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), g.w[ii + r]));
  }
  out[t] = -tmp;
}

static int ensure_tw64() {
  if (g_tw64_ready.cur()) return 0;
  double2 tw[SEG_NFFT / 2];
  for (int k = 0; k < SEG_NFFT / 2; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)SEG_NFFT;
    tw[k] = make_double2(cos(a), sin(a));
  }
  SCB_CUDA_CHECK(cudaMemcpyToSymbol(c_tw64, tw, sizeof(tw)));
  g_tw64_ready.cur() = 1;
  return 0;
}

static long long seg_num_frames(long long n_samples) {
  if (n_samples <= SEG_WIN) return 1;
  return 1 + (n_samples - SEG_WIN + SEG_HOP - 1) / SEG_HOP;     // 1 + ceil((n - 400) / 160)
}

// ------------------------------------------------------------------------------------------ host cut search
struct CutNode { int parent; long long cut; };

static int cut_search(const double* smoothed, long long n_frames, const ScSegmentParams& p, std::vector<long long>& best) {
  struct Seq { int node; double score; };
  struct Cand { int parent_node; long long cut; double score; };
  std::vector<CutNode> arena;
  arena.push_back({-1, 0});
  std::vector<Seq> seqs{{0, 0.0}};
  std::vector<Cand> cands;
  const double ideal = (double)p.ideal_segment_len;
  const double factor = p.len_reward_weight / (double)p.ideal_segment_len;
  for (;;) {
    cands.clear();
    bool expand = false;
    const double worst = seqs.back().score;
    for (const Seq& s : seqs) {
      const long long last = arena[s.node].cut;
      const long long hi = std::min<long long>(p.max_lookahead, n_frames - last - 1);
      for (long long j = p.min_len; j < hi; j += p.step) {
        // cost_function (simple_endpointing.py:36-41); every product/sum rounded separately like the Python floats
        volatile double reward = factor * (ideal - fabs(ideal - (double)j));
        volatile double lw = p.len_reward_weight * reward;
        volatile double ew = p.energy_weight * smoothed[last + j];
        volatile double cost = lw + ew;
        const double nw = s.score + cost;
        if (nw > s.score) cands.push_back({s.node, last + j + 1, nw});
        if (nw > worst) expand = true;
      }
    }
    if (cands.empty() || !expand) break;
    // sorted(..., reverse=True)[:beam]: stable, descending
    std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });
    const size_t keep = std::min<size_t>(cands.size(), (size_t)p.beam_size);
    seqs.clear();
    for (size_t i = 0; i < keep; ++i) {
      arena.push_back({cands[i].parent_node, cands[i].cut});
      seqs.push_back({(int)arena.size() - 1, cands[i].score});
    }
  }
  best.clear();
  for (int n = seqs[0].node; n >= 0; n = arena[n].parent) best.push_back(arena[n].cut);
  std::reverse(best.begin(), best.end());
  if (best.size() == 1) best.push_back(n_frames);            // [0] -> [0, fbank_feat_len]  (:68)
  return 0;
}

}  // namespace scb

using namespace scb;

extern "C" {

int sc_segment_num_frames(int64_t n_samples, int64_t* n_frames) {
  if (!n_frames || n_samples < 1) { set_last_error("segment_num_frames: bad argument"); return SC_ERR_ARG; }
  *n_frames = seg_num_frames(n_samples);
  return SC_OK;
}

int sc_segment_filterbank_bins(double* bins28) {
  if (!bins28) { set_last_error("segment_filterbank_bins: null argument"); return SC_ERR_ARG; }
  host_filterbank_bins(bins28);
  return SC_OK;
}

int sc_segment_energy(const int16_t* pcm_dev, int64_t n_samples, double* energy_dev, double* smoothed_dev,
                      int64_t n_frames, double sigma, void* stream) {
  if (!pcm_dev || !energy_dev || !smoothed_dev || n_samples < 1) {
    set_last_error("segment_energy: null argument"); return SC_ERR_ARG;
  }
  if (n_frames != seg_num_frames(n_samples)) {
    set_last_error("segment_energy: n_frames=%lld but %lld samples give %lld frames", (long long)n_frames,
                   (long long)n_samples, seg_num_frames(n_samples));
    return SC_ERR_ARG;
  }
  SegGauss g;
  g.radius = (int)(4.0 * sigma + 0.5);                          // scipy: int(truncate * sd + 0.5), truncate = 4
  if (sigma <= 0.0 || g.radius > SEG_MAX_RADIUS) {
    set_last_error("segment_energy: sigma %.3f out of range (radius <= %d)", sigma, SEG_MAX_RADIUS); return SC_ERR_ARG;
  }
  double sum = 0.0;
  for (int i = -g.radius; i <= g.radius; ++i) {                 // scipy _gaussian_kernel1d: exp(-0.5 / sigma^2 * x^2)
    g.w[i + g.radius] = exp(-0.5 / (sigma * sigma) * (double)(i * i));
    sum += g.w[i + g.radius];
  }
  for (int i = 0; i <= 2 * g.radius; ++i) g.w[i] /= sum;
  SegBins bins;
  host_filterbank_bins(bins.b);
  if (ensure_tw64()) return SC_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  seg_energy_kernel<<<(unsigned)((n_frames + 3) / 4), 128, 0, st>>>(pcm_dev, n_samples, n_frames, bins, energy_dev);
  SCB_LAUNCH_CHECK();
  seg_smooth_kernel<<<(unsigned)((n_frames + 255) / 256), 256, 0, st>>>(energy_dev, n_frames, g, smoothed_dev);
  SCB_LAUNCH_CHECK();
  return SC_OK;
}

int sc_segment_search(const double* smoothed_host, int64_t n_frames, const ScSegmentParams* p, int64_t* cuts,
                      int32_t max_cuts, int32_t* n_cuts) {
  if (!smoothed_host || !p || !cuts || !n_cuts || n_frames < 1) {
    set_last_error("segment_search: null argument"); return SC_ERR_ARG;
  }
  if (p->beam_size < 1 || p->step < 1 || p->ideal_segment_len < 1 || p->min_len < 0) {
    set_last_error("segment_search: bad parameters"); return SC_ERR_ARG;
  }
  std::vector<long long> best;
  cut_search(smoothed_host, n_frames, *p, best);
  if ((int64_t)best.size() > max_cuts) {
    set_last_error("segment_search: %zu cuts do not fit max_cuts=%d", best.size(), max_cuts); return SC_ERR_CAPACITY;
  }
  for (size_t i = 0; i < best.size(); ++i) cuts[i] = best[i];
  *n_cuts = (int32_t)best.size();
  return SC_OK;
}

}  // extern "C"
