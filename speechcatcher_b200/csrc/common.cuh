// Shared device/host helpers for the speechcatcher_b200 CUDA path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace scb {

constexpr float kLogZero = -10000000000.0f;   // reference: ctc_prefix_score_full.py:58
constexpr int kBlock = 40;                    // encoder block size   (beam_search.py:285)
constexpr int kHopB = 16;                     // encoder block hop    (beam_search.py:286)
constexpr int kLook = 16;                     // look-ahead           (beam_search.py:287)
constexpr int kSlots = kBlock + 2;            // ctx-in + 40 frames + ctx-out
constexpr int kPreBeam = 40;                  // beam_search.py:75
constexpr int kMaxLength = 500;               // beam_search.py:289
constexpr int kNumSMs = 148;

// error plumbing: kernels never throw; launchers record the first failure
void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define SCB_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      scb::set_last_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,        \
                          cudaGetErrorString(_e));                                    \
      return -1;                                                                      \
    }                                                                                 \
  } while (0)

#define SCB_LAUNCH_CHECK()                                                            \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      scb::set_last_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,    \
                          cudaGetErrorString(_e));                                    \
      return -1;                                                                      \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// torch.logsumexp over two values: m = max; log(exp(a-m)+exp(b-m)) + m  (m forced to 0 if inf)
__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (isinf(m)) m = 0.f;
  return logf(expf(a - m) + expf(b - m)) + m;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Function attributes (cudaFuncSetAttribute) and __constant__ / __device__ symbols are per DEVICE, not per process: a
// "done" flag (or the largest dynamic shared-memory size set so far) is therefore kept per device ordinal.
struct PerDeviceMark {
  size_t v[64] = {};
  size_t& cur() { int d = 0; cudaGetDevice(&d); return v[d & 63]; }
};

// Programmatic dependent launch (PDL): the decode step is a chain of ~170 small dependent kernels, so the launch
// and CTA-scheduling latency of kernel N+1 is overlapped with the execution of kernel N.  Every kernel launched
// through launch_k() MUST call pdl_sync() as its first statement (before any early exit): griddepcontrol.wait
// blocks until the preceding grid has completed and its writes are visible, so ordering is exactly that of a
// plain stream; launch_dependents then lets the next grid's CTAs be scheduled while this one runs.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
extern bool g_use_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace scb
