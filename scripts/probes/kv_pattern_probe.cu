// Bandwidth of the decoder cross-attention's K|V read pattern, without any arithmetic (timing experiment, round 2):
//   layout A (current): row = [hi K(256) | hi V(256) | lo K(256) | lo V(256)] fp16 per frame (2 KB); a (stream, head) CTA
//                       reads four 64-byte slices of every row
//   layout B:           [head][frame][K hi | K lo | V hi | V lo][32] : 256 contiguous bytes per (head, frame)
// Same grid / warp structure as dec_attn_x3_kernel: one CTA of 4 warps per (stream, head), a warp takes 16 frames of
// every 64-frame step and double-buffers them with cp.async.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o kv_pattern_probe kv_pattern_probe.cu && ./kv_pattern_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void cp16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}

template <int LAYOUT>
__global__ void __launch_bounds__(128, 4) probe(const __half* __restrict__ kv, int T, int Tcap, int H, float* out) {
  const int head = blockIdx.x, s = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ __align__(16) unsigned char sm[];
  __half* st = reinterpret_cast<__half*>(sm) + warp * 2 * (4 * 16 * 40);
  const int lr = lane >> 2, ch8 = (lane & 3) * 8;              // 8 rows per pass, 4 chunks of 16 B per 64-byte slice
  const size_t row_stride = 1024;
  const __half* base = LAYOUT == 0 ? kv + (size_t)s * Tcap * row_stride + head * 32
                                   : kv + ((size_t)s * H + head) * Tcap * 128;
  auto issue = [&](int t, int buf) {
    const int u0 = t * 64 + 16 * warp;
    __half* d = st + buf * (4 * 16 * 40);
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int r = lr + 8 * ps, u = u0 + r;
      if (u < T) {
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
          const __half* src = LAYOUT == 0 ? base + (size_t)u * row_stride + pl * 256 + ch8 : base + (size_t)u * 128 + pl * 32 + ch8;
          cp16(d + (pl * 16 + r) * 40 + ch8, src);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int n_steps = (T + 63) / 64;
  float acc = 0.f;
  if (n_steps > 0) issue(0, 0);
  for (int t = 0; t < n_steps; ++t) {
    if (t + 1 < n_steps) { issue(t + 1, (t + 1) & 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    acc += __half2float(st[(t & 1) * (4 * 16 * 40) + lane]);
    __syncwarp();
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main() {
  const int S = 256, H = 8, Tcap = 1600;
  const size_t elems = (size_t)S * Tcap * 1024;
  __half* kv; float* out;
  cudaMalloc(&kv, elems * 2); cudaMalloc(&out, 4);
  cudaMemset(kv, 0, elems * 2);
  const size_t smem = 4 * 2 * (4 * 16 * 40) * 2;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int T : {96, 256, 512, 750, 1500}) {
    for (int layout = 0; layout < 2; ++layout) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (layout == 0) probe<0><<<dim3(H, S), 128, smem>>>(kv, T, Tcap, H, out);
        else probe<1><<<dim3(H, S), 128, smem>>>(kv, T, Tcap, H, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const double bytes = (double)S * T * 2048.0;
      printf("{\"T\": %d, \"layout\": \"%s\", \"us\": %.1f, \"GB_per_s\": %.0f}\n", T, layout == 0 ? "row [hi K|V][lo K|V] (64-byte slices)" : "head-major (256 contiguous bytes per key)", best * 1e3, bytes / (best * 1e-3) / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
