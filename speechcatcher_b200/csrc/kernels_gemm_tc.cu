// bf16 tensor-core GEMM (tcgen05 + TMEM + TMA) -- placeholder until the tcgen05 kernel lands.
#include "kernels.h"
namespace scb {
int launch_gemm_bf16(const __nv_bfloat16*, int, const __nv_bfloat16*, const float*, const float*, int, float*, int,
                     __nv_bfloat16*, int, int, int, int, int, const int*, cudaStream_t) {
  set_last_error("bf16 tensor-core GEMM is not built in this revision");
  return -1;
}
}  // namespace scb
