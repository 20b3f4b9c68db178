// Shared helpers for the robir_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace robir {

void set_last_error(const char* fmt, ...);

#define RB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      robir::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

#define RB_REQUIRE(cond, msg)                                            \
  do {                                                                   \
    if (!(cond)) {                                                       \
      robir::set_last_error("%s (%s:%d)", msg, __FILE__, __LINE__);      \
      return 2;                                                          \
    }                                                                    \
  } while (0)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY02 = 2, ACT_SOFTPLUS100 = 3 };

// torch.nn.Softplus(beta=100, threshold=20) and its derivative
__device__ __forceinline__ float softplus100(float x) {
  const float bx = x * 100.f;
  return bx > 20.f ? x : log1pf(expf(bx)) / 100.f;
}
__device__ __forceinline__ float softplus100_grad(float x) {
  const float bx = x * 100.f;
  if (bx > 20.f) return 1.f;
  const float e = expf(bx);
  return e / (e + 1.f);
}

}  // namespace robir
