// Shared helpers for the textgcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/textgcn_b200.h"

namespace tgcn {

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of this library bumps tgcn_launch_count()

#define TGCN_CHECK_ARG(cond, ...)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      ::tgcn::set_error(__VA_ARGS__);                     \
      return TGCN_EINVAL;                                 \
    }                                                     \
  } while (0)

#define TGCN_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::tgcn::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                        cudaGetErrorString(_e));                                          \
      return TGCN_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define TGCN_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    ::tgcn::count_launch();                                                               \
    cudaError_t _e = cudaPeekAtLastError();                                               \
    if (_e != cudaSuccess) {                                                              \
      ::tgcn::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                        cudaGetErrorString(_e));                                          \
      return TGCN_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// ---- Philox4x32-10, counter = element index, key = seed (same stream in fwd and bwd) ----
__device__ __forceinline__ uint4 philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t seed) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep decision for 4 consecutive elements starting at element index `e4*4`
// (one Philox call yields the four 32-bit lanes of elements 4*e4 .. 4*e4+3).
__device__ __forceinline__ uint4 philox_quad(uint64_t e4, uint64_t seed, uint64_t offset) {
  return philox4x32_10(e4, offset, seed);
}
// uniform in [0,1) from 32 random bits, 24-bit mantissa (same convention as curand_uniform shifted)
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// Stores through an NVSwitch multicast mapping (one store, replicated to every rank's copy of a
// symmetric buffer).  Used by producer kernels to fuse the exchange step of the row partition into
// their own output stores (SASS: STG.E[.128].STRONG.SYS on the multicast address).
__device__ __forceinline__ void multimem_st_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void multimem_st_f32(float* addr, float a) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

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

}  // namespace tgcn
