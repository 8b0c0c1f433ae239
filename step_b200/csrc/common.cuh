// Shared device/host helpers for libstep_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/step_b200.h"

namespace stepk {

// ---- error plumbing -------------------------------------------------------
extern thread_local char g_last_error[512];
extern unsigned long long g_launch_count;      // kernels enqueued by this library (every launch site calls check_launch)

// fmt may use up to two %lld
inline int fail(int code, const char *fmt, long long x = 0, long long y = 0) {
  snprintf(g_last_error, sizeof(g_last_error), fmt, x, y);
  return code;
}
inline int fail_msg(int code, const char *msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}

inline int check_launch(const char *what) {
  __sync_fetch_and_add(&g_launch_count, 1ull);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return STEP_OK;
}

#define STEP_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) return stepk::fail_msg(STEP_EINVAL, msg " (" #cond ")");    \
  } while (0)

#define STEP_LAUNCH_CHECK(what)                        \
  do {                                                 \
    int _rc = stepk::check_launch(what);               \
    if (_rc != STEP_OK) return _rc;                    \
  } while (0)

// opt a kernel into > 48 KB dynamic shared memory once per process
template <typename K>
inline int allow_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return STEP_OK;
}

// ---- warp helpers ----------------------------------------------------------
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
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- counter-based RNG (Philox4x32-10) ---------------------------------------
// Used for every stochastic site (dropout, Gumbel draws) so that forward and backward can
// regenerate the same stream from (seed, site, element index) without storing masks.
__device__ __forceinline__ uint4 philox4x32(uint64_t counter, uint64_t key) {
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__host__ __device__ __forceinline__ uint64_t rng_key(uint64_t seed, uint32_t site) {
  return seed ^ ((uint64_t)site * 0x9E3779B97F4A7C15ull);
}
// U[0,1) with 24 random bits, the granularity of torch.rand for float32
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
// keep-mask: true with probability 1-p; thr = (uint32)(p * 2^32)
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}


// Packed fp32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): (c0, c1) += (a0, a1) * (b0, b1) in one issue slot.
__device__ __forceinline__ void ffma2(float &c0, float &c1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(c0), "+f"(c1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
}  // namespace stepk
