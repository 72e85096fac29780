// Shared helpers for the diffgfdn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "diffgfdn_b200.h"

namespace dgfdn {

void set_error(const char* fmt, ...);

#define DGFDN_CHECK(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      dgfdn::set_error(__VA_ARGS__);  \
      return 1;                       \
    }                                 \
  } while (0)

#define DGFDN_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      dgfdn::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

#define DGFDN_LAUNCH_CHECK() DGFDN_CUDA(cudaGetLastError())

int sm_count();

// ---- complex arithmetic on double2 / float2 -------------------------------------------------
__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__host__ __device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  double s = 1.0 / (b.x * b.x + b.y * b.y);
  return make_double2((a.x * b.x + a.y * b.y) * s, (a.y * b.x - a.x * b.y) * s);
}
__host__ __device__ __forceinline__ double2 cinv(double2 b) {
  double s = 1.0 / (b.x * b.x + b.y * b.y);
  return make_double2(b.x * s, -b.y * s);
}
__host__ __device__ __forceinline__ double cnorm(double2 a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

__host__ __device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ float2 cmulcf(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

}  // namespace dgfdn
