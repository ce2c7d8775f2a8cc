// Shared helpers: status codes, storage-type traits, vector loads.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cplxk.h"

namespace cplxk {

#define CPLXK_CUDA_TRY(expr)                    \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) {                    \
      set_last_cuda_error(_e);                  \
      return CPLXK_ERR_CUDA;                    \
    }                                           \
  } while (0)

void set_last_cuda_error(cudaError_t e);

// sqrt(max(s2, 1e-8)) of the reparameterisation: one MUFU.SQRT.  The IEEE sqrtf() expands to
// rsqrt + two Newton steps + a BRANCH to a slow path per element, which serialises the epilogue
// (no interleaving across the run) while the tensor pipe waits for its accumulators back; the
// argument is >= 1e-8 (normal), and the result differs from the correctly rounded one by at most
// a couple of ulp (~2e-7 relative; the acceptance bound is 1e-3).
__device__ __forceinline__ float sd_of(float s2) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s2, 1e-8f)));
  return r;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of T viewed as floats
template <typename T>
struct Vec16 {
  static constexpr int N = Elem<T>::kVec;
  float v[N];
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (N == 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
      uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x, v[2 * i + 1] = f.y;
      }
    }
  }
  // generic (shared or global) non-read-only load
  __device__ __forceinline__ void load_smem(const T* p) {
    if constexpr (N == 4) {
      float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
      uint4 t = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x, v[2 * i + 1] = f.y;
      }
    }
  }
  __device__ __forceinline__ void store(T* p) const {
    if constexpr (N == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint4 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint4*>(p) = t;
    }
  }
};

}  // namespace cplxk
