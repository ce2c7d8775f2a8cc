// Counter-based Gaussian noise for the local-reparameterisation epilogue.
//
// PHILOX_TORCH reproduces, element by element, what torch's CUDA
// `normal_` kernel writes for a contiguous tensor of `numel` floats
// (ATen/native/cuda/DistributionTemplates.h, distribution_elementwise_grid_stride_kernel):
//   T = 256 * grid threads; thread idx runs curand_init(seed, idx, offset) and
//   its k-th curand_normal4() lands on elements  idx + T*(4k + ii), ii = 0..3.
// Hence element li  ->  subsequence idx = li % T, slot j = li / T,
//   Philox counter = (offset/4 + j/4, idx), Box-Muller component j % 4.
// The reference consumes this stream in cplx.randn (cplxmodule/cplx.py:544-550)
// as ONE randn(2, M, N) / sqrt(2) call: plane 0 -> real, plane 1 -> imag.
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>  // _curand_box_muller: the exact Box-Muller torch's kernel uses
#include <stdint.h>

namespace cplxk {

struct PhiloxKey {
  uint32_t k0, k1;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, PhiloxKey key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

struct NoiseParams {
  int mode;            // cplxk_noise
  uint32_t seed_lo, seed_hi;
  uint64_t ctr_base;   // offset / 4
  uint32_t threads;    // T (PHILOX_TORCH)
  float scale;         // 1/sqrt(2) as torch computes it for complex, 1 for real
};

// standard normal that torch would write to linear element `li` (see header)
__device__ __forceinline__ float philox_torch_normal(uint32_t idx, uint64_t slot,
                                                     const NoiseParams& np) {
  uint64_t ctr = np.ctr_base + (slot >> 2);
  uint4 c = make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u);
  uint4 r = philox4x32_10(c, PhiloxKey{np.seed_lo, np.seed_hi});
  uint32_t comp = static_cast<uint32_t>(slot) & 3u;
  float2 g = (comp < 2u) ? _curand_box_muller(r.x, r.y) : _curand_box_muller(r.z, r.w);
  return (comp & 1u) ? g.y : g.x;
}

// Same normal with the slot-dependent parts (Philox counter, Box-Muller component) supplied by
// the caller: a run of consecutive elements that does not wrap the thread index shares them, so
// they are computed once per run.  Branch-free (selects): the eight chains of a run interleave.
// Bit-identical to philox_torch_normal: _curand_box_muller is
//   u = x * 2^-32 + 2^-33,  v = y * (2^-32 * 2 pi) + (2^-32 * 2 pi / 2),  s = sqrtf(-2 logf(u)),
//   __sincosf(v, &sin, &cos)  ->  (sin * s, cos * s).
__device__ __forceinline__ float philox_torch_normal_at(uint64_t ctr, uint32_t idx, uint32_t comp,
                                                        const NoiseParams& np) {
  uint4 c = make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u);
  uint4 r = philox4x32_10(c, PhiloxKey{np.seed_lo, np.seed_hi});
  const uint32_t ux = comp < 2u ? r.x : r.z, uy = comp < 2u ? r.y : r.w;
  const float u = ux * CURAND_2POW32_INV + (CURAND_2POW32_INV / 2);
  const float v = uy * CURAND_2POW32_INV_2PI + (CURAND_2POW32_INV_2PI / 2);
  const float s = sqrtf(-2.0f * logf(u));
  const float t = (comp & 1u) ? __cosf(v) : __sinf(v);
  return t * s;
}

// Walks consecutive linear elements li, li+1, ... of torch's layout without a
// division per element.
struct TorchNoiseCursor {
  uint32_t idx;
  uint64_t slot;
  __device__ __forceinline__ void seek(uint64_t li, uint32_t T) {
    slot = li / T;
    idx = static_cast<uint32_t>(li - slot * T);
  }
  __device__ __forceinline__ float next(const NoiseParams& np) {
    float v = philox_torch_normal(idx, slot, np);
    if (++idx == np.threads) {
      idx = 0;
      ++slot;
    }
    return v;
  }
};

// PHILOX_FAST: one Philox call -> 4 normals. Counter = (ctr_base + q, plane | 0x80000000)
// where q indexes groups of four consecutive elements of one plane.
__device__ __forceinline__ float4 philox_fast_normal4(uint64_t quad, uint32_t plane,
                                                      const NoiseParams& np) {
  uint64_t ctr = np.ctr_base + quad;
  uint4 c = make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), plane,
                       0x80000000u);
  uint4 r = philox4x32_10(c, PhiloxKey{np.seed_lo, np.seed_hi});
  float2 a = _curand_box_muller(r.x, r.y);
  float2 b = _curand_box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

// PHILOX_FAST, complex conv outputs: neighbouring elements of one output plane belong to
// different threads there (a thread walks channels), so the unit is one complex element:
// counter = (ctr_base + element, 0, 0x80000001) and ONE Box-Muller gives its (re, im) pair.
__device__ __forceinline__ float2 philox_fast_pair(uint64_t elem, const NoiseParams& np) {
  uint64_t ctr = np.ctr_base + elem;
  uint4 c = make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u,
                       0x80000001u);
  uint4 r = philox4x32_10(c, PhiloxKey{np.seed_lo, np.seed_hi});
  return _curand_box_muller(r.x, r.y);
}

}  // namespace cplxk
