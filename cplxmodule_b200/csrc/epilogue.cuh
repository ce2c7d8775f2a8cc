// Shared epilogue of the (variational) linear / conv forward kernels:
//   y = mu + bias + eps * sqrt(max(s2, 1e-8))
// (cplxmodule/nn/relevance/complex/base.py:56, real/base.py:49; bias add
// cplx.py:644-646).  One thread owns a run of C consecutive output columns of
// one output row.
#pragma once
#include "common.cuh"
#include "noise.cuh"

namespace cplxk {

struct EpiParams {
  const void* b_re;
  const void* b_im;
  const void* eps_re;
  const void* eps_im;
  void* y_re;
  void* y_im;
  int64_t M, N;        // logical output matrix [M, N] (rows = samples / output pixels)
  int64_t plane_elems; // elements of one noise plane in torch's randn call (= M*N)
  NoiseParams noise;
};

// `row_off` = linear offset of (m, 0) inside a plane; consecutive n are contiguous.
template <typename T, bool kCplx, bool kVD, int C>
__device__ __forceinline__ void epilogue_run(const EpiParams& p, int64_t m, int64_t n0,
                                             float (&re)[C], float (&im)[C], float (&s2)[C]) {
  static_assert(C % 4 == 0, "runs are whole noise quads");
  if (m >= p.M || n0 >= p.N) return;
  const int64_t row_off = m * p.N;
  const int nvalid = (p.N - n0) < C ? static_cast<int>(p.N - n0) : C;

  if (p.b_re) {
    const T* br = static_cast<const T*>(p.b_re);
    const T* bi = static_cast<const T*>(p.b_im);
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) {
        re[j] += Elem<T>::to_f(__ldg(br + n0 + j));
        if constexpr (kCplx) im[j] += Elem<T>::to_f(__ldg(bi + n0 + j));
      }
  }

  if constexpr (kVD) {
    float sd[C];
#pragma unroll
    for (int j = 0; j < C; ++j) sd[j] = sqrtf(fmaxf(s2[j], 1e-8f));

    if (p.noise.mode == CPLXK_NOISE_INJECT) {
      const T* er = static_cast<const T*>(p.eps_re) + row_off + n0;
      const T* ei = kCplx ? static_cast<const T*>(p.eps_im) + row_off + n0 : nullptr;
#pragma unroll
      for (int j = 0; j < C; ++j)
        if (j < nvalid) {
          re[j] = fmaf(Elem<T>::to_f(__ldg(er + j)), sd[j], re[j]);
          if constexpr (kCplx) im[j] = fmaf(Elem<T>::to_f(__ldg(ei + j)), sd[j], im[j]);
        }
    } else if (p.noise.mode == CPLXK_NOISE_PHILOX_TORCH) {
      TorchNoiseCursor cur;
      cur.seek(static_cast<uint64_t>(row_off + n0), p.noise.threads);
#pragma unroll
      for (int j = 0; j < C; ++j)
        if (j < nvalid) re[j] = fmaf(cur.next(p.noise) * p.noise.scale, sd[j], re[j]);
      if constexpr (kCplx) {
        cur.seek(static_cast<uint64_t>(p.plane_elems + row_off + n0), p.noise.threads);
#pragma unroll
        for (int j = 0; j < C; ++j)
          if (j < nvalid) im[j] = fmaf(cur.next(p.noise) * p.noise.scale, sd[j], im[j]);
      }
    } else {  // CPLXK_NOISE_PHILOX_FAST: quads are per-row groups of four columns
      const uint64_t quads_per_row = static_cast<uint64_t>((p.N + 3) >> 2);
      const uint64_t q0 = static_cast<uint64_t>(m) * quads_per_row + static_cast<uint64_t>(n0 >> 2);
#pragma unroll
      for (int q = 0; q < C / 4; ++q) {
        float4 g = philox_fast_normal4(q0 + q, 0u, p.noise);
        re[4 * q + 0] = fmaf(g.x * p.noise.scale, sd[4 * q + 0], re[4 * q + 0]);
        re[4 * q + 1] = fmaf(g.y * p.noise.scale, sd[4 * q + 1], re[4 * q + 1]);
        re[4 * q + 2] = fmaf(g.z * p.noise.scale, sd[4 * q + 2], re[4 * q + 2]);
        re[4 * q + 3] = fmaf(g.w * p.noise.scale, sd[4 * q + 3], re[4 * q + 3]);
        if constexpr (kCplx) {
          float4 h = philox_fast_normal4(q0 + q, 1u, p.noise);
          im[4 * q + 0] = fmaf(h.x * p.noise.scale, sd[4 * q + 0], im[4 * q + 0]);
          im[4 * q + 1] = fmaf(h.y * p.noise.scale, sd[4 * q + 1], im[4 * q + 1]);
          im[4 * q + 2] = fmaf(h.z * p.noise.scale, sd[4 * q + 2], im[4 * q + 2]);
          im[4 * q + 3] = fmaf(h.w * p.noise.scale, sd[4 * q + 3], im[4 * q + 3]);
        }
      }
    }
  }

  // ---- store
  constexpr int V = Elem<T>::kVec;
  T* yr = static_cast<T*>(p.y_re) + row_off + n0;
  T* yi = kCplx ? static_cast<T*>(p.y_im) + row_off + n0 : nullptr;
  const bool vec = (nvalid == C) && ((reinterpret_cast<uintptr_t>(yr) & 15u) == 0) &&
                   (!kCplx || (reinterpret_cast<uintptr_t>(yi) & 15u) == 0) && (C % V == 0);
  if (vec) {
#pragma unroll
    for (int c = 0; c < C / V; ++c) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = re[c * V + j];
      o.store(yr + c * V);
      if constexpr (kCplx) {
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = im[c * V + j];
        o.store(yi + c * V);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) {
        yr[j] = Elem<T>::from_f(re[j]);
        if constexpr (kCplx) yi[j] = Elem<T>::from_f(im[j]);
      }
  }
}

}  // namespace cplxk
