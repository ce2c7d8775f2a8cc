// Outer-product features of the bilinear layers.
//
//   y_j = x1^{H or T} A_j x2 + b_j = sum_{p,q} z[(p,q)] A_j[(p,q)],   z[(p,q)] = conj?(x1_p) x2_q
//
// (cplxmodule/cplx.py:1062-1090 bilinear_naive: four F.bilinear calls on the real planes; real
// layers: torch.nn.functional.bilinear as used by nn/relevance/real/base.py:52-80).  Flattening
// (p,q) turns the bilinear map into the AFFINE map of cplx.linear on z with the weight viewed as
// [out, in1 * in2], and the variance of its local-reparameterisation forward
// F.bilinear(|x1|^2, |x2|^2, exp(log_sigma2)) (complex/base.py:77-82) into the variance GEMM on
// |z|^2 = |x1|^2 |x2|^2 -- so every bilinear layer runs on the linear tensor-core kernels; this
// file only builds z (HBM-bound elementwise) and propagates gradients through it.
#include "common.cuh"
#include "knobs.cuh"

namespace cplxk {

template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
outer_fwd_kernel(const T* __restrict__ x1_re, const T* __restrict__ x1_im, const T* __restrict__ x2_re,
                 const T* __restrict__ x2_im, T* __restrict__ z_re, T* __restrict__ z_im, int64_t B,
                 int d1, int d2, int conj1) {
  const int64_t per_row = static_cast<int64_t>(d1) * d2, n = B * per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / per_row;
    const int r = static_cast<int>(i - b * per_row);
    const int p = r / d2, q = r - p * d2;
    const float a = Elem<T>::to_f(x1_re[b * d1 + p]);
    const float u = Elem<T>::to_f(x2_re[b * d2 + q]);
    if constexpr (kCplx) {
      float bb = Elem<T>::to_f(x1_im[b * d1 + p]);
      if (conj1) bb = -bb;
      const float v = Elem<T>::to_f(x2_im[b * d2 + q]);
      z_re[i] = Elem<T>::from_f(fmaf(a, u, -bb * v));
      z_im[i] = Elem<T>::from_f(fmaf(a, v, bb * u));
    } else {
      z_re[i] = Elem<T>::from_f(a * u);
    }
  }
}

// one block per batch row: dx1_p = sum_q g_pq * d z_pq / d x1_p, dx2_q = sum_p ...
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
outer_bwd_kernel(const T* __restrict__ g_re, const T* __restrict__ g_im, const T* __restrict__ x1_re,
                 const T* __restrict__ x1_im, const T* __restrict__ x2_re, const T* __restrict__ x2_im,
                 T* __restrict__ d1_re, T* __restrict__ d1_im, T* __restrict__ d2_re,
                 T* __restrict__ d2_im, int64_t B, int d1, int d2, int conj1) {
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const T* gr = g_re + b * d1 * d2;
    const T* gi = kCplx ? g_im + b * d1 * d2 : nullptr;
    const float sgn = conj1 ? -1.f : 1.f;   // z uses (a, sgn * b) for x1
    if (d1_re) {
      // one warp per p: lanes stride over q (coalesced), shuffle-reduce
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (int p = warp; p < d1; p += 8) {
        float sa = 0.f, sb = 0.f;
        for (int q = lane; q < d2; q += 32) {
          const float g0 = Elem<T>::to_f(gr[p * d2 + q]);
          const float u = Elem<T>::to_f(x2_re[b * d2 + q]);
          if constexpr (kCplx) {
            const float g1 = Elem<T>::to_f(gi[p * d2 + q]);
            const float v = Elem<T>::to_f(x2_im[b * d2 + q]);
            // z_re = a u - (s b) v, z_im = a v + (s b) u
            sa += fmaf(g0, u, g1 * v);
            sb += sgn * fmaf(g1, u, -g0 * v);
          } else {
            sa = fmaf(g0, u, sa);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          if constexpr (kCplx) sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (lane == 0) {
          d1_re[b * d1 + p] = Elem<T>::from_f(sa);
          if constexpr (kCplx) d1_im[b * d1 + p] = Elem<T>::from_f(sb);
        }
      }
    }
    if (d2_re) {
      for (int q = threadIdx.x; q < d2; q += blockDim.x) {
        float su = 0.f, sv = 0.f;
        for (int p = 0; p < d1; ++p) {
          const float g0 = Elem<T>::to_f(gr[p * d2 + q]);
          const float a = Elem<T>::to_f(x1_re[b * d1 + p]);
          if constexpr (kCplx) {
            const float g1 = Elem<T>::to_f(gi[p * d2 + q]);
            const float bb = sgn * Elem<T>::to_f(x1_im[b * d1 + p]);
            su += fmaf(g0, a, g1 * bb);
            sv += fmaf(g1, a, -g0 * bb);
          } else {
            su = fmaf(g0, a, su);
          }
        }
        d2_re[b * d2 + q] = Elem<T>::from_f(su);
        if constexpr (kCplx) d2_im[b * d2 + q] = Elem<T>::from_f(sv);
      }
    }
  }
}

}  // namespace cplxk

using namespace cplxk;

extern "C" int cplxk_outer_fwd(const void* x1_re, const void* x1_im, const void* x2_re,
                               const void* x2_im, void* z_re, void* z_im, int64_t B, int64_t d1,
                               int64_t d2, int conjugate, int dtype, void* stream) {
  if (!x1_re || !x2_re || !z_re || B < 0 || d1 < 0 || d2 < 0) return CPLXK_ERR_BADARG;
  const bool cplx = x1_im != nullptr;
  if (cplx != (x2_im != nullptr) || cplx != (z_im != nullptr)) return CPLXK_ERR_BADARG;
  if (d1 > 0x7fffffff || d2 > 0x7fffffff || d1 * d2 > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;
  const int64_t n = B * d1 * d2;
  if (n == 0) return CPLXK_OK;
  int sms = 148;
  if (current_device_sm_count(&sms) != CPLXK_OK) sms = 148;
  const int64_t want = (n + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want > 16LL * sms ? 16LL * sms : want);
  auto st = static_cast<cudaStream_t>(stream);
#define CPLXK_OUTER(T)                                                                                  \
  if (cplx)                                                                                             \
    outer_fwd_kernel<T, true><<<grid, 256, 0, st>>>(                                                    \
        static_cast<const T*>(x1_re), static_cast<const T*>(x1_im), static_cast<const T*>(x2_re),       \
        static_cast<const T*>(x2_im), static_cast<T*>(z_re), static_cast<T*>(z_im), B,                  \
        static_cast<int>(d1), static_cast<int>(d2), conjugate);                                         \
  else                                                                                                  \
    outer_fwd_kernel<T, false><<<grid, 256, 0, st>>>(static_cast<const T*>(x1_re), nullptr,             \
                                                     static_cast<const T*>(x2_re), nullptr,             \
                                                     static_cast<T*>(z_re), nullptr, B,                 \
                                                     static_cast<int>(d1), static_cast<int>(d2), 0);
  if (dtype == CPLXK_F32) { CPLXK_OUTER(float) }
  else if (dtype == CPLXK_BF16) { CPLXK_OUTER(__nv_bfloat16) }
  else return CPLXK_ERR_BADARG;
#undef CPLXK_OUTER
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_outer_bwd(const void* g_re, const void* g_im, const void* x1_re,
                               const void* x1_im, const void* x2_re, const void* x2_im,
                               void* d1_re, void* d1_im, void* d2_re, void* d2_im, int64_t B,
                               int64_t d1, int64_t d2, int conjugate, int dtype, void* stream) {
  if (!g_re || !x1_re || !x2_re || B < 0 || d1 < 0 || d2 < 0) return CPLXK_ERR_BADARG;
  const bool cplx = x1_im != nullptr;
  if (cplx != (x2_im != nullptr) || cplx != (g_im != nullptr)) return CPLXK_ERR_BADARG;
  if (cplx && ((d1_re != nullptr) != (d1_im != nullptr) || (d2_re != nullptr) != (d2_im != nullptr)))
    return CPLXK_ERR_BADARG;
  if (d1 > 0x7fffffff || d2 > 0x7fffffff || d1 * d2 > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;
  if (B * d1 * d2 == 0 || (!d1_re && !d2_re)) return CPLXK_OK;
  int sms = 148;
  if (current_device_sm_count(&sms) != CPLXK_OK) sms = 148;
  const unsigned grid = static_cast<unsigned>(B > 8LL * sms ? 8LL * sms : B);
  auto st = static_cast<cudaStream_t>(stream);
#define CPLXK_OUTERB(T, C)                                                                              \
  outer_bwd_kernel<T, C><<<grid, 256, 0, st>>>(                                                         \
      static_cast<const T*>(g_re), static_cast<const T*>(g_im), static_cast<const T*>(x1_re),           \
      static_cast<const T*>(x1_im), static_cast<const T*>(x2_re), static_cast<const T*>(x2_im),         \
      static_cast<T*>(d1_re), static_cast<T*>(d1_im), static_cast<T*>(d2_re), static_cast<T*>(d2_im),   \
      B, static_cast<int>(d1), static_cast<int>(d2), conjugate);
  if (dtype == CPLXK_F32) { if (cplx) { CPLXK_OUTERB(float, true) } else { CPLXK_OUTERB(float, false) } }
  else if (dtype == CPLXK_BF16) {
    if (cplx) { CPLXK_OUTERB(__nv_bfloat16, true) } else { CPLXK_OUTERB(__nv_bfloat16, false) }
  } else return CPLXK_ERR_BADARG;
#undef CPLXK_OUTERB
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}
