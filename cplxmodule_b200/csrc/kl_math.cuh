// Per-element KL penalties of the variational layers and the deterministic block / grid
// reduction used by kl.cu (stand-alone pass) and by the operand pre-pass of the fused forward
// (fwd_tc3.cu), which evaluates the same penalty on the weight rows it converts.
//
// Reference formulas (cplxmodule/nn/relevance/...):
//   log_alpha           real/base.py:23-26, complex/base.py:27-31 (|w| = sqrt(re^2+im^2), cplx.py:183-192)
//   REAL_VD  penalty    real/vd.py:74-76     0.5 softplus(-la) + 0.63576 sigmoid(-1.48695 la - 1.87320)
//   REAL_ARD penalty    real/ard.py:39       0.5 softplus(-la)
//   CPLX_VD  penalty    complex/vd.py:95-99  gamma - la - Ei(-exp(-la))      (host scipy in the reference)
//   CPLX_ARD penalty    complex/ard.py:39    softplus(-la)
#pragma once
#include "common.cuh"

namespace cplxk {

constexpr int kKlThreads = 256;
constexpr int kKlMaxBlocks = 2048;

struct KlWorkspace {
  unsigned int ticket;
  unsigned int pad;
  unsigned long long fp;     // running parameter fingerprint of the launch in flight (0 between launches)
  double partial[kKlMaxBlocks];
};

// ---- parameter fingerprint (guard of the fused KL by-product, see kl.cu: cplxk_kl_guard) ------
// wrapping 64-bit sum over the first 8 entries of every row of every plane of
// (bits ^ position tag) * odd constant: a change detector (any whole-tensor edit moves it), not a
// cryptographic hash -- two multiply-adds per entry, so the guard kernel stays a few microseconds
__device__ __forceinline__ unsigned long long fp_term(float v, unsigned int tag) {
  return static_cast<unsigned long long>(__float_as_uint(v) ^ (tag * 0x9E3779B1u)) * 0xD6E8FEB86659FD93ull;
}
__device__ __forceinline__ unsigned long long fingerprint_elem(float w_re, float w_im, bool cplx,
                                                               float ls2, int64_t row, int j) {
  const unsigned int tag = (static_cast<unsigned int>(row) * 8u + static_cast<unsigned int>(j)) * 3u;
  unsigned long long acc = fp_term(w_re, tag);
  if (cplx) acc += fp_term(w_im, tag + 1u);
  return acc + fp_term(ls2, tag + 2u);
}
// MUFU approximations without the denormal/range fix-up code of __logf/__expf/__fdividef:
// every argument in this kernel is a normal number well inside the fast range.
__device__ __forceinline__ float f_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float f_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float f_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float f_log(float x) { return 0.69314718056f * f_lg2(x); }
__device__ __forceinline__ float f_exp(float x) { return f_ex2(1.44269504089f * x); }

// log(1 + e) for e in (0, 1]: 4-term series below 0.03 (rel. err < 2e-7), fast log above
// (abs. err ~2^-22 on a result >= 0.0296).  log1pf() costs ~4x as many instructions and this
// kernel has to stay under ~40 instructions per element to remain HBM bound.
__device__ __forceinline__ float log1p_unit(float e) {
  if (e < 0.03f) return e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.0f);
  return f_log(1.0f + e);
}

__device__ __forceinline__ float softplus_f(float x) {
  // log(1 + e^x), stable for both signs (torch switches to identity above 20: same fp32 value)
  return fmaxf(x, 0.f) + log1p_unit(f_exp(-fabsf(x)));
}

// Ein(t) = gamma + ln t + E1(t) = gamma - la - Ei(-exp(-la)),  t = exp(-la) = 1/alpha > 0.
// t <= 1: t * P8(t)  (minimax fit of the entire series sum (-1)^(k+1) t^k / (k k!),
//         rel. err 1e-9, tools/fit_ein.py) -- cancellation free, unlike the
//         reference's fp32 "gamma + n - Ei" which returns 0 for la >= 15.
// t >  1: gamma + ln t + exp(-t)/t * R44(t), Abramowitz & Stegun 5.1.56 (|eps| < 2e-8).
__device__ __forceinline__ float ein_of(float t, float n) {
  if (t <= 1.0f) {
    float p = 2.055084504e-07f;  // deg-8 fit, highest power first
    p = fmaf(p, t, -2.924913139e-06f);
    p = fmaf(p, t, 2.817395093e-05f);
    p = fmaf(p, t, -2.313831111e-04f);
    p = fmaf(p, t, 1.666633120e-03f);
    p = fmaf(p, t, -1.041666021e-02f);
    p = fmaf(p, t, 5.555555493e-02f);
    p = fmaf(p, t, -2.500000000e-01f);
    p = fmaf(p, t, 1.0f);
    return t * p;
  }
  const float kGamma = 0.57721566490153286f;
  float e1 = 0.f;
  if (t < 60.f) {
    float num = (((t + 8.5733287401f) * t + 18.0590169730f) * t + 8.6347608925f) * t + 0.2677737343f;
    float den = (((t + 9.5733223454f) * t + 25.6329561486f) * t + 21.0996530827f) * t + 3.9584969228f;
    e1 = f_exp(-t) * num * f_rcp(den * t);
  }
  return kGamma + n + e1;
}

// log_alpha = log_sigma2 - 2 log(|w| + 1e-12).  The device form folds the modulus into the
// logarithm:  2 log(|w| + 1e-12) ~= log(|w|^2 + 1e-24)  (equal at |w| = 0 and for |w| >> 1e-12;
// in between, |w| ~ 1e-9, log_alpha moves by < 2e-3 -- weights that are pruned anyway), which
// saves the square root and lets t = 1/alpha = |w|^2 exp(-log_sigma2) come without a log/exp pair.
constexpr bool kl_kind_is_cplx(int kind) { return kind >= CPLXK_KL_CPLX_VD; }
constexpr int kKlKinds = 6;

template <int kKind>
__device__ __forceinline__ float modulus2_of(float wr, float wi) {
  if constexpr (kl_kind_is_cplx(kKind)) {
    return fmaf(wr, wr, wi * wi) + 1e-24f;
  } else {
    return fmaf(wr, wr, 1e-24f);
  }
}

template <int kKind>
__device__ __forceinline__ float log_alpha_of(float wr, float wi, float ls2) {
  return ls2 - f_log(modulus2_of<kKind>(wr, wi));
}

template <int kKind>
__device__ __forceinline__ float penalty_of(float wr, float wi, float ls2) {
  const float n = f_log(modulus2_of<kKind>(wr, wi)) - ls2;  // -log_alpha
  if constexpr (kKind == CPLXK_KL_REAL_VD) {
    float z = fmaf(1.48695f, n, -1.87320f);
    float sig = f_rcp(1.0f + f_exp(-z));
    return fmaf(0.63576f, sig, 0.5f * softplus_f(n));
  } else if constexpr (kKind == CPLXK_KL_REAL_ARD) {
    return 0.5f * softplus_f(n);
  } else if constexpr (kKind == CPLXK_KL_CPLX_VD) {
    return ein_of(f_exp(n), n);
  } else if constexpr (kKind == CPLXK_KL_CPLX_VD_APPROX) {
    // extensions/complex.py:97-99
    float z = fmaf(1.36526f, n, -1.45926f);
    float sig = f_rcp(1.0f + f_exp(-z));
    return fmaf(0.57810f, sig, softplus_f(n));
  } else if constexpr (kKind == CPLXK_KL_CPLX_VD_SCALEFREE) {
    // extensions/complex.py:40-43: log|w| - ls2 - Ei(-t)/2 with t = 1/alpha; since
    // -Ei(-t) = E1(t) = Ein(t) - gamma - ln t and ln t = 2 log|w| - ls2 this is
    // (Ein(t) - gamma - ls2) / 2, again free of the cancellation at large log_alpha
    return 0.5f * (ein_of(f_exp(n), n) - 0.57721566490153286f - ls2);
  } else {
    return softplus_f(n);
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < kKlThreads / 32) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// finish of a grid-wide sum: every block deposits one double, the last one (ticket) adds them in
// index order -- deterministic, no float atomics.  `sh` holds kKlThreads / 32 doubles.
__device__ __forceinline__ void grid_sum_finish(double bsum, KlWorkspace* ws, float* out_sum,
                                                double scale, double* sh, bool* is_last,
                                                unsigned long long* fp_out = nullptr) {
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.x] = bsum;
    __threadfence();
    unsigned int t = atomicAdd(&ws->ticket, 1u);
    *is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!*is_last) return;
  __threadfence();
  double v = 0.0;
  for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += kKlThreads)
    v += *(volatile double*)&ws->partial[i];
  __syncthreads();  // sh reuse
  double total = block_sum(v, sh);
  if (threadIdx.x == 0) {
    *out_sum = static_cast<float>(total * scale);
    if (fp_out) {    // every block added its share before taking its ticket
      *fp_out = *reinterpret_cast<volatile unsigned long long*>(&ws->fp);
      ws->fp = 0ull;
    }
    ws->ticket = 0;  // restore the workspace for the next call on this stream
  }
}

// runtime-kind form (uniform branch) for kernels that are not specialised per penalty
__device__ __forceinline__ float penalty_any(int kind, float wr, float wi, float ls2) {
  switch (kind) {
    case CPLXK_KL_REAL_VD: return penalty_of<CPLXK_KL_REAL_VD>(wr, wi, ls2);
    case CPLXK_KL_REAL_ARD: return penalty_of<CPLXK_KL_REAL_ARD>(wr, wi, ls2);
    case CPLXK_KL_CPLX_VD: return penalty_of<CPLXK_KL_CPLX_VD>(wr, wi, ls2);
    case CPLXK_KL_CPLX_VD_APPROX: return penalty_of<CPLXK_KL_CPLX_VD_APPROX>(wr, wi, ls2);
    case CPLXK_KL_CPLX_VD_SCALEFREE: return penalty_of<CPLXK_KL_CPLX_VD_SCALEFREE>(wr, wi, ls2);
    default: return penalty_of<CPLXK_KL_CPLX_ARD>(wr, wi, ls2);
  }
}

}  // namespace cplxk
