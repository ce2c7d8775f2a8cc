// One row of the fp32 -> 16-bit operand conversion in front of the persistent tensor-core GEMMs
// (variational forward: fwd_tc3.cu; plain / masked affine map: fwd_lin3.cu), as a device function
// of kThreads cooperating threads.
//
// Row maximum -> power-of-two scale that puts it in [2^13, 2^14) -> fp16 planes; the row's
// inverse scale goes to isx / isw.  The same pass writes the variance-GEMM operands |x|^2 and
// exp(log_sigma2) as bf16 (unscaled: bf16 has fp32's range), multiplies a fixed-sparsity mask
// into the weights (kMask), and evaluates the layer's KL penalty on the weight row it holds in
// registers.  K % 8 == 0.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "kl_math.cuh"

namespace cplxk {

struct PrepArgs {
  const float *x_re, *x_im;            // [M, K]
  const float *w_re, *w_im, *ls2;      // [N, K]  (ls2 nullable: plain affine map)
  const float* w_mask;                 // [N, K] or nullptr
  int64_t M, N, K;
  __half *xh_re, *xh_im, *wh_re, *wh_im;
  __nv_bfloat16 *q, *e;                // nullable together
  float *isx, *isw;
  int kl_kind;                         // < 0: no KL
  int64_t kl_row0, kl_row1;            // weight rows that enter the KL sum
  int want_fp;                         // also fingerprint the parameters (kl_math.cuh; KL requests only)
  int prefetch_next;                   // prefetch the group's next row to L2 (knob CPLXK_PREP_PREFETCH)
};

// `sync()` is a barrier over the kThreads cooperating threads (a block: __syncthreads); `red`
// points at kThreads / 32 floats of shared
// memory.  kCache = 8-element groups per thread kept in registers between the two passes (the
// rest of the row is re-read, from L2).  Returns this thread's share of the row's KL penalty
// (0 for x rows / rows outside [kl_row0, kl_row1)).
template <bool kCplx, int kThreads, int kCache, bool kMask, typename Sync>
__device__ __forceinline__ float prep_convert_row(const PrepArgs& a, bool is_x, int64_t r, int tid,
                                                  float* red, Sync sync, unsigned long long& fp_acc) {
  const int64_t K = a.K;
  // (the planes never alias: without __restrict__ the loads of the write pass could not be
  // hoisted above its stores)
  const float* __restrict__ pr = (is_x ? a.x_re : a.w_re) + r * K;
  const float* __restrict__ pi = kCplx ? (is_x ? a.x_im : a.w_im) + r * K : nullptr;
  __half* __restrict__ hr = (is_x ? a.xh_re : a.wh_re) + r * K;
  __half* __restrict__ hi = kCplx ? (is_x ? a.xh_im : a.wh_im) + r * K : nullptr;
  const bool has_var = a.q != nullptr;       // plain affine map: no variance operands (fwd_lin3.cu)
  __nv_bfloat16* __restrict__ dv = has_var ? (is_x ? a.q : a.e) + r * K : nullptr;
  const float* __restrict__ pl = (is_x || !has_var) ? nullptr : a.ls2 + r * K;
  // fixed-sparsity layers (nn/masked): W enters the GEMM as W * mask, applied here where every
  // weight is read anyway (no materialised masked copy, no extra launch)
  const float* __restrict__ pm = (!kMask || is_x || a.w_mask == nullptr) ? nullptr : a.w_mask + r * K;
  float kl_acc = 0.f;

  if (pl != nullptr) {   // log_sigma2 is needed after the group-wide max: pull it towards L2 now
    for (int64_t k = static_cast<int64_t>(tid) * 32; k < K; k += kThreads * 32)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + k));
  }
  float cr[kCache][8], ci[kCache][8];
  float amax = 0.f;
#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const int64_t k = (static_cast<int64_t>(it) * kThreads + tid) * 8;
    if (k < K) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(pr + k));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(pr + k + 4));
      cr[it][0] = a0.x, cr[it][1] = a0.y, cr[it][2] = a0.z, cr[it][3] = a0.w;
      cr[it][4] = a1.x, cr[it][5] = a1.y, cr[it][6] = a1.z, cr[it][7] = a1.w;
      if constexpr (kCplx) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(pi + k));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(pi + k + 4));
        ci[it][0] = b0.x, ci[it][1] = b0.y, ci[it][2] = b0.z, ci[it][3] = b0.w;
        ci[it][4] = b1.x, ci[it][5] = b1.y, ci[it][6] = b1.z, ci[it][7] = b1.w;
      }
      if (kMask && pm != nullptr) {
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(pm + k));
        const float4 m1 = __ldg(reinterpret_cast<const float4*>(pm + k + 4));
        const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cr[it][j] *= mk[j];
          if constexpr (kCplx) ci[it][j] *= mk[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        amax = fmaxf(amax, fabsf(cr[it][j]));
        if constexpr (kCplx) amax = fmaxf(amax, fabsf(ci[it][j]));
      }
    }
  }
  for (int64_t k = (static_cast<int64_t>(kCache) * kThreads + tid) * 8; k < K; k += kThreads * 8) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 v = __ldg(reinterpret_cast<const float4*>(pr + k + 4 * h));
      float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
      if (kMask && pm != nullptr) m = __ldg(reinterpret_cast<const float4*>(pm + k + 4 * h));
      v.x *= m.x, v.y *= m.y, v.z *= m.z, v.w *= m.w;
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      if constexpr (kCplx) {
        float4 b = __ldg(reinterpret_cast<const float4*>(pi + k + 4 * h));
        b.x *= m.x, b.y *= m.y, b.z *= m.z, b.w *= m.w;
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(b.x), fabsf(b.y))), fmaxf(fabsf(b.z), fabsf(b.w)));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  sync();                                   // red[] of the previous row fully consumed
  if ((tid & 31) == 0) red[tid >> 5] = amax;
  sync();
  amax = red[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) amax = fmaxf(amax, red[w]);
  // scale = 2^s with amax * 2^s in [2^13, 2^14); s clamped so that 2^s and 2^-s are normal
  const int ex = static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 127;
  int s = 13 - ex;
  if (amax == 0.f || ex == 128) s = 0;      // empty row, or inf / nan: leave as is
  s = s > 126 ? 126 : s;
  const float scale = __uint_as_float(static_cast<uint32_t>(s + 127) << 23);
  if (tid == 0) (is_x ? a.isx : a.isw)[r] = __uint_as_float(static_cast<uint32_t>(127 - s) << 23);

  float vcarry = 0.f;                       // rounding error carried along the variance operand
  auto emit = [&](int64_t k, const float (&vr)[8], const float (&vi)[8]) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(vr[2 * j] * scale, vr[2 * j + 1] * scale);
    *reinterpret_cast<uint4*>(hr + k) = o;
    if constexpr (kCplx) {
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(vi[2 * j] * scale, vi[2 * j + 1] * scale);
      *reinterpret_cast<uint4*>(hi + k) = o;
    }
    if (!has_var) return;
    __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&o);
    float var[8];
    if (is_x) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        var[j] = vr[j] * vr[j];
        if constexpr (kCplx) var[j] = fmaf(vi[j], vi[j], var[j]);
      }
    } else {
      const float4 l0 = __ldg(reinterpret_cast<const float4*>(pl + k));
      const float4 l1 = __ldg(reinterpret_cast<const float4*>(pl + k + 4));
      const float l[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      if (a.kl_kind >= 0 && r >= a.kl_row0 && r < a.kl_row1) {   // weights and log_sigma2 are in registers anyway
#pragma unroll
        for (int j = 0; j < 8; ++j) kl_acc += penalty_any(a.kl_kind, vr[j], kCplx ? vi[j] : 0.f, l[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) var[j] = __expf(l[j]);
      if (!kMask && a.want_fp && k == 0) {   // thread 0 holds the head of the row
#pragma unroll
        for (int j = 0; j < 8; ++j) fp_acc += fingerprint_elem(vr[j], kCplx ? vi[j] : 0.f, kCplx, l[j], r, j);
      }
    }
    uint32_t pk[4];
    if (is_x) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(var[2 * j], var[2 * j + 1]);
        pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    } else {
      // bf16 keeps 8 significant bits: rounding a CONSTANT row of exp(log_sigma2) (log_sigma2 at
      // its initial value: every freshly constructed layer) is a systematic relative error of up
      // to 2^-9 in s2 -- it does not average out over K like the rounding of |x|^2 does.  The
      // rounding error of each element is therefore carried into the next one of this thread's
      // run (error diffusion, round-to-nearest): a run's sum is preserved to half an ulp of ONE
      // element.  On rows with scattered values this is as accurate as plain rounding
      // (simulated and measured: profiles/README.md), on constant rows 8x more.
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = var[j] + vcarry;
        const __nv_bfloat16 h = __float2bfloat16_rn(fmaxf(t, 0.f));
        const float c = t - __bfloat162float(h);
        vcarry = fabsf(c) <= 3.0e38f ? c : 0.f;          // inf / nan stay in their element
        const uint32_t bits = static_cast<uint32_t>(__bfloat16_as_ushort(h));
        if (j & 1) pk[j >> 1] |= bits << 16; else pk[j >> 1] = bits;
      }
    }
    o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    (void)b;
    *reinterpret_cast<uint4*>(dv + k) = o;
  };
#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const int64_t k = (static_cast<int64_t>(it) * kThreads + tid) * 8;
    if (k < K) emit(k, cr[it], ci[it]);
  }
  for (int64_t k = (static_cast<int64_t>(kCache) * kThreads + tid) * 8; k < K; k += kThreads * 8) {
    float vr[8], vi[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 v = __ldg(reinterpret_cast<const float4*>(pr + k + 4 * h));
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (kCplx) b = __ldg(reinterpret_cast<const float4*>(pi + k + 4 * h));
      if (kMask && pm != nullptr) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(pm + k + 4 * h));
        v.x *= m.x, v.y *= m.y, v.z *= m.z, v.w *= m.w;
        b.x *= m.x, b.y *= m.y, b.z *= m.z, b.w *= m.w;
      }
      vr[4 * h] = v.x, vr[4 * h + 1] = v.y, vr[4 * h + 2] = v.z, vr[4 * h + 3] = v.w;
      vi[4 * h] = b.x, vi[4 * h + 1] = b.y, vi[4 * h + 2] = b.z, vi[4 * h + 3] = b.w;
    }
    emit(k, vr, vi);
  }
  return kl_acc;
}

// Pull the row this group converts NEXT towards L2 while the current one is being processed: the
// pass is bound by the latency of its global loads (ncu: long-scoreboard stalls, 55 % of the DRAM
// peak, half the issue slots idle), and an L2 hit costs a third of a DRAM round trip.
template <bool kCplx, int kThreads>
__device__ __forceinline__ void prep_prefetch_row(const PrepArgs& a, bool is_x, int64_t r, int tid) {
  const int64_t K = a.K;
  const float* pr = (is_x ? a.x_re : a.w_re) + r * K;
  const float* pi = kCplx ? (is_x ? a.x_im : a.w_im) + r * K : nullptr;
  const float* pl = (is_x || a.q == nullptr) ? nullptr : a.ls2 + r * K;
  for (int64_t k = static_cast<int64_t>(tid) * 32; k < K; k += kThreads * 32) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + k));
    if (kCplx) asm volatile("prefetch.global.L2 [%0];" ::"l"(pi + k));
    if (pl) asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + k));
  }
}

// Sum of `v` over the kThreads cooperating threads (deterministic: fixed shuffle tree, then the
// warp sums in index order); valid in thread 0.  `sh` = kThreads / 32 doubles of shared memory.
template <int kThreads, typename Sync>
__device__ __forceinline__ double prep_group_sum(double v, int tid, double* sh, Sync sync) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  sync();
  if ((tid & 31) == 0) sh[tid >> 5] = v;
  sync();
  double r = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) r += sh[w];
  }
  return r;
}

}  // namespace cplxk
