// tcgen05 + TMA implicit-GEMM complex convolution (NCHW in / NCHW out, groups == 1).
//
//   re = x_re * U - x_im * V ,  im = x_re * V + x_im * U      (cplx.py:729-742, convnd_quick)
//   s2 = |x|^2 * exp(log_sigma2) ,  y = mu + b + eps sqrt(max(s2, 1e-8))   (complex/base.py:120-135)
//
// Like the reference's convnd_quick, the real and imaginary kernels are STACKED along the
// output-channel axis: one 128-wide B operand [U(64 rows); V(64 rows)] per k-block, so a
// 128-pixel x 64-channel complex output tile costs two N=128 MMAs per k-step
//   D1 += x_re . [U;V]^T   ->  [ x_re*U | x_re*V ]
//   D2 += x_im . [U;V]^T   ->  [ x_im*U | x_im*V ]
// and the epilogue combines  re = D1[:, :64] - D2[:, 64:],  im = D1[:, 64:] + D2[:, :64].
//
// GEMM view: M = output pixels (a Ht x Wt patch of one image, Ht*Wt = 128), N = channels,
// K = (r, s, c): for each kernel tap (r, s) and 128-byte channel chunk one 4-d TMA box
// {channels, Wt, Ht, 1} of the channels-last copy of the input lands as a K-major 128B-swizzled
// [128 pixels x 128 B] tile -- zero padding is TMA's out-of-bounds fill, stride is the tensor
// map's element stride, dilation an offset of the box origin.  The channels-last copies (and
// |x|^2, exp(log_sigma2), and the tap-major weight planes) are written once per call by
// elementwise pre-pass kernels into a caller-provided workspace.
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "knobs.cuh"
#include "noise.cuh"
#include "ptx.cuh"
#include "conv_tc.cuh"
#include "tc3_common.cuh"

namespace cplxk {

template <typename T>
__device__ __forceinline__ float round_mma_operand(float v) {
  if constexpr (std::is_same<T, float>::value) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
  } else {
    return v;
  }
}

// ------------------------------------------------------------------ pre-pass kernels
// NCHW -> channels-last (NHWC, channels padded to Cp) for x_re, x_im and (VD) |x|^2.
// 32 x 32 smem transpose: reads coalesced along W, writes coalesced along C.
template <typename T, bool kVD>
__global__ void __launch_bounds__(256)
conv_nhwc_kernel(const T* __restrict__ x_re, const T* __restrict__ x_im, T* __restrict__ o_re,
                 T* __restrict__ o_im, T* __restrict__ o_q, int C, int Cp, int H, int W) {
  __shared__ float s_re[32][33], s_im[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int w0 = blockIdx.z * 32, c0 = blockIdx.y * 32;
  const int64_t bh = blockIdx.x;  // b * H + h
  const int64_t b = bh / H, h = bh - b * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, w = w0 + tx;
    float vr = 0.f, vi = 0.f;
    if (c < C && w < W) {
      const int64_t off = ((b * C + c) * H + h) * W + w;
      vr = Elem<T>::to_f(x_re[off]);
      if (x_im) vi = Elem<T>::to_f(x_im[off]);        // real planes: x_im == nullptr
    }
    s_re[ty + 8 * i][tx] = vr;
    s_im[ty + 8 * i][tx] = vi;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int w = w0 + ty + 8 * i, c = c0 + tx;
    if (w < W && c < Cp) {
      const float vr = s_re[tx][ty + 8 * i], vi = s_im[tx][ty + 8 * i];
      const int64_t off = ((b * H + h) * W + w) * Cp + c;
      o_re[off] = Elem<T>::from_f(round_mma_operand<T>(vr));
      if (o_im) o_im[off] = Elem<T>::from_f(round_mma_operand<T>(vi));
      if constexpr (kVD) o_q[off] = Elem<T>::from_f(round_mma_operand<T>(fmaf(vr, vr, vi * vi)));
    }
  }
}

// bf16 NCHW -> channels-last with 16-byte loads (8 pixels per lane) and 128-byte channel rows:
// 64 channels x 64 pixels per block (W % 8 == 0, 16-byte aligned planes, Cp % 2 == 0)
__global__ void __launch_bounds__(256)
conv_nhwc_bf16_v8_kernel(const __nv_bfloat16* __restrict__ x_re, const __nv_bfloat16* __restrict__ x_im,
                         __nv_bfloat16* __restrict__ o_re, __nv_bfloat16* __restrict__ o_im, int C,
                         int Cp, int H, int W, bool abs2 = false) {
  __shared__ __nv_bfloat16 s_re[64][66], s_im[64][66];
  const int tid = threadIdx.x;
  const int w0 = blockIdx.z * 64, c0 = blockIdx.y * 64;
  const int64_t bh = blockIdx.x;
  const int64_t b = bh / H, h = bh - b * H;
  // 64 channel rows x 8 uint4 (8 pixels each): 512 vector loads per plane, 2 per thread
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + 256 * i;
    const int cl = idx >> 3, q = idx & 7;
    const int c = c0 + cl, w = w0 + 8 * q;
    uint4 vr = make_uint4(0u, 0u, 0u, 0u), vi = vr;
    if (c < C && w < W) {
      const int64_t off = ((b * C + c) * H + h) * W + w;
      vr = __ldg(reinterpret_cast<const uint4*>(x_re + off));
      if (x_im) vi = __ldg(reinterpret_cast<const uint4*>(x_im + off));     // real planes: x_im == nullptr
    }
    const __nv_bfloat16* pr = reinterpret_cast<const __nv_bfloat16*>(&vr);
    const __nv_bfloat16* pi = reinterpret_cast<const __nv_bfloat16*>(&vi);
    if (abs2) {      // variance operand |x|^2 = x_re^2 + x_im^2 of a complex input (o_im == nullptr)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = __bfloat162float(pr[j]), c2 = __bfloat162float(pi[j]);
        s_re[cl][8 * q + j] = __float2bfloat16_rn(fmaf(a, a, c2 * c2));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_re[cl][8 * q + j] = pr[j], s_im[cl][8 * q + j] = pi[j];
    }
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  const int c = c0 + 2 * tx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int wl = ty + 8 * i, w = w0 + wl;
    if (w < W && c < Cp) {
      const int64_t off = ((b * H + h) * W + w) * Cp + c;
      __nv_bfloat162 r2, i2;
      r2.x = s_re[2 * tx][wl], r2.y = s_re[2 * tx + 1][wl];
      i2.x = s_im[2 * tx][wl], i2.y = s_im[2 * tx + 1][wl];
      *reinterpret_cast<__nv_bfloat162*>(o_re + off) = r2;
      if (o_im) *reinterpret_cast<__nv_bfloat162*>(o_im + off) = i2;
    }
  }
}

// fp32 NCHW -> channels-last fp32 (tf32-rounded MMA operands) with 16-byte loads: 64 channels x
// 64 pixels per block, float2 stores (256-byte channel rows per warp).  x_im / o_im nullable
// (real planes).  W % 4 == 0, 16-byte aligned planes, Cp % 2 == 0.
__global__ void __launch_bounds__(256)
conv_nhwc_f32_v4_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im,
                        float* __restrict__ o_re, float* __restrict__ o_im, int C, int Cp, int H, int W,
                        bool abs2 = false) {
  __shared__ float s_re[64][65], s_im[64][65];
  const int tid = threadIdx.x;
  const int w0 = blockIdx.z * 64, c0 = blockIdx.y * 64;
  const int64_t bh = blockIdx.x;
  const int64_t b = bh / H, h = bh - b * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + 256 * i;
    const int cl = idx >> 4, q = idx & 15;
    const int c = c0 + cl, w = w0 + 4 * q;
    float4 vr = make_float4(0.f, 0.f, 0.f, 0.f), vi = vr;
    if (c < C && w < W) {     // W % 4 == 0: a float4 never straddles the row end
      const int64_t off = ((b * C + c) * H + h) * W + w;
      vr = __ldg(reinterpret_cast<const float4*>(x_re + off));
      if (x_im) vi = __ldg(reinterpret_cast<const float4*>(x_im + off));
    }
    if (abs2) {      // variance operand |x|^2 of a complex input (o_im == nullptr)
      vr.x = fmaf(vr.x, vr.x, vi.x * vi.x), vr.y = fmaf(vr.y, vr.y, vi.y * vi.y);
      vr.z = fmaf(vr.z, vr.z, vi.z * vi.z), vr.w = fmaf(vr.w, vr.w, vi.w * vi.w);
    }
    s_re[cl][4 * q] = vr.x, s_re[cl][4 * q + 1] = vr.y, s_re[cl][4 * q + 2] = vr.z, s_re[cl][4 * q + 3] = vr.w;
    s_im[cl][4 * q] = vi.x, s_im[cl][4 * q + 1] = vi.y, s_im[cl][4 * q + 2] = vi.z, s_im[cl][4 * q + 3] = vi.w;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  const int c = c0 + 2 * tx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int wl = ty + 8 * i, w = w0 + wl;
    if (w < W && c < Cp) {
      const int64_t off = ((b * H + h) * W + w) * Cp + c;
      *reinterpret_cast<float2*>(o_re + off) =
          make_float2(round_mma_operand<float>(s_re[2 * tx][wl]), round_mma_operand<float>(s_re[2 * tx + 1][wl]));
      if (o_im)
        *reinterpret_cast<float2*>(o_im + off) =
            make_float2(round_mma_operand<float>(s_im[2 * tx][wl]), round_mma_operand<float>(s_im[2 * tx + 1][wl]));
    }
  }
}

// weights [O, tCg, kh, kw] -> tap-major planes [(r*kw+s) * Op + grp * Ogp + o][Cgp], grp / o / c
// counting (super-)groups of Og output and Cg input channels; a super-group packs Og / tOg true
// groups block-diagonally (tCg x tOg blocks, zeros elsewhere).  E = exp(log_sigma2) (zero off the
// diagonal: exp is only taken of entries that exist).  Real planes: w_im == v == nullptr.
template <typename T, bool kVD>
__global__ void __launch_bounds__(256)
conv_wprep_kernel(const T* __restrict__ w_re, const T* __restrict__ w_im, const T* __restrict__ ls2,
                  T* __restrict__ u, T* __restrict__ v, T* __restrict__ e, int Og, int Ogp, int Op,
                  int Cg, int Cgp, int tOg, int tCg, int khw) {
  const int64_t total = static_cast<int64_t>(khw) * Op * Cgp;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cgp);
    const int64_t t = i / Cgp;
    const int op = static_cast<int>(t % Op);
    const int rs = static_cast<int>(t / Op);
    const int grp = op / Ogp, o = op - grp * Ogp;
    float fu = 0.f, fv = 0.f, fe = 0.f;
    if (o < Og && c < Cg && o / tOg == c / tCg) {
      const int64_t src = (static_cast<int64_t>(grp * Og + o) * tCg + (c % tCg)) * khw + rs;
      fu = Elem<T>::to_f(w_re[src]);
      if (w_im) fv = Elem<T>::to_f(w_im[src]);
      if constexpr (kVD) fe = __expf(Elem<T>::to_f(ls2[src]));
    }
    u[i] = Elem<T>::from_f(round_mma_operand<T>(fu));
    if (v) v[i] = Elem<T>::from_f(round_mma_operand<T>(fv));
    if constexpr (kVD) e[i] = Elem<T>::from_f(round_mma_operand<T>(fe));
  }
}

// Output address of (image b, channel o, pixel oh,ow).  The noise / eps planes always follow
// torch's logical NCHW element order; only y follows the activations' memory format.
__device__ __forceinline__ int64_t conv_y_offset(const ConvTcGeom& g, bool nhwc, int64_t b,
                                                 int64_t o, int64_t oh, int64_t ow) {
  return nhwc ? ((b * g.Ho + oh) * g.Wo + ow) * g.O + o : ((b * g.O + o) * g.Ho + oh) * g.Wo + ow;
}

// kN consecutive output channels of one pixel in an NCHW plane: `q` points at the first channel,
// planes are `hw` elements apart, `left` channels remain before O.  The address is ONE pointer
// walked by the plane stride and the channel bound is tested once per run: the per-store 64-bit
// multiply + 64-bit compare + branch this replaces cost 18 instructions per store (ncu:
// profiles/prof_convpair_r1) and made the epilogue, not the MMAs, the critical path.
template <typename T, int kN>
__device__ __forceinline__ void conv_store_nchw(T* __restrict__ q, int64_t hw, int left,
                                                const float (&v)[kN]) {
  if (left >= kN) {
#pragma unroll
    for (int j = 0; j < kN; ++j, q += hw) *q = Elem<T>::from_f(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < kN; ++j, q += hw)
      if (j < left) *q = Elem<T>::from_f(v[j]);
  }
}

// Eight consecutive output channels of one pixel.  NCHW: one scalar store per channel (the 32
// lanes of a warp hold 32 consecutive pixels, so each store instruction is one 128 B line).
// NHWC: the eight values are contiguous, written as one 32 B (fp32) / 16 B (bf16) vector.
template <typename T>
__device__ __forceinline__ void conv_store8(T* __restrict__ y, const ConvTcGeom& g, bool nhwc,
                                            int64_t nchw_off, int64_t hw, int64_t nhwc_off, int o0,
                                            const float (&v)[8], int o_end = -1) {
  if (o_end < 0) o_end = static_cast<int>(g.O);      // grouped: end of the group's channels
  if (!nhwc) {
    conv_store_nchw<T, 8>(y + nchw_off + static_cast<int64_t>(o0) * hw, hw, o_end - o0, v);
    return;
  }
  T* dst = y + nhwc_off + o0;
  if (o0 + 8 <= o_end && ((g.O | o0) & 7) == 0) {
    if constexpr (std::is_same<T, float>::value) {
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]),
                   "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                   : "memory");
    } else {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0), pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2), pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = pk;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (o0 + j < o_end) dst[j] = Elem<T>::from_f(v[j]);
}

// Sixteen consecutive output channels of one pixel: NHWC bf16 rows are written as ONE 32-byte
// store (a full sector; two 16-byte stores are two partial-sector writes), everything else goes
// through conv_store8.
template <typename T>
__device__ __forceinline__ void conv_store16(T* __restrict__ y, const ConvTcGeom& g, bool nhwc,
                                             int64_t nchw_off, int64_t hw, int64_t nhwc_off, int o0,
                                             const float (&v)[16]) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    T* dst = y + nhwc_off + o0;
    if (nhwc && o0 + 16 <= g.O && (g.O & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w[0]),
                   "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
      return;
    }
  }
  if (!nhwc) {
    conv_store_nchw<T, 16>(y + nchw_off + static_cast<int64_t>(o0) * hw, hw, static_cast<int>(g.O) - o0, v);
    return;
  }
  float lo[8], hi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) lo[j] = v[j], hi[j] = v[8 + j];
  conv_store8<T>(y, g, nhwc, nchw_off, hw, nhwc_off, o0, lo);
  conv_store8<T>(y, g, nhwc, nchw_off, hw, nhwc_off, o0 + 8, hi);
}

// channels-last |x|^2 for the variational forward when the input needs no transposition
template <typename T>
__global__ void __launch_bounds__(256)
conv_abs2_kernel(const T* __restrict__ x_re, const T* __restrict__ x_im, T* __restrict__ q, int64_t n) {
  constexpr int V = Elem<T>::kVec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n / V;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    Vec16<T> a, b, o;
    a.load(x_re + i * V), b.load(x_im + i * V);
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = round_mma_operand<T>(fmaf(a.v[j], a.v[j], b.v[j] * b.v[j]));
    o.store(q + i * V);
  }
}

// ---- fp32 planes on fp16 operands (NCHW inputs: a transposing pre-pass exists anyway) --------
// The im2col rows of one image overlap, so only a per-IMAGE power-of-two scale factors out of the
// GEMM (per output channel on the weight side).  Same 11-bit significand as tf32, twice the MMA
// rate, half the operand bytes.  power-of-two scale that puts `amax` into [2^13, 2^14).
__device__ __forceinline__ int f16_scale_exp(float amax) {
  const int ex = static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 127;
  int s = 13 - ex;
  if (amax == 0.f || ex == 128) s = 0;      // empty, or inf / nan: leave as is
  return s > 126 ? 126 : s;
}
__device__ __forceinline__ float pow2f(int s) { return __uint_as_float(static_cast<uint32_t>(s + 127) << 23); }

// amax[b] = max |x_re|, |x_im| over image b (non-negative floats order like their bit patterns;
// NaN bits compare above inf and end up as "leave as is")
__global__ void __launch_bounds__(256)
conv_amax_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im, int64_t per_image,
                 unsigned int* __restrict__ amax) {
  const int64_t b = blockIdx.y;
  const float* xr = x_re + b * per_image;
  const float* xi = x_im + b * per_image;
  unsigned int mbits = 0u;      // largest |value| as a bit pattern
  auto upd = [&](float v) {
    const unsigned int bits = __float_as_uint(v) & 0x7fffffffu;
    mbits = bits > mbits ? bits : mbits;
  };
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // 16-byte loads when every image starts on a 16-byte boundary
  const bool vec = (per_image % 4 == 0) && (((reinterpret_cast<uintptr_t>(x_re) | reinterpret_cast<uintptr_t>(x_im)) & 15u) == 0);
  if (vec) {
    const float4* pr = reinterpret_cast<const float4*>(xr);
    const float4* pi = reinterpret_cast<const float4*>(xi);
    for (int64_t i = tid; i < per_image / 4; i += nth) {
      const float4 a = __ldg(pr + i), c = __ldg(pi + i);
      upd(a.x), upd(a.y), upd(a.z), upd(a.w), upd(c.x), upd(c.y), upd(c.z), upd(c.w);
    }
  } else {
    for (int64_t i = tid; i < per_image; i += nth) upd(xr[i]), upd(xi[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned int other = __shfl_xor_sync(0xffffffffu, mbits, o);
    mbits = other > mbits ? other : mbits;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(amax + b, mbits);
}

// Scale exponent of one IMAGE's fp16 copy.  Images whose largest magnitude already lies in
// [2^-2, 2^15) are copied unscaled: every element above 2^-14 keeps its 11-bit significand and
// the absolute error of the ones below (<= 2^-25) is <= 2^-23 of the image's maximum.  That is what
// lets the transposing pass run BEFORE the amax is known (kMode 1 below): the common case needs
// no separate amax pass over the 2 x 4-byte planes.
__device__ __forceinline__ int f16_image_scale_exp(float amax) {
  const int ex = static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 127;
  if (amax == 0.f || (ex >= -2 && ex <= 14)) return 0;
  return f16_scale_exp(amax);
}

// block-wide max of |value| bit patterns -> atomicMax(amax + b) (skipped when it cannot raise it)
__device__ __forceinline__ void conv_block_amax(unsigned int mbits, unsigned int* red, unsigned int* amax_b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned int other = __shfl_xor_sync(0xffffffffu, mbits, o);
    mbits = other > mbits ? other : mbits;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mbits;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int m = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = red[k] > m ? red[k] : m;
    if (m > *reinterpret_cast<volatile unsigned int*>(amax_b)) atomicMax(amax_b, m);
  }
}

// NCHW fp32 -> channels-last fp16 (channels padded to Cp), scaled by the image's power of two.
// One block = 64 channels x 32 pixels of `rows` image rows: reads are 128-byte rows along W, writes
// are 128-byte rows along C (each lane two channels as one half2).
//   kMode 0: amax[] is known (conv_amax_kernel ran): scale by 2^f16_image_scale_exp
//   kMode 1: optimistic single pass: copy UNSCALED and collect amax[] on the way
//   kMode 2: fix-up after a kMode-1 pass: images whose amax asks for a scale are converted again
//            (every block of the others returns at once)
template <int kMode>
__global__ void __launch_bounds__(256, kMode == 2 ? 1 : 6)
conv_nhwc_f16_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im,
                     __half* __restrict__ o_re, __half* __restrict__ o_im,
                     unsigned int* __restrict__ amax, int C, int Cp, int H, int W, int rows) {
  __shared__ float s_re[64][33], s_im[64][33];
  __shared__ unsigned int red[8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int w0 = blockIdx.z * 32, c0 = blockIdx.y * 64;
  // only the fix-up blocks (kMode 2) walk several rows
  const int chunks = kMode == 2 ? (H + rows - 1) / rows : H;
  const int64_t b = blockIdx.x / chunks;
  const int h_lo = static_cast<int>(blockIdx.x - b * chunks) * (kMode == 2 ? rows : 1);
  const int h_hi = kMode == 2 ? (h_lo + rows < H ? h_lo + rows : H) : h_lo + 1;
  float scale = 1.f;
  if constexpr (kMode != 1) {
    const int s = f16_image_scale_exp(__uint_as_float(amax[b]));
    if (kMode == 2 && s == 0) return;
    scale = pow2f(s);
  }
  unsigned int mbits = 0u;
  for (int h = h_lo; h < h_hi; ++h) {
    if (h != h_lo) __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + ty + 8 * i, w = w0 + tx;
      float vr = 0.f, vi = 0.f;
      if (c < C && w < W) {
        const int64_t off = ((b * C + c) * H + h) * W + w;
        vr = x_re[off], vi = x_im[off];
      }
      if constexpr (kMode == 1) {
        const unsigned int a = __float_as_uint(vr) & 0x7fffffffu, d = __float_as_uint(vi) & 0x7fffffffu;
        mbits = a > mbits ? a : mbits;
        mbits = d > mbits ? d : mbits;
      }
      s_re[ty + 8 * i][tx] = vr, s_im[ty + 8 * i][tx] = vi;
    }
    __syncthreads();
    const int c = c0 + 2 * tx;          // Cp is a multiple of 16: channel pairs never straddle it
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int wl = ty + 8 * i, w = w0 + wl;
      if (w < W && c < Cp) {
        const int64_t off = ((b * H + h) * W + w) * Cp + c;
        *reinterpret_cast<__half2*>(o_re + off) =
            __floats2half2_rn(s_re[2 * tx][wl] * scale, s_re[2 * tx + 1][wl] * scale);
        *reinterpret_cast<__half2*>(o_im + off) =
            __floats2half2_rn(s_im[2 * tx][wl] * scale, s_im[2 * tx + 1][wl] * scale);
      }
    }
  }
  if constexpr (kMode == 1) conv_block_amax(mbits, red, amax + b);
}

// Same with 16-byte loads: 64 channels x 64 pixels per block (W % 4 == 0, 16-byte aligned planes)
template <int kMode>
__global__ void __launch_bounds__(256, kMode == 2 ? 1 : 6)
conv_nhwc_f16_v4_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im,
                        __half* __restrict__ o_re, __half* __restrict__ o_im,
                        unsigned int* __restrict__ amax, int C, int Cp, int H, int W, int rows) {
  __shared__ float s_re[64][65], s_im[64][65];
  __shared__ unsigned int red[8];
  const int tid = threadIdx.x;
  const int w0 = blockIdx.z * 64, c0 = blockIdx.y * 64;
  // only the fix-up blocks (kMode 2) walk several rows
  const int chunks = kMode == 2 ? (H + rows - 1) / rows : H;
  const int64_t b = blockIdx.x / chunks;
  const int h_lo = static_cast<int>(blockIdx.x - b * chunks) * (kMode == 2 ? rows : 1);
  const int h_hi = kMode == 2 ? (h_lo + rows < H ? h_lo + rows : H) : h_lo + 1;
  float scale = 1.f;
  if constexpr (kMode != 1) {
    const int s = f16_image_scale_exp(__uint_as_float(amax[b]));
    if (kMode == 2 && s == 0) return;
    scale = pow2f(s);
  }
  unsigned int mbits = 0u;
  auto upd = [&](float v) {
    const unsigned int bits = __float_as_uint(v) & 0x7fffffffu;
    mbits = bits > mbits ? bits : mbits;
  };
  for (int h = h_lo; h < h_hi; ++h) {
    if (h != h_lo) __syncthreads();
    // 64 channel rows x 16 float4: 1024 vector loads per plane, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      const int cl = idx >> 4, q = idx & 15;
      const int c = c0 + cl, w = w0 + 4 * q;
      float4 vr = make_float4(0.f, 0.f, 0.f, 0.f), vi = vr;
      if (c < C && w < W) {     // W % 4 == 0: a float4 never straddles the row end
        const int64_t off = ((b * C + c) * H + h) * W + w;
        vr = __ldg(reinterpret_cast<const float4*>(x_re + off));
        vi = __ldg(reinterpret_cast<const float4*>(x_im + off));
      }
      if constexpr (kMode == 1) {
        upd(vr.x), upd(vr.y), upd(vr.z), upd(vr.w), upd(vi.x), upd(vi.y), upd(vi.z), upd(vi.w);
      }
      s_re[cl][4 * q] = vr.x, s_re[cl][4 * q + 1] = vr.y, s_re[cl][4 * q + 2] = vr.z, s_re[cl][4 * q + 3] = vr.w;
      s_im[cl][4 * q] = vi.x, s_im[cl][4 * q + 1] = vi.y, s_im[cl][4 * q + 2] = vi.z, s_im[cl][4 * q + 3] = vi.w;
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;
    const int c = c0 + 2 * tx;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int wl = ty + 8 * i, w = w0 + wl;
      if (w < W && c < Cp) {
        const int64_t off = ((b * H + h) * W + w) * Cp + c;
        *reinterpret_cast<__half2*>(o_re + off) =
            __floats2half2_rn(s_re[2 * tx][wl] * scale, s_re[2 * tx + 1][wl] * scale);
        *reinterpret_cast<__half2*>(o_im + off) =
            __floats2half2_rn(s_im[2 * tx][wl] * scale, s_im[2 * tx + 1][wl] * scale);
      }
    }
  }
  if constexpr (kMode == 1) conv_block_amax(mbits, red, amax + b);
}

// weights [O, tCg, kh, kw] fp32 -> tap-major fp16 planes [(r*kw+s) * Op + o][Cp], one block per
// output channel: its own power-of-two scale, inverse to isw[o].  tCg < C: grouped layer packed
// block-diagonally (output channel o reads input channels [o / tOg * tCg, + tCg)).
__global__ void __launch_bounds__(256)
conv_wprep_f16_kernel(const float* __restrict__ w_re, const float* __restrict__ w_im,
                      __half* __restrict__ u, __half* __restrict__ v, float* __restrict__ isw, int O,
                      int Op, int C, int Cp, int tOg, int tCg, int khw) {
  __shared__ unsigned int red[8];
  const int o = blockIdx.x;
  const int n = tCg * khw;
  const int c_lo = (o / tOg) * tCg;
  unsigned int m = 0u;
  if (o < O) {
    for (int i = threadIdx.x; i < n; i += 256) {
      const unsigned int a = __float_as_uint(w_re[static_cast<int64_t>(o) * n + i]) & 0x7fffffffu;
      const unsigned int b = __float_as_uint(w_im[static_cast<int64_t>(o) * n + i]) & 0x7fffffffu;
      m = a > m ? a : m;
      m = b > m ? b : m;
    }
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    const unsigned int other = __shfl_xor_sync(0xffffffffu, m, k);
    m = other > m ? other : m;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) m = red[k] > m ? red[k] : m;
  const int s = f16_scale_exp(__uint_as_float(m));
  const float scale = pow2f(s);
  if (threadIdx.x == 0) isw[o] = pow2f(-s);
  for (int i = threadIdx.x; i < khw * Cp; i += 256) {
    const int rs = i / Cp, c = i - rs * Cp;
    float fu = 0.f, fv = 0.f;
    if (o < O && c >= c_lo && c < c_lo + tCg) {
      const int64_t src = (static_cast<int64_t>(o) * tCg + (c - c_lo)) * khw + rs;
      fu = w_re[src] * scale, fv = w_im[src] * scale;
    }
    const int64_t dst = (static_cast<int64_t>(rs) * Op + o) * Cp + c;
    u[dst] = __float2half_rn(fu), v[dst] = __float2half_rn(fv);
  }
}

// ------------------------------------------------------------------------ main kernel
// kReal: real planes.  One A tile (x) per k-block, the 128-row B tile holds 128 REAL output
// channels (two 64-row boxes of the one weight plane), one accumulator (+ one for the variance):
// the tensor pipe does exactly the multiplies a real convolution has.
template <typename T, bool kVD, bool kReal = false>
struct ConvCfg {
  static constexpr bool kBF16 = std::is_same<T, __nv_bfloat16>::value;
  static constexpr int BM = 128, BNO = kReal ? 128 : 64;   // pixels x output channels (complex: stacked [U;V])
  static constexpr int BKC = 128 / static_cast<int>(sizeof(T));  // channels per k-block
  static constexpr int KSTEPS = 4;
  static constexpr int A_TILE = 128 * 128;                 // 16 KB
  static constexpr int OFF_XR = 0, OFF_XI = A_TILE, OFF_Q = (kReal ? 1 : 2) * A_TILE;
  static constexpr int OFF_UV = ((kReal ? 1 : 2) + (kVD ? 1 : 0)) * A_TILE;   // 128 rows = 16 KB
  static constexpr int OFF_E = OFF_UV + A_TILE;            // BNO rows
  static constexpr int STAGE_BYTES = OFF_E + (kVD ? BNO * 128 : 0);
  static constexpr int STAGES = kVD ? 3 : 2;               // VD: 1 CTA/SM, plain: 2 (complex) / 3 (real) CTAs/SM
  // complex: D1 128 | D2 128 | (s2 64);  real: D1 128 | (s2 128)
  static constexpr int TMEM_COLS = kReal ? (kVD ? 256 : 128) : (kVD ? 512 : 256);
  static constexpr int OFF_S2 = kReal ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 1024;
  // VD: sixteen epilogue warps (four per TMEM lane quarter).  The kernel is
  // bound by the torch-exact noise (one Philox-10 + Box-Muller per normal, branchy libm inside),
  // and only more resident warps hide its fixed-latency dependency stalls.
  static constexpr int THREADS = kVD ? 576 : 192;
  static constexpr int CH_PER_WARP = kVD ? BNO / 4 : BNO;
};

template <typename T, bool kVD, bool kReal = false>
__global__ void __launch_bounds__((kVD ? 576 : 192), (kVD ? 1 : (kReal ? 3 : 2)))
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_xr, const __grid_constant__ CUtensorMap tm_xi,
               const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_u,
               const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_e,
               const ConvTcGeom g, const ConvTcEpi ep) {
  using C = ConvCfg<T, kVD, kReal>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t aux = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_full = aux, bar_empty = aux + 8 * C::STAGES, bar_accum = aux + 16 * C::STAGES;
  const uint32_t tmem_slot = bar_accum + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 16 * C::STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile -> (n block [group, block inside the group], w block, h block, image)
  int t = blockIdx.x;
  const int n_blk = t % g.tiles_n;
  t /= g.tiles_n;
  const int w_blk = t % g.tiles_w;
  t /= g.tiles_w;
  const int h_blk = t % g.tiles_h;
  const int b = t / g.tiles_h;
  const int grp = n_blk / g.tiles_ng;
  const int ow0 = w_blk * g.Wt, oh0 = h_blk * g.Ht;
  const int n0 = (n_blk - grp * g.tiles_ng) * C::BNO;   // first output channel INSIDE the group
  const int obase = grp * g.Og;                          // the group's first output channel
  const int cchunks = (g.Cgp + C::BKC - 1) / C::BKC;
  const int num_kb = g.kh * g.kw * cchunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_xr);
    ptx::prefetch_tensormap(&tm_u);
    if constexpr (!kReal) {
      ptx::prefetch_tensormap(&tm_xi);
      ptx::prefetch_tensormap(&tm_v);
    }
    if constexpr (kVD) {
      ptx::prefetch_tensormap(&tm_q);
      ptx::prefetch_tensormap(&tm_e);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Producer and MMA issuer: the WHOLE warp walks the loop and one elected lane issues -- with
  // warp-uniform control flow the addresses / descriptors stay in uniform registers (a
  // tcgen05.mma costs ~3 issue slots instead of ELECT + 4 R2UR + ... on a scheduler shared with
  // the noise warps).
  if (warp == 0) {
    {
      const bool elected = ptx::elect_one();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t fb = bar_full + 8 * s;
        const uint32_t st = base + s * C::STAGE_BYTES;
        const int rs = kb / cchunks, cc = kb - rs * cchunks;
        const int r = rs / g.kw, sx = rs - r * g.kw;
        const int32_t cw = cc * C::BKC;                 // channel inside the group (weight planes)
        const int32_t c0 = grp * g.Cg + cw;             // channel of the activation planes
        // box origin in INPUT coordinates (may be negative: zero padding = OOB fill)
        const int32_t iw = ow0 * g.sw - g.pw + sx * g.dw;
        const int32_t ih = oh0 * g.sh - g.ph + r * g.dh;
        const int32_t wrow = rs * g.Op + grp * g.Ogp + n0;
        if (elected) {
          ptx::mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
          ptx::tma_load_4d(st + C::OFF_XR, &tm_xr, fb, c0, iw, ih, b);
          if constexpr (!kReal) ptx::tma_load_4d(st + C::OFF_XI, &tm_xi, fb, c0, iw, ih, b);
          if constexpr (kVD) ptx::tma_load_4d(st + C::OFF_Q, &tm_q, fb, c0, iw, ih, b);
          ptx::tma_load_2d(st + C::OFF_UV, &tm_u, fb, cw, wrow);
          if constexpr (kReal) {
            ptx::tma_load_2d(st + C::OFF_UV + C::A_TILE / 2, &tm_u, fb, cw, wrow + 64);
            if constexpr (kVD) {
              ptx::tma_load_2d(st + C::OFF_E, &tm_e, fb, cw, wrow);
              ptx::tma_load_2d(st + C::OFF_E + C::A_TILE / 2, &tm_e, fb, cw, wrow + 64);
            }
          } else {
            ptx::tma_load_2d(st + C::OFF_UV + C::A_TILE / 2, &tm_v, fb, cw, wrow);
            if constexpr (kVD) ptx::tma_load_2d(st + C::OFF_E, &tm_e, fb, cw, wrow);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      const bool elected = ptx::elect_one();
      constexpr uint32_t idesc128 = ptx::make_idesc<C::kBF16>(128, 128, false, false);
      constexpr uint32_t idesc_s2 = ptx::make_idesc<C::kBF16>(128, C::BNO, false, false);
      const uint32_t t_d1 = tmem_base, t_d2 = tmem_base + 128, t_s2 = tmem_base + C::OFF_S2;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        const uint32_t st = base + s * C::STAGE_BYTES;
        ptx::mbar_wait(bar_full + 8 * s, ph);
        ptx::tcgen05_fence_after();
        const uint64_t a_r = ptx::make_kmajor_desc<128>(st + C::OFF_XR);
        const uint64_t a_i = ptx::make_kmajor_desc<128>(st + C::OFF_XI);
        const uint64_t a_q = ptx::make_kmajor_desc<128>(st + C::OFF_Q);
        const uint64_t b_uv = ptx::make_kmajor_desc<128>(st + C::OFF_UV);
        const uint64_t b_e = ptx::make_kmajor_desc<128>(st + C::OFF_E);
        if (elected) {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            const uint32_t off = k * 32;
            ptx::umma_ss<C::kBF16>(t_d1, ptx::desc_advance(a_r, off), ptx::desc_advance(b_uv, off), idesc128, acc);
            if constexpr (!kReal)
              ptx::umma_ss<C::kBF16>(t_d2, ptx::desc_advance(a_i, off), ptx::desc_advance(b_uv, off), idesc128, acc);
            if constexpr (kVD)
              ptx::umma_ss<C::kBF16>(t_s2, ptx::desc_advance(a_q, off), ptx::desc_advance(b_e, off), idesc_s2, acc);
          }
          ptx::umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
      }
      if (elected) ptx::umma_commit(bar_accum);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int p = quarter * 32 + lane;            // pixel of the tile = TMEM lane
    const int hh = p / g.Wt, ww = p - hh * g.Wt;
    const int64_t oh = oh0 + hh, ow = ow0 + ww;
    const bool pix_ok = oh < g.Ho && ow < g.Wo;
    const int64_t hw = g.Ho * g.Wo;
    const int64_t pix_off = static_cast<int64_t>(b) * g.O * hw + oh * g.Wo + ow;  // + o * hw
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int o_end = obase + g.Og;               // one past the group's last output channel

    // VD: the normals of this pixel for this warp's share of the tile's channels (one per output
    // channel and plane) are generated / fetched while the MMAs run and stay in registers.
    // Consecutive channels are `hw` apart in torch's linear element order, so (subsequence, slot)
    // advance by a constant -- no division.
    constexpr int NCH = C::CH_PER_WARP;            // channels this thread owns
    const int cbase = kVD ? ((warp - 2) >> 2) * NCH : 0;   // first of them inside the tile
    const int ofirst = obase + n0 + cbase;         // ... as a channel of the layer
    float nre[kVD ? NCH : 1], nim[(kVD && !kReal) ? NCH : 1];
    if constexpr (kVD) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        nre[j] = 0.f;
        if constexpr (!kReal) nim[j] = 0.f;
      }
      if (pix_ok) {
        if (ep.noise.mode == CPLXK_NOISE_INJECT) {
#pragma unroll
          for (int j = 0; j < NCH; ++j)
            if (ofirst + j < o_end) {
              const int64_t off = pix_off + static_cast<int64_t>(ofirst + j) * hw;
              nre[j] = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.eps_re) + off));
              if constexpr (!kReal) nim[j] = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.eps_im) + off));
            }
        } else if (ep.noise.mode == CPLXK_NOISE_PHILOX_TORCH) {
          const uint32_t T_ = ep.noise.threads;
          const uint64_t li0 = static_cast<uint64_t>(pix_off + static_cast<int64_t>(ofirst) * hw);
          uint64_t slot_re = li0 / T_;
          uint32_t idx_re = static_cast<uint32_t>(li0 - slot_re * T_);
          const uint64_t li1 = li0 + static_cast<uint64_t>(ep.plane_elems);
          uint64_t slot_im = li1 / T_;
          uint32_t idx_im = static_cast<uint32_t>(li1 - slot_im * T_);
          const uint64_t dq = static_cast<uint64_t>(hw) / T_;
          const uint32_t dr = static_cast<uint32_t>(static_cast<uint64_t>(hw) - dq * T_);
#pragma unroll 1
          for (int trip = 0; trip < NCH / 8; ++trip) {
#pragma unroll
            for (int j = 0; j < NCH - 8; ++j) {
              nre[j] = nre[j + 8];
              if constexpr (!kReal) nim[j] = nim[j + 8];
            }
            uint32_t ir[8], ii[8];
            uint64_t sr[8], si[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              ir[j] = idx_re, sr[j] = slot_re, ii[j] = idx_im, si[j] = slot_im;
              idx_re += dr, slot_re += dq;
              if (idx_re >= T_) idx_re -= T_, ++slot_re;
              idx_im += dr, slot_im += dq;
              if (idx_im >= T_) idx_im -= T_, ++slot_im;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              nre[NCH - 8 + j] = philox_torch_normal(ir[j], sr[j], ep.noise) * ep.noise.scale;
              if constexpr (!kReal)
                nim[NCH - 8 + j] = philox_torch_normal(ii[j], si[j], ep.noise) * ep.noise.scale;
            }
          }
        } else {
#pragma unroll 1
          for (int trip = 0; trip < NCH / 8; ++trip) {
#pragma unroll
            for (int j = 0; j < NCH - 8; ++j) {
              nre[j] = nre[j + 8];
              if constexpr (!kReal) nim[j] = nim[j + 8];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int64_t off = pix_off + static_cast<int64_t>(ofirst + trip * 8 + j) * hw;
              if constexpr (kReal) {
                // the real layout of the CUDA-core kernel (conv.cu): element -> quad, component
                const float4 a = philox_fast_normal4(static_cast<uint64_t>(off) >> 2, 0u, ep.noise);
                const int comp = static_cast<int>(off & 3);
                nre[NCH - 8 + j] = (comp == 0 ? a.x : comp == 1 ? a.y : comp == 2 ? a.z : a.w) * ep.noise.scale;
              } else {
                const float2 z = philox_fast_pair(static_cast<uint64_t>(off), ep.noise);
                nre[NCH - 8 + j] = z.x * ep.noise.scale;
                nim[NCH - 8 + j] = z.y * ep.noise.scale;
              }
            }
          }
        }
      }
    }

    ptx::mbar_wait(bar_accum, 0);
    ptx::tcgen05_fence_after();
    const int64_t cl_off = ((static_cast<int64_t>(b) * g.Ho + oh) * g.Wo + ow) * g.O;
    if constexpr (kReal && !kVD) {
      // plain real convolution: 128 channels per thread in 16-column chunks, the loads of the next
      // chunk in flight while this one is biased and stored (one-tile-per-CTA kernel: the drain
      // is on the critical path of every CTA)
      uint32_t d[2][16];
      ptx::tmem_ld_32x32b_x16(lane_base, d[0]);
#pragma unroll
      for (int q = 0; q < NCH / 16; ++q) {
        const int cur = q & 1;
        ptx::tmem_ld_wait();
        if (q + 1 < NCH / 16) ptx::tmem_ld_32x32b_x16(lane_base + 16 * (q + 1), d[cur ^ 1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int o0 = obase + n0 + q * 16 + h * 8;
          float re8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float re = __uint_as_float(d[cur][h * 8 + j]);
            if (ep.b_re && o0 + j < o_end) re += Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o0 + j));
            re8[j] = re;
          }
          if (pix_ok && o0 < o_end)
            conv_store8<T>(static_cast<T*>(ep.y_re), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, re8, o_end);
        }
      }
    } else
#pragma unroll
    for (int cc = 0; cc < NCH / 8; ++cc) {
      const int c = cbase / 8 + cc;                             // 8-channel chunk inside the tile
      const int o0 = obase + n0 + c * 8;
      if constexpr (kReal) {
        uint32_t d1[8], s2r[8];
        ptx::tmem_ld_32x32b_x8(lane_base + c * 8, d1);
        if constexpr (kVD) ptx::tmem_ld_32x32b_x8(lane_base + C::OFF_S2 + c * 8, s2r);
        ptx::tmem_ld_wait();
        float re8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float re = __uint_as_float(d1[j]);
          if (ep.b_re && o0 + j < o_end) re += Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o0 + j));
          if constexpr (kVD) {
            re = fmaf(nre[cc * 8 + j], sd_of(__uint_as_float(s2r[j])), re);
          }
          re8[j] = re;
        }
        if (pix_ok && o0 < o_end)
          conv_store8<T>(static_cast<T*>(ep.y_re), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, re8, o_end);
      } else {
        uint32_t d1a[8], d1b[8], d2a[8], d2b[8], s2r[8];
        ptx::tmem_ld_32x32b_x8(lane_base + c * 8, d1a);          // x_re * U
        ptx::tmem_ld_32x32b_x8(lane_base + 64 + c * 8, d1b);     // x_re * V
        ptx::tmem_ld_32x32b_x8(lane_base + 128 + c * 8, d2a);    // x_im * U
        ptx::tmem_ld_32x32b_x8(lane_base + 192 + c * 8, d2b);    // x_im * V
        if constexpr (kVD) ptx::tmem_ld_32x32b_x8(lane_base + C::OFF_S2 + c * 8, s2r);
        ptx::tmem_ld_wait();
        float re8[8], im8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int o = o0 + j;
          float re = __uint_as_float(d1a[j]) - __uint_as_float(d2b[j]);
          float im = __uint_as_float(d1b[j]) + __uint_as_float(d2a[j]);
          if (ep.b_re && o < o_end) {
            re += Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o));
            im += Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_im) + o));
          }
          if constexpr (kVD) {
            const float sd = sd_of(__uint_as_float(s2r[j]));
            re = fmaf(nre[cc * 8 + j], sd, re);
            im = fmaf(nim[cc * 8 + j], sd, im);
          }
          re8[j] = re, im8[j] = im;
        }
        if (pix_ok && o0 < o_end) {
          conv_store8<T>(static_cast<T*>(ep.y_re), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, re8, o_end);
          conv_store8<T>(static_cast<T*>(ep.y_im), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, im8, o_end);
        }
      }
    }
    ptx::tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------- persistent variant
// Plain (non-variational) convolution: one CTA per SM walks its share of the tiles.  The TMA
// producer and the MMA issuer run ahead across tile boundaries (the smem ring never drains)
// and the two 256-column halves of TMEM are used alternately, so the epilogue of tile i
// (TMEM -> registers -> NCHW stores) overlaps the MMAs of tile i+1.
template <typename T>
struct ConvPCfg {
  static constexpr bool kBF16 = std::is_same<T, __nv_bfloat16>::value;
  static constexpr int BNO = 64;
  static constexpr int BKC = 128 / static_cast<int>(sizeof(T));
  static constexpr int A_TILE = 128 * 128;
  static constexpr int OFF_XR = 0, OFF_XI = A_TILE, OFF_UV = 2 * A_TILE;
  static constexpr int STAGE_BYTES = 3 * A_TILE;   // 48 KB
  static constexpr int STAGES = 4;
  static constexpr int OFF_BIAS = 256;             // after the barriers: 8 warps x 64 floats
  static constexpr int THREADS = 320;              // TMA warp, MMA warp, 8 epilogue warps
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + OFF_BIAS + 8 * 64 * 4;
};

__device__ __forceinline__ void conv_tile_coords(const ConvTcGeom& g, int tile, int& b, int& oh0,
                                                 int& ow0, int& n0) {
  const int n_blk = tile % g.tiles_n;
  tile /= g.tiles_n;
  const int w_blk = tile % g.tiles_w;
  tile /= g.tiles_w;
  const int h_blk = tile % g.tiles_h;
  b = tile / g.tiles_h;
  ow0 = w_blk * g.Wt, oh0 = h_blk * g.Ht, n0 = n_blk * 64;
}

template <typename T>
__global__ void __launch_bounds__(320, 1)
conv_tc_persistent_kernel(const __grid_constant__ CUtensorMap tm_xr,
                          const __grid_constant__ CUtensorMap tm_xi,
                          const __grid_constant__ CUtensorMap tm_u,
                          const __grid_constant__ CUtensorMap tm_v, const ConvTcGeom g,
                          const ConvTcEpi ep, const int total_tiles) {
  using C = ConvPCfg<T>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t aux = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_full = aux, bar_empty = aux + 8 * C::STAGES;
  const uint32_t bar_tfull = aux + 16 * C::STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + C::STAGES * C::STAGE_BYTES + 16 * C::STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cchunks = (g.Cp + C::BKC - 1) / C::BKC;
  const int num_kb = g.kh * g.kw * cchunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_xr);
    ptx::prefetch_tensormap(&tm_xi);
    ptx::prefetch_tensormap(&tm_u);
    ptx::prefetch_tensormap(&tm_v);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_tfull + 8 * i, 1);
      ptx::mbar_init(bar_tempty + 8 * i, 8);   // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    {
      const bool elected = ptx::elect_one();
      uint32_t kbg = 0;  // k-blocks issued so far, across tiles
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b, oh0, ow0, n0;
        conv_tile_coords(g, tile, b, oh0, ow0, n0);
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const uint32_t s = kbg % C::STAGES, ph = (kbg / C::STAGES) & 1u;
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t fb = bar_full + 8 * s, st = base + s * C::STAGE_BYTES;
          const int rs = kb / cchunks, cc = kb - rs * cchunks;
          const int r = rs / g.kw, sx = rs - r * g.kw;
          const int32_t c0 = cc * C::BKC;
          const int32_t iw = ow0 * g.sw - g.pw + sx * g.dw;
          const int32_t ih = oh0 * g.sh - g.ph + r * g.dh;
          const int32_t wrow = rs * g.Op + n0;
          if (elected) {
            ptx::mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
            ptx::tma_load_4d(st + C::OFF_XR, &tm_xr, fb, c0, iw, ih, b);
            ptx::tma_load_4d(st + C::OFF_XI, &tm_xi, fb, c0, iw, ih, b);
            ptx::tma_load_2d(st + C::OFF_UV, &tm_u, fb, c0, wrow);
            ptx::tma_load_2d(st + C::OFF_UV + C::A_TILE / 2, &tm_v, fb, c0, wrow);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool elected = ptx::elect_one();
      constexpr uint32_t idesc = ptx::make_idesc<C::kBF16>(128, 128, false, false);
      uint32_t kbg = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u, tph = (it >> 1) & 1u;
        ptx::mbar_wait(bar_tempty + 8 * buf, tph ^ 1u);   // epilogue drained this half
        ptx::tcgen05_fence_after();
        const uint32_t t_d1 = tmem_base + buf * 256, t_d2 = t_d1 + 128;
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const uint32_t s = kbg % C::STAGES, ph = (kbg / C::STAGES) & 1u;
          const uint32_t st = base + s * C::STAGE_BYTES;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tcgen05_fence_after();
          const uint64_t a_r = ptx::make_kmajor_desc<128>(st + C::OFF_XR);
          const uint64_t a_i = ptx::make_kmajor_desc<128>(st + C::OFF_XI);
          const uint64_t b_uv = ptx::make_kmajor_desc<128>(st + C::OFF_UV);
          if (elected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              const uint32_t off = k * 32;
              ptx::umma_ss<C::kBF16>(t_d1, ptx::desc_advance(a_r, off), ptx::desc_advance(b_uv, off), idesc, acc);
              ptx::umma_ss<C::kBF16>(t_d2, ptx::desc_advance(a_i, off), ptx::desc_advance(b_uv, off), idesc, acc);
            }
            ptx::umma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
        }
        if (elected) ptx::umma_commit(bar_tfull + 8 * buf);
        __syncwarp();
      }
    }
  } else {
    // 8 epilogue warps: TMEM lane quarter = warp % 4, channel half = (warp - 2) / 4
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int p = quarter * 32 + lane;
    const int hh = p / g.Wt, ww = p - hh * g.Wt;
    const int64_t hw = g.Ho * g.Wo;
    // this warp's 32 + 32 bias values, staged once per n-block and read back as broadcasts
    float* sbias = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + C::OFF_BIAS) +
                   (warp - 2) * 64;
    int bias_n0 = -1;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int b, oh0, ow0, n0;
      conv_tile_coords(g, tile, b, oh0, ow0, n0);
      if (n0 != bias_n0) {
        __syncwarp();
        const int o = n0 + half * 32 + lane;
        float br = 0.f, bi = 0.f;
        if (ep.b_re && o < g.O) {
          br = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o));
          bi = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_im) + o));
        }
        sbias[lane] = br, sbias[32 + lane] = bi;
        __syncwarp();
        bias_n0 = n0;
      }
      const uint32_t buf = it & 1u, tph = (it >> 1) & 1u;
      const int64_t oh = oh0 + hh, ow = ow0 + ww;
      const bool pix_ok = oh < g.Ho && ow < g.Wo;
      const int64_t pix_off = static_cast<int64_t>(b) * g.O * hw + oh * g.Wo + ow;
      const int64_t cl_off = ((static_cast<int64_t>(b) * g.Ho + oh) * g.Wo + ow) * g.O;
      ptx::mbar_wait(bar_tfull + 8 * buf, tph);
      ptx::tcgen05_fence_after();
      const uint32_t lane_base = tmem_base + buf * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = half * 32 + c * 16;
        uint32_t d1a[16], d1b[16], d2a[16], d2b[16];
        ptx::tmem_ld_32x32b_x16(lane_base + col, d1a);          // x_re * U
        ptx::tmem_ld_32x32b_x16(lane_base + 64 + col, d1b);     // x_re * V
        ptx::tmem_ld_32x32b_x16(lane_base + 128 + col, d2a);    // x_im * U
        ptx::tmem_ld_32x32b_x16(lane_base + 192 + col, d2b);    // x_im * V
        ptx::tmem_ld_wait();
        float re16[16], im16[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          re16[k] = __uint_as_float(d1a[k]) - __uint_as_float(d2b[k]) + sbias[c * 16 + k];
          im16[k] = __uint_as_float(d1b[k]) + __uint_as_float(d2a[k]) + sbias[32 + c * 16 + k];
        }
        if (pix_ok) {
          const int o0 = n0 + col;
          conv_store16<T>(static_cast<T*>(ep.y_re), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, re16);
          conv_store16<T>(static_cast<T*>(ep.y_im), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, im16);
        }
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);   // this half may be overwritten
    }
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// CTA-pair (cta_group::2) version of the persistent conv kernel: the two CTAs of a cluster take
// two pixel tiles of the same n-block; the stacked [U;V] B operand is split along N across the pair
// (U in the leader's shared memory, V in the peer's), so every M = 256 MMA reads 4 KB of A and
// 2 KB of B per CTA instead of 4 KB + 4 KB -- the single-CTA kernel runs at the 128 B/clk
// shared-memory read limit.  Stages shrink to 40 KB (5 stages).
// kRow (wide images: one output row of 128 pixels per tile, unit stride along W): a k-block is one
// kernel ROW and channel chunk -- the 128 + (kw-1)*dw input pixels are loaded ONCE and the kw taps
// read them through shared-memory descriptors shifted by whole 128-byte pixel rows, so the
// activation traffic L2 -> SM drops kw-fold
// (at 40 KB per 512 MMA cycles the per-tap version sits on the ~64 B/clk per-SM L2 ingest limit).
// kReal: real planes (round 2): one activation plane, the pair's B operand is 128 REAL output
// channels (rows [n0, n0+64) of the one weight plane in the leader, [n0+64, n0+128) in the peer),
// one accumulator of 128 columns per TMEM half, one MMA per k-step.
template <typename T, bool kHalf = false, bool kRow = false, bool kReal = false>
struct ConvPairCfg {
  static_assert(!kHalf || std::is_same<T, float>::value, "fp16 operand copies: fp32 planes");
  static_assert(!(kHalf && kReal), "real planes: tf32 / bf16 operands");
  static constexpr bool kBF16 = std::is_same<T, __nv_bfloat16>::value || kHalf;   // kind::f16
  static constexpr int BKC = kHalf ? 64 : 128 / static_cast<int>(sizeof(T));
  static constexpr int MAX_TAPS = kRow ? 3 : 1;           // taps served by one k-block
  static constexpr int MAX_HALO = 8;                      // (kw - 1) * dw pixels beyond the 128
  static constexpr int A_TILE = (kRow ? 128 + MAX_HALO : 128) * 128;
  static constexpr int B_TILE = 64 * 128;                 // this CTA's half of one tap's [U;V]
  static constexpr int PLANES = kReal ? 1 : 2;
  static constexpr int OFF_XR = 0, OFF_XI = A_TILE, OFF_UV = PLANES * A_TILE;
  static constexpr int STAGE_BYTES = PLANES * A_TILE + MAX_TAPS * B_TILE;   // 40 KB / 58 KB (real: 24 / 41)
  static constexpr int STAGES = kRow ? (kReal ? 4 : 3) : 5;
  static constexpr int BN = kReal ? 128 : 64;             // output channels per n-block
  static constexpr int OFF_BIAS = 256;
  static constexpr int THREADS = 320;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + OFF_BIAS + 8 * 96 * 4;
};

template <typename T, bool kHalf = false, bool kRow = false, bool kReal = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tm_xr,
                          const __grid_constant__ CUtensorMap tm_xi,
                          const __grid_constant__ CUtensorMap tm_u,
                          const __grid_constant__ CUtensorMap tm_v, const ConvTcGeom g,
                          const ConvTcEpi ep, const int total_tiles) {
  using C = ConvPairCfg<T, kHalf, kRow, kReal>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t aux = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_full = aux, bar_empty = aux + 8 * C::STAGES;
  const uint32_t bar_tfull = aux + 16 * C::STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + C::STAGES * C::STAGE_BYTES + 16 * C::STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  // work item = (pair of pixel tiles, n-block); this CTA's pixel tile is 2 * pair + rank
  const int pix_tiles = total_tiles / g.tiles_n;
  const int items = ((pix_tiles + 1) / 2) * g.tiles_n;
  auto item_tile = [&](int item) { return ((item / g.tiles_n) * 2 + static_cast<int>(rank)) * g.tiles_n + item % g.tiles_n; };
  const int cchunks = (g.Cp + C::BKC - 1) / C::BKC;
  const int num_kb = (kRow ? g.kh : g.kh * g.kw) * cchunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_xr);
    ptx::prefetch_tensormap(&tm_xi);
    ptx::prefetch_tensormap(&tm_u);
    ptx::prefetch_tensormap(&tm_v);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_tfull + 8 * i, 1);
      ptx::mbar_init(bar_tempty + 8 * i, 16);  // one arrive per epilogue warp of BOTH CTAs (leader's is used)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    {
      const bool elected = ptx::elect_one();
      uint32_t kbg = 0;  // k-blocks issued so far, across tiles
      for (int item = cluster_id; item < items; item += num_clusters) {
        int b, oh0, ow0, n0;
        conv_tile_coords(g, item_tile(item), b, oh0, ow0, n0);   // b >= B for the odd tile out: OOB zero fill
        if constexpr (kReal) n0 = 2 * n0 + (leader ? 0 : 64);   // 128-channel n-blocks, this CTA's 64 rows
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const uint32_t s = kbg % C::STAGES, ph = (kbg / C::STAGES) & 1u;
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t fb = bar_full + 8 * s, st = base + s * C::STAGE_BYTES;
          if constexpr (kRow) {
            // one kernel row r, one channel chunk: 128 + halo input pixels, kw weight tiles
            const int r = kb / cchunks, cc = kb - r * cchunks;
            const int32_t c0 = cc * C::BKC;
            const int32_t iw = ow0 - g.pw;                       // sw == 1
            const int32_t ih = oh0 * g.sh - g.ph + r * g.dh;
            const uint32_t tx = C::PLANES * static_cast<uint32_t>(128 + (g.kw - 1) * g.dw) * 128u +
                                static_cast<uint32_t>(g.kw) * C::B_TILE;
            if (elected) {
              if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * tx);   // both CTAs' bytes
              ptx::tma_load_4d_pair(st + C::OFF_XR, &tm_xr, fb, c0, iw, ih, b);
              if constexpr (!kReal) ptx::tma_load_4d_pair(st + C::OFF_XI, &tm_xi, fb, c0, iw, ih, b);
              for (int sx = 0; sx < g.kw; ++sx)
                ptx::tma_load_2d_pair(st + C::OFF_UV + sx * C::B_TILE, (leader || kReal) ? &tm_u : &tm_v, fb,
                                      c0, (r * g.kw + sx) * g.Op + n0);
            }
          } else {
            const int rs = kb / cchunks, cc = kb - rs * cchunks;
            const int r = rs / g.kw, sx = rs - r * g.kw;
            const int32_t c0 = cc * C::BKC;
            const int32_t iw = ow0 * g.sw - g.pw + sx * g.dw;
            const int32_t ih = oh0 * g.sh - g.ph + r * g.dh;
            const int32_t wrow = rs * g.Op + n0;
            if (elected) {
              if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * C::STAGE_BYTES);   // both CTAs' bytes
              ptx::tma_load_4d_pair(st + C::OFF_XR, &tm_xr, fb, c0, iw, ih, b);
              if constexpr (!kReal) ptx::tma_load_4d_pair(st + C::OFF_XI, &tm_xi, fb, c0, iw, ih, b);
              // B = [U(64 rows); V(64 rows)] is split along N across the pair: U here, V in the peer
              // (real planes: the two 64-row halves of the n-block's 128 weight rows)
              ptx::tma_load_2d_pair(st + C::OFF_UV, (leader || kReal) ? &tm_u : &tm_v, fb, c0, wrow);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool elected = ptx::elect_one();
      // a/b format bits: 1 = bf16, 0 = fp16 (scaled fp32 planes), 2 = tf32
      constexpr uint32_t idesc = ptx::make_idesc<C::kBF16>(256, 128, false, false) ^
                                 (kHalf ? ((1u << 7) | (1u << 10)) : 0u);
      uint32_t kbg = 0, it = 0;
      for (int item = cluster_id; item < items; item += num_clusters, ++it) {
        const uint32_t buf = it & 1u, tph = (it >> 1) & 1u;
        ptx::mbar_wait_cluster(bar_tempty + 8 * buf, tph ^ 1u);   // both CTAs drained this half
        ptx::tcgen05_fence_after();
        const uint32_t t_d1 = tmem_base + buf * 256, t_d2 = t_d1 + 128;
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const uint32_t s = kbg % C::STAGES, ph = (kbg / C::STAGES) & 1u;
          const uint32_t st = base + s * C::STAGE_BYTES;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tcgen05_fence_after();
          if constexpr (kRow) {
            for (int sx = 0; sx < g.kw; ++sx) {
              // tap sx reads pixels [sx*dw, sx*dw + 128) of the row: the start address moves by whole
              // 128-byte rows.  The tensor core applies the 128B swizzle to ABSOLUTE shared-memory
              // address bits (as TMA did when it wrote the tile), so the descriptor's base-offset
              // field [49,52) stays 0 -- measured: with (start >> 7) & 7 there the results are wrong
              // (profiles/conv_row_ab_r2.jsonl, gpurun log pytest_row_bo / pytest_row_nobo)
              const uint32_t rows = static_cast<uint32_t>(sx * g.dw);
              const uint64_t a_r = ptx::make_kmajor_desc<128>(st + C::OFF_XR + rows * 128u);
              const uint64_t a_i = ptx::make_kmajor_desc<128>(st + C::OFF_XI + rows * 128u);
              const uint64_t b_uv = ptx::make_kmajor_desc<128>(st + C::OFF_UV + sx * C::B_TILE);
              if (elected) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t acc = (kb > 0 || sx > 0 || k > 0) ? 1u : 0u;
                  const uint32_t off = k * 32;
                  ptx::umma_ss_pair<C::kBF16>(t_d1, ptx::desc_advance(a_r, off), ptx::desc_advance(b_uv, off), idesc, acc);
                  if constexpr (!kReal)
                    ptx::umma_ss_pair<C::kBF16>(t_d2, ptx::desc_advance(a_i, off), ptx::desc_advance(b_uv, off), idesc, acc);
                }
              }
            }
            if (elected) ptx::umma_commit_pair(bar_empty + 8 * s);
          } else {
            const uint64_t a_r = ptx::make_kmajor_desc<128>(st + C::OFF_XR);
            const uint64_t a_i = ptx::make_kmajor_desc<128>(st + C::OFF_XI);
            const uint64_t b_uv = ptx::make_kmajor_desc<128>(st + C::OFF_UV);
            if (elected) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                const uint32_t off = k * 32;
                ptx::umma_ss_pair<C::kBF16>(t_d1, ptx::desc_advance(a_r, off), ptx::desc_advance(b_uv, off), idesc, acc);
                if constexpr (!kReal)
                  ptx::umma_ss_pair<C::kBF16>(t_d2, ptx::desc_advance(a_i, off), ptx::desc_advance(b_uv, off), idesc, acc);
              }
              ptx::umma_commit_pair(bar_empty + 8 * s);
            }
          }
          __syncwarp();
        }
        if (elected) ptx::umma_commit_pair(bar_tfull + 8 * buf);
        __syncwarp();
      }
    }
  } else {
    // 8 epilogue warps: TMEM lane quarter = warp % 4, channel half = (warp - 2) / 4
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int p = quarter * 32 + lane;
    const int hh = p / g.Wt, ww = p - hh * g.Wt;
    const int64_t hw = g.Ho * g.Wo;
    // this warp's 32 + 32 bias values, staged once per n-block and read back as broadcasts
    float* sbias = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + C::OFF_BIAS) +
                   (warp - 2) * 96;
    int bias_n0 = -1;
    uint32_t it = 0;
    const uint32_t tempty_remote0 = ptx::mapa_u32(bar_tempty, 0);
    for (int item = cluster_id; item < items; item += num_clusters, ++it) {
      int b, oh0, ow0, n0;
      conv_tile_coords(g, item_tile(item), b, oh0, ow0, n0);
      if constexpr (kReal) n0 *= 2;                  // 128-channel n-blocks
      if (n0 != bias_n0) {
        __syncwarp();
        if constexpr (kReal) {
          // this warp's 64 channels: half * 64 + [0, 64)
          const int o = n0 + half * 64 + lane;
          float b0 = 0.f, b1 = 0.f;
          if (ep.b_re && o < g.O) b0 = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o));
          if (ep.b_re && o + 32 < g.O) b1 = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o + 32));
          sbias[lane] = b0, sbias[32 + lane] = b1;
        } else {
          const int o = n0 + half * 32 + lane;
          float br = 0.f, bi = 0.f;
          if (ep.b_re && o < g.O) {
            br = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_re) + o));
            bi = Elem<T>::to_f(__ldg(static_cast<const T*>(ep.b_im) + o));
          }
          sbias[lane] = br, sbias[32 + lane] = bi;
          if constexpr (kHalf) sbias[64 + lane] = o < g.O ? __ldg(ep.isw + o) : 1.f;
        }
        __syncwarp();
        bias_n0 = n0;
      }
      [[maybe_unused]] float isx = 1.f;     // inverse of this image's power-of-two scale
      if constexpr (kHalf) isx = b < g.B ? pow2f(-f16_image_scale_exp(__uint_as_float(__ldg(ep.amax + b)))) : 1.f;
      const uint32_t buf = it & 1u, tph = (it >> 1) & 1u;
      const int64_t oh = oh0 + hh, ow = ow0 + ww;
      const bool pix_ok = oh < g.Ho && ow < g.Wo && b < g.B;
      const int64_t pix_off = static_cast<int64_t>(b) * g.O * hw + oh * g.Wo + ow;
      const int64_t cl_off = ((static_cast<int64_t>(b) * g.Ho + oh) * g.Wo + ow) * g.O;
      ptx::mbar_wait(bar_tfull + 8 * buf, tph);
      ptx::tcgen05_fence_after();
      const uint32_t lane_base = tmem_base + buf * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      if constexpr (kReal) {
        // one accumulator: columns [0, 128) = the n-block's real output channels; the next
        // chunk's tcgen05.ld is in flight while this one is stored
        uint32_t d[2][16];
        ptx::tmem_ld_32x32b_x16(lane_base + half * 64, d[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = half * 64 + c * 16;
          ptx::tmem_ld_wait();
          if (c < 3) ptx::tmem_ld_32x32b_x16(lane_base + col + 16, d[(c + 1) & 1]);
          float v16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v16[k] = __uint_as_float(d[c & 1][k]) + sbias[c * 16 + k];
          if (pix_ok)
            conv_store16<T>(static_cast<T*>(ep.y_re), g, ep.nhwc != 0, pix_off, hw, cl_off, n0 + col, v16);
        }
      } else
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = half * 32 + c * 16;
        uint32_t d1a[16], d1b[16], d2a[16], d2b[16];
        ptx::tmem_ld_32x32b_x16(lane_base + col, d1a);          // x_re * U
        ptx::tmem_ld_32x32b_x16(lane_base + 64 + col, d1b);     // x_re * V
        ptx::tmem_ld_32x32b_x16(lane_base + 128 + col, d2a);    // x_im * U
        ptx::tmem_ld_32x32b_x16(lane_base + 192 + col, d2b);    // x_im * V
        ptx::tmem_ld_wait();
        float re16[16], im16[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if constexpr (kHalf) {
            const float sc = isx * sbias[64 + c * 16 + k];
            re16[k] = fmaf(__uint_as_float(d1a[k]) - __uint_as_float(d2b[k]), sc, sbias[c * 16 + k]);
            im16[k] = fmaf(__uint_as_float(d1b[k]) + __uint_as_float(d2a[k]), sc, sbias[32 + c * 16 + k]);
          } else {
            re16[k] = __uint_as_float(d1a[k]) - __uint_as_float(d2b[k]) + sbias[c * 16 + k];
            im16[k] = __uint_as_float(d1b[k]) + __uint_as_float(d2a[k]) + sbias[32 + c * 16 + k];
          }
        }
        if (pix_ok) {
          const int o0 = n0 + col;
          conv_store16<T>(static_cast<T*>(ep.y_re), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, re16);
          conv_store16<T>(static_cast<T*>(ep.y_im), g, ep.nhwc != 0, pix_off, hw, cl_off, o0, im16);
        }
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty_remote0 + 8 * buf);   // this half may be overwritten
    }
  }

  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------- host side
template <typename T>
static CUtensorMapDataType conv_dt() {
  return std::is_same<T, float>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
         : std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                          : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

// channels-last activation plane [B, H, W, Cp]: box {BKC channels, Wt px (stride sw), Ht rows (stride sh), 1}
// (+ `halo` pixels for the row mode of the CTA-pair kernel, which serves every tap of a kernel row from one load)
template <typename T>
static int make_act_map(CUtensorMap* out, const void* ptr, const ConvTcGeom& g, bool round_tf32 = false,
                        int halo = 0) {
  auto enc = tensor_map_encode_fn();
  if (!enc) return CPLXK_ERR_CUDA;
  const cuuint64_t es = sizeof(T);
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(g.Cp), static_cast<cuuint64_t>(g.W),
                        static_cast<cuuint64_t>(g.H), static_cast<cuuint64_t>(g.B)};
  cuuint64_t gstr[3] = {g.Cp * es, static_cast<cuuint64_t>(g.W) * g.Cp * es,
                        static_cast<cuuint64_t>(g.H) * g.W * g.Cp * es};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(128 / sizeof(T)),
                       static_cast<cuuint32_t>((g.Wt - 1) * g.sw + 1 + halo),   // halo: row mode (sw == 1)
                       static_cast<cuuint32_t>((g.Ht - 1) * g.sh + 1), 1u};
  cuuint32_t estr[4] = {1u, static_cast<cuuint32_t>(g.sw), static_cast<cuuint32_t>(g.sh), 1u};
  const CUtensorMapDataType dt = (round_tf32 && std::is_same<T, float>::value)
                                     ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : conv_dt<T>();
  CUresult r = enc(out, dt, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CPLXK_OK : CPLXK_ERR_CUDA;
}

// tap-major weight plane [khw * Op, Cgp]: box {BKC channels, 64 rows}
template <typename T>
static int make_w_map(CUtensorMap* out, const void* ptr, const ConvTcGeom& g) {
  auto enc = tensor_map_encode_fn();
  if (!enc) return CPLXK_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(g.Cgp),
                        static_cast<cuuint64_t>(g.kh) * g.kw * g.Op};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(g.Cgp) * sizeof(T)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / sizeof(T)), 64u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, conv_dt<T>(), 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CPLXK_OK : CPLXK_ERR_CUDA;
}

static inline size_t up256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static void conv_tc_plan(ConvTcGeom& g, int dtype, int groups = 1, bool real = false) {
  const int gran = dtype == CPLXK_F32 ? 8 : 16;   // pitch % 32 B, whole k-steps
  const int bn = real ? 128 : 64;                 // output channels per n-block
  g.tCg = static_cast<int>(g.C / groups), g.tOg = static_cast<int>(g.O / groups);
  if (g.tOg < 1) g.tOg = 1;
  if (g.tCg < 1) g.tCg = 1;
  // groups narrower than an n-block: pack `ng` of them (a divisor of `groups`) into a super-group
  int ng = 1;
  if (groups > 1 && g.tOg < bn)
    for (int d = 1; d <= groups && d * g.tOg <= bn; ++d)
      if (groups % d == 0) ng = d;
  groups /= ng;
  g.groups = groups;
  g.Cg = ng * g.tCg, g.Og = ng * g.tOg;
  g.Cp = static_cast<int>((g.C + gran - 1) / gran * gran);
  g.Cgp = groups == 1 ? g.Cp : (g.Cg + gran - 1) / gran * gran;
  g.Ogp = (g.Og + bn - 1) / bn * bn;
  g.Op = groups * g.Ogp;
  int wt = 128;
  while (wt > 8 && wt / 2 >= g.Wo) wt /= 2;       // smallest power of two covering Wo (>= 8)
  g.Wt = wt;
  g.Ht = 128 / wt;
  g.tiles_w = static_cast<int>((g.Wo + g.Wt - 1) / g.Wt);
  g.tiles_h = static_cast<int>((g.Ho + g.Ht - 1) / g.Ht);
  g.tiles_ng = g.Ogp / bn;
  g.tiles_n = groups * g.tiles_ng;
}

size_t conv_tc_workspace_bytes(int dtype, bool vd, int64_t B, int64_t C, int64_t H, int64_t W,
                               int64_t O, int64_t kh, int64_t kw, int groups, bool real) {
  ConvTcGeom g{};
  g.B = B, g.C = C, g.H = H, g.W = W, g.O = O, g.kh = static_cast<int>(kh), g.kw = static_cast<int>(kw);
  g.Wo = g.Ho = 128;
  conv_tc_plan(g, dtype, groups, real);
  const size_t es = dtype == CPLXK_F32 ? 4 : 2;
  const size_t act = up256(static_cast<size_t>(B) * H * W * g.Cp * es);
  const size_t wgt = up256(static_cast<size_t>(kh) * kw * g.Op * g.Cgp * es);
  const size_t planes = (real ? 1 : 2) + (vd ? 1 : 0);
  // + per-image maxima and per-output-channel scales of the fp16 operand path
  return planes * act + planes * wgt + up256(static_cast<size_t>(B) * 4) +
         up256(static_cast<size_t>(g.Op) * 4);
}

bool conv_tc_supported(int dtype, int64_t B, int64_t C, int64_t H, int64_t W, int64_t O, int64_t Ho,
                       int64_t Wo, int kh, int kw, int sh, int sw, int groups, bool real) {
  if (B < 1 || B > 0x7fffffff || H > 0x7fffffff || W > 0x7fffffff) return false;
  if (groups < 1 || C % groups || O % groups || C > 0x3fffffff || O > 0x3fffffff) return false;
  ConvTcGeom g{};
  g.C = C, g.O = O, g.Ho = Ho, g.Wo = Wo;
  conv_tc_plan(g, dtype, groups, real);
  if ((g.Wt - 1) * sw + 1 > 256 || (g.Ht - 1) * sh + 1 > 256) return false;  // TMA box limit
  // the channel coordinate of a (super-)group's first k-block must be 16-byte aligned
  if (g.groups > 1 && (g.Cg * (dtype == CPLXK_F32 ? 4 : 2)) % 16 != 0) return false;
  const int64_t tiles = B * g.tiles_h * g.tiles_w * g.tiles_n;
  return tiles > 0 && tiles <= 0x7fffffff && kh * kw <= 4096 &&
         static_cast<int64_t>(kh) * kw * g.Op <= 0x7fffffff;
}

// Row mode of the CTA-pair kernel (ConvPairCfg<.., kRow>): tiles are single output rows of 128
// pixels, unit stride along W, 2 or 3 taps per kernel row within 8 pixels
static bool conv_row_mode(const ConvTcGeom& g) {
  return knobs().conv_row && g.Ht == 1 && g.Wt == 128 && g.sw == 1 && g.kw >= 2 && g.kw <= 3 &&
         (g.kw - 1) * g.dw <= 8;
}
static int conv_row_halo(const ConvTcGeom& g) { return conv_row_mode(g) ? (g.kw - 1) * g.dw : 0; }

template <typename T, bool kHalf, bool kReal = false>
static int launch_conv_pair(const CUtensorMap& tm_xr, const CUtensorMap& tm_xi, const CUtensorMap& tm_u,
                            const CUtensorMap& tm_v, const ConvTcGeom& g, const ConvTcEpi& ep,
                            int64_t tiles, cudaStream_t st) {
  int sms = 148, rc;
  if ((rc = current_device_sm_count(&sms))) return rc;
  const int64_t items = ((tiles / g.tiles_n + 1) / 2) * g.tiles_n;
  int64_t clusters = sms / 2;
  if (clusters > items) clusters = items;
  auto go = [&](auto kern, int threads, int smem) -> int {
    CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<static_cast<unsigned>(2 * clusters), threads, smem, st>>>(tm_xr, tm_xi, tm_u, tm_v, g, ep,
                                                                    static_cast<int>(tiles));
    CPLXK_CUDA_TRY(cudaGetLastError());
    return CPLXK_OK;
  };
  if (conv_row_mode(g)) {
    using PC = ConvPairCfg<T, kHalf, true, kReal>;
    return go(conv_tc_pair_kernel<T, kHalf, true, kReal>, PC::THREADS, PC::SMEM_BYTES);
  }
  using PC = ConvPairCfg<T, kHalf, false, kReal>;
  return go(conv_tc_pair_kernel<T, kHalf, false, kReal>, PC::THREADS, PC::SMEM_BYTES);
}

// Side stream of the chunked fp32 NCHW path (one per device, HIGHEST priority: it runs the
// persistent GEMM launches, whose CTAs are then placed before the pending conversion blocks of the
// caller's stream -- which take what a GEMM CTA leaves free on an SM)
struct ConvSide {
  cudaStream_t stream = nullptr;
  cudaEvent_t join = nullptr;
  cudaEvent_t done[16] = {};
  std::mutex mu;        // one enqueue sequence at a time (the events are shared)
  bool ok = false;
};
static ConvSide* conv_side_stream() {
  static ConvSide sides[64];
  static std::mutex init_mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  ConvSide& s = sides[dev];
  std::lock_guard<std::mutex> lock(init_mu);
  if (!s.ok) {
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (auto& e : s.done)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.ok = true;
  }
  return &s;
}

// Pre-pass (HBM bound) and GEMM (tensor bound) of an NCHW convolution over CHUNKS of images: the
// pre-pass of chunk c + 1 runs on the caller's stream while the GEMM of chunk c runs on the
// device's high-priority side stream (fork / join through events; nothing synchronises with the
// host).  prepass(b0, nb, stream) / gemm(b0, nb, stream) work on images [b0, b0 + nb).  Chunks hold
// >= 2 waves of pixel-tile pairs each, at most 16 (knob CPLXK_CONV_OVERLAP, default 4); a stream
// that is being captured into a graph, or a small batch, keeps the simple serial form.
template <typename Pre, typename Gemm>
static int conv_run_chunked(const ConvTcGeom& g, cudaStream_t st, Pre&& prepass, Gemm&& gemm) {
  int sms = 148, rc;
  if ((rc = current_device_sm_count(&sms))) return rc;
  const int64_t tiles_img = static_cast<int64_t>(g.tiles_h) * g.tiles_w;
  int64_t per = (2 * static_cast<int64_t>(sms) + tiles_img - 1) / tiles_img;   // images per chunk, lower bound
  int64_t nchunk = knobs().conv_overlap > 0 ? knobs().conv_overlap : 0;
  if (nchunk > 16) nchunk = 16;
  if (nchunk * per > g.B) nchunk = g.B / per;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (nchunk >= 2 && cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
    (void)cudaGetLastError();          // cannot tell: keep the serial form
    cap = cudaStreamCaptureStatusActive;
  }
  ConvSide* side = (nchunk >= 2 && cap == cudaStreamCaptureStatusNone) ? conv_side_stream() : nullptr;
  if (!side) {
    if ((rc = prepass(0, g.B, st))) return rc;
    return gemm(0, g.B, st);
  }
  std::lock_guard<std::mutex> lock(side->mu);
  const int64_t nb = (g.B + nchunk - 1) / nchunk;
  int64_t c = 0;
#ifdef CPLXK_CONV_TRACE_BUILD      // side builds: begin / end of every chunk's pre-pass and GEMM on stderr
  static cudaEvent_t tr[4][16];
  static bool tr_ok = false;
  if (!tr_ok) {
    for (auto& row : tr)
      for (auto& e : row) cudaEventCreate(&e);
    tr_ok = true;
  }
#endif
  for (int64_t b0 = 0; b0 < g.B; ++c, b0 += nb) {
    const int64_t n = b0 + nb <= g.B ? nb : g.B - b0;
#ifdef CPLXK_CONV_TRACE_BUILD
    cudaEventRecord(tr[0][c], st);
#endif
    if ((rc = prepass(b0, n, st))) break;
#ifdef CPLXK_CONV_TRACE_BUILD
    cudaEventRecord(tr[1][c], st);
#endif
    if (cudaEventRecord(side->done[c], st) != cudaSuccess ||
        cudaStreamWaitEvent(side->stream, side->done[c], 0) != cudaSuccess) {
      rc = CPLXK_ERR_CUDA;
      break;
    }
#ifdef CPLXK_CONV_TRACE_BUILD
    cudaEventRecord(tr[2][c], side->stream);
#endif
    if ((rc = gemm(b0, n, side->stream))) break;
#ifdef CPLXK_CONV_TRACE_BUILD
    cudaEventRecord(tr[3][c], side->stream);
#endif
  }
#ifdef CPLXK_CONV_TRACE_BUILD
  if (rc == CPLXK_OK && std::getenv("CPLXK_CONV_TRACE")) {
    cudaStreamSynchronize(side->stream);
    cudaStreamSynchronize(st);
    for (int64_t i = 0; i < c; ++i) {
      float t[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tr[0][0], tr[k][i]);
      std::fprintf(stderr, "conv_trace chunk %d: prepass %.3f .. %.3f ms, gemm %.3f .. %.3f ms\n",
                   static_cast<int>(i), t[0], t[1], t[2], t[3]);
    }
  }
#endif
  // join (also on the error paths: the caller's stream never runs ahead of the side stream)
  if (cudaEventRecord(side->join, side->stream) != cudaSuccess ||
      cudaStreamWaitEvent(st, side->join, 0) != cudaSuccess)
    return CPLXK_ERR_CUDA;
  return rc;
}

// fp32 NCHW planes on fp16 operands: transposing pre-pass (optimistic per-image scale, see
// f16_image_scale_exp) -> weights with per-output-channel scales -> CTA-pair kernel on kind::f16.
// The pre-pass is HBM bound and the GEMM tensor bound, so large batches run in CHUNKS of images:
// chunk c + 1 is converted on the caller's stream while the GEMM of chunk c runs on a
// high-priority side stream (fork / join through events; nothing synchronises with the host).
static int launch_conv_f16(const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                           void* workspace, ConvTcGeom g, const ConvTcEpi& ep_in, cudaStream_t st) {
  const size_t act32 = up256(static_cast<size_t>(g.B) * g.H * g.W * g.Cp * 4);
  const size_t wgt32 = up256(static_cast<size_t>(g.kh) * g.kw * g.Op * g.Cp * 4);
  g.Cp = g.Cgp = static_cast<int>((g.C + 15) / 16 * 16);  // whole 32-byte k-steps of fp16
  const size_t act = up256(static_cast<size_t>(g.B) * g.H * g.W * g.Cp * 2);
  const size_t wgt = up256(static_cast<size_t>(g.kh) * g.kw * g.Op * g.Cp * 2);
  if (2 * act > 2 * act32 || 2 * wgt > 2 * wgt32) return CPLXK_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* a_re = reinterpret_cast<__half*>(ws);
  __half* a_im = reinterpret_cast<__half*>(ws + act);
  __half* u = reinterpret_cast<__half*>(ws + 2 * act32);
  __half* v = reinterpret_cast<__half*>(ws + 2 * act32 + wgt);
  unsigned int* amax = reinterpret_cast<unsigned int*>(ws + 2 * act32 + 2 * wgt32);
  float* isw = reinterpret_cast<float*>(ws + 2 * act32 + 2 * wgt32 + up256(static_cast<size_t>(g.B) * 4));

  if (g.B > 65535) return CPLXK_ERR_UNSUPPORTED;
  const bool v4 = (g.W % 4 == 0) &&
                  (((reinterpret_cast<uintptr_t>(x_re) | reinterpret_cast<uintptr_t>(x_im)) & 15u) == 0);
  const unsigned cy = static_cast<unsigned>((g.Cp + 63) / 64);
  const unsigned cz = static_cast<unsigned>(v4 ? (g.W + 63) / 64 : (g.W + 31) / 32);
  if (cy > 65535u || cz > 65535u || g.B * g.H > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;
  const float* xr = static_cast<const float*>(x_re);
  const float* xi = static_cast<const float*>(x_im);
  const int Ci = static_cast<int>(g.C), Hi = static_cast<int>(g.H), Wi = static_cast<int>(g.W);
  const int64_t img_in = g.C * g.H * g.W, img_a = g.H * g.W * g.Cp;
  // the conversion blocks are to share SMs with a resident GEMM CTA (chunked path below): same
  // (maximal) shared-memory carve-out as the GEMM kernel
  static const bool carve_set = [] {
    const int mx = cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(conv_nhwc_f16_v4_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(conv_nhwc_f16_v4_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(conv_nhwc_f16_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(conv_nhwc_f16_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    return true;
  }();
  (void)carve_set;
  // kMode (see conv_nhwc_f16_kernel), image rows per block, images [b0, b0 + nb), stream
  auto convert = [&](int mode, int rows, int64_t b0, int64_t nb, cudaStream_t cs) {
    const dim3 tg(static_cast<unsigned>(nb * ((Hi + rows - 1) / rows)), cy, cz);
    const float *pr = xr + b0 * img_in, *pi = xi + b0 * img_in;
    __half *or_ = a_re + b0 * img_a, *oi = a_im + b0 * img_a;
    unsigned int* am = amax + b0;
#define CPLXK_CONV_CVT(MODE)                                                                        \
    if (v4) conv_nhwc_f16_v4_kernel<MODE><<<tg, 256, 0, cs>>>(pr, pi, or_, oi, am, Ci, g.Cp, Hi, Wi, rows); \
    else conv_nhwc_f16_kernel<MODE><<<tg, 256, 0, cs>>>(pr, pi, or_, oi, am, Ci, g.Cp, Hi, Wi, rows);
    if (mode == 0) { CPLXK_CONV_CVT(0) } else if (mode == 1) { CPLXK_CONV_CVT(1) } else { CPLXK_CONV_CVT(2) }
#undef CPLXK_CONV_CVT
  };
  // the whole pre-pass of images [b0, b0 + nb) on stream cs
  auto prepass = [&](int64_t b0, int64_t nb, cudaStream_t cs) -> int {
    if (knobs().conv_amax_pass) {
      // CPLXK_CONV_AMAX_PASS=1 (A/B): per-image amax first, then ONE scaled conversion
      int64_t chunks = img_in / (4 * 256 * 8) + 1;
      if (chunks > 64) chunks = 64;
      conv_amax_kernel<<<dim3(static_cast<unsigned>(chunks), static_cast<unsigned>(nb)), 256, 0, cs>>>(
          xr + b0 * img_in, xi + b0 * img_in, img_in, amax + b0);
      CPLXK_CUDA_TRY(cudaGetLastError());
      convert(0, 1, b0, nb, cs);
    } else {
      // optimistic: convert unscaled while collecting the amax; images that do need a scale (largest
      // magnitude outside [2^-2, 2^15)) are converted again by the fix-up launch, whose blocks (an
      // eighth of an image each) return at once otherwise
      convert(1, 1, b0, nb, cs);
      CPLXK_CUDA_TRY(cudaGetLastError());
      convert(2, (Hi + 7) / 8, b0, nb, cs);
    }
    CPLXK_CUDA_TRY(cudaGetLastError());
    return CPLXK_OK;
  };
  // the GEMM of images [b0, b0 + nb) on stream gs
  const int halo = conv_row_halo(g);
  CUtensorMap tm_u, tm_v;
  int rc;
  if ((rc = make_w_map<__half>(&tm_u, u, g))) return rc;
  if ((rc = make_w_map<__half>(&tm_v, v, g))) return rc;
  const int64_t img_out = g.O * g.Ho * g.Wo;
  auto gemm = [&](int64_t b0, int64_t nb, cudaStream_t gs) -> int {
    ConvTcGeom gc = g;
    gc.B = nb;
    CUtensorMap tm_xr, tm_xi;
    int r;
    if ((r = make_act_map<__half>(&tm_xr, a_re + b0 * img_a, gc, false, halo))) return r;
    if ((r = make_act_map<__half>(&tm_xi, a_im + b0 * img_a, gc, false, halo))) return r;
    ConvTcEpi ep = ep_in;
    ep.amax = amax + b0, ep.isw = isw;
    ep.y_re = static_cast<float*>(ep_in.y_re) + b0 * img_out;
    ep.y_im = static_cast<float*>(ep_in.y_im) + b0 * img_out;
    const int64_t tiles = nb * gc.tiles_h * gc.tiles_w * gc.tiles_n;
    return launch_conv_pair<float, true>(tm_xr, tm_xi, tm_u, tm_v, gc, ep, tiles, gs);
  };

  CPLXK_CUDA_TRY(cudaMemsetAsync(amax, 0, static_cast<size_t>(g.B) * 4, st));
  conv_wprep_f16_kernel<<<static_cast<unsigned>(g.Op), 256, 0, st>>>(
      static_cast<const float*>(w_re), static_cast<const float*>(w_im), u, v, isw, static_cast<int>(g.O),
      g.Op, static_cast<int>(g.C), g.Cp, g.tOg, g.tCg, g.kh * g.kw);
  CPLXK_CUDA_TRY(cudaGetLastError());

  return conv_run_chunked(g, st, prepass, gemm);
}

template <typename T, bool kVD>
static int launch_conv_tc(const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                          const void* ls2, void* workspace, ConvTcGeom g, const ConvTcEpi& ep,
                          cudaStream_t st) {
  using C = ConvCfg<T, kVD>;
  const size_t es = sizeof(T);
  const size_t act = up256(static_cast<size_t>(g.B) * g.H * g.W * g.Cp * es);
  const size_t wgt = up256(static_cast<size_t>(g.kh) * g.kw * g.Op * g.Cp * es);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  T* a_re = reinterpret_cast<T*>(ws);
  T* a_im = reinterpret_cast<T*>(ws + act);
  T* a_q = kVD ? reinterpret_cast<T*>(ws + 2 * act) : nullptr;
  uint8_t* wbase = ws + (kVD ? 3 : 2) * act;
  T* u = reinterpret_cast<T*>(wbase);
  T* v = reinterpret_cast<T*>(wbase + wgt);
  T* e = kVD ? reinterpret_cast<T*>(wbase + 2 * wgt) : nullptr;

  const bool nhwc = ep.nhwc != 0;
  if constexpr (std::is_same<T, float>::value && !kVD) {
    // NCHW fp32: the transposing pre-pass exists anyway -- let it write per-image scaled fp16
    // (MATH_TENSOR_TF32: tf32 operands as below)
    const int64_t pix_tiles = g.B * g.tiles_h * g.tiles_w;
    if (!nhwc && ep.f16_ok && knobs().conv_pair && pix_tiles >= 2)
      return launch_conv_f16(x_re, x_im, w_re, w_im, workspace, g, ep, st);
  }
  if (nhwc) {
    // activations already channels-last: TMA reads them in place (rounding to tf32 on load)
    if (g.Cp != g.C || ((reinterpret_cast<uintptr_t>(x_re) | reinterpret_cast<uintptr_t>(x_im)) & 15u))
      return CPLXK_ERR_UNSUPPORTED;
    a_re = const_cast<T*>(static_cast<const T*>(x_re));
    a_im = const_cast<T*>(static_cast<const T*>(x_im));
    if (kVD) {
      const int64_t n = g.B * g.H * g.W * g.C;
      int sms_q = 148;
      if (current_device_sm_count(&sms_q) != CPLXK_OK) sms_q = 148;
      const int64_t want_q = n / (256 * Elem<T>::kVec) + 1;
      conv_abs2_kernel<T><<<static_cast<unsigned>(want_q > sms_q * 16 ? sms_q * 16 : want_q), 256, 0, st>>>(
          a_re, a_im, a_q, n);
      CPLXK_CUDA_TRY(cudaGetLastError());
    }
  } else {
    dim3 tg(static_cast<unsigned>(g.B * g.H), static_cast<unsigned>((g.Cp + 31) / 32),
            static_cast<unsigned>((g.W + 31) / 32));
    if (tg.y > 65535u || tg.z > 65535u || g.B * g.H > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;
    bool done = false;
    if constexpr (std::is_same<T, __nv_bfloat16>::value && !kVD) {
      if (g.W % 8 == 0 && g.Cp % 2 == 0 &&
          (((reinterpret_cast<uintptr_t>(x_re) | reinterpret_cast<uintptr_t>(x_im)) & 15u) == 0)) {
        dim3 t8(static_cast<unsigned>(g.B * g.H), static_cast<unsigned>((g.Cp + 63) / 64),
                static_cast<unsigned>((g.W + 63) / 64));
        conv_nhwc_bf16_v8_kernel<<<t8, 256, 0, st>>>(static_cast<const T*>(x_re), static_cast<const T*>(x_im),
                                                     a_re, a_im, static_cast<int>(g.C), g.Cp,
                                                     static_cast<int>(g.H), static_cast<int>(g.W));
        done = true;
      }
    }
    if (!done)
      conv_nhwc_kernel<T, kVD><<<tg, 256, 0, st>>>(static_cast<const T*>(x_re), static_cast<const T*>(x_im),
                                                   a_re, a_im, a_q, static_cast<int>(g.C), g.Cp,
                                                   static_cast<int>(g.H), static_cast<int>(g.W));
    CPLXK_CUDA_TRY(cudaGetLastError());
  }
  const int64_t wtotal = static_cast<int64_t>(g.kh) * g.kw * g.Op * g.Cp;
  conv_wprep_kernel<T, kVD><<<static_cast<unsigned>(wtotal / 256 + 1 > 1184 ? 1184 : wtotal / 256 + 1), 256, 0, st>>>(
      static_cast<const T*>(w_re), static_cast<const T*>(w_im), static_cast<const T*>(ls2), u, v, e,
      g.Og, g.Ogp, g.Op, g.Cg, g.Cgp, g.tOg, g.tCg, g.kh * g.kw);
  CPLXK_CUDA_TRY(cudaGetLastError());

  CUtensorMap tm_xr, tm_xi, tm_q, tm_u, tm_v, tm_e;
  int rc;
  const int64_t tiles = g.B * g.tiles_h * g.tiles_w * g.tiles_n;
  const bool pair = !kVD && knobs().conv_persistent && knobs().conv_pair && tiles / g.tiles_n >= 2;
  const int halo = pair ? conv_row_halo(g) : 0;
  if ((rc = make_act_map<T>(&tm_xr, a_re, g, nhwc, halo))) return rc;
  if ((rc = make_act_map<T>(&tm_xi, a_im, g, nhwc, halo))) return rc;
  if ((rc = make_w_map<T>(&tm_u, u, g))) return rc;
  if ((rc = make_w_map<T>(&tm_v, v, g))) return rc;
  tm_q = tm_xr, tm_e = tm_u;
  if (kVD) {
    if ((rc = make_act_map<T>(&tm_q, a_q, g))) return rc;
    if ((rc = make_w_map<T>(&tm_e, e, g))) return rc;
  }
  if constexpr (!kVD) {
    if (knobs().conv_persistent) {
      int sms = 148;
      if ((rc = current_device_sm_count(&sms))) return rc;
      if (pair) return launch_conv_pair<T, false>(tm_xr, tm_xi, tm_u, tm_v, g, ep, tiles, st);
      auto pk = conv_tc_persistent_kernel<T>;
      CPLXK_CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvPCfg<T>::SMEM_BYTES));
      const unsigned grid = static_cast<unsigned>(tiles < sms ? tiles : sms);
      pk<<<grid, ConvPCfg<T>::THREADS, ConvPCfg<T>::SMEM_BYTES, st>>>(tm_xr, tm_xi, tm_u, tm_v, g, ep,
                                                     static_cast<int>(tiles));
      CPLXK_CUDA_TRY(cudaGetLastError());
      return CPLXK_OK;
    }
  }
  auto kern = conv_tc_kernel<T, kVD>;
  CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  kern<<<static_cast<unsigned>(tiles), C::THREADS, C::SMEM_BYTES, st>>>(tm_xr, tm_xi, tm_q, tm_u, tm_v,
                                                                      tm_e, g, ep);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

// Real planes and / or (super-)groups > 1 after packing: conv_tc_kernel<T, kVD, kReal> (one tile
// per CTA), a group is one more factor of the n-block index.  NCHW planes only.  (Grouped complex
// layers whose groups all fit ONE n-block are a dense layer with block-diagonal weights and take
// the CTA-pair / persistent kernels.)  A k-block of a group whose channel
// count is not a multiple of the block's width also loads the first channels of the NEXT group;
// their weight rows are zero in the prepared planes, so they contribute 0 (0 * inf = nan aside).
template <typename T, bool kVD, bool kReal>
static int launch_conv_rg(const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                          const void* ls2, void* workspace, const ConvTcGeom& g, const ConvTcEpi& ep,
                          cudaStream_t st, bool abs2 = false) {
  using C = ConvCfg<T, kVD, kReal>;
  if (ep.nhwc) return CPLXK_ERR_UNSUPPORTED;
  const size_t es = sizeof(T);
  const size_t act = up256(static_cast<size_t>(g.B) * g.H * g.W * g.Cp * es);
  const size_t wgt = up256(static_cast<size_t>(g.kh) * g.kw * g.Op * g.Cgp * es);
  constexpr int NP = kReal ? 1 : 2;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  T* a_re = reinterpret_cast<T*>(ws);
  T* a_im = kReal ? nullptr : reinterpret_cast<T*>(ws + act);
  T* a_q = kVD ? reinterpret_cast<T*>(ws + NP * act) : nullptr;
  uint8_t* wbase = ws + (NP + (kVD ? 1 : 0)) * act;
  T* u = reinterpret_cast<T*>(wbase);
  T* v = kReal ? nullptr : reinterpret_cast<T*>(wbase + wgt);
  T* e = kVD ? reinterpret_cast<T*>(wbase + NP * wgt) : nullptr;

  dim3 tg(static_cast<unsigned>(g.B * g.H), static_cast<unsigned>((g.Cp + 31) / 32),
          static_cast<unsigned>((g.W + 31) / 32));
  if (tg.y > 65535u || tg.z > 65535u || g.B * g.H > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;
  if constexpr (kReal && !kVD) {
    // ungrouped real planes: the persistent CTA-pair kernel (ConvPairCfg<.., kReal>; row mode for
    // wide images) -- the one-tile-per-CTA kernel below pulls 32 KB per 256 MMA cycles through L2,
    // twice what an SM can ingest -- with the transposing pre-pass and the GEMM over image chunks
    // (conv_run_chunked)
    const int64_t pix_tiles = g.B * g.tiles_h * g.tiles_w;
    if (g.groups == 1 && knobs().conv_real_pair && knobs().conv_persistent && knobs().conv_pair &&
        pix_tiles >= 2) {
      const bool al = (reinterpret_cast<uintptr_t>(x_re) & 15u) == 0 &&
                      (!x_im || (reinterpret_cast<uintptr_t>(x_im) & 15u) == 0) && g.Cp % 2 == 0;
      const bool fast = al && (std::is_same<T, float>::value ? g.W % 4 == 0 : g.W % 8 == 0);
      if (abs2 && !fast) return CPLXK_ERR_UNSUPPORTED;    // |x|^2 is formed by the 16-byte-load transposers only
      const int64_t wtotal = static_cast<int64_t>(g.kh) * g.kw * g.Op * g.Cgp;
      conv_wprep_kernel<T, kVD><<<static_cast<unsigned>(wtotal / 256 + 1 > 1184 ? 1184 : wtotal / 256 + 1), 256, 0, st>>>(
          static_cast<const T*>(w_re), static_cast<const T*>(w_im), static_cast<const T*>(ls2), u, v, e,
          g.Og, g.Ogp, g.Op, g.Cg, g.Cgp, g.tOg, g.tCg, g.kh * g.kw);
      CPLXK_CUDA_TRY(cudaGetLastError());
      CUtensorMap tm_w;
      int rc;
      if ((rc = make_w_map<T>(&tm_w, u, g))) return rc;
      const int64_t img_in = g.C * g.H * g.W, img_a = g.H * g.W * g.Cp, img_out = g.O * g.Ho * g.Wo;
      const int Ci = static_cast<int>(g.C), Hi = static_cast<int>(g.H), Wi = static_cast<int>(g.W);
      auto prepass = [&](int64_t b0, int64_t nb, cudaStream_t cs) -> int {
        const T* pr = static_cast<const T*>(x_re) + b0 * img_in;
        const T* pi = x_im ? static_cast<const T*>(x_im) + b0 * img_in : nullptr;
        T* or_ = a_re + b0 * img_a;
        if (fast) {
          const dim3 tf(static_cast<unsigned>(nb * g.H), static_cast<unsigned>((g.Cp + 63) / 64),
                        static_cast<unsigned>((g.W + 63) / 64));
          if constexpr (std::is_same<T, float>::value)
            conv_nhwc_f32_v4_kernel<<<tf, 256, 0, cs>>>(pr, pi, or_, static_cast<float*>(nullptr), Ci, g.Cp, Hi, Wi, abs2);
          else
            conv_nhwc_bf16_v8_kernel<<<tf, 256, 0, cs>>>(pr, pi, or_, static_cast<__nv_bfloat16*>(nullptr), Ci, g.Cp, Hi, Wi, abs2);
        } else {
          const dim3 tgc(static_cast<unsigned>(nb * g.H), tg.y, tg.z);
          conv_nhwc_kernel<T, kVD><<<tgc, 256, 0, cs>>>(pr, pi, or_, static_cast<T*>(nullptr), static_cast<T*>(nullptr),
                                                        Ci, g.Cp, Hi, Wi);
        }
        CPLXK_CUDA_TRY(cudaGetLastError());
        return CPLXK_OK;
      };
      const int halo = conv_row_halo(g);
      auto gemm = [&](int64_t b0, int64_t nb, cudaStream_t gs) -> int {
        ConvTcGeom gc = g;
        gc.B = nb;
        CUtensorMap tm_x;
        int r;
        if ((r = make_act_map<T>(&tm_x, a_re + b0 * img_a, gc, false, halo))) return r;
        ConvTcEpi ec = ep;
        ec.y_re = static_cast<T*>(ep.y_re) + b0 * img_out;
        return launch_conv_pair<T, false, true>(tm_x, tm_x, tm_w, tm_w, gc, ec, nb * gc.tiles_h * gc.tiles_w * gc.tiles_n, gs);
      };
      return conv_run_chunked(g, st, prepass, gemm);
    }
  }
  bool done = false;
  if constexpr (std::is_same<T, __nv_bfloat16>::value && !kVD) {
    // 16-byte loads, 128-byte channel rows (the complex layers' fast transposer; x_im may be null)
    if (g.W % 8 == 0 && g.Cp % 2 == 0 && (reinterpret_cast<uintptr_t>(x_re) & 15u) == 0 &&
        (!x_im || (reinterpret_cast<uintptr_t>(x_im) & 15u) == 0)) {
      dim3 t8(static_cast<unsigned>(g.B * g.H), static_cast<unsigned>((g.Cp + 63) / 64),
              static_cast<unsigned>((g.W + 63) / 64));
      conv_nhwc_bf16_v8_kernel<<<t8, 256, 0, st>>>(static_cast<const T*>(x_re), static_cast<const T*>(x_im),
                                                   a_re, a_im, static_cast<int>(g.C), g.Cp,
                                                   static_cast<int>(g.H), static_cast<int>(g.W), abs2);
      done = true;
    }
  }
  if constexpr (std::is_same<T, float>::value && !kVD) {
    if (g.W % 4 == 0 && g.Cp % 2 == 0 && (reinterpret_cast<uintptr_t>(x_re) & 15u) == 0 &&
        (!x_im || (reinterpret_cast<uintptr_t>(x_im) & 15u) == 0)) {
      dim3 t4(static_cast<unsigned>(g.B * g.H), static_cast<unsigned>((g.Cp + 63) / 64),
              static_cast<unsigned>((g.W + 63) / 64));
      conv_nhwc_f32_v4_kernel<<<t4, 256, 0, st>>>(static_cast<const float*>(x_re), static_cast<const float*>(x_im),
                                                  a_re, a_im, static_cast<int>(g.C), g.Cp,
                                                  static_cast<int>(g.H), static_cast<int>(g.W), abs2);
      done = true;
    }
  }
  if (abs2 && !done) return CPLXK_ERR_UNSUPPORTED;    // |x|^2 is formed by the 16-byte-load transposers only
  if (!done)
    conv_nhwc_kernel<T, kVD><<<tg, 256, 0, st>>>(static_cast<const T*>(x_re), static_cast<const T*>(x_im),
                                                 a_re, a_im, a_q, static_cast<int>(g.C), g.Cp,
                                                 static_cast<int>(g.H), static_cast<int>(g.W));
  CPLXK_CUDA_TRY(cudaGetLastError());
  const int64_t wtotal = static_cast<int64_t>(g.kh) * g.kw * g.Op * g.Cgp;
  conv_wprep_kernel<T, kVD><<<static_cast<unsigned>(wtotal / 256 + 1 > 1184 ? 1184 : wtotal / 256 + 1), 256, 0, st>>>(
      static_cast<const T*>(w_re), static_cast<const T*>(w_im), static_cast<const T*>(ls2), u, v, e,
      g.Og, g.Ogp, g.Op, g.Cg, g.Cgp, g.tOg, g.tCg, g.kh * g.kw);
  CPLXK_CUDA_TRY(cudaGetLastError());

  CUtensorMap tm_xr, tm_xi, tm_q, tm_u, tm_v, tm_e;
  int rc;
  if ((rc = make_act_map<T>(&tm_xr, a_re, g))) return rc;
  if ((rc = make_w_map<T>(&tm_u, u, g))) return rc;
  tm_xi = tm_xr, tm_v = tm_u, tm_q = tm_xr, tm_e = tm_u;
  if (!kReal) {
    if ((rc = make_act_map<T>(&tm_xi, a_im, g))) return rc;
    if ((rc = make_w_map<T>(&tm_v, v, g))) return rc;
  }
  if (kVD) {
    if ((rc = make_act_map<T>(&tm_q, a_q, g))) return rc;
    if ((rc = make_w_map<T>(&tm_e, e, g))) return rc;
  }
  const int64_t tiles = g.B * g.tiles_h * g.tiles_w * g.tiles_n;
  auto kern = conv_tc_kernel<T, kVD, kReal>;
  CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  kern<<<static_cast<unsigned>(tiles), C::THREADS, C::SMEM_BYTES, st>>>(tm_xr, tm_xi, tm_q, tm_u, tm_v,
                                                                      tm_e, g, ep);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

int conv_tc_dispatch(int dtype, bool vd, bool nhwc, const void* x_re, const void* x_im, const void* w_re,
                     const void* w_im, const void* ls2, void* workspace, int64_t B, int64_t C,
                     int64_t H, int64_t W, int64_t O, int64_t Ho, int64_t Wo, int kh, int kw, int sh,
                     int sw, int ph, int pw, int dh, int dw, const ConvTcEpi& ep_in, cudaStream_t st,
                     int groups, bool abs2) {
  ConvTcEpi ep = ep_in;
  ep.nhwc = nhwc ? 1 : 0;
  const bool real = x_im == nullptr || abs2;      // abs2: real conv of |x_re + i x_im|^2
  if (abs2 && (vd || nhwc)) return CPLXK_ERR_UNSUPPORTED;
  ConvTcGeom g{};
  g.B = B, g.C = C, g.H = H, g.W = W, g.O = O, g.Ho = Ho, g.Wo = Wo;
  g.kh = kh, g.kw = kw, g.sh = sh, g.sw = sw, g.ph = ph, g.pw = pw, g.dh = dh, g.dw = dw;
  conv_tc_plan(g, dtype, groups, real);
#define CPLXK_RG(T)                                                                                       \
  if (real && vd) return launch_conv_rg<T, true, true>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st);   \
  if (real) return launch_conv_rg<T, false, true>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st, abs2);  \
  if (g.groups > 1 && vd) return launch_conv_rg<T, true, false>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st); \
  if (g.groups > 1) return launch_conv_rg<T, false, false>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st);
  if (dtype == CPLXK_F32) {
    CPLXK_RG(float)
    if (vd) return launch_conv_tc<float, true>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st);
    return launch_conv_tc<float, false>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st);
  }
  if (dtype == CPLXK_BF16) {
    CPLXK_RG(__nv_bfloat16)
    if (vd) return launch_conv_tc<__nv_bfloat16, true>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st);
    return launch_conv_tc<__nv_bfloat16, false>(x_re, x_im, w_re, w_im, ls2, workspace, g, ep, st);
  }
#undef CPLXK_RG
  return CPLXK_ERR_BADARG;
}

}  // namespace cplxk
