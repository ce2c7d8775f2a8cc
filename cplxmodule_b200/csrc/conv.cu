// Complex / real 2-d cross-correlation (NCHW, groups == 1, zero padding) as an
// implicit GEMM:  M = B*Ho*Wo output pixels, N = O output channels,
// K = C*kh*kw, A[m,k] gathered on the fly (im2col never materialised).
//   re = x_re * U - x_im * V ; im = x_re * V + x_im * U      (cplx.py:729-742, convnd_quick)
//   s2 = |x|^2 * exp(log_sigma2)                              (complex/base.py:125-133)
//   y  = mu + bias + eps sqrt(max(s2, 1e-8))                  (complex/base.py:135, cplx.py:796-798)
// This file holds the exact-fp32 CUDA-core version (any shape / stride /
// dilation / padding); conv1d is the H == kh == 1 case.
#include "common.cuh"
#include "noise.cuh"
#include "conv_tc.cuh"

namespace cplxk {

struct ConvGeom {
  int64_t B, C, H, W, O, Ho, Wo;
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int64_t M, K;  // B*Ho*Wo, C*kh*kw
};

struct ConvEpi {
  const void* b_re;
  const void* b_im;
  const void* eps_re;
  const void* eps_im;
  void* y_re;
  void* y_im;
  int64_t plane_elems;  // B*O*Ho*Wo
  NoiseParams noise;
};

constexpr int CB = 64;
constexpr int CK = 16;

template <typename T, bool kCplx, bool kVD>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T* __restrict__ x_re, const T* __restrict__ x_im,
                 const T* __restrict__ w_re, const T* __restrict__ w_im,
                 const T* __restrict__ ls2, ConvGeom g, ConvEpi ep) {
  __shared__ float As[kCplx ? 2 : 1][CK][CB + 1];
  __shared__ float Aq[kVD ? CK : 1][CB + 1];
  __shared__ float Bs[kCplx ? 2 : 1][CK][CB + 1];
  __shared__ float Be[kVD ? CK : 1][CB + 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx -> pixels (coalesced stores), ty -> channels
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * CB;
  const int64_t n0 = static_cast<int64_t>(blockIdx.y) * CB;

  float acc_re[4][4] = {}, acc_im[4][4] = {}, acc_s2[4][4] = {};

  // A loader: thread -> (pixel lm = tid & 63, 4 k's starting at (tid >> 6) * 4)
  const int lm = tid & 63;
  const int lka = (tid >> 6) * 4;
  const int64_t am = m0 + lm;
  int64_t ab = 0, aoh = 0, aow = 0;
  const bool am_ok = am < g.M;
  if (am_ok) {
    ab = am / (g.Ho * g.Wo);
    int64_t r = am - ab * g.Ho * g.Wo;
    aoh = r / g.Wo;
    aow = r - aoh * g.Wo;
  }
  // B loader: thread -> (channel row = tid >> 2, 4 consecutive k)
  const int lr = tid >> 2;
  const int lkb = (tid & 3) * 4;
  const int khw = g.kh * g.kw;

  for (int64_t k0 = 0; k0 < g.K; k0 += CK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + lka + j;
      float xr = 0.f, xi = 0.f;
      if (am_ok && k < g.K) {
        const int64_t c = k / khw;
        const int rs = static_cast<int>(k - c * khw);
        const int r = rs / g.kw, s = rs - r * g.kw;
        const int64_t ih = aoh * g.sh - g.ph + static_cast<int64_t>(r) * g.dh;
        const int64_t iw = aow * g.sw - g.pw + static_cast<int64_t>(s) * g.dw;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
          const int64_t off = ((ab * g.C + c) * g.H + ih) * g.W + iw;
          xr = Elem<T>::to_f(x_re[off]);
          if constexpr (kCplx) xi = Elem<T>::to_f(x_im[off]);
        }
      }
      As[0][lka + j][lm] = xr;
      if constexpr (kCplx) As[1][lka + j][lm] = xi;
      if constexpr (kVD) Aq[lka + j][lm] = fmaf(xr, xr, xi * xi);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + lkb + j;
      const int64_t n = n0 + lr;
      float wr = 0.f, wi = 0.f, e = 0.f;
      if (k < g.K && n < g.O) {
        wr = Elem<T>::to_f(w_re[n * g.K + k]);
        if constexpr (kCplx) wi = Elem<T>::to_f(w_im[n * g.K + k]);
        if constexpr (kVD) e = expf(Elem<T>::to_f(ls2[n * g.K + k]));
      }
      Bs[0][lkb + j][lr] = wr;
      if constexpr (kCplx) Bs[1][lkb + j][lr] = wi;
      if constexpr (kVD) Be[lkb + j][lr] = e;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; ++kk) {
      float ar[4], ai[4], aq[4], br[4], bi[4], be[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ar[i] = As[0][kk][tx + 16 * i];
        br[i] = Bs[0][kk][ty * 4 + i];
        if constexpr (kCplx) {
          ai[i] = As[1][kk][tx + 16 * i];
          bi[i] = Bs[1][kk][ty * 4 + i];
        }
        if constexpr (kVD) {
          aq[i] = Aq[kk][tx + 16 * i];
          be[i] = Be[kk][ty * 4 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc_re[i][j] = fmaf(ar[i], br[j], acc_re[i][j]);
          if constexpr (kCplx) {
            acc_re[i][j] = fmaf(-ai[i], bi[j], acc_re[i][j]);
            acc_im[i][j] = fmaf(ar[i], bi[j], acc_im[i][j]);
            acc_im[i][j] = fmaf(ai[i], br[j], acc_im[i][j]);
          }
          if constexpr (kVD) acc_s2[i][j] = fmaf(aq[i], be[j], acc_s2[i][j]);
        }
    }
    __syncthreads();
  }

  // ---- epilogue: NCHW scatter; for fixed (i, j) the 16 tx lanes write 16 consecutive pixels
  const int64_t hw = g.Ho * g.Wo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tx + 16 * i;
    if (m >= g.M) continue;
    const int64_t b = m / hw;
    const int64_t pix = m - b * hw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t o = n0 + ty * 4 + j;
      if (o >= g.O) continue;
      const int64_t off = (b * g.O + o) * hw + pix;
      float re = acc_re[i][j], im = acc_im[i][j];
      if (ep.b_re) {
        re += Elem<T>::to_f(static_cast<const T*>(ep.b_re)[o]);
        if constexpr (kCplx) im += Elem<T>::to_f(static_cast<const T*>(ep.b_im)[o]);
      }
      if constexpr (kVD) {
        const float sd = sd_of(acc_s2[i][j]);
        float er, ei = 0.f;
        if (ep.noise.mode == CPLXK_NOISE_INJECT) {
          er = Elem<T>::to_f(static_cast<const T*>(ep.eps_re)[off]);
          if constexpr (kCplx) ei = Elem<T>::to_f(static_cast<const T*>(ep.eps_im)[off]);
        } else if (ep.noise.mode == CPLXK_NOISE_PHILOX_TORCH) {
          TorchNoiseCursor cur;
          cur.seek(static_cast<uint64_t>(off), ep.noise.threads);
          er = cur.next(ep.noise) * ep.noise.scale;
          if constexpr (kCplx) {
            cur.seek(static_cast<uint64_t>(ep.plane_elems + off), ep.noise.threads);
            ei = cur.next(ep.noise) * ep.noise.scale;
          }
        } else {
          if constexpr (kCplx) {
            const float2 z = philox_fast_pair(static_cast<uint64_t>(off), ep.noise);
            er = z.x * ep.noise.scale, ei = z.y * ep.noise.scale;
          } else {
            const uint64_t q = static_cast<uint64_t>(off) >> 2;
            const int comp = static_cast<int>(off & 3);
            float4 a = philox_fast_normal4(q, 0u, ep.noise);
            er = (comp == 0 ? a.x : comp == 1 ? a.y : comp == 2 ? a.z : a.w) * ep.noise.scale;
          }
        }
        re = fmaf(er, sd, re);
        if constexpr (kCplx) im = fmaf(ei, sd, im);
      }
      static_cast<T*>(ep.y_re)[off] = Elem<T>::from_f(re);
      if constexpr (kCplx) static_cast<T*>(ep.y_im)[off] = Elem<T>::from_f(im);
    }
  }
}

template <typename T, bool kCplx, bool kVD>
static int launch_conv(const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                       const void* ls2, const ConvGeom& g, const ConvEpi& ep, cudaStream_t st) {
  const int64_t gm = (g.M + CB - 1) / CB, gn = (g.O + CB - 1) / CB;
  if (gm > 0x7fffffff || gn > 65535) return CPLXK_ERR_UNSUPPORTED;
  dim3 grid(static_cast<unsigned>(gm), static_cast<unsigned>(gn));
  conv_simt_kernel<T, kCplx, kVD><<<grid, 256, 0, st>>>(
      static_cast<const T*>(x_re), static_cast<const T*>(x_im), static_cast<const T*>(w_re),
      static_cast<const T*>(w_im), static_cast<const T*>(ls2), g, ep);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

}  // namespace cplxk

using namespace cplxk;

extern "C" int cplxk_conv2d_fwd_g(const void* x_re, const void* x_im, const void* w_re,
                                  const void* w_im, const void* b_re, const void* b_im,
                                  const void* log_sigma2, const void* eps_re, const void* eps_im,
                                  int noise, uint64_t seed, uint64_t offset, uint32_t philox_threads,
                                  void* y_re, void* y_im, int64_t B, int64_t C, int64_t H, int64_t W,
                                  int64_t O, int64_t kh, int64_t kw, int64_t stride_h,
                                  int64_t stride_w, int64_t pad_h, int64_t pad_w, int64_t dil_h,
                                  int64_t dil_w, int64_t groups, int dtype, int math, int channels_last,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || C < 1 || H < 1 || W < 1 || O < 0 || kh < 1 || kw < 1 || stride_h < 1 ||
      stride_w < 1 || pad_h < 0 || pad_w < 0 || dil_h < 1 || dil_w < 1)
    return CPLXK_ERR_BADARG;
  if (B == 0 || O == 0) return CPLXK_OK;     // empty batch / layer (pointers may be null)
  if (!x_re || !w_re || !y_re) return CPLXK_ERR_BADARG;
  if (groups < 1 || groups > 0x7fffffff || C % groups || O % groups) return CPLXK_ERR_BADARG;
  // x_im with REAL weights and output: the real convolution of |x_re + i x_im|^2 (variance operand of
  // a complex variational layer; squared inside the transposing pre-pass, tensor-core path only)
  const bool abs2 = x_im != nullptr && w_im == nullptr && y_im == nullptr;
  const bool cplx = x_im != nullptr && !abs2;
  if (cplx != (w_im != nullptr) || cplx != (y_im != nullptr)) return CPLXK_ERR_BADARG;
  if (b_re && cplx && !b_im) return CPLXK_ERR_BADARG;
  const bool vd = log_sigma2 != nullptr;
  if (abs2 && (vd || channels_last || !aligned16(x_im))) return CPLXK_ERR_UNSUPPORTED;
  if (vd) {
    if (noise < CPLXK_NOISE_INJECT || noise > CPLXK_NOISE_PHILOX_FAST) return CPLXK_ERR_BADARG;
    if (noise == CPLXK_NOISE_INJECT && (!eps_re || (cplx && !eps_im))) return CPLXK_ERR_BADARG;
    if (noise == CPLXK_NOISE_PHILOX_TORCH && (philox_threads == 0 || (offset & 3u)))
      return CPLXK_ERR_BADARG;
  }
  ConvGeom g;
  g.B = B, g.C = C, g.H = H, g.W = W, g.O = O;
  g.kh = static_cast<int>(kh), g.kw = static_cast<int>(kw);
  g.sh = static_cast<int>(stride_h), g.sw = static_cast<int>(stride_w);
  g.ph = static_cast<int>(pad_h), g.pw = static_cast<int>(pad_w);
  g.dh = static_cast<int>(dil_h), g.dw = static_cast<int>(dil_w);
  const int64_t eh = H + 2 * pad_h - dil_h * (kh - 1) - 1;
  const int64_t ew = W + 2 * pad_w - dil_w * (kw - 1) - 1;
  if (eh < 0 || ew < 0) return CPLXK_ERR_BADARG;
  g.Ho = eh / stride_h + 1, g.Wo = ew / stride_w + 1;
  g.M = B * g.Ho * g.Wo, g.K = C * kh * kw;
  if (g.M == 0 || O == 0) return CPLXK_OK;

  ConvEpi ep;
  ep.b_re = b_re, ep.b_im = b_im, ep.eps_re = eps_re, ep.eps_im = eps_im;
  ep.y_re = y_re, ep.y_im = y_im, ep.plane_elems = B * O * g.Ho * g.Wo;
  ep.noise.mode = vd ? noise : CPLXK_NOISE_INJECT;
  ep.noise.seed_lo = static_cast<uint32_t>(seed);
  ep.noise.seed_hi = static_cast<uint32_t>(seed >> 32);
  ep.noise.ctr_base = offset >> 2;
  ep.noise.threads = philox_threads ? philox_threads : 1u;
  ep.noise.scale = cplx ? (1.0f / static_cast<float>(1.4142135623730951)) : 1.0f;

  auto st = static_cast<cudaStream_t>(stream);
  // tensor-core implicit GEMM: workspace supplied, geometry within the TMA box limits.  Complex
  // planes with groups == 1 take the CTA-pair / persistent kernels; real planes and grouped
  // convolutions the one-tile-per-CTA kernel (a group = one more factor of the n-block index).
  const int ng = static_cast<int>(groups);
  const bool tc_ok = workspace != nullptr && aligned16(workspace) && aligned16(x_re) &&
                     (cplx || !channels_last) &&
                     conv_tc_supported(dtype, B, C, H, W, O, g.Ho, g.Wo, g.kh, g.kw, g.sh, g.sw, ng, !cplx) &&
                     workspace_bytes >= conv_tc_workspace_bytes(dtype, vd, B, C, H, W, O, kh, kw, ng, !cplx);
  if (math < CPLXK_MATH_AUTO || math > CPLXK_MATH_TENSOR_TF32) return CPLXK_ERR_BADARG;
  if (math == CPLXK_MATH_TENSOR && !tc_ok) return workspace ? CPLXK_ERR_UNSUPPORTED : CPLXK_ERR_WORKSPACE;
  if (channels_last && !(tc_ok && math != CPLXK_MATH_SIMT)) return CPLXK_ERR_UNSUPPORTED;  // NHWC: TC path only
  if (channels_last && ng > 1) return CPLXK_ERR_UNSUPPORTED;
  if (math != CPLXK_MATH_SIMT && tc_ok) {
    ConvTcEpi te;
    te.b_re = b_re, te.b_im = b_im, te.eps_re = eps_re, te.eps_im = eps_im;
    te.y_re = y_re, te.y_im = y_im, te.plane_elems = ep.plane_elems, te.noise = ep.noise;
    te.nhwc = channels_last ? 1 : 0;
    te.f16_ok = math != CPLXK_MATH_TENSOR_TF32;
    return conv_tc_dispatch(dtype, vd, channels_last != 0, x_re, x_im, w_re, w_im, log_sigma2, workspace, B, C, H, W, O,
                            g.Ho, g.Wo, g.kh, g.kw, g.sh, g.sw, g.ph, g.pw, g.dh, g.dw, te, st, ng, abs2);
  }
  if (abs2) return CPLXK_ERR_UNSUPPORTED;
  // the exact-fp32 CUDA-core kernel covers one group per launch: the caller loops
  if (ng > 1) return CPLXK_ERR_UNSUPPORTED;
#define CPLXK_CONV_CASE(T)                                                                  \
  if (cplx && vd) return launch_conv<T, true, true>(x_re, x_im, w_re, w_im, log_sigma2, g, ep, st);   \
  if (cplx && !vd) return launch_conv<T, true, false>(x_re, x_im, w_re, w_im, log_sigma2, g, ep, st); \
  if (!cplx && vd) return launch_conv<T, false, true>(x_re, x_im, w_re, w_im, log_sigma2, g, ep, st); \
  return launch_conv<T, false, false>(x_re, x_im, w_re, w_im, log_sigma2, g, ep, st);
  if (dtype == CPLXK_F32) { CPLXK_CONV_CASE(float) }
  if (dtype == CPLXK_BF16) { CPLXK_CONV_CASE(__nv_bfloat16) }
#undef CPLXK_CONV_CASE
  return CPLXK_ERR_BADARG;
}

extern "C" int cplxk_conv2d_fwd(const void* x_re, const void* x_im, const void* w_re,
                                const void* w_im, const void* b_re, const void* b_im,
                                const void* log_sigma2, const void* eps_re, const void* eps_im,
                                int noise, uint64_t seed, uint64_t offset, uint32_t philox_threads,
                                void* y_re, void* y_im, int64_t B, int64_t C, int64_t H, int64_t W,
                                int64_t O, int64_t kh, int64_t kw, int64_t stride_h,
                                int64_t stride_w, int64_t pad_h, int64_t pad_w, int64_t dil_h,
                                int64_t dil_w, int dtype, int math, int channels_last,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return cplxk_conv2d_fwd_g(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise, seed,
                            offset, philox_threads, y_re, y_im, B, C, H, W, O, kh, kw, stride_h, stride_w,
                            pad_h, pad_w, dil_h, dil_w, 1, dtype, math, channels_last, workspace,
                            workspace_bytes, stream);
}

extern "C" size_t cplxk_conv2d_workspace_bytes_g(int64_t B, int64_t C, int64_t H, int64_t W, int64_t O,
                                                 int64_t kh, int64_t kw, int64_t groups, int is_complex,
                                                 int dtype, int variational) {
  if (B < 0 || C < 1 || H < 1 || W < 1 || O < 0 || kh < 1 || kw < 1 || groups < 1 || C % groups ||
      O % groups)
    return 0;
  return conv_tc_workspace_bytes(dtype, variational != 0, B, C, H, W, O, kh, kw,
                                 static_cast<int>(groups), is_complex == 0);
}

extern "C" size_t cplxk_conv2d_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t O,
                                               int64_t kh, int64_t kw, int dtype, int variational) {
  return cplxk_conv2d_workspace_bytes_g(B, C, H, W, O, kh, kw, 1, 1, dtype, variational);
}
