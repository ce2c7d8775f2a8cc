// Exact-fp32 (CUDA-core FMA) forward for any shape / alignment:
//   mu = x W^T (+b)   [complex: 4 real products, cplx.py:641-642]
//   s2 = |x|^2 . exp(log_sigma2)^T          [complex/base.py:50-54, real/base.py:48]
//   y  = mu + eps sqrt(max(s2,1e-8))        [complex/base.py:56]
// Used for shapes the TMA path cannot take (K*sizeof(T) % 16 != 0, unaligned
// planes) and as the on-device fp32 cross-check of the tcgen05 kernel.
#include "epilogue.cuh"

namespace cplxk {

constexpr int SB = 64;   // tile rows (M) and cols (N)
constexpr int SK = 16;   // k-slab

template <typename T, bool kCplx, bool kVD>
__global__ void __launch_bounds__(256)
fwd_simt_kernel(const T* __restrict__ x_re, const T* __restrict__ x_im,
                const T* __restrict__ w_re, const T* __restrict__ w_im,
                const T* __restrict__ ls2, int64_t M, int64_t N, int64_t K, EpiParams ep) {
  // [plane][k][row] with +1 padding: conflict-free transposed stores and row reads
  __shared__ float As[kCplx ? 2 : 1][SK][SB + 1];
  __shared__ float Aq[kVD ? SK : 1][SB + 1];
  __shared__ float Bs[kCplx ? 2 : 1][SK][SB + 1];
  __shared__ float Be[kVD ? SK : 1][SB + 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * SB;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * SB;

  float acc_re[4][4] = {}, acc_im[4][4] = {}, acc_s2[4][4] = {};

  // loader mapping: thread -> (row r, 4 consecutive k)
  const int lr = tid >> 2;          // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12

  for (int64_t k0 = 0; k0 < K; k0 += SK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + lk + j;
      const int64_t m = m0 + lr, n = n0 + lr;
      const bool ka = k < K;
      float xr = 0.f, xi = 0.f, wr = 0.f, wi = 0.f, e = 0.f;
      if (ka && m < M) {
        xr = Elem<T>::to_f(x_re[m * K + k]);
        if constexpr (kCplx) xi = Elem<T>::to_f(x_im[m * K + k]);
      }
      if (ka && n < N) {
        wr = Elem<T>::to_f(w_re[n * K + k]);
        if constexpr (kCplx) wi = Elem<T>::to_f(w_im[n * K + k]);
        if constexpr (kVD) e = expf(Elem<T>::to_f(ls2[n * K + k]));
      }
      As[0][lk + j][lr] = xr;
      Bs[0][lk + j][lr] = wr;
      if constexpr (kCplx) {
        As[1][lk + j][lr] = xi;
        Bs[1][lk + j][lr] = wi;
      }
      if constexpr (kVD) {
        Aq[lk + j][lr] = fmaf(xr, xr, xi * xi);
        Be[lk + j][lr] = e;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float ar[4], ai[4], aq[4], br[4], bi[4], be[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ar[i] = As[0][kk][ty * 4 + i];
        br[i] = Bs[0][kk][tx * 4 + i];
        if constexpr (kCplx) {
          ai[i] = As[1][kk][ty * 4 + i];
          bi[i] = Bs[1][kk][tx * 4 + i];
        }
        if constexpr (kVD) {
          aq[i] = Aq[kk][ty * 4 + i];
          be[i] = Be[kk][tx * 4 + i];
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

#pragma unroll
  for (int i = 0; i < 4; ++i)
    epilogue_run<T, kCplx, kVD, 4>(ep, m0 + ty * 4 + i, n0 + tx * 4, acc_re[i], acc_im[i], acc_s2[i]);
}

template <typename T, bool kCplx, bool kVD>
static int launch_simt(const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                       const void* ls2, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                       cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((N + SB - 1) / SB), static_cast<unsigned>((M + SB - 1) / SB));
  if (grid.y > 65535u) return CPLXK_ERR_UNSUPPORTED;
  fwd_simt_kernel<T, kCplx, kVD><<<grid, 256, 0, st>>>(
      static_cast<const T*>(x_re), static_cast<const T*>(x_im), static_cast<const T*>(w_re),
      static_cast<const T*>(w_im), static_cast<const T*>(ls2), M, N, K, ep);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

int fwd_simt_dispatch(int dtype, bool cplx, bool vd, const void* x_re, const void* x_im,
                      const void* w_re, const void* w_im, const void* ls2, int64_t M, int64_t N,
                      int64_t K, const EpiParams& ep, cudaStream_t st) {
#define CPLXK_SIMT_CASE(T)                                                                       \
  if (cplx && vd) return launch_simt<T, true, true>(x_re, x_im, w_re, w_im, ls2, M, N, K, ep, st);   \
  if (cplx && !vd) return launch_simt<T, true, false>(x_re, x_im, w_re, w_im, ls2, M, N, K, ep, st); \
  if (!cplx && vd) return launch_simt<T, false, true>(x_re, x_im, w_re, w_im, ls2, M, N, K, ep, st); \
  return launch_simt<T, false, false>(x_re, x_im, w_re, w_im, ls2, M, N, K, ep, st);
  if (dtype == CPLXK_F32) { CPLXK_SIMT_CASE(float) }
  if (dtype == CPLXK_BF16) { CPLXK_SIMT_CASE(__nv_bfloat16) }
#undef CPLXK_SIMT_CASE
  return CPLXK_ERR_BADARG;
}

}  // namespace cplxk
