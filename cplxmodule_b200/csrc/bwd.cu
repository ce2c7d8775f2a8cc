// Backward-pass support kernels.  The heavy lifting of every gradient is a GEMM that runs
// through the SAME tcgen05 forward kernels (cplxk_linear_fwd):
//   dx = g . conj(W)           ->  linear(g, (U^T, -V^T))          [M,K]
//   dW = g^T . conj(x)         ->  linear((g_re^T, g_im^T), (x_re^T, -x_im^T))   [N,K]
//   variational part:  g_s2 = (g_re eps_re + g_im eps_im) / (2 sqrt(s2)) [s2 > 1e-8]
//                      dq = g_s2 . E,  dE = g_s2^T . q,  d log_sigma2 = dE * E,
//                      dx += 2 x * dq
// (derivatives of cplxmodule/nn/relevance/complex/base.py:43-56; the reference gets them
// from torch autograd).  Both GEMM operands must be K-major for the TMA path, so the
// kernels here produce the transposed / derived operand planes, plus the small elementwise
// pieces and the closed-form KL gradients (complex/vd.py:38-41: dEi(x)/dx = e^x / x).
#include "common.cuh"
#include "knobs.cuh"
#include "noise.cuh"

namespace cplxk {

enum { TR_COPY = 0, TR_NEG = 1, TR_EXP = 2, TR_ABS2 = 3, TR_SQR = 4, TR_MUL = 5 };

// out[c, r] = op(in[r, c])  (TR_ABS2: in^2 + in2^2), 32 x 32 smem tiles, coalesced both ways
template <typename T, int kOp>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ in, const T* __restrict__ in2, T* __restrict__ out,
                 int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 32, r0 = static_cast<int64_t>(blockIdx.y) * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = Elem<T>::to_f(in[r * cols + c]);
      if constexpr (kOp == TR_NEG) v = -v;
      if constexpr (kOp == TR_EXP) v = __expf(v);
      if constexpr (kOp == TR_SQR) v = v * v;
      if constexpr (kOp == TR_ABS2) {
        const float w = Elem<T>::to_f(in2[r * cols + c]);
        v = fmaf(v, v, w * w);
      }
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < cols && r < rows) out[c * rows + r] = Elem<T>::from_f(tile[tx][ty + 8 * i]);
  }
}

// out[n] = sum_m g[m, n]  (bias gradient); one 1024-thread block per 32 columns, four independent
// loads in flight per thread (the 256-thread, one-load-per-thread version ran at 0.8 TB/s: 83 us
// for a 4096 x 4096 plane, 8 % of a training step); fixed summation order: deterministic
template <typename T>
__global__ void __launch_bounds__(1024)
colsum_kernel(const T* __restrict__ g, T* __restrict__ out, int64_t M, int64_t N) {
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (n < N) {
    const T* p = g + n;
    int64_t m = ty;
    for (; m + 96 < M; m += 128) {
      const float v0 = Elem<T>::to_f(p[m * N]), v1 = Elem<T>::to_f(p[(m + 32) * N]);
      const float v2 = Elem<T>::to_f(p[(m + 64) * N]), v3 = Elem<T>::to_f(p[(m + 96) * N]);
      a0 += v0, a1 += v1, a2 += v2, a3 += v3;
    }
    for (; m < M; m += 32) a0 += Elem<T>::to_f(p[m * N]);
  }
  part[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += part[i][tx];
    out[n] = Elem<T>::from_f(s);
  }
}

// g_s2 = (g_re eps_re + g_im eps_im) * 0.5 / sqrt(s2) where s2 > 1e-8, else 0
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
vd_grad_s2_kernel(const T* __restrict__ g_re, const T* __restrict__ g_im, const T* __restrict__ s2,
                  const T* __restrict__ eps_re, const T* __restrict__ eps_im, T* __restrict__ out,
                  int64_t M, int64_t N, NoiseParams np) {
  // one thread = 8 consecutive columns of one row (noise quads / cursor stay cheap)
  const int64_t runs_per_row = (N + 7) / 8;
  const int64_t total = M * runs_per_row;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = t / runs_per_row;
    const int64_t n0 = (t - m * runs_per_row) * 8;
    const int64_t off = m * N + n0;
    const int nv = (N - n0) < 8 ? static_cast<int>(N - n0) : 8;
    float er[8], ei[8];
    if (np.mode == CPLXK_NOISE_INJECT) {
      for (int j = 0; j < nv; ++j) {
        er[j] = Elem<T>::to_f(eps_re[off + j]);
        ei[j] = kCplx ? Elem<T>::to_f(eps_im[off + j]) : 0.f;
      }
    } else if (np.mode == CPLXK_NOISE_PHILOX_TORCH) {
      TorchNoiseCursor cur;
      cur.seek(static_cast<uint64_t>(off), np.threads);
      for (int j = 0; j < nv; ++j) er[j] = cur.next(np) * np.scale;
      if constexpr (kCplx) {
        cur.seek(static_cast<uint64_t>(M * N + off), np.threads);
        for (int j = 0; j < nv; ++j) ei[j] = cur.next(np) * np.scale;
      }
    } else {
      const uint64_t q0 = static_cast<uint64_t>(m) * static_cast<uint64_t>((N + 3) >> 2) +
                          static_cast<uint64_t>(n0 >> 2);
      for (int q = 0; q < 2; ++q) {
        float4 a = philox_fast_normal4(q0 + q, 0u, np);
        er[4 * q] = a.x * np.scale, er[4 * q + 1] = a.y * np.scale;
        er[4 * q + 2] = a.z * np.scale, er[4 * q + 3] = a.w * np.scale;
        if constexpr (kCplx) {
          float4 b = philox_fast_normal4(q0 + q, 1u, np);
          ei[4 * q] = b.x * np.scale, ei[4 * q + 1] = b.y * np.scale;
          ei[4 * q + 2] = b.z * np.scale, ei[4 * q + 3] = b.w * np.scale;
        }
      }
    }
    for (int j = 0; j < nv; ++j) {
      const float v = Elem<T>::to_f(s2[off + j]);
      float acc = Elem<T>::to_f(g_re[off + j]) * er[j];
      if constexpr (kCplx) acc = fmaf(Elem<T>::to_f(g_im[off + j]), ei[j], acc);
      out[off + j] = Elem<T>::from_f(v > 1e-8f ? acc * 0.5f * rsqrtf(v) : 0.f);
    }
  }
}

// The same for the torch-exact noise layout, organised by PHILOX CALL instead of by output run:
// one Philox4x32-10 call (counter c, subsequence idx) yields the four normals of the elements
// idx + T (4c + k), k = 0..3 (rows 74 apart at the headline shape).  A thread owns those four
// outputs: one call gives their real-plane noise, the imaginary-plane noise (elements M N further
// on in torch's ONE randn(2, M, N) stream) sits in at most two more calls of one other
// subsequence, and every Box-Muller evaluation yields two normals that are both used.  Three
// Philox calls instead of eight per four complex outputs (the run-by-run kernel regenerated the
// 33.5 M normals of the headline layer in 192 us, 9 % of a training step).  Bit-identical values.
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
vd_grad_s2_torch_kernel(const T* __restrict__ g_re, const T* __restrict__ g_im, const T* __restrict__ s2,
                        T* __restrict__ out, int64_t MN, uint32_t r0, uint64_t q0, uint64_t calls,
                        NoiseParams np) {
  const uint32_t Tn = np.threads;
  const PhiloxKey key{np.seed_lo, np.seed_hi};
  const uint64_t total = static_cast<uint64_t>(Tn) * calls;
  for (uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t c = t / Tn;
    const uint32_t idx = static_cast<uint32_t>(t - c * Tn);
    const uint64_t e0 = idx + static_cast<uint64_t>(Tn) * 4u * c;
    if (e0 >= static_cast<uint64_t>(MN)) continue;
    float er[4], ei[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const uint64_t ctr = np.ctr_base + c;
      const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u), key);
      const float2 a = _curand_box_muller(r.x, r.y), b = _curand_box_muller(r.z, r.w);
      er[0] = a.x * np.scale, er[1] = a.y * np.scale, er[2] = b.x * np.scale, er[3] = b.y * np.scale;
    }
    if constexpr (kCplx) {
      uint32_t idx2 = idx + r0;          // subsequence of the imaginary-plane elements (r0 < T, idx < T)
      uint64_t sl0 = 4u * c + q0;        // their first slot
      if (idx2 >= Tn) idx2 -= Tn, ++sl0;
      const uint32_t c0 = static_cast<uint32_t>(sl0) & 3u;
      const uint64_t ctr = np.ctr_base + (sl0 >> 2);
      float n8[8];
      {
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx2, 0u), key);
        const float2 a = _curand_box_muller(r.x, r.y), b = _curand_box_muller(r.z, r.w);
        n8[0] = a.x, n8[1] = a.y, n8[2] = b.x, n8[3] = b.y;
      }
      n8[4] = n8[5] = n8[6] = n8[7] = 0.f;
      if (c0 != 0u) {
        const uint64_t ctr1 = ctr + 1;
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr1), static_cast<uint32_t>(ctr1 >> 32), idx2, 0u), key);
        const float2 a = _curand_box_muller(r.x, r.y);
        n8[4] = a.x, n8[5] = a.y;
        if (c0 == 3u) {
          const float2 b = _curand_box_muller(r.z, r.w);
          n8[6] = b.x;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) v = (static_cast<uint32_t>(j) == c0 + k) ? n8[j] : v;
        ei[k] = v * np.scale;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t e = e0 + static_cast<uint64_t>(Tn) * k;
      if (e < static_cast<uint64_t>(MN)) {
        const float v = Elem<T>::to_f(s2[e]);
        float acc = Elem<T>::to_f(g_re[e]) * er[k];
        if constexpr (kCplx) acc = fmaf(Elem<T>::to_f(g_im[e]), ei[k], acc);
        out[e] = Elem<T>::from_f(v > 1e-8f ? acc * 0.5f * rsqrtf(v) : 0.f);
      }
    }
  }
}

// ---- y += eps * sqrt(max(s2, 1e-8)) in place (variational conv forward built from the mean conv
// and the variance conv of the fast kernels; complex/base.py:120-135, real/base.py:149-163).
// Torch-exact noise is generated BY PHILOX CALL as in vd_grad_s2_torch_kernel: a thread owns the
// four elements one call serves (T apart), the imaginary-plane noise of those elements comes from
// at most two more calls, every Box-Muller evaluation yields two used normals.
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
vd_combine_torch_kernel(T* __restrict__ y_re, T* __restrict__ y_im, const T* __restrict__ s2, int64_t MN,
                        uint32_t r0, uint64_t q0, uint64_t calls, NoiseParams np) {
  const uint32_t Tn = np.threads;
  const PhiloxKey key{np.seed_lo, np.seed_hi};
  const uint64_t total = static_cast<uint64_t>(Tn) * calls;
  for (uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t c = t / Tn;
    const uint32_t idx = static_cast<uint32_t>(t - c * Tn);
    const uint64_t e0 = idx + static_cast<uint64_t>(Tn) * 4u * c;
    if (e0 >= static_cast<uint64_t>(MN)) continue;
    float er[4], ei[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const uint64_t ctr = np.ctr_base + c;
      const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u), key);
      const float2 a = _curand_box_muller(r.x, r.y), b = _curand_box_muller(r.z, r.w);
      er[0] = a.x * np.scale, er[1] = a.y * np.scale, er[2] = b.x * np.scale, er[3] = b.y * np.scale;
    }
    if constexpr (kCplx) {
      uint32_t idx2 = idx + r0;
      uint64_t sl0 = 4u * c + q0;
      if (idx2 >= Tn) idx2 -= Tn, ++sl0;
      const uint32_t c0 = static_cast<uint32_t>(sl0) & 3u;
      const uint64_t ctr = np.ctr_base + (sl0 >> 2);
      float n8[8];
      {
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx2, 0u), key);
        const float2 a = _curand_box_muller(r.x, r.y), b = _curand_box_muller(r.z, r.w);
        n8[0] = a.x, n8[1] = a.y, n8[2] = b.x, n8[3] = b.y;
      }
      n8[4] = n8[5] = n8[6] = n8[7] = 0.f;
      if (c0 != 0u) {
        const uint64_t ctr1 = ctr + 1;
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr1), static_cast<uint32_t>(ctr1 >> 32), idx2, 0u), key);
        const float2 a = _curand_box_muller(r.x, r.y);
        n8[4] = a.x, n8[5] = a.y;
        if (c0 == 3u) n8[6] = _curand_box_muller(r.z, r.w).x;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) v = (static_cast<uint32_t>(j) == c0 + k) ? n8[j] : v;
        ei[k] = v * np.scale;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t e = e0 + static_cast<uint64_t>(Tn) * k;
      if (e < static_cast<uint64_t>(MN)) {
        const float sd = sd_of(Elem<T>::to_f(s2[e]));
        y_re[e] = Elem<T>::from_f(fmaf(er[k], sd, Elem<T>::to_f(y_re[e])));
        if constexpr (kCplx) y_im[e] = Elem<T>::from_f(fmaf(ei[k], sd, Elem<T>::to_f(y_im[e])));
      }
    }
  }
}

// Complex planes, flat form: torch draws the 2 MN normals of a complex tensor as ONE stream (real
// plane first, cplx.py:544-550), so a thread may simply own one Philox call of that stream and add
// its four normals wherever they land -- real or imaginary plane.  Two calls and four Box-Muller
// evaluations per four complex outputs (the minimum; the form above needs three and five to six,
// plus the component selects), paid for with a second read of s2 (the two planes of one element are
// MN normals apart in the stream).
template <typename T>
__global__ void __launch_bounds__(256)
vd_combine_torch_flat_kernel(T* __restrict__ y_re, T* __restrict__ y_im, const T* __restrict__ s2,
                             int64_t MN, uint64_t calls, NoiseParams np) {
  // grid: x over torch's Tn generator threads, y (strided) over the Philox calls -- no division
  const uint32_t Tn = np.threads;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Tn) return;
  const PhiloxKey key{np.seed_lo, np.seed_hi};
  const uint64_t n2 = 2u * static_cast<uint64_t>(MN);
  for (uint64_t c = blockIdx.y; c < calls; c += gridDim.y) {
    const uint64_t l0 = idx + static_cast<uint64_t>(Tn) * 4u * c;
    if (l0 >= n2) return;
    const uint64_t ctr = np.ctr_base + c;
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u), key);
    const float2 a = _curand_box_muller(r.x, r.y), b = _curand_box_muller(r.z, r.w);
    const float n[4] = {a.x * np.scale, a.y * np.scale, b.x * np.scale, b.y * np.scale};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t l = l0 + static_cast<uint64_t>(Tn) * k;
      if (l < n2) {
        const bool im = l >= static_cast<uint64_t>(MN);
        const uint64_t e = im ? l - static_cast<uint64_t>(MN) : l;
        T* y = im ? y_im : y_re;
        const float sd = sd_of(Elem<T>::to_f(s2[e]));
        y[e] = Elem<T>::from_f(fmaf(n[k], sd, Elem<T>::to_f(y[e])));
      }
    }
  }
}

// injected noise / the private `fast` layout of the conv kernels (one element per thread)
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
vd_combine_kernel(T* __restrict__ y_re, T* __restrict__ y_im, const T* __restrict__ s2,
                  const T* __restrict__ eps_re, const T* __restrict__ eps_im, int64_t MN, NoiseParams np) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < MN;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float er, ei = 0.f;
    if (np.mode == CPLXK_NOISE_INJECT) {
      er = Elem<T>::to_f(eps_re[e]);
      if constexpr (kCplx) ei = Elem<T>::to_f(eps_im[e]);
    } else if constexpr (kCplx) {
      const float2 z = philox_fast_pair(static_cast<uint64_t>(e), np);
      er = z.x * np.scale, ei = z.y * np.scale;
    } else {
      const float4 a = philox_fast_normal4(static_cast<uint64_t>(e) >> 2, 0u, np);
      const int comp = static_cast<int>(e & 3);
      er = (comp == 0 ? a.x : comp == 1 ? a.y : comp == 2 ? a.z : a.w) * np.scale;
    }
    const float sd = sd_of(Elem<T>::to_f(s2[e]));
    y_re[e] = Elem<T>::from_f(fmaf(er, sd, Elem<T>::to_f(y_re[e])));
    if constexpr (kCplx) y_im[e] = Elem<T>::from_f(fmaf(ei, sd, Elem<T>::to_f(y_im[e])));
  }
}

// dx_re += 2 x_re dq ; dx_im += 2 x_im dq
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
vd_grad_input_kernel(T* __restrict__ dx_re, T* __restrict__ dx_im, const T* __restrict__ x_re,
                     const T* __restrict__ x_im, const T* __restrict__ dq, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float d = 2.f * Elem<T>::to_f(dq[i]);
    dx_re[i] = Elem<T>::from_f(fmaf(Elem<T>::to_f(x_re[i]), d, Elem<T>::to_f(dx_re[i])));
    if constexpr (kCplx)
      dx_im[i] = Elem<T>::from_f(fmaf(Elem<T>::to_f(x_im[i]), d, Elem<T>::to_f(dx_im[i])));
  }
}

// out = a * exp(b) (+ out if accumulate)
template <typename T>
__global__ void __launch_bounds__(256)
mul_exp_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n,
               int accumulate) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = Elem<T>::to_f(a[i]) * __expf(Elem<T>::to_f(b[i]));
    if (accumulate) v += Elem<T>::to_f(out[i]);
    out[i] = Elem<T>::from_f(v);
  }
}

// d penalty / d (w_re, w_im, log_sigma2), times the upstream gradient (scalar or per element)
template <typename T, int kKind>
__global__ void __launch_bounds__(256)
kl_bwd_kernel(const T* __restrict__ w_re, const T* __restrict__ w_im, const T* __restrict__ ls2,
              int64_t n, const void* __restrict__ grad, int grad_is_tensor, int grad_is_f32,
              float scale, T* __restrict__ d_w_re, T* __restrict__ d_w_im, T* __restrict__ d_ls2) {
  constexpr bool kCplx = kKind >= CPLXK_KL_CPLX_VD;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float wr = Elem<T>::to_f(w_re[i]);
    const float wi = kCplx ? Elem<T>::to_f(w_im[i]) : 0.f;
    const float r2 = fmaf(wr, wr, wi * wi) + 1e-24f;
    const float nla = __logf(r2) - Elem<T>::to_f(ls2[i]);  // -log_alpha
    float dn;                                             // d penalty / d (-log_alpha)
    if constexpr (kKind == CPLXK_KL_REAL_VD) {
      const float s = __fdividef(1.f, 1.f + __expf(-fmaf(1.48695f, nla, -1.87320f)));
      dn = 0.5f * __fdividef(1.f, 1.f + __expf(-nla)) + 0.63576f * 1.48695f * s * (1.f - s);
    } else if constexpr (kKind == CPLXK_KL_REAL_ARD) {
      dn = 0.5f * __fdividef(1.f, 1.f + __expf(-nla));
    } else if constexpr (kKind == CPLXK_KL_CPLX_VD) {
      dn = -expm1f(-__expf(nla));                          // 1 - exp(-1/alpha)
    } else if constexpr (kKind == CPLXK_KL_CPLX_VD_APPROX) {
      const float s = __fdividef(1.f, 1.f + __expf(-fmaf(1.36526f, nla, -1.45926f)));
      dn = __fdividef(1.f, 1.f + __expf(-nla)) + 0.57810f * 1.36526f * s * (1.f - s);
    } else if constexpr (kKind == CPLXK_KL_CPLX_VD_SCALEFREE) {
      dn = -0.5f * expm1f(-__expf(nla));                   // (Ein(t) - gamma - ls2) / 2 through t
    } else {
      dn = __fdividef(1.f, 1.f + __expf(-nla));
    }
    float g;
    if (grad_is_tensor)
      g = grad_is_f32 ? static_cast<const float*>(grad)[i] : Elem<T>::to_f(static_cast<const T*>(grad)[i]);
    else
      g = grad_is_f32 ? static_cast<const float*>(grad)[0] : Elem<T>::to_f(static_cast<const T*>(grad)[0]);
    const float g0 = g * scale;
    g = g0 * dn;
    // -log_alpha = log(r2) - ls2  =>  d/dw = 2 w / r2 ,  d/dls2 = -1
    const float k = 2.f * __fdividef(g, r2);
    d_w_re[i] = Elem<T>::from_f(k * wr);
    if constexpr (kCplx) d_w_im[i] = Elem<T>::from_f(k * wi);
    // the scale-free penalty also depends on log_sigma2 directly (-ls2 / 2)
    const float direct = kKind == CPLXK_KL_CPLX_VD_SCALEFREE ? -0.5f * g0 : 0.f;
    d_ls2[i] = Elem<T>::from_f(direct - g);
  }
}

// out = op(a [, b]) elementwise: same op codes as the transposing kernel, no transposition
template <typename T>
__global__ void __launch_bounds__(256)
eltwise_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n, int op) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = Elem<T>::to_f(a[i]);
    if (op == TR_NEG) v = -v;
    else if (op == TR_EXP) v = __expf(v);
    else if (op == TR_SQR) v = v * v;
    else if (op == TR_ABS2) {
      const float w = Elem<T>::to_f(b[i]);
      v = fmaf(v, v, w * w);
    } else if (op == TR_MUL) {
      v *= Elem<T>::to_f(b[i]);
    }
    out[i] = Elem<T>::from_f(v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mul2_kernel(const T* __restrict__ a0, const T* __restrict__ a1, const T* __restrict__ b,
            T* __restrict__ o0, T* __restrict__ o1, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float m = Elem<T>::to_f(b[i]);
    o0[i] = Elem<T>::from_f(Elem<T>::to_f(a0[i]) * m);
    if (a1) o1[i] = Elem<T>::from_f(Elem<T>::to_f(a1[i]) * m);
  }
}

static inline int ew_grid(int64_t n) {
  int sms = 148;
  if (current_device_sm_count(&sms) != CPLXK_OK || sms < 1) sms = 148;
  const int64_t b = (n + 255) / 256, cap = static_cast<int64_t>(sms) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

int eltwise_mul2(const void* a0, const void* a1, const void* b, void* o0, void* o1, int64_t n, int dtype,
                 cudaStream_t st) {
  if (!a0 || !b || !o0 || (a1 && !o1) || n < 0) return CPLXK_ERR_BADARG;
  if (n == 0) return CPLXK_OK;
  if (dtype == CPLXK_F32)
    mul2_kernel<float><<<ew_grid(n), 256, 0, st>>>(static_cast<const float*>(a0), static_cast<const float*>(a1),
                                                  static_cast<const float*>(b), static_cast<float*>(o0),
                                                  static_cast<float*>(o1), n);
  else if (dtype == CPLXK_BF16)
    mul2_kernel<__nv_bfloat16><<<ew_grid(n), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(a0), static_cast<const __nv_bfloat16*>(a1),
        static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(o0), static_cast<__nv_bfloat16*>(o1), n);
  else
    return CPLXK_ERR_BADARG;
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

}  // namespace cplxk

using namespace cplxk;

#define CPLXK_BY_DTYPE(dtype, ...)                                    \
  if ((dtype) == CPLXK_F32) { using T = float; __VA_ARGS__; }         \
  else if ((dtype) == CPLXK_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else return CPLXK_ERR_BADARG;

extern "C" int cplxk_transpose2d(const void* in, const void* in2, void* out, int64_t rows,
                                 int64_t cols, int dtype, int op, void* stream) {
  if (!in || !out || rows < 0 || cols < 0 || op < TR_COPY || op > TR_SQR) return CPLXK_ERR_BADARG;
  if (op == TR_ABS2 && !in2) return CPLXK_ERR_BADARG;
  if (rows == 0 || cols == 0) return CPLXK_OK;
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  if (grid.y > 65535u) return CPLXK_ERR_UNSUPPORTED;
  auto st = static_cast<cudaStream_t>(stream);
  CPLXK_BY_DTYPE(dtype, {
    auto a = static_cast<const T*>(in);
    auto b = static_cast<const T*>(in2);
    auto o = static_cast<T*>(out);
    switch (op) {
      case TR_COPY: transpose_kernel<T, TR_COPY><<<grid, 256, 0, st>>>(a, b, o, rows, cols); break;
      case TR_NEG: transpose_kernel<T, TR_NEG><<<grid, 256, 0, st>>>(a, b, o, rows, cols); break;
      case TR_EXP: transpose_kernel<T, TR_EXP><<<grid, 256, 0, st>>>(a, b, o, rows, cols); break;
      case TR_SQR: transpose_kernel<T, TR_SQR><<<grid, 256, 0, st>>>(a, b, o, rows, cols); break;
      default: transpose_kernel<T, TR_ABS2><<<grid, 256, 0, st>>>(a, b, o, rows, cols); break;
    }
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_eltwise(int op, const void* a, const void* b, void* out, int64_t n, int dtype,
                             void* stream) {
  if (!a || !out || n < 0 || op < TR_COPY || op > TR_MUL) return CPLXK_ERR_BADARG;
  if ((op == TR_ABS2 || op == TR_MUL) && !b) return CPLXK_ERR_BADARG;
  if (n == 0) return CPLXK_OK;
  auto st = static_cast<cudaStream_t>(stream);
  CPLXK_BY_DTYPE(dtype, {
    eltwise_kernel<T><<<ew_grid(n), 256, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b),
                                                  static_cast<T*>(out), n, op);
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_colsum(const void* g, void* out, int64_t M, int64_t N, int dtype, void* stream) {
  if (!g || !out || M < 0 || N < 0) return CPLXK_ERR_BADARG;
  if (N == 0) return CPLXK_OK;
  auto st = static_cast<cudaStream_t>(stream);
  CPLXK_BY_DTYPE(dtype, {
    colsum_kernel<T><<<static_cast<unsigned>((N + 31) / 32), 1024, 0, st>>>(
        static_cast<const T*>(g), static_cast<T*>(out), M, N);
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_vd_grad_s2(const void* g_re, const void* g_im, const void* s2,
                                const void* eps_re, const void* eps_im, int noise, uint64_t seed,
                                uint64_t offset, uint32_t philox_threads, void* out, int64_t M,
                                int64_t N, int dtype, void* stream) {
  if (!g_re || !s2 || !out || M < 0 || N < 0) return CPLXK_ERR_BADARG;
  if (noise < CPLXK_NOISE_INJECT || noise > CPLXK_NOISE_PHILOX_FAST) return CPLXK_ERR_BADARG;
  const bool cplx = g_im != nullptr;
  if (noise == CPLXK_NOISE_INJECT && (!eps_re || (cplx && !eps_im))) return CPLXK_ERR_BADARG;
  if (noise == CPLXK_NOISE_PHILOX_TORCH && (philox_threads == 0 || (offset & 3u))) return CPLXK_ERR_BADARG;
  if (M == 0 || N == 0) return CPLXK_OK;
  NoiseParams np;
  np.mode = noise, np.seed_lo = static_cast<uint32_t>(seed), np.seed_hi = static_cast<uint32_t>(seed >> 32);
  np.ctr_base = offset >> 2, np.threads = philox_threads ? philox_threads : 1u;
  np.scale = cplx ? (1.0f / static_cast<float>(1.4142135623730951)) : 1.0f;
  auto st = static_cast<cudaStream_t>(stream);
  const int grid = ew_grid(M * ((N + 7) / 8));
  CPLXK_BY_DTYPE(dtype, {
    auto a = static_cast<const T*>(g_re); auto b = static_cast<const T*>(g_im);
    auto c = static_cast<const T*>(s2); auto e1 = static_cast<const T*>(eps_re);
    auto e2 = static_cast<const T*>(eps_im); auto o = static_cast<T*>(out);
    if (noise == CPLXK_NOISE_PHILOX_TORCH) {
      const int64_t MN = M * N;
      const uint64_t Tn = np.threads;
      const uint64_t calls = (static_cast<uint64_t>(MN) + 4 * Tn - 1) / (4 * Tn);
      const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(MN) % Tn);
      const uint64_t q0 = static_cast<uint64_t>(MN) / Tn;
      const int grid2 = ew_grid(static_cast<int64_t>(Tn * calls));
      if (cplx) vd_grad_s2_torch_kernel<T, true><<<grid2, 256, 0, st>>>(a, b, c, o, MN, r0, q0, calls, np);
      else vd_grad_s2_torch_kernel<T, false><<<grid2, 256, 0, st>>>(a, b, c, o, MN, r0, q0, calls, np);
    } else if (cplx) {
      vd_grad_s2_kernel<T, true><<<grid, 256, 0, st>>>(a, b, c, e1, e2, o, M, N, np);
    } else {
      vd_grad_s2_kernel<T, false><<<grid, 256, 0, st>>>(a, b, c, e1, e2, o, M, N, np);
    }
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_vd_grad_input(void* dx_re, void* dx_im, const void* x_re, const void* x_im,
                                   const void* dq, int64_t n, int dtype, void* stream) {
  if (!dx_re || !x_re || !dq || n < 0) return CPLXK_ERR_BADARG;
  if ((dx_im != nullptr) != (x_im != nullptr)) return CPLXK_ERR_BADARG;
  if (n == 0) return CPLXK_OK;
  auto st = static_cast<cudaStream_t>(stream);
  CPLXK_BY_DTYPE(dtype, {
    if (dx_im)
      vd_grad_input_kernel<T, true><<<ew_grid(n), 256, 0, st>>>(
          static_cast<T*>(dx_re), static_cast<T*>(dx_im), static_cast<const T*>(x_re),
          static_cast<const T*>(x_im), static_cast<const T*>(dq), n);
    else
      vd_grad_input_kernel<T, false><<<ew_grid(n), 256, 0, st>>>(
          static_cast<T*>(dx_re), nullptr, static_cast<const T*>(x_re), nullptr,
          static_cast<const T*>(dq), n);
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_mul_exp(const void* a, const void* b, void* out, int64_t n, int dtype,
                             int accumulate, void* stream) {
  if (!a || !b || !out || n < 0) return CPLXK_ERR_BADARG;
  if (n == 0) return CPLXK_OK;
  auto st = static_cast<cudaStream_t>(stream);
  CPLXK_BY_DTYPE(dtype, {
    mul_exp_kernel<T><<<ew_grid(n), 256, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b),
                                                  static_cast<T*>(out), n, accumulate);
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_kl_bwd(int kind, const void* w_re, const void* w_im, const void* log_sigma2,
                            int64_t n, int dtype, const void* grad, int grad_is_tensor,
                            int grad_is_f32, double scale, void* d_w_re, void* d_w_im,
                            void* d_log_sigma2, void* stream) {
  if (kind < 0 || kind > CPLXK_KL_CPLX_VD_SCALEFREE || n < 0) return CPLXK_ERR_BADARG;
  if (n == 0) return CPLXK_OK;
  const bool cplx = kind >= CPLXK_KL_CPLX_VD;
  if (!w_re || !log_sigma2 || !grad || !d_w_re || !d_log_sigma2) return CPLXK_ERR_BADARG;
  if (cplx != (w_im != nullptr) || cplx != (d_w_im != nullptr)) return CPLXK_ERR_BADARG;
  auto st = static_cast<cudaStream_t>(stream);
  const float sc = static_cast<float>(scale);
#define CPLXK_KLB(K)                                                                              \
  kl_bwd_kernel<T, K><<<ew_grid(n), 256, 0, st>>>(                                                \
      static_cast<const T*>(w_re), static_cast<const T*>(w_im), static_cast<const T*>(log_sigma2), \
      n, grad, grad_is_tensor, grad_is_f32, sc, static_cast<T*>(d_w_re), static_cast<T*>(d_w_im), \
      static_cast<T*>(d_log_sigma2))
  CPLXK_BY_DTYPE(dtype, {
    switch (kind) {
      case CPLXK_KL_REAL_VD: CPLXK_KLB(CPLXK_KL_REAL_VD); break;
      case CPLXK_KL_REAL_ARD: CPLXK_KLB(CPLXK_KL_REAL_ARD); break;
      case CPLXK_KL_CPLX_VD: CPLXK_KLB(CPLXK_KL_CPLX_VD); break;
      case CPLXK_KL_CPLX_VD_APPROX: CPLXK_KLB(CPLXK_KL_CPLX_VD_APPROX); break;
      case CPLXK_KL_CPLX_VD_SCALEFREE: CPLXK_KLB(CPLXK_KL_CPLX_VD_SCALEFREE); break;
      default: CPLXK_KLB(CPLXK_KL_CPLX_ARD); break;
    }
  })
#undef CPLXK_KLB
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_vd_combine(void* y_re, void* y_im, const void* s2, const void* eps_re,
                                const void* eps_im, int noise, uint64_t seed, uint64_t offset,
                                uint32_t philox_threads, int64_t numel, int dtype, void* stream) {
  if (!y_re || !s2 || numel < 0) return CPLXK_ERR_BADARG;
  if (noise < CPLXK_NOISE_INJECT || noise > CPLXK_NOISE_PHILOX_FAST) return CPLXK_ERR_BADARG;
  const bool cplx = y_im != nullptr;
  if (noise == CPLXK_NOISE_INJECT && (!eps_re || (cplx && !eps_im))) return CPLXK_ERR_BADARG;
  if (noise == CPLXK_NOISE_PHILOX_TORCH && (philox_threads == 0 || (offset & 3u))) return CPLXK_ERR_BADARG;
  if (numel == 0) return CPLXK_OK;
  NoiseParams np;
  np.mode = noise, np.seed_lo = static_cast<uint32_t>(seed), np.seed_hi = static_cast<uint32_t>(seed >> 32);
  np.ctr_base = offset >> 2, np.threads = philox_threads ? philox_threads : 1u;
  np.scale = cplx ? (1.0f / static_cast<float>(1.4142135623730951)) : 1.0f;
  auto st = static_cast<cudaStream_t>(stream);
  CPLXK_BY_DTYPE(dtype, {
    auto yr = static_cast<T*>(y_re); auto yi = static_cast<T*>(y_im);
    auto v = static_cast<const T*>(s2);
    if (noise == CPLXK_NOISE_PHILOX_TORCH) {
      const uint64_t Tn = np.threads;
      const uint64_t calls = (static_cast<uint64_t>(numel) + 4 * Tn - 1) / (4 * Tn);
      const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(numel) % Tn);
      const uint64_t q0 = static_cast<uint64_t>(numel) / Tn;
      if (cplx && knobs().combine_flat) {
        const uint64_t calls2 = (2u * static_cast<uint64_t>(numel) + 4 * Tn - 1) / (4 * Tn);
        const dim3 fg(static_cast<unsigned>((Tn + 255) / 256), static_cast<unsigned>(calls2 < 65535 ? calls2 : 65535));
        vd_combine_torch_flat_kernel<T><<<fg, 256, 0, st>>>(yr, yi, v, numel, calls2, np);
      } else {
        const int grid = ew_grid(static_cast<int64_t>(Tn * calls));
        if (cplx) vd_combine_torch_kernel<T, true><<<grid, 256, 0, st>>>(yr, yi, v, numel, r0, q0, calls, np);
        else vd_combine_torch_kernel<T, false><<<grid, 256, 0, st>>>(yr, yi, v, numel, r0, q0, calls, np);
      }
    } else {
      auto e1 = static_cast<const T*>(eps_re); auto e2 = static_cast<const T*>(eps_im);
      const int grid = ew_grid(numel);
      if (cplx) vd_combine_kernel<T, true><<<grid, 256, 0, st>>>(yr, yi, v, e1, e2, numel, np);
      else vd_combine_kernel<T, false><<<grid, 256, 0, st>>>(yr, yi, v, e1, e2, numel, np);
    }
  })
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}
