// kl_reduce: one HBM pass over (weight planes, log_sigma2) -> per-element KL
// penalty and/or its sum.  HBM-bandwidth bound: 128-bit coalesced loads, a
// few dozen flops per element, warp-shuffle + block reduction, one double
// partial per CTA and a deterministic last-CTA finish (no float atomics).
//
// Reference formulas (cplxmodule/nn/relevance/...):
//   log_alpha           real/base.py:23-26, complex/base.py:27-31 (|w| = sqrt(re^2+im^2), cplx.py:183-192)
//   REAL_VD  penalty    real/vd.py:74-76     0.5 softplus(-la) + 0.63576 sigmoid(-1.48695 la - 1.87320)
//   REAL_ARD penalty    real/ard.py:39       0.5 softplus(-la)
//   CPLX_VD  penalty    complex/vd.py:95-99  gamma - la - Ei(-exp(-la))      (host scipy in the reference)
//   CPLX_ARD penalty    complex/ard.py:39    softplus(-la)
#include "kl_math.cuh"
#include "knobs.cuh"

namespace cplxk {

template <typename T, int kKind, bool kVecOK>
__global__ void __launch_bounds__(kKlThreads)
kl_kernel(const T* __restrict__ w_re, const T* __restrict__ w_im, const T* __restrict__ ls2,
          int64_t n, T* __restrict__ out_elem, float* __restrict__ out_sum, double scale,
          KlWorkspace* __restrict__ ws, T* __restrict__ out_mask, float threshold) {
  constexpr bool kCplx = kl_kind_is_cplx(kKind);
  constexpr int V = Elem<T>::kVec;
  __shared__ double sh[kKlThreads / 32];
  __shared__ bool is_last;

  float acc = 0.f;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kKlThreads + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * kKlThreads;

  int64_t done = 0;
  if constexpr (kVecOK) {
    const int64_t nvec = n / V;
    // two independent 16-byte streams per plane in flight per thread
    for (int64_t i = tid; i < nvec; i += 2 * nthreads) {
      const int64_t i2 = i + nthreads;
      const bool has2 = i2 < nvec;
      Vec16<T> a0, b0, c0, a1, b1, c1;
      a0.load(w_re + i * V);
      if constexpr (kCplx) b0.load(w_im + i * V);
      c0.load(ls2 + i * V);
      if (has2) {
        a1.load(w_re + i2 * V);
        if constexpr (kCplx) b1.load(w_im + i2 * V);
        c1.load(ls2 + i2 * V);
      }
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float p = penalty_of<kKind>(a0.v[j], kCplx ? b0.v[j] : 0.f, c0.v[j]);
        acc += p;
        o.v[j] = p;
      }
      if (out_elem) o.store(out_elem + i * V);
      if (out_mask) {   // relevance mask (log_alpha <= threshold) from the same loads
#pragma unroll
        for (int j = 0; j < V; ++j)
          o.v[j] = log_alpha_of<kKind>(a0.v[j], kCplx ? b0.v[j] : 0.f, c0.v[j]) <= threshold ? 1.f : 0.f;
        o.store(out_mask + i * V);
      }
      if (has2) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float p = penalty_of<kKind>(a1.v[j], kCplx ? b1.v[j] : 0.f, c1.v[j]);
          acc += p;
          o.v[j] = p;
        }
        if (out_elem) o.store(out_elem + i2 * V);
        if (out_mask) {
#pragma unroll
          for (int j = 0; j < V; ++j)
            o.v[j] = log_alpha_of<kKind>(a1.v[j], kCplx ? b1.v[j] : 0.f, c1.v[j]) <= threshold ? 1.f : 0.f;
          o.store(out_mask + i2 * V);
        }
      }
    }
    done = nvec * V;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    float wr = Elem<T>::to_f(w_re[i]);
    float wi = kCplx ? Elem<T>::to_f(w_im[i]) : 0.f;
    float p = penalty_of<kKind>(wr, wi, Elem<T>::to_f(ls2[i]));
    acc += p;
    if (out_elem) out_elem[i] = Elem<T>::from_f(p);
    if (out_mask)
      out_mask[i] = Elem<T>::from_f(log_alpha_of<kKind>(wr, wi, Elem<T>::to_f(ls2[i])) <= threshold ? 1.f : 0.f);
  }

  if (out_sum == nullptr) return;

  double bsum = block_sum(static_cast<double>(acc), sh);
  grid_sum_finish(bsum, ws, out_sum, scale, sh, &is_last);
}

template <typename T, bool kCplx, bool kVecOK>
__global__ void __launch_bounds__(256)
log_alpha_kernel(const T* __restrict__ w_re, const T* __restrict__ w_im,
                 const T* __restrict__ ls2, int64_t n, T* __restrict__ out_la, float threshold,
                 T* __restrict__ out_mask) {
  constexpr int V = Elem<T>::kVec;
  constexpr int kKind = kCplx ? CPLXK_KL_CPLX_ARD : CPLXK_KL_REAL_ARD;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t done = 0;
  if constexpr (kVecOK) {
    const int64_t nvec = n / V;
    for (int64_t i = tid; i < nvec; i += nthreads) {
      Vec16<T> a, b, c, la, mk;
      a.load(w_re + i * V);
      if constexpr (kCplx) b.load(w_im + i * V);
      c.load(ls2 + i * V);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        la.v[j] = log_alpha_of<kKind>(a.v[j], kCplx ? b.v[j] : 0.f, c.v[j]);
        mk.v[j] = la.v[j] <= threshold ? 1.f : 0.f;
      }
      if (out_la) la.store(out_la + i * V);
      if (out_mask) mk.store(out_mask + i * V);
    }
    done = nvec * V;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    float la = log_alpha_of<kKind>(Elem<T>::to_f(w_re[i]), kCplx ? Elem<T>::to_f(w_im[i]) : 0.f,
                                   Elem<T>::to_f(ls2[i]));
    if (out_la) out_la[i] = Elem<T>::from_f(la);
    if (out_mask) out_mask[i] = Elem<T>::from_f(la <= threshold ? 1.f : 0.f);
  }
}

// ------------------------------------------------------------------ host side
static int sm_count_cached() {   // per device (knobs.cuh); 148 if the query fails
  int sms = 148;
  if (current_device_sm_count(&sms) != CPLXK_OK || sms < 1) sms = 148;
  return sms;
}

template <typename T, int kKind>
static int launch_kl(const void* w_re, const void* w_im, const void* ls2, int64_t n, void* out_elem,
                     float* out_sum, double scale, KlWorkspace* ws, cudaStream_t st,
                     void* out_mask = nullptr, float threshold = 0.f) {
  constexpr int V = Elem<T>::kVec;
  const bool vec_ok = aligned16(w_re) && aligned16(ls2) && (w_im == nullptr || aligned16(w_im)) &&
                      (out_elem == nullptr || aligned16(out_elem)) &&
                      (out_mask == nullptr || aligned16(out_mask));
  auto mk = static_cast<T*>(out_mask);
  int64_t work = (n + V - 1) / V;
  int64_t want = (work + 2 * kKlThreads - 1) / (2 * kKlThreads);
  int grid = static_cast<int>(want < 1 ? 1 : (want > 8 * sm_count_cached() ? 8 * sm_count_cached() : want));
  if (grid > kKlMaxBlocks) grid = kKlMaxBlocks;
  auto a = static_cast<const T*>(w_re);
  auto b = static_cast<const T*>(w_im);
  auto c = static_cast<const T*>(ls2);
  auto o = static_cast<T*>(out_elem);
  if (vec_ok)
    kl_kernel<T, kKind, true><<<grid, kKlThreads, 0, st>>>(a, b, c, n, o, out_sum, scale, ws, mk, threshold);
  else
    kl_kernel<T, kKind, false><<<grid, kKlThreads, 0, st>>>(a, b, c, n, o, out_sum, scale, ws, mk, threshold);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

template <typename T>
static int dispatch_kl(int kind, const void* w_re, const void* w_im, const void* ls2, int64_t n,
                       void* out_elem, float* out_sum, double scale, KlWorkspace* ws,
                       cudaStream_t st, void* out_mask = nullptr, float threshold = 0.f) {
  switch (kind) {
    case CPLXK_KL_REAL_VD:
      return launch_kl<T, CPLXK_KL_REAL_VD>(w_re, w_im, ls2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
    case CPLXK_KL_REAL_ARD:
      return launch_kl<T, CPLXK_KL_REAL_ARD>(w_re, w_im, ls2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
    case CPLXK_KL_CPLX_VD:
      return launch_kl<T, CPLXK_KL_CPLX_VD>(w_re, w_im, ls2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
    case CPLXK_KL_CPLX_ARD:
      return launch_kl<T, CPLXK_KL_CPLX_ARD>(w_re, w_im, ls2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
    case CPLXK_KL_CPLX_VD_APPROX:
      return launch_kl<T, CPLXK_KL_CPLX_VD_APPROX>(w_re, w_im, ls2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
    case CPLXK_KL_CPLX_VD_SCALEFREE:
      return launch_kl<T, CPLXK_KL_CPLX_VD_SCALEFREE>(w_re, w_im, ls2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
  }
  return CPLXK_ERR_BADARG;
}

}  // namespace cplxk

using namespace cplxk;

extern "C" size_t cplxk_kl_workspace_bytes(void) { return sizeof(KlWorkspace); }

static int kl_entry(int kind, const void* w_re, const void* w_im, const void* log_sigma2, int64_t n,
                    int dtype, void* out_elem, float* out_sum, double scale, void* workspace,
                    size_t workspace_bytes, void* stream, void* out_mask, float threshold);

extern "C" int cplxk_kl(int kind, const void* w_re, const void* w_im, const void* log_sigma2,
                        int64_t n, int dtype, void* out_elem, float* out_sum, double scale,
                        void* workspace, size_t workspace_bytes, void* stream) {
  return kl_entry(kind, w_re, w_im, log_sigma2, n, dtype, out_elem, out_sum, scale, workspace,
                  workspace_bytes, stream, nullptr, 0.f);
}

extern "C" int cplxk_kl_mask(int kind, const void* w_re, const void* w_im, const void* log_sigma2,
                             int64_t n, int dtype, float threshold, void* out_mask, float* out_sum,
                             double scale, void* workspace, size_t workspace_bytes, void* stream) {
  if (!out_mask && n != 0) return CPLXK_ERR_BADARG;
  return kl_entry(kind, w_re, w_im, log_sigma2, n, dtype, nullptr, out_sum, scale, workspace,
                  workspace_bytes, stream, out_mask, threshold);
}

static int kl_entry(int kind, const void* w_re, const void* w_im, const void* log_sigma2, int64_t n,
                    int dtype, void* out_elem, float* out_sum, double scale, void* workspace,
                    size_t workspace_bytes, void* stream, void* out_mask, float threshold) {
  if (n == 0) {  // empty layer: the sum of nothing
    if (out_sum) CPLXK_CUDA_TRY(cudaMemsetAsync(out_sum, 0, sizeof(float), static_cast<cudaStream_t>(stream)));
    return (kind < 0 || kind >= kKlKinds) ? CPLXK_ERR_BADARG : CPLXK_OK;
  }
  if (!w_re || !log_sigma2 || n < 0 || (!out_elem && !out_sum && !out_mask)) return CPLXK_ERR_BADARG;
  const bool cplx = kl_kind_is_cplx(kind);
  if (kind < 0 || kind >= kKlKinds || cplx != (w_im != nullptr)) return CPLXK_ERR_BADARG;
  if (out_sum) {
    if (!workspace || workspace_bytes < sizeof(KlWorkspace)) return CPLXK_ERR_WORKSPACE;
    if (!aligned16(workspace)) return CPLXK_ERR_ALIGN;
  }
  auto st = static_cast<cudaStream_t>(stream);
  auto ws = static_cast<KlWorkspace*>(workspace);
  if (dtype == CPLXK_F32)
    return dispatch_kl<float>(kind, w_re, w_im, log_sigma2, n, out_elem, out_sum, scale, ws, st, out_mask, threshold);
  if (dtype == CPLXK_BF16)
    return dispatch_kl<__nv_bfloat16>(kind, w_re, w_im, log_sigma2, n, out_elem, out_sum, scale, ws, st, out_mask,
                                      threshold);
  return CPLXK_ERR_BADARG;
}

// ---- guard of the fused KL by-product ------------------------------------------------------
// The forward's pre-pass leaves a layer's KL sum behind; penalties() hands it out if the
// parameters are unchanged.  Host-side bookkeeping (tensor identity, autograd version counters)
// cannot see writes through `param.data`, so the hand-out is checked ON THE DEVICE: a 64-bit
// fingerprint of a strided sample of the parameters (the first 8 entries of every row of every
// plane: any whole-tensor edit -- mul_, clamp_, copy_, fill_ -- changes it) is taken right after
// the forward and again when the sum is asked for.  One tiny launch each, no synchronisation.
// One row head per thread (first 8 entries of each plane), many small blocks: the loads are
// scattered 32-byte sectors a row pitch apart, and one SM cannot keep enough of them in flight (a
// single 1024-thread block needed 45 us for 4096 rows).  Block sums go to ws->fp, the last block
// (ticket) finishes.
constexpr int kGuardThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kGuardThreads)
kl_guard_kernel(const T* __restrict__ w_re, const T* __restrict__ w_im, const T* __restrict__ ls2,
                int64_t N, int64_t K, KlWorkspace* __restrict__ ws, unsigned long long* __restrict__ fp_out,
                const unsigned long long* __restrict__ fp_ref, const float* __restrict__ fused,
                float* __restrict__ out, int* __restrict__ stale_flag) {
  __shared__ unsigned long long sh[kGuardThreads / 32];
  __shared__ bool is_last;
  const int ncol = K < 8 ? static_cast<int>(K) : 8;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec = ncol == 8 && (K % 4) == 0 && std::is_same<T, float>::value && al16(w_re) &&
                   al16(ls2) && (!w_im || al16(w_im));
  unsigned long long acc = 0ull;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kGuardThreads + threadIdx.x; r < N;
       r += static_cast<int64_t>(gridDim.x) * kGuardThreads) {
    if (vec) {
      const float* a = reinterpret_cast<const float*>(w_re) + r * K;
      const float* b = w_im ? reinterpret_cast<const float*>(w_im) + r * K : nullptr;
      const float* c = reinterpret_cast<const float*>(ls2) + r * K;
      float4 va[2], vb[2], vc[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        va[h] = __ldg(reinterpret_cast<const float4*>(a + 4 * h));
        vb[h] = b ? __ldg(reinterpret_cast<const float4*>(b + 4 * h)) : make_float4(0.f, 0.f, 0.f, 0.f);
        vc[h] = __ldg(reinterpret_cast<const float4*>(c + 4 * h));
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float fa[4] = {va[h].x, va[h].y, va[h].z, va[h].w};
        const float fb[4] = {vb[h].x, vb[h].y, vb[h].z, vb[h].w};
        const float fc[4] = {vc[h].x, vc[h].y, vc[h].z, vc[h].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += fingerprint_elem(fa[j], fb[j], b != nullptr, fc[j], r, 4 * h + j);
      }
    } else {
      for (int j = 0; j < ncol; ++j) {
        const int64_t i = r * K + j;
        acc += fingerprint_elem(Elem<T>::to_f(w_re[i]), w_im ? Elem<T>::to_f(w_im[i]) : 0.f, w_im != nullptr,
                                Elem<T>::to_f(ls2[i]), r, j);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long total = 0ull;
    for (int w = 0; w < kGuardThreads / 32; ++w) total += sh[w];
    if (total != 0ull) atomicAdd(&ws->fp, total);
    __threadfence();
    is_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  const unsigned long long fp = *reinterpret_cast<volatile unsigned long long*>(&ws->fp);
  ws->fp = 0ull, ws->ticket = 0;     // restore the workspace for the next call on this stream
  if (fp_ref == nullptr) {           // record
    fp_out[0] = fp;
    return;
  }
  const bool same = fp == fp_ref[0];
  out[0] = same ? fused[0] : __int_as_float(0x7fc00000);
  if (!same && stale_flag) {
    *reinterpret_cast<volatile int*>(stale_flag) = 1;
    __threadfence_system();
  }
}

extern "C" int cplxk_kl_guard(const void* w_re, const void* w_im, const void* log_sigma2, int64_t N,
                              int64_t K, int dtype, void* fp_out, const void* fp_ref,
                              const float* fused_sum, float* out_sum, int* stale_flag,
                              void* kl_workspace, size_t kl_workspace_bytes, void* stream) {
  if (!w_re || !log_sigma2 || N < 0 || K < 0) return CPLXK_ERR_BADARG;
  if (fp_ref ? (!fused_sum || !out_sum) : !fp_out) return CPLXK_ERR_BADARG;
  if (!kl_workspace || kl_workspace_bytes < sizeof(KlWorkspace)) return CPLXK_ERR_WORKSPACE;
  if (!aligned16(kl_workspace)) return CPLXK_ERR_ALIGN;
  auto st = static_cast<cudaStream_t>(stream);
  auto ws = static_cast<KlWorkspace*>(kl_workspace);
  auto fo = static_cast<unsigned long long*>(fp_out);
  auto fr = static_cast<const unsigned long long*>(fp_ref);
  int64_t want = (N + kGuardThreads - 1) / kGuardThreads;
  const int grid = static_cast<int>(want < 1 ? 1 : (want > 1024 ? 1024 : want));
  // Launched as a programmatic dependent: behind the persistent forward GEMM (which triggers
  // its dependents once the pre-pass in front of it is complete) the check runs under the GEMM's
  // mainloop instead of behind its last tile; behind any other kernel it starts when that kernel
  // has finished, like an ordinary launch.  It never calls griddepcontrol.wait: nothing it reads
  // is written by the kernel right before it.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid)), cfg.blockDim = dim3(kGuardThreads);
  cfg.dynamicSmemBytes = 0, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = knobs().pdl ? 1 : 0;
  if (dtype == CPLXK_F32) {
    CPLXK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kl_guard_kernel<float>, static_cast<const float*>(w_re),
                                      static_cast<const float*>(w_im), static_cast<const float*>(log_sigma2), N,
                                      K, ws, fo, fr, fused_sum, out_sum, stale_flag));
  } else if (dtype == CPLXK_BF16) {
    CPLXK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kl_guard_kernel<__nv_bfloat16>,
                                      static_cast<const __nv_bfloat16*>(w_re),
                                      static_cast<const __nv_bfloat16*>(w_im),
                                      static_cast<const __nv_bfloat16*>(log_sigma2), N, K, ws, fo, fr, fused_sum,
                                      out_sum, stale_flag));
  } else {
    return CPLXK_ERR_BADARG;
  }
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

template <typename T>
static int launch_log_alpha(const void* w_re, const void* w_im, const void* ls2, int64_t n,
                            void* out_la, float thr, void* out_mask, cudaStream_t st) {
  constexpr int V = Elem<T>::kVec;
  const bool vec_ok = aligned16(w_re) && aligned16(ls2) && (!w_im || aligned16(w_im)) &&
                      (!out_la || aligned16(out_la)) && (!out_mask || aligned16(out_mask));
  int64_t work = (n + V - 1) / V;
  int64_t want = (work + 255) / 256;
  int grid = static_cast<int>(want < 1 ? 1 : (want > 16 * sm_count_cached() ? 16 * sm_count_cached() : want));
  auto a = static_cast<const T*>(w_re);
  auto b = static_cast<const T*>(w_im);
  auto c = static_cast<const T*>(ls2);
  auto o = static_cast<T*>(out_la);
  auto m = static_cast<T*>(out_mask);
  if (w_im) {
    if (vec_ok) log_alpha_kernel<T, true, true><<<grid, 256, 0, st>>>(a, b, c, n, o, thr, m);
    else log_alpha_kernel<T, true, false><<<grid, 256, 0, st>>>(a, b, c, n, o, thr, m);
  } else {
    if (vec_ok) log_alpha_kernel<T, false, true><<<grid, 256, 0, st>>>(a, b, c, n, o, thr, m);
    else log_alpha_kernel<T, false, false><<<grid, 256, 0, st>>>(a, b, c, n, o, thr, m);
  }
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

extern "C" int cplxk_log_alpha(const void* w_re, const void* w_im, const void* log_sigma2,
                               int64_t n, int dtype, void* out_log_alpha, float threshold,
                               void* out_mask, void* stream) {
  if (n == 0) return CPLXK_OK;
  if (!w_re || !log_sigma2 || n < 0 || (!out_log_alpha && !out_mask)) return CPLXK_ERR_BADARG;
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == CPLXK_F32)
    return launch_log_alpha<float>(w_re, w_im, log_sigma2, n, out_log_alpha, threshold, out_mask, st);
  if (dtype == CPLXK_BF16)
    return launch_log_alpha<__nv_bfloat16>(w_re, w_im, log_sigma2, n, out_log_alpha, threshold, out_mask, st);
  return CPLXK_ERR_BADARG;
}
