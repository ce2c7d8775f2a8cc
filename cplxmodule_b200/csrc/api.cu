// C-ABI entry points (include/cplxk.h): argument validation, path selection,
// error reporting.  No torch types, no allocation, asynchronous on `stream`.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "epilogue.cuh"
#include "knobs.cuh"

namespace cplxk {

static thread_local cudaError_t g_last_cuda = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda = e; }

int fwd_simt_dispatch(int dtype, bool cplx, bool vd, const void* x_re, const void* x_im,
                      const void* w_re, const void* w_im, const void* ls2, int64_t M, int64_t N,
                      int64_t K, const EpiParams& ep, cudaStream_t st);
bool fwd_tc_supported(int dtype, bool cplx, const void* x_re, const void* x_im, const void* w_re,
                      const void* w_im, const void* ls2, int64_t M, int64_t N, int64_t K);
int fwd_tc_dispatch(int dtype, bool cplx, bool vd, int swz, bool f16_ok, const void* x_re,
                    const void* x_im, const void* w_re, const void* w_im, const void* ls2,
                    void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                    cudaStream_t st, const KlFuse& kl);
bool fwd_tc_fuses_kl(int dtype, bool f16_ok, int64_t M, int64_t N, int64_t K);
// persistent double-buffered affine map on 16-bit operands (fwd_lin3.cu)
size_t fwd_lin3_workspace_bytes(int dtype, int64_t M, int64_t N, int64_t K);
bool fwd_lin3_supported(int64_t M, int64_t N, int64_t K);
int fwd_lin3_f32(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                 void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t st,
                 const void* w_mask);
int fwd_lin3_bf16(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                  int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t st);
size_t fwd_tc_workspace_bytes(int dtype, int64_t M, int64_t N, int64_t K);

// ---- process-level switches: the environment is read once, on first use
static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e && e[0] ? std::atoi(e) : dflt;
}

const Knobs& knobs() {
  static const Knobs k = [] {
    Knobs v;
    const int swz = env_int("CPLXK_TC_SWIZZLE", 0);
    v.tc_swizzle = (swz == 64 || swz == 128) ? swz : 0;
    v.raster = env_int("CPLXK_RASTER", 6);
    if (v.raster < 1) v.raster = 6;
    v.lin3 = env_int("CPLXK_LIN3", 1) != 0;
    v.tma_raw_f32 = env_int("CPLXK_TMA_RAW_F32", 0) == 1;
    v.conv_pair = env_int("CPLXK_CONV_PAIR", 1) != 0;
    v.conv_persistent = env_int("CPLXK_CONV_NONPERSISTENT", 0) != 1;
    v.conv_row = env_int("CPLXK_CONV_ROW", 1) != 0;
    v.combine_flat = env_int("CPLXK_COMBINE_FLAT", 1) != 0;
    v.conv_real_pair = env_int("CPLXK_CONV_REAL_PAIR", 1) != 0;
    v.conv_overlap = env_int("CPLXK_CONV_OVERLAP", 4);
    v.conv_amax_pass = env_int("CPLXK_CONV_AMAX_PASS", 0) == 1;
    v.pdl = env_int("CPLXK_PDL", 1) != 0;
    v.prep_prefetch = env_int("CPLXK_PREP_PREFETCH", 0) != 0;   // measured: slower (profiles/prep_prefetch_ab_r2.jsonl)
#ifdef CPLXK_DEBUG
    v.dbg = env_int("CPLXK_DBG", 0);
#else
    v.dbg = 0;
#endif
    return v;
  }();
  return k;
}

static std::atomic<int> g_sm_reserve{-1};
int sm_reserve() {
  int v = g_sm_reserve.load(std::memory_order_relaxed);
  if (v < 0) {
    v = env_int("CPLXK_SM_RESERVE", 0);
    if (v < 0) v = 0;
    g_sm_reserve.store(v, std::memory_order_relaxed);
  }
  return v;
}

// ---- per-device facts (a process may drive several GPUs: nothing is cached per process)
static constexpr int kMaxDevices = 64;
static std::atomic<int> g_sm_count[kMaxDevices];   // 0: unknown
static std::atomic<int> g_cc_major[kMaxDevices];   // 0: unknown

int current_device_sm_count(int* sms) {
  int dev = 0;
  CPLXK_CUDA_TRY(cudaGetDevice(&dev));
  int v = (dev >= 0 && dev < kMaxDevices) ? g_sm_count[dev].load(std::memory_order_relaxed) : 0;
  if (!v) {
    CPLXK_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < kMaxDevices) g_sm_count[dev].store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return CPLXK_OK;
}

static int check_arch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_last_cuda_error(cudaGetLastError());
    return CPLXK_ERR_CUDA;
  }
  int major = (dev >= 0 && dev < kMaxDevices) ? g_cc_major[dev].load(std::memory_order_relaxed) : 0;
  if (!major) {
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      set_last_cuda_error(cudaGetLastError());
      return CPLXK_ERR_CUDA;
    }
    if (dev >= 0 && dev < kMaxDevices) g_cc_major[dev].store(major, std::memory_order_relaxed);
  }
  return major == 10 ? CPLXK_OK : CPLXK_ERR_ARCH;
}

static NoiseParams make_noise(int mode, uint64_t seed, uint64_t offset, uint32_t threads, bool cplx) {
  NoiseParams np;
  np.mode = mode;
  np.seed_lo = static_cast<uint32_t>(seed);
  np.seed_hi = static_cast<uint32_t>(seed >> 32);
  np.ctr_base = offset >> 2;
  np.threads = threads ? threads : 1u;
  // torch: randn(...) / sqrt(2) on CUDA multiplies by the fp32 reciprocal of fp32(sqrt 2)
  np.scale = cplx ? (1.0f / static_cast<float>(1.4142135623730951)) : 1.0f;
  return np;
}

static int forward_common(bool vd, const void* x_re, const void* x_im, const void* w_re,
                          const void* w_im, const void* b_re, const void* b_im, const void* ls2,
                          const void* eps_re, const void* eps_im, int noise, uint64_t seed,
                          uint64_t offset, uint32_t threads, void* y_re, void* y_im, int64_t M,
                          int64_t N, int64_t K, int dtype, int math, void* s2_out, void* workspace,
                          size_t workspace_bytes, void* stream, KlFuse kl = KlFuse{-1, nullptr, nullptr, 0, -1, nullptr},
                          int* kl_done = nullptr, const void* w_mask = nullptr, int* mask_done = nullptr) {
  // w_mask (plain affine map only): applied by the path taken when *mask_done is set to 1 on
  // return; otherwise the caller must have passed already-masked weights
  if (mask_done) *mask_done = 0;
  if (kl_done) *kl_done = 0;
  if (M < 0 || N < 0 || K < 0) return CPLXK_ERR_BADARG;
  if (M == 0 || N == 0) return CPLXK_OK;     // empty batch / layer: nothing to write (pointers may be null)
  if (!x_re || !w_re || !y_re) return CPLXK_ERR_BADARG;
  const bool cplx = (x_im != nullptr);
  if (cplx != (w_im != nullptr) || cplx != (y_im != nullptr)) return CPLXK_ERR_BADARG;
  if ((b_re != nullptr) && cplx && !b_im) return CPLXK_ERR_BADARG;
  if (dtype != CPLXK_F32 && dtype != CPLXK_BF16) return CPLXK_ERR_BADARG;
  if (math < CPLXK_MATH_AUTO || math > CPLXK_MATH_TENSOR_TF32) return CPLXK_ERR_BADARG;
  if (vd) {
    if (!ls2) return CPLXK_ERR_BADARG;
    if (noise < CPLXK_NOISE_INJECT || noise > CPLXK_NOISE_PHILOX_FAST) return CPLXK_ERR_BADARG;
    if (noise == CPLXK_NOISE_INJECT && (!eps_re || (cplx && !eps_im))) return CPLXK_ERR_BADARG;
    if (noise == CPLXK_NOISE_PHILOX_TORCH && (threads == 0 || (offset & 3u))) return CPLXK_ERR_BADARG;
  }
  int rc = check_arch();
  if (rc) return rc;
  if (M == 0 || N == 0) return CPLXK_OK;

  EpiParams ep;
  ep.b_re = b_re, ep.b_im = b_im, ep.eps_re = eps_re, ep.eps_im = eps_im;
  ep.y_re = y_re, ep.y_im = y_im, ep.s2_out = vd ? s2_out : nullptr;
  ep.M = M, ep.N = N, ep.plane_elems = M * N;
  ep.noise = make_noise(vd ? noise : CPLXK_NOISE_INJECT, seed, offset, threads, cplx);

  auto st = static_cast<cudaStream_t>(stream);
  const bool tc_ok = fwd_tc_supported(dtype, cplx, x_re, x_im, w_re, w_im, vd ? ls2 : nullptr, M, N, K);
  if (math == CPLXK_MATH_TENSOR && !tc_ok) return CPLXK_ERR_ALIGN;
  // F32 planes: per-row scaled fp16 operands unless the caller asked for tf32 ones
  const bool f16_ok = math != CPLXK_MATH_TENSOR_TF32;
  if (math != CPLXK_MATH_SIMT && tc_ok) {
    int swz = knobs().tc_swizzle;
    if (!swz) swz = (cplx && vd) ? 64 : 128;
    if (workspace) {
      if (!aligned16(workspace)) return CPLXK_ERR_ALIGN;
      const size_t need = vd ? fwd_tc_workspace_bytes(dtype, M, N, K) : fwd_lin3_workspace_bytes(dtype, M, N, K);
      if (workspace_bytes < need) return CPLXK_ERR_WORKSPACE;
    }
    if (!vd && fwd_lin3_supported(M, N, K)) {
      // plain affine map: persistent CTA-pair kernel with double-buffered accumulators; fp32 planes
      // need the workspace for their row-scaled fp16 copies (MATH_TENSOR_TF32 / no workspace: tf32 below)
      if (dtype == CPLXK_F32 && workspace && K >= 64 && f16_ok) {
        if (mask_done) *mask_done = 1;
        return fwd_lin3_f32(cplx, x_re, x_im, w_re, w_im, workspace, M, N, K, ep, st, w_mask);
      }
      if (dtype == CPLXK_BF16 && knobs().lin3)   // CPLXK_LIN3=0: one tile per CTA (fwd_tc.cu)
        return fwd_lin3_bf16(cplx, x_re, x_im, w_re, w_im, M, N, K, ep, st);
    }
    const bool fuse = vd && workspace && kl.kind >= 0 && kl.sum && kl.ws && fwd_tc_fuses_kl(dtype, f16_ok, M, N, K);
    if (!fuse) kl.kind = -1;
    rc = fwd_tc_dispatch(dtype, cplx, vd, swz, f16_ok, x_re, x_im, w_re, w_im, ls2, vd ? workspace : nullptr,
                         M, N, K, ep, st, kl);
    if (rc == CPLXK_OK && fuse && kl_done) *kl_done = 1;
    return rc;
  }
  return fwd_simt_dispatch(dtype, cplx, vd, x_re, x_im, w_re, w_im, ls2, M, N, K, ep, st);
}

__global__ void randn_philox_torch_kernel(float* out, int64_t n, NoiseParams np) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 16;
  for (int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16; i0 < n;
       i0 += stride) {
    TorchNoiseCursor cur;
    cur.seek(static_cast<uint64_t>(i0), np.threads);
    for (int j = 0; j < 16 && i0 + j < n; ++j) out[i0 + j] = cur.next(np) * np.scale;
  }
}

}  // namespace cplxk

using namespace cplxk;

extern "C" int cplxk_abi_version(void) { return CPLXK_ABI_VERSION; }

extern "C" const char* cplxk_strerror(int status) {
  static thread_local char buf[256];
  switch (status) {
    case CPLXK_OK: return "ok";
    case CPLXK_ERR_BADARG: return "cplxk: bad argument (null pointer, bad enum or negative size)";
    case CPLXK_ERR_ALIGN: return "cplxk: pointer/pitch violates the 16-byte alignment contract of the tensor-core path";
    case CPLXK_ERR_ARCH: return "cplxk: current device is not compute capability 10.x (sm_100a kernels only; there is no CPU or other-arch path)";
    case CPLXK_ERR_CUDA:
      std::snprintf(buf, sizeof(buf), "cplxk: CUDA error: %s", cudaGetErrorString(g_last_cuda));
      return buf;
    case CPLXK_ERR_UNSUPPORTED: return "cplxk: unsupported configuration";
    case CPLXK_ERR_WORKSPACE: return "cplxk: workspace missing or too small";
  }
  return "cplxk: unknown status";
}

extern "C" int cplxk_set_sm_reserve(int n_sms) {
  if (n_sms < 0) return CPLXK_ERR_BADARG;
  g_sm_reserve.store(n_sms, std::memory_order_relaxed);
  return CPLXK_OK;
}

extern "C" int cplxk_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  CPLXK_CUDA_TRY(cudaGetDevice(&dev));
  if (sm_count) CPLXK_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) CPLXK_CUDA_TRY(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) CPLXK_CUDA_TRY(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return CPLXK_OK;
}

extern "C" int cplxk_linear_fwd(const void* x_re, const void* x_im, const void* w_re,
                                const void* w_im, const void* b_re, const void* b_im, void* y_re,
                                void* y_im, int64_t M, int64_t N, int64_t K, int dtype, int math,
                                void* stream) {
  return forward_common(false, x_re, x_im, w_re, w_im, b_re, b_im, nullptr, nullptr, nullptr, 0, 0,
                        0, 0, y_re, y_im, M, N, K, dtype, math, nullptr, nullptr, 0, stream);
}

extern "C" size_t cplxk_linear_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype) {
  if (M < 0 || N < 0 || K < 0) return 0;
  return fwd_lin3_workspace_bytes(dtype, M, N, K);
}

extern "C" int cplxk_linear_fwd_ws(const void* x_re, const void* x_im, const void* w_re,
                                   const void* w_im, const void* b_re, const void* b_im, void* y_re,
                                   void* y_im, int64_t M, int64_t N, int64_t K, int dtype, int math,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  return forward_common(false, x_re, x_im, w_re, w_im, b_re, b_im, nullptr, nullptr, nullptr, 0, 0,
                        0, 0, y_re, y_im, M, N, K, dtype, math, nullptr, workspace, workspace_bytes,
                        stream);
}

extern "C" int cplxk_linear_vd_fwd(const void* x_re, const void* x_im, const void* w_re,
                                   const void* w_im, const void* b_re, const void* b_im,
                                   const void* log_sigma2, const void* eps_re, const void* eps_im,
                                   int noise, uint64_t seed, uint64_t offset,
                                   uint32_t philox_threads, void* y_re, void* y_im, int64_t M,
                                   int64_t N, int64_t K, int dtype, int math, void* s2_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  return forward_common(true, x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise,
                        seed, offset, philox_threads, y_re, y_im, M, N, K, dtype, math, s2_out,
                        workspace, workspace_bytes, stream);
}

extern "C" int cplxk_linear_vd_fwd_kl(const void* x_re, const void* x_im, const void* w_re,
                                      const void* w_im, const void* b_re, const void* b_im,
                                      const void* log_sigma2, const void* eps_re, const void* eps_im,
                                      int noise, uint64_t seed, uint64_t offset,
                                      uint32_t philox_threads, void* y_re, void* y_im, int64_t M,
                                      int64_t N, int64_t K, int dtype, int math, void* s2_out,
                                      void* workspace, size_t workspace_bytes, int kl_kind,
                                      float* kl_sum, void* kl_workspace, size_t kl_workspace_bytes,
                                      int64_t kl_row_begin, int64_t kl_row_end, void* kl_event,
                                      void* kl_fingerprint, int* kl_done, void* stream) {
  if (kl_done) *kl_done = 0;
  KlFuse kl{-1, nullptr, nullptr, 0, -1, nullptr};
  if (kl_kind >= 0) {
    const bool cplx_kind = kl_kind >= CPLXK_KL_CPLX_VD;
    if (kl_kind > CPLXK_KL_CPLX_VD_SCALEFREE || cplx_kind != (w_im != nullptr) || !kl_sum || !kl_done)
      return CPLXK_ERR_BADARG;
    if (!kl_workspace || kl_workspace_bytes < cplxk_kl_workspace_bytes()) return CPLXK_ERR_WORKSPACE;
    if (!aligned16(kl_workspace)) return CPLXK_ERR_ALIGN;
    if (kl_row_begin < 0 || kl_row_end > N || (kl_row_end >= 0 && kl_row_end < kl_row_begin))
      return CPLXK_ERR_BADARG;
    kl = KlFuse{kl_kind, kl_sum, kl_workspace, kl_row_begin, kl_row_end, kl_event, kl_fingerprint};
  }
  return forward_common(true, x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise,
                        seed, offset, philox_threads, y_re, y_im, M, N, K, dtype, math, s2_out,
                        workspace, workspace_bytes, stream, kl, kl_done);
}

extern "C" int cplxk_linear_vd_fuses_kl(int64_t M, int64_t N, int64_t K, int dtype, int math) {
  return (math != CPLXK_MATH_SIMT && fwd_tc_fuses_kl(dtype, math != CPLXK_MATH_TENSOR_TF32, M, N, K)) ? 1 : 0;
}

// ---- fixed-sparsity layers: y = x (W * mask)^T + b
namespace cplxk {
int eltwise_mul2(const void* a0, const void* a1, const void* b, void* o0, void* o1, int64_t n, int dtype,
                 cudaStream_t st);   // bwd.cu: o0 = a0 * b, o1 = a1 * b (a1/o1 nullable), one launch
}

extern "C" size_t cplxk_linear_masked_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype) {
  if (M < 0 || N < 0 || K < 0) return 0;
  const size_t es = dtype == CPLXK_F32 ? 4 : 2;
  const size_t plane = (static_cast<size_t>(N) * K * es + 255) & ~static_cast<size_t>(255);
  const size_t base = (fwd_lin3_workspace_bytes(dtype, M, N, K) + 255) & ~static_cast<size_t>(255);
  return base + 2 * plane;
}

extern "C" int cplxk_linear_masked_fwd(const void* x_re, const void* x_im, const void* w_re,
                                       const void* w_im, const void* mask, const void* b_re,
                                       const void* b_im, void* y_re, void* y_im, int64_t M,
                                       int64_t N, int64_t K, int dtype, int math, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (!mask) return CPLXK_ERR_BADARG;
  if (!workspace || !aligned16(workspace)) return workspace ? CPLXK_ERR_ALIGN : CPLXK_ERR_WORKSPACE;
  if (M < 0 || N < 0 || K < 0) return CPLXK_ERR_BADARG;
  if (workspace_bytes < cplxk_linear_masked_workspace_bytes(M, N, K, dtype)) return CPLXK_ERR_WORKSPACE;
  if (M == 0 || N == 0) return CPLXK_OK;
  const size_t es = dtype == CPLXK_F32 ? 4 : 2;
  const size_t plane = (static_cast<size_t>(N) * K * es + 255) & ~static_cast<size_t>(255);
  const size_t base = (fwd_lin3_workspace_bytes(dtype, M, N, K) + 255) & ~static_cast<size_t>(255);
  // first choice: the mask rides along in the operand pre-pass
  int mask_done = 0;
  const bool fusable = dtype == CPLXK_F32 && math != CPLXK_MATH_SIMT && math != CPLXK_MATH_TENSOR_TF32 &&
                       K >= 64 && fwd_lin3_supported(M, N, K) &&
                       fwd_tc_supported(dtype, x_im != nullptr, x_re, x_im, w_re, w_im, mask, M, N, K);
  if (fusable) {
    int rc = forward_common(false, x_re, x_im, w_re, w_im, b_re, b_im, nullptr, nullptr, nullptr, 0, 0,
                            0, 0, y_re, y_im, M, N, K, dtype, math, nullptr, workspace, base, stream,
                            KlFuse{-1, nullptr, nullptr, 0, -1, nullptr}, nullptr, mask, &mask_done);
    if (rc != CPLXK_OK || mask_done) return rc;
    return CPLXK_ERR_UNSUPPORTED;   // unreachable by construction: the fused path was predicted
  }
  if (!x_re || !w_re || !y_re) return CPLXK_ERR_BADARG;
  if ((x_im != nullptr) != (w_im != nullptr)) return CPLXK_ERR_BADARG;
  if (dtype != CPLXK_F32 && dtype != CPLXK_BF16) return CPLXK_ERR_BADARG;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* m_re = ws + base;
  void* m_im = w_im ? ws + base + plane : nullptr;
  int rc = check_arch();
  if (rc) return rc;
  if (N > 0 && K > 0) {
    rc = eltwise_mul2(w_re, w_im, mask, m_re, m_im, N * K, dtype, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
  }
  return forward_common(false, x_re, x_im, m_re, m_im, b_re, b_im, nullptr, nullptr, nullptr, 0, 0, 0,
                        0, y_re, y_im, M, N, K, dtype, math, nullptr, base ? workspace : nullptr, base,
                        stream);
}

// operand pre-pass of the fp32-plane tensor-core path on its own (no GEMM launch)
namespace cplxk {
bool fwd_tc3_supported(int dtype, int64_t M, int64_t N, int64_t K);
int fwd_tc3_f32(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                const void* ls2, void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                cudaStream_t st, const KlFuse& kl);
}  // namespace cplxk

extern "C" int cplxk_linear_vd_prepare(const void* x_re, const void* x_im, const void* w_re,
                                       const void* w_im, const void* log_sigma2, int64_t M,
                                       int64_t N, int64_t K, int dtype, void* workspace,
                                       size_t workspace_bytes, int kl_kind, float* kl_sum,
                                       void* kl_workspace, size_t kl_workspace_bytes, void* stream) {
  if (!x_re || !w_re || !log_sigma2 || !workspace || M < 0 || N < 0 || K < 0) return CPLXK_ERR_BADARG;
  const bool cplx = x_im != nullptr;
  if (cplx != (w_im != nullptr)) return CPLXK_ERR_BADARG;
  int rc = check_arch();
  if (rc) return rc;
  if (dtype != CPLXK_F32 || K < 64 || !fwd_tc3_supported(dtype, M, N, K) ||
      !fwd_tc_supported(dtype, cplx, x_re, x_im, w_re, w_im, log_sigma2, M, N, K))
    return CPLXK_ERR_UNSUPPORTED;
  if (!aligned16(workspace)) return CPLXK_ERR_ALIGN;
  if (workspace_bytes < fwd_tc_workspace_bytes(dtype, M, N, K)) return CPLXK_ERR_WORKSPACE;
  KlFuse kl{-1, nullptr, nullptr, 0, -1, nullptr};
  if (kl_kind >= 0) {
    const bool cplx_kind = kl_kind >= CPLXK_KL_CPLX_VD;
    if (kl_kind > CPLXK_KL_CPLX_VD_SCALEFREE || cplx_kind != cplx || !kl_sum) return CPLXK_ERR_BADARG;
    if (!kl_workspace || kl_workspace_bytes < cplxk_kl_workspace_bytes()) return CPLXK_ERR_WORKSPACE;
    if (!aligned16(kl_workspace)) return CPLXK_ERR_ALIGN;
    kl = KlFuse{kl_kind, kl_sum, kl_workspace, 0, -1, nullptr};
  }
  EpiParams ep{};          // y_re == nullptr: stop after the pre-pass
  ep.M = M, ep.N = N;
  return fwd_tc3_f32(cplx, x_re, x_im, w_re, w_im, log_sigma2, workspace, M, N, K, ep,
                     static_cast<cudaStream_t>(stream), kl);
}

extern "C" size_t cplxk_linear_vd_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype) {
  if (M < 0 || N < 0 || K < 0) return 0;
  return fwd_tc_workspace_bytes(dtype, M, N, K);
}

extern "C" int cplxk_randn_philox_torch(float* out, int64_t n, uint64_t seed, uint64_t offset,
                                        uint32_t philox_threads, float scale, void* stream) {
  if (!out || n < 0 || philox_threads == 0 || (offset & 3u)) return CPLXK_ERR_BADARG;
  if (n == 0) return CPLXK_OK;
  NoiseParams np = make_noise(CPLXK_NOISE_PHILOX_TORCH, seed, offset, philox_threads, false);
  np.scale = scale;
  int64_t want = (n + 16 * 256 - 1) / (16 * 256);
  int grid = static_cast<int>(want > 4096 ? 4096 : want);
  randn_philox_torch_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, np);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}
