// CTA-pair (cta_group::2) version of the fused variational forward, used when the derived
// operands come precomputed through the workspace (see fwd_tc.cu for the algorithm).
//
// Two CTAs on the two SMs of a TPC share one 256 x 128 output tile: each loads the A tiles
// (x_re, x_im, |x|^2) of ITS 128 rows and only HALF (64 rows) of every B tile
// (U, V, exp(log_sigma2)); the leader issues M = 256 tcgen05.mma that read both halves.
// Per CTA a pipeline stage is 3 x 16 KB + 3 x 8 KB = 72 KB instead of 96 KB, which lets the
// 128-byte-swizzle / 3-stage schedule of the plain complex GEMM fit, and the B operands
// cross L2 -> SM once per pair instead of once per CTA.
#include <cstdlib>
#include <type_traits>

#include "epilogue.cuh"
#include "ptx.cuh"

namespace cplxk {

// kMixVar (fp32 planes only): the two operands of the variance GEMM, |x|^2 and exp(log_sigma2),
// are all-positive and only feed sqrt(s2) * eps, so they travel as bf16 (round-to-nearest in
// the pre-pass; the rounding errors average out over K) and that GEMM runs as kind::f16 at
// twice the tf32 rate: 4.5 instead of 5 tf32-MMA-equivalents per k-step, smaller stages.
// kOp: 0 = operands are the planes themselves (tf32 / bf16), 1 = kMixVar, 2 = fp32 planes whose
// mean operands arrive as per-row-scaled fp16 and variance operands as bf16 (pre-pass
// vd_prepare_f16_kernel, fwd_tc3.cu): everything runs as kind::f16, the epilogue undoes the scales.
template <typename T, bool kCplx, int kOp>
struct Tc2Cfg {
  static constexpr bool kMixVar = kOp == 1;
  static constexpr bool kHalfOps = kOp == 2;
  static_assert(kOp == 0 || std::is_same<T, float>::value, "derived 16-bit operands: fp32 planes");
  static constexpr bool kBF16 = std::is_same<T, __nv_bfloat16>::value || kHalfOps;   // kind::f16 MMAs
  static constexpr int BM = 128, BN = 128;            // per-CTA rows; N of the pair tile
  static constexpr int BK = kHalfOps ? 64 : 128 / static_cast<int>(sizeof(T));
  static constexpr int KSTEPS = 4;
  static constexpr int A_TILE = 128 * 128, B_HALF = 64 * 128;
  static constexpr int NA = kCplx ? 2 : 1;
  static constexpr int Q_TILE = kMixVar ? A_TILE / 2 : A_TILE;     // bf16: 128 rows x 64 B (SW64)
  static constexpr int E_HALF = kMixVar ? B_HALF / 2 : B_HALF;
  static constexpr int VAR_SWZ = kMixVar ? 64 : 128;
  static constexpr int VAR_KSTEPS = kMixVar ? 2 : 4;               // K = 16 bf16 per MMA
  static constexpr int OFF_A0 = 0, OFF_A1 = A_TILE, OFF_Q = NA * A_TILE;
  static constexpr int OFF_B0 = OFF_Q + Q_TILE, OFF_B1 = OFF_B0 + B_HALF;
  static constexpr int OFF_E = OFF_B0 + NA * B_HALF;
  static constexpr int STAGE_BYTES = OFF_E + E_HALF;               // 72 KB complex (60 KB mixed)
  static constexpr int STAGES = (227 * 1024 - 2048) / STAGE_BYTES > 6 ? 6 : (227 * 1024 - 2048) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2048;
  static constexpr int NACC = NA + 1;
  static constexpr int TMEM_COLS = NACC * BN <= 256 ? 256 : 512;
  static constexpr int THREADS = 320;                  // TMA, MMA, 8 epilogue warps
};

struct Tc2Params {
  int64_t M, N, K;
  int tiles_m2, tiles_n;   // tiles of 256 rows, 128 columns
  int dbg;                 // CPLXK_DBG: 1 = MMAs without operand loads, 2 = loads without MMAs, 3 = no mainloop
  const float* sx;         // kOp == 2: inverse row scales of x [M] and of W [N]
  const float* sw;
  EpiParams ep;
};

template <typename T, bool kCplx, int kOp>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_xr, const __grid_constant__ CUtensorMap tm_xi,
               const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_wr,
               const __grid_constant__ CUtensorMap tm_wi, const __grid_constant__ CUtensorMap tm_e,
               const Tc2Params p) {
  using C = Tc2Cfg<T, kCplx, kOp>;
  constexpr bool kMixVar = C::kMixVar;
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
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  // tile pair: grouped raster over (256-row tiles) x (128-column tiles)
  int tile_m, tile_n;
  {
    constexpr int kGroup = 6;
    const int t = blockIdx.x >> 1;
    const int per_group = kGroup * p.tiles_n;
    const int g = t / per_group;
    const int first_m = g * kGroup;
    const int gsize = (p.tiles_m2 - first_m) < kGroup ? (p.tiles_m2 - first_m) : kGroup;
    const int r = t - g * per_group;
    tile_m = first_m + r % gsize;
    tile_n = r / gsize;
  }
  const int32_t m0 = tile_m * 256 + static_cast<int32_t>(rank) * 128;   // this CTA's rows
  const int32_t n0 = tile_n * C::BN;
  const int32_t nb0 = n0 + static_cast<int32_t>(rank) * 64;             // this CTA's half of B
  const int num_kb = p.dbg == 3 ? 0 : static_cast<int>((p.K + C::BK - 1) / C::BK);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_xr);
    ptx::prefetch_tensormap(&tm_wr);
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_e);
    if constexpr (kCplx) {
      ptx::prefetch_tensormap(&tm_xi);
      ptx::prefetch_tensormap(&tm_wi);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);    // only the leader's is used
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / TMA
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------------------------------------- TMA producer (one per CTA of the pair)
    // (whole warp in the loop, one elected lane issues: see the MMA issuer)
    if (p.dbg != 1) {
      const bool elected = ptx::elect_one();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t fb = bar_full + 8 * s;
        const uint32_t st = base + s * C::STAGE_BYTES;
        const int32_t k0 = kb * C::BK;
        if (elected) {
          if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * C::STAGE_BYTES);   // both CTAs' bytes
          ptx::tma_load_2d_pair(st + C::OFF_A0, &tm_xr, fb, k0, m0);
          if constexpr (kCplx) ptx::tma_load_2d_pair(st + C::OFF_A1, &tm_xi, fb, k0, m0);
          ptx::tma_load_2d_pair(st + C::OFF_Q, &tm_q, fb, k0, m0);
          ptx::tma_load_2d_pair(st + C::OFF_B0, &tm_wr, fb, k0, nb0);
          if constexpr (kCplx) ptx::tma_load_2d_pair(st + C::OFF_B1, &tm_wi, fb, k0, nb0);
          ptx::tma_load_2d_pair(st + C::OFF_E, &tm_e, fb, k0, nb0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: leader only
    // The whole warp walks the loop, one elected lane issues: warp-uniform control flow keeps the
    // descriptors in uniform registers (~3 issue slots per tcgen05.mma instead of ~12).
    if (leader) {
      const bool elected = ptx::elect_one();
      // a/b format field: 0 = fp16 (scaled operands), 1 = bf16, 2 = tf32
      constexpr uint32_t fmt_fix = C::kHalfOps ? (1u << 7) | (1u << 10) : 0u;   // clears bf16 -> fp16
      constexpr uint32_t idesc = ptx::make_idesc<C::kBF16>(256, C::BN, false, false) ^ fmt_fix;
      constexpr uint32_t idesc_na = ptx::make_idesc<C::kBF16>(256, C::BN, true, false) ^ fmt_fix;
      constexpr uint32_t idesc_v = ptx::make_idesc<C::kBF16>(256, C::BN, false, false);
      const uint32_t t_re = tmem_base, t_im = tmem_base + C::BN, t_s2 = tmem_base + C::NA * C::BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        const uint32_t st = base + s * C::STAGE_BYTES;
        if (p.dbg != 1) ptx::mbar_wait(bar_full + 8 * s, ph);
        ptx::tcgen05_fence_after();
        if (p.dbg == 2) {
          if (elected) ptx::umma_commit_pair(bar_empty + 8 * s);
          __syncwarp();
          continue;
        }
        const uint64_t a0 = ptx::make_kmajor_desc<128>(st + C::OFF_A0);
        const uint64_t a1 = ptx::make_kmajor_desc<128>(st + C::OFF_A1);
        const uint64_t aq = ptx::make_kmajor_desc<C::VAR_SWZ>(st + C::OFF_Q);
        const uint64_t b0 = ptx::make_kmajor_desc<128>(st + C::OFF_B0);
        const uint64_t b1 = ptx::make_kmajor_desc<128>(st + C::OFF_B1);
        const uint64_t be = ptx::make_kmajor_desc<C::VAR_SWZ>(st + C::OFF_E);
        if (elected) {
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k) {
          const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
          const uint32_t off = k * 32;
          ptx::umma_ss_pair<C::kBF16>(t_re, ptx::desc_advance(a0, off), ptx::desc_advance(b0, off), idesc, acc);
          if constexpr (kCplx) {
            ptx::umma_ss_pair<C::kBF16>(t_re, ptx::desc_advance(a1, off), ptx::desc_advance(b1, off), idesc_na, 1u);
            ptx::umma_ss_pair<C::kBF16>(t_im, ptx::desc_advance(a0, off), ptx::desc_advance(b1, off), idesc, acc);
            ptx::umma_ss_pair<C::kBF16>(t_im, ptx::desc_advance(a1, off), ptx::desc_advance(b0, off), idesc, 1u);
          }
          if constexpr (!kMixVar)
            ptx::umma_ss_pair<C::kBF16>(t_s2, ptx::desc_advance(aq, off), ptx::desc_advance(be, off), idesc_v, acc);
        }
        if constexpr (kMixVar) {
          constexpr uint32_t idesc_bf = ptx::make_idesc<true>(256, C::BN, false, false);
#pragma unroll
          for (int k = 0; k < C::VAR_KSTEPS; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            ptx::umma_ss_pair<true>(t_s2, ptx::desc_advance(aq, k * 32), ptx::desc_advance(be, k * 32),
                                    idesc_bf, acc);
          }
        }
        ptx::umma_commit_pair(bar_empty + 8 * s);   // frees the stage in BOTH CTAs
        }
        __syncwarp();
      }
      if (elected) ptx::umma_commit_pair(bar_accum);
      __syncwarp();
    }
  } else {
    // -------------------------------------- 8 epilogue warps: noise prefetch, then drain
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int64_t m = static_cast<int64_t>(m0) + quarter * 32 + lane;
    const int64_t nb = static_cast<int64_t>(n0) + half * 64;
    float nre[64], nim[kCplx ? 64 : 1];
    noise_prefetch<T, kCplx, 64>(p.ep, m, nb, nre, nim);
    [[maybe_unused]] float sxm = 1.f;
    if constexpr (C::kHalfOps) sxm = m < p.M ? __ldg(p.sx + m) : 1.f;
    ptx::mbar_wait(bar_accum, 0);
    ptx::tcgen05_fence_after();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t r_re[8], r_im[8], r_s2[8];
      ptx::tmem_ld_32x32b_x8(lane_base + c * 8, r_re);
      if constexpr (kCplx) ptx::tmem_ld_32x32b_x8(lane_base + C::BN + c * 8, r_im);
      ptx::tmem_ld_32x32b_x8(lane_base + C::NA * C::BN + c * 8, r_s2);
      ptx::tmem_ld_wait();
      float f_re[8], f_im[8], f_s2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f_re[j] = __uint_as_float(r_re[j]);
        f_im[j] = kCplx ? __uint_as_float(r_im[j]) : 0.f;
        f_s2[j] = __uint_as_float(r_s2[j]);
        if constexpr (C::kHalfOps) {   // undo the power-of-two row scales of x and W (exact)
          const int64_t n = nb + c * 8 + j;
          const float sc = sxm * (n < p.N ? __ldg(p.sw + n) : 1.f);
          f_re[j] *= sc;
          f_im[j] *= sc;
        }
      }
      epilogue_finish<T, kCplx, 8>(p.ep, m, nb + c * 8, f_re, f_im, f_s2, &nre[c * 8],
                                   &nim[kCplx ? c * 8 : 0]);
    }
    ptx::tcgen05_fence_before();
  }

  ptx::cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still use it
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------ host side
// plane [rows, K] row-major of T -> box {box_k elements of K, box_rows}, swizzle = box bytes
template <typename T>
static int plane_map2(CUtensorMap* out, const void* ptr, int64_t rows, int64_t K, int box_rows,
                      bool round_tf32, int box_k = 128 / static_cast<int>(sizeof(T))) {
  auto enc = tensor_map_encode_fn();
  if (!enc) return CPLXK_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * sizeof(T)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMapDataType dt = std::is_same<T, float>::value
                               ? (round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)
                               : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = box_k * sizeof(T) == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                         : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CPLXK_OK : CPLXK_ERR_CUDA;
}

// fp16 planes written by the pre-pass: box {64 elements of K, rows}, 128B swizzle
static int half_map2(CUtensorMap* out, const void* ptr, int64_t rows, int64_t K, int box_rows) {
  auto enc = tensor_map_encode_fn();
  if (!enc) return CPLXK_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CPLXK_OK : CPLXK_ERR_CUDA;
}

template <typename T, bool kCplx, int kOp>
static int launch_tc2(const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                      const void* q, const void* e, int64_t M, int64_t N, int64_t K,
                      const EpiParams& ep, cudaStream_t st, const float* sx = nullptr,
                      const float* sw = nullptr) {
  using C = Tc2Cfg<T, kCplx, kOp>;
  constexpr bool kMixVar = kOp == 1;
  CUtensorMap tm_xr, tm_xi, tm_q, tm_wr, tm_wi, tm_e;
  int rc;
  if (kOp == 2) {
    if ((rc = half_map2(&tm_xr, x_re, M, K, 128))) return rc;
    if ((rc = half_map2(&tm_wr, w_re, N, K, 64))) return rc;
    if ((rc = plane_map2<__nv_bfloat16>(&tm_q, q, M, K, 128, false))) return rc;
    if ((rc = plane_map2<__nv_bfloat16>(&tm_e, e, N, K, 64, false))) return rc;
    tm_xi = tm_xr, tm_wi = tm_wr;
    if (kCplx) {
      if ((rc = half_map2(&tm_xi, x_im, M, K, 128))) return rc;
      if ((rc = half_map2(&tm_wi, w_im, N, K, 64))) return rc;
    }
  } else {
  if ((rc = plane_map2<T>(&tm_xr, x_re, M, K, 128, true))) return rc;
  if ((rc = plane_map2<T>(&tm_wr, w_re, N, K, 64, true))) return rc;
  if (kMixVar) {   // q / E were written as bf16 by the pre-pass: 32 K-elements = 64-byte rows
    if ((rc = plane_map2<__nv_bfloat16>(&tm_q, q, M, K, 128, false, 32))) return rc;
    if ((rc = plane_map2<__nv_bfloat16>(&tm_e, e, N, K, 64, false, 32))) return rc;
  } else {
    if ((rc = plane_map2<T>(&tm_q, q, M, K, 128, false))) return rc;
    if ((rc = plane_map2<T>(&tm_e, e, N, K, 64, false))) return rc;
  }
  tm_xi = tm_xr, tm_wi = tm_wr;
  if (kCplx) {
    if ((rc = plane_map2<T>(&tm_xi, x_im, M, K, 128, true))) return rc;
    if ((rc = plane_map2<T>(&tm_wi, w_im, N, K, 64, true))) return rc;
  }
  }
  Tc2Params p;
  p.sx = sx, p.sw = sw;
  p.M = M, p.N = N, p.K = K;
  p.tiles_m2 = static_cast<int>((M + 255) / 256);
  p.tiles_n = static_cast<int>((N + C::BN - 1) / C::BN);
  p.ep = ep;
  const char* dbg_env = std::getenv("CPLXK_DBG");
  p.dbg = dbg_env ? std::atoi(dbg_env) : 0;
  const int64_t pairs = static_cast<int64_t>(p.tiles_m2) * p.tiles_n;
  if (2 * pairs > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;
  auto kern = fwd_tc2_kernel<T, kCplx, kOp>;
  CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  kern<<<static_cast<unsigned>(2 * pairs), C::THREADS, C::SMEM_BYTES, st>>>(tm_xr, tm_xi, tm_q, tm_wr,
                                                                          tm_wi, tm_e, p);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

int fwd_tc2_dispatch(int dtype, bool cplx, bool mix_var, const void* x_re, const void* x_im,
                     const void* w_re, const void* w_im, const void* q, const void* e, int64_t M,
                     int64_t N, int64_t K, const EpiParams& ep, cudaStream_t st) {
  if (dtype == CPLXK_F32) {
    if (mix_var) {
      if (cplx) return launch_tc2<float, true, 1>(x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
      return launch_tc2<float, false, 1>(x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
    }
    if (cplx) return launch_tc2<float, true, 0>(x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
    return launch_tc2<float, false, 0>(x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
  }
  if (dtype == CPLXK_BF16) {
    if (cplx) return launch_tc2<__nv_bfloat16, true, 0>(x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
    return launch_tc2<__nv_bfloat16, false, 0>(x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
  }
  return CPLXK_ERR_BADARG;
}

// fp32 planes, operands already converted to per-row-scaled fp16 (+ bf16 variance operands)
int fwd_tc2_half_dispatch(bool cplx, const void* xh_re, const void* xh_im, const void* wh_re,
                          const void* wh_im, const void* q, const void* e, const float* sx,
                          const float* sw, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                          cudaStream_t st) {
  if (cplx) return launch_tc2<float, true, 2>(xh_re, xh_im, wh_re, wh_im, q, e, M, N, K, ep, st, sx, sw);
  return launch_tc2<float, false, 2>(xh_re, xh_im, wh_re, wh_im, q, e, M, N, K, ep, st, sx, sw);
}

}  // namespace cplxk
