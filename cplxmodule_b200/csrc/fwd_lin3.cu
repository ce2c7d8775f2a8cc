// Persistent CTA-pair (cta_group::2) complex / real affine map  y = x W^T + b  on 16-bit operands
// (cplxmodule/cplx.py:634-648: re = x_re U^T - x_im V^T, im = x_re V^T + x_im U^T, + bias).
//
// Same machinery as fwd_tc3.cu without the variational part: fp32 planes arrive as per-row
// power-of-two scaled fp16 copies (pre-pass vd_prepare_f16_kernel with the variance operands
// switched off) and run on kind::f16 at twice the tf32 rate; bf16 planes are consumed as they are.
// With only two accumulators (256 TMEM columns) the accumulators are DOUBLE BUFFERED: the
// epilogue of tile i drains one half of TMEM while the MMAs of tile i+1 fill the other, so the
// tensor pipe never waits for a drain.  Used by eval-mode forwards and by every gradient GEMM of
// the backward pass (dx = g conj(W), dW = g^T conj(x), dq, dE -- ops.py:_gemm).
//
// Warps (320 threads / CTA): 0 = TMA producer, 1 = MMA issuer (leader CTA), 2..9 = epilogue.
#include <cstdlib>
#include <type_traits>

#include "epilogue.cuh"
#include "knobs.cuh"
#include "prep_row.cuh"
#include "ptx.cuh"
#include "tc3_common.cuh"

namespace cplxk {

template <typename OutT, bool kCplx>
struct Lin3Cfg {
  static constexpr int BN = 128, BK = 64, KSTEPS = 4;
  static constexpr int A_TILE = 128 * 128, B_HALF = 64 * 128;
  static constexpr int NA = kCplx ? 2 : 1;
  static constexpr int OFF_A0 = 0, OFF_A1 = A_TILE, OFF_B0 = NA * A_TILE, OFF_B1 = OFF_B0 + B_HALF;
  static constexpr int STAGE_BYTES = NA * (A_TILE + B_HALF);   // 48 KB complex, 24 KB real
  static constexpr int EPI_WARPS = 8;
  static constexpr int AUX_BYTES = 4096;
  static constexpr int AVAIL = 227 * 1024 - 1024 - AUX_BYTES;
  static constexpr int STAGES = AVAIL / STAGE_BYTES > 8 ? 8 : AVAIL / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
  static constexpr int ACC_COLS = NA * BN;                     // one accumulator set
  static constexpr int TMEM_COLS = 2 * ACC_COLS;               // two sets: 512 / 256
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

struct Lin3Params {
  int64_t M, N, K;
  int tiles_m2, tiles_n;
  int f16;
  const float* sx;
  const float* sw;
  const void *b_re, *b_im;
  void *y_re, *y_im;
};

template <typename OutT, bool kCplx>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
lin_tc3_kernel(const __grid_constant__ CUtensorMap tm_xr, const __grid_constant__ CUtensorMap tm_xi,
               const __grid_constant__ CUtensorMap tm_wr, const __grid_constant__ CUtensorMap tm_wi,
               const Lin3Params p) {
  using C = Lin3Cfg<OutT, kCplx>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t aux = base + C::STAGES * C::STAGE_BYTES;
  // aux: full[8] empty[8] accum_full[2] tmem_empty[2] tmem_slot | colvec[2][3][128] floats at +1024
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_accum = aux + 128, bar_tfree = aux + 144;
  const uint32_t tmem_slot = aux + 160;
  uint8_t* aux_ptr = smem + C::STAGES * C::STAGE_BYTES;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(aux_ptr + 160);
  float* colvec = reinterpret_cast<float*>(aux_ptr + 1024);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.tiles_m2 * p.tiles_n;
  const int num_kb = static_cast<int>((p.K + C::BK - 1) / C::BK);

  auto decode_tile = [&](int t, int& tile_m, int& tile_n) {
    raster_tile(t, p.tiles_m2, p.tiles_n, 6, tile_m, tile_n);
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_xr);
    ptx::prefetch_tensormap(&tm_wr);
    if constexpr (kCplx) {
      ptx::prefetch_tensormap(&tm_xi);
      ptx::prefetch_tensormap(&tm_wi);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_accum + 8 * b, 1);
      ptx::mbar_init(bar_tfree + 8 * b, 2 * C::EPI_WARPS);   // only the leader's is used
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (whole warp, elected issue)
    const bool elected = ptx::elect_one();
    int s = 0;
    uint32_t ph = 0;
    ptx::grid_dep_wait();   // fp32 planes: the operands come from the pre-pass launched just before
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int tile_m, tile_n;
      decode_tile(t, tile_m, tile_n);
      const int32_t m0 = tile_m * 256 + static_cast<int32_t>(rank) * 128;
      const int32_t nb0 = tile_n * C::BN + static_cast<int32_t>(rank) * 64;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t fb = bar_full + 8 * s;
        const uint32_t st = base + s * C::STAGE_BYTES;
        const int32_t k0 = kb * C::BK;
        if (elected) {
          if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * C::STAGE_BYTES);
          ptx::tma_load_2d_pair(st + C::OFF_A0, &tm_xr, fb, k0, m0);
          if constexpr (kCplx) ptx::tma_load_2d_pair(st + C::OFF_A1, &tm_xi, fb, k0, m0);
          ptx::tma_load_2d_pair(st + C::OFF_B0, &tm_wr, fb, k0, nb0);
          if constexpr (kCplx) ptx::tma_load_2d_pair(st + C::OFF_B1, &tm_wi, fb, k0, nb0);
        }
        __syncwarp();
        if (++s == C::STAGES) s = 0, ph ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer: leader CTA, whole warp in the loop, elected issue
    if (leader) {
      const bool elected = ptx::elect_one();
      const uint32_t fmt = p.f16 ? 0u : 1u;
      const uint32_t idesc = ptx::make_idesc_f16(fmt, 256, C::BN, false);
      const uint32_t idesc_na = ptx::make_idesc_f16(fmt, 256, C::BN, true);
      int s = 0;
      uint32_t ph = 0, it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const uint32_t buf = it & 1u, tph = (it >> 1) & 1u;
        ptx::mbar_wait_cluster(bar_tfree + 8 * buf, tph ^ 1u);   // this half drained by both CTAs
        ptx::tcgen05_fence_after();
        const uint32_t t_re = tmem_base + buf * C::ACC_COLS, t_im = t_re + C::BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t st = base + s * C::STAGE_BYTES;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tcgen05_fence_after();
          const uint64_t a0 = ptx::make_kmajor_desc<128>(st + C::OFF_A0);
          const uint64_t a1 = ptx::make_kmajor_desc<128>(st + C::OFF_A1);
          const uint64_t b0 = ptx::make_kmajor_desc<128>(st + C::OFF_B0);
          const uint64_t b1 = ptx::make_kmajor_desc<128>(st + C::OFF_B1);
          const uint32_t acc0 = kb > 0 ? 1u : 0u;
          if (elected) {
#pragma unroll
            for (int k = 0; k < C::KSTEPS; ++k) {
              const uint32_t acc = k > 0 ? 1u : acc0;
              const uint32_t off = k * 32;
              ptx::umma_ss_pair<true>(t_re, ptx::desc_advance(a0, off), ptx::desc_advance(b0, off), idesc, acc);
              if constexpr (kCplx) {
                ptx::umma_ss_pair<true>(t_re, ptx::desc_advance(a1, off), ptx::desc_advance(b1, off), idesc_na, 1u);
                ptx::umma_ss_pair<true>(t_im, ptx::desc_advance(a0, off), ptx::desc_advance(b1, off), idesc, acc);
                ptx::umma_ss_pair<true>(t_im, ptx::desc_advance(a1, off), ptx::desc_advance(b0, off), idesc, 1u);
              }
            }
            ptx::umma_commit_pair(bar_empty + 8 * s);
          }
          __syncwarp();
          if (++s == C::STAGES) s = 0, ph ^= 1u;
        }
        if (elected) ptx::umma_commit_pair(bar_accum + 8 * buf);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ 8 epilogue warps
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int te = threadIdx.x - 64;
    uint32_t it = 0;
    ptx::grid_dep_wait();   // row / column scales are written by the pre-pass
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const uint32_t buf = it & 1u, tph = (it >> 1) & 1u;
      int tile_m, tile_n;
      decode_tile(t, tile_m, tile_n);
      const int32_t m0 = tile_m * 256 + static_cast<int32_t>(rank) * 128;
      const int32_t n0 = tile_n * C::BN;
      const int64_t m = static_cast<int64_t>(m0) + quarter * 32 + lane;
      const int64_t nb = static_cast<int64_t>(n0) + half * 64;
      float* cv = colvec + buf * 3 * 128;
      if (te < 128) {
        const int64_t n = static_cast<int64_t>(n0) + te;
        const bool ok = n < p.N;
        const OutT* br = static_cast<const OutT*>(p.b_re);
        const OutT* bi = static_cast<const OutT*>(p.b_im);
        cv[te] = (ok && br) ? Elem<OutT>::to_f(__ldg(br + n)) : 0.f;
        cv[128 + te] = (kCplx && ok && bi) ? Elem<OutT>::to_f(__ldg(bi + n)) : 0.f;
        cv[256 + te] = (ok && p.sw) ? __ldg(p.sw + n) : 1.f;
      }
      const float sxm = (p.sx && m < p.M) ? __ldg(p.sx + m) : 1.f;
      ptx::named_bar_sync(1, 32 * C::EPI_WARPS);

      ptx::mbar_wait(bar_accum + 8 * buf, tph);
      ptx::tcgen05_fence_after();
      const uint32_t lane_base = tmem_base + buf * C::ACC_COLS +
                                 (static_cast<uint32_t>(quarter * 32) << 16) + half * 64;
      const float* cvb = cv + half * 64;
      const bool row_ok = m < p.M;
      OutT* yr = static_cast<OutT*>(p.y_re) + m * p.N + nb;
      OutT* yi = kCplx ? static_cast<OutT*>(p.y_im) + m * p.N + nb : nullptr;
      const bool vec = (p.N - nb) >= 64 && ((reinterpret_cast<uintptr_t>(yr) & 31u) == 0) &&
                       (!kCplx || (reinterpret_cast<uintptr_t>(yi) & 31u) == 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = c * 16;
        uint32_t r_re[16], r_im[16];
        ptx::tmem_ld_32x32b_x16(lane_base + col, r_re);
        if constexpr (kCplx) ptx::tmem_ld_32x32b_x16(lane_base + C::BN + col, r_im);
        ptx::tmem_ld_wait();
        if (c == 3) {   // last TMEM read of this warp: this half may be overwritten
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(bar_tfree + 8 * buf, 0));
        }
        float f_re[16], f_im[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float sc = sxm * cvb[256 + col + j];
          f_re[j] = fmaf(__uint_as_float(r_re[j]), sc, cvb[col + j]);
          f_im[j] = kCplx ? fmaf(__uint_as_float(r_im[j]), sc, cvb[128 + col + j]) : 0.f;
        }
        if (row_ok && nb + col < p.N) {
          const int nvalid = (p.N - (nb + col)) < 16 ? static_cast<int>(p.N - (nb + col)) : 16;
          store_run16<OutT>(yr + col, f_re, vec, nvalid);
          if constexpr (kCplx) store_run16<OutT>(yi + col, f_im, vec, nvalid);
        }
      }
    }
    ptx::tcgen05_fence_before();
  }

  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------ host side
struct Lin3Operands {
  const void *a_re, *a_im, *b_re, *b_im;   // 16-bit planes [M,K] / [N,K]
  const float *sx, *sw;
  bool f16;
};

template <typename OutT, bool kCplx>
static int launch_lin3(const Lin3Operands& o, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                       cudaStream_t st) {
  using C = Lin3Cfg<OutT, kCplx>;
  CUtensorMap tm_xr, tm_xi, tm_wr, tm_wi;
  const CUtensorMapDataType dt_op = o.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc;
  if ((rc = map2d(&tm_xr, dt_op, 2, o.a_re, M, K, C::BK, 128))) return rc;
  if ((rc = map2d(&tm_wr, dt_op, 2, o.b_re, N, K, C::BK, 64))) return rc;
  tm_xi = tm_xr, tm_wi = tm_wr;
  if (kCplx) {
    if ((rc = map2d(&tm_xi, dt_op, 2, o.a_im, M, K, C::BK, 128))) return rc;
    if ((rc = map2d(&tm_wi, dt_op, 2, o.b_im, N, K, C::BK, 64))) return rc;
  }
  Lin3Params p;
  p.M = M, p.N = N, p.K = K;
  p.tiles_m2 = static_cast<int>((M + 255) / 256);
  p.tiles_n = static_cast<int>((N + C::BN - 1) / C::BN);
  p.f16 = o.f16 ? 1 : 0;
  p.sx = o.sx, p.sw = o.sw;
  p.b_re = ep.b_re, p.b_im = ep.b_im, p.y_re = ep.y_re, p.y_im = ep.y_im;
  const int64_t pairs = static_cast<int64_t>(p.tiles_m2) * p.tiles_n;
  if (pairs > 0x3fffffff) return CPLXK_ERR_UNSUPPORTED;
  int sm_count = 0;
  if ((rc = current_device_sm_count(&sm_count))) return rc;
  int64_t clusters = (sm_count - sm_reserve()) / 2;
  if (clusters < 1) clusters = 1;
  if (clusters > pairs) clusters = pairs;
  auto kern = lin_tc3_kernel<OutT, kCplx>;
  CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters)), cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = (o.f16 && knobs().pdl) ? 1 : 0;   // f16: launched right after the pre-pass
  CPLXK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm_xr, tm_xi, tm_wr, tm_wi, p));
  return CPLXK_OK;
}

// workspace of the fp32-plane path: xh_re, xh_im [M,K]; wh_re, wh_im [N,K] fp16; isx [M], isw [N]
size_t fwd_lin3_workspace_bytes(int dtype, int64_t M, int64_t N, int64_t K) {
  if (dtype != CPLXK_F32) return 0;
  return 2 * align256(static_cast<size_t>(M) * K * 2) + 2 * align256(static_cast<size_t>(N) * K * 2) +
         align256(static_cast<size_t>(M) * 4) + align256(static_cast<size_t>(N) * 4);
}

bool fwd_lin3_supported(int64_t M, int64_t N, int64_t K) { return M > 128 && K % 8 == 0 && K >= 8 && N >= 1; }

int vd_prepare_f16_launch(bool cplx, const PrepArgs& a, const KlFuse& kl, cudaStream_t st);

int fwd_lin3_f32(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                 void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t st,
                 const void* w_mask) {
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t xb = align256(static_cast<size_t>(M) * K * 2), wb = align256(static_cast<size_t>(N) * K * 2);
  auto f = [](const void* p) { return static_cast<const float*>(p); };
  PrepArgs a{};
  a.x_re = f(x_re), a.x_im = f(x_im), a.w_re = f(w_re), a.w_im = f(w_im), a.ls2 = nullptr, a.w_mask = f(w_mask);
  a.M = M, a.N = N, a.K = K;
  a.xh_re = reinterpret_cast<__half*>(ws);
  a.xh_im = reinterpret_cast<__half*>(ws + xb);
  a.wh_re = reinterpret_cast<__half*>(ws + 2 * xb);
  a.wh_im = reinterpret_cast<__half*>(ws + 2 * xb + wb);
  a.q = nullptr, a.e = nullptr;
  a.isx = reinterpret_cast<float*>(ws + 2 * xb + 2 * wb);
  a.isw = reinterpret_cast<float*>(ws + 2 * xb + 2 * wb + align256(static_cast<size_t>(M) * 4));
  a.kl_kind = -1, a.kl_row0 = 0, a.kl_row1 = 0;
  const KlFuse none{-1, nullptr, nullptr, 0, -1, nullptr};
  int rc = vd_prepare_f16_launch(cplx, a, none, st);
  if (rc) return rc;
  Lin3Operands o{a.xh_re, a.xh_im, a.wh_re, a.wh_im, a.isx, a.isw, true};
  return cplx ? launch_lin3<float, true>(o, M, N, K, ep, st) : launch_lin3<float, false>(o, M, N, K, ep, st);
}

int fwd_lin3_bf16(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                  int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t st) {
  Lin3Operands o{x_re, x_im, w_re, w_im, nullptr, nullptr, false};
  return cplx ? launch_lin3<__nv_bfloat16, true>(o, M, N, K, ep, st)
              : launch_lin3<__nv_bfloat16, false>(o, M, N, K, ep, st);
}

}  // namespace cplxk
