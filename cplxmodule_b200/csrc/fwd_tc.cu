// tcgen05 + TMA forward of the (variational) complex / real linear layer.
//
// One CTA computes a 128 x 128 tile of  y = mu + eps * sqrt(max(s2, 1e-8)):
//   re  += x_re U^T - x_im V^T         (cplxmodule/cplx.py:641, negation via the
//   im  += x_re V^T + x_im U^T          instruction descriptor's negate-A bit)  (:642)
//   s2  += (x_re^2 + x_im^2) exp(log_sigma2)^T   (nn/relevance/complex/base.py:50-54)
// i.e. up to five tcgen05.mma per 32-byte k-step into three fp32 accumulators
// held in TMEM (3 x 128 columns).  Operand planes are streamed by TMA into a
// ring of 128B/64B-swizzled K-major tiles; the two DERIVED operands (|x|^2 and
// exp(log_sigma2)) are produced in shared memory by the four transform warps
// from the tiles TMA just landed (generic-proxy stores + fence.proxy.async),
// so neither ever exists in HBM.  The same four warps then run the epilogue:
// tcgen05.ld -> bias -> Philox/injected noise -> global stores.
//
// Warp roles (192 threads): 0 = TMA producer, 1 = MMA issuer (+TMEM alloc),
// 2..5 = transform + epilogue (TMEM lane quarter = warp % 4).
#include <cstdlib>
#include <type_traits>

#include "epilogue.cuh"
#include "knobs.cuh"
#include "ptx.cuh"

namespace cplxk {

template <typename T, bool kCplx, bool kVD, bool kXform, int kSwz>
struct TcCfg {
  static constexpr bool kBF16 = std::is_same<T, __nv_bfloat16>::value;
  static constexpr int BM = 128, BN = 128;
  static constexpr int BK = kSwz / static_cast<int>(sizeof(T));  // elements of K per stage
  static constexpr int KSTEPS = kSwz / 32;                        // one UMMA consumes 32 B of K
  static constexpr int TILE_BYTES = 128 * kSwz;
  static constexpr int NA = kCplx ? 2 : 1;
  static constexpr int NB = kCplx ? 2 : 1;
  static constexpr int NTILES = NA + NB + (kVD ? 2 : 0);
  // stage layout: A0 [A1] [Q]  B0 [B1] [E]
  static constexpr int OFF_A0 = 0;
  static constexpr int OFF_A1 = TILE_BYTES;
  static constexpr int OFF_Q = NA * TILE_BYTES;
  static constexpr int OFF_B0 = (NA + (kVD ? 1 : 0)) * TILE_BYTES;
  static constexpr int OFF_B1 = OFF_B0 + TILE_BYTES;
  static constexpr int OFF_E = OFF_B0 + NB * TILE_BYTES;
  static constexpr int STAGE_BYTES = NTILES * TILE_BYTES;
  // kXform: |x|^2 and exp(log_sigma2) are made in smem by the transform warps (TMA brings
  // log_sigma2 into the E slot); otherwise both arrive precomputed through TMA.
  static constexpr int LOAD_BYTES = (kVD && !kXform) ? NTILES * TILE_BYTES
                                                     : (NA + NB + (kVD ? 1 : 0)) * TILE_BYTES;
  static constexpr int NACC = NA + (kVD ? 1 : 0);
  static constexpr int TMEM_COLS = NACC * BN <= 128 ? 128 : (NACC * BN <= 256 ? 256 : 512);
  static constexpr int AUX_BYTES = 1024;  // barriers + tmem slot
  static constexpr int SMEM_LIMIT = 227 * 1024;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - AUX_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
  static_assert(STAGES >= 2, "need a double-buffered ring at least");
  // With precomputed operands no warp has mainloop duties besides TMA / MMA issue, so the
  // epilogue is spread over EIGHT warps (two per TMEM lane quarter, 64 columns each) that
  // spend the mainloop generating their noise into registers.
  static constexpr bool kPrefetchNoise = kVD && !kXform;
  static constexpr int EPI_WARPS = kPrefetchNoise ? 8 : 4;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

// derived operands are rounded to the MMA operand precision with round-to-nearest (the
// bf16 store already does; fp32 tiles are consumed as tf32, which would otherwise truncate)
template <typename T>
__device__ __forceinline__ float round_operand(float v) {
  if constexpr (std::is_same<T, float>::value) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
  } else {
    return v;
  }
}

constexpr int kRasterGroup = 12;  // m-tiles per raster group: a wave covers ~12x12 tiles (L2 reuse)

struct TcParams {
  int64_t M, N, K;
  int tiles_m, tiles_n;
  EpiParams ep;
};

template <typename T, bool kCplx, bool kVD, bool kXform, int kSwz>
__global__ void __launch_bounds__((TcCfg<T, kCplx, kVD, kXform, kSwz>::THREADS), 1)
fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_xr, const __grid_constant__ CUtensorMap tm_xi,
              const __grid_constant__ CUtensorMap tm_wr, const __grid_constant__ CUtensorMap tm_wi,
              const __grid_constant__ CUtensorMap tm_ls, const __grid_constant__ CUtensorMap tm_q,
              const TcParams p) {
  using C = TcCfg<T, kCplx, kVD, kXform, kSwz>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms need 1024-B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t aux = base + C::STAGES * C::STAGE_BYTES;
  // aux: full[S] xf[S] empty[S] accum tmem_slot
  const uint32_t bar_full = aux;
  const uint32_t bar_xf = aux + 8 * C::STAGES;
  const uint32_t bar_empty = aux + 16 * C::STAGES;
  const uint32_t bar_accum = aux + 24 * C::STAGES;
  const uint32_t tmem_slot = bar_accum + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 24 * C::STAGES + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // grouped rasterisation: consecutive CTAs walk kRasterGroup m-tiles before the next n-tile
  int tile_m, tile_n;
  {
    const int t = blockIdx.x;
    const int per_group = kRasterGroup * p.tiles_n;
    const int g = t / per_group;
    const int first_m = g * kRasterGroup;
    const int gsize = (p.tiles_m - first_m) < kRasterGroup ? (p.tiles_m - first_m) : kRasterGroup;
    const int r = t - g * per_group;
    tile_m = first_m + r % gsize;
    tile_n = r / gsize;
  }
  const int32_t m0 = tile_m * C::BM;
  const int32_t n0 = tile_n * C::BN;
  const int num_kb = static_cast<int>((p.K + C::BK - 1) / C::BK);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_xr);
    ptx::prefetch_tensormap(&tm_wr);
    if constexpr (kCplx) {
      ptx::prefetch_tensormap(&tm_xi);
      ptx::prefetch_tensormap(&tm_wi);
    }
    if constexpr (kVD) ptx::prefetch_tensormap(&tm_ls);
    if constexpr (kVD && !kXform) ptx::prefetch_tensormap(&tm_q);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_xf + 8 * s, 4);  // one arrive per transform warp
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp in the loop, one elected lane issues: keeps addresses in uniform registers)
    {
      const bool elected = ptx::elect_one();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t fb = bar_full + 8 * s;
        const uint32_t st = base + s * C::STAGE_BYTES;
        const int32_t k0 = kb * C::BK;
        if (elected) {
          ptx::mbar_arrive_expect_tx(fb, C::LOAD_BYTES);
          ptx::tma_load_2d(st + C::OFF_A0, &tm_xr, fb, k0, m0);
          if constexpr (kCplx) ptx::tma_load_2d(st + C::OFF_A1, &tm_xi, fb, k0, m0);
          ptx::tma_load_2d(st + C::OFF_B0, &tm_wr, fb, k0, n0);
          if constexpr (kCplx) ptx::tma_load_2d(st + C::OFF_B1, &tm_wi, fb, k0, n0);
          if constexpr (kVD) ptx::tma_load_2d(st + C::OFF_E, &tm_ls, fb, k0, n0);  // log_sigma2 or exp() of it
          if constexpr (kVD && !kXform) ptx::tma_load_2d(st + C::OFF_Q, &tm_q, fb, k0, m0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    // The whole warp walks the loop, one elected lane issues: with warp-uniform control flow the
    // descriptors live in uniform registers (~3 issue slots per tcgen05.mma instead of ~12).
    {
      const bool elected = ptx::elect_one();
      constexpr uint32_t idesc = ptx::make_idesc<C::kBF16>(C::BM, C::BN, false, false);
      constexpr uint32_t idesc_na = ptx::make_idesc<C::kBF16>(C::BM, C::BN, true, false);
      const uint32_t t_re = tmem_base;
      const uint32_t t_im = tmem_base + C::BN;
      const uint32_t t_s2 = tmem_base + C::NA * C::BN;
      // The variance MMAs of k-block kb-1 are issued AFTER the mean MMAs of k-block kb: the
      // transform warps get a whole stage time to produce |x|^2 / exp(log_sigma2) and the
      // tensor pipe never waits for them.
      auto issue_var = [&](int kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        const uint32_t st = base + s * C::STAGE_BYTES;
        if constexpr (kXform) {
          ptx::mbar_wait(bar_xf + 8 * s, ph);
          ptx::tcgen05_fence_after();
        }
        const uint64_t aq = ptx::make_kmajor_desc<kSwz>(st + C::OFF_Q);
        const uint64_t be = ptx::make_kmajor_desc<kSwz>(st + C::OFF_E);
        if (elected) {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            const uint32_t off = k * 32;
            ptx::umma_ss<C::kBF16>(t_s2, ptx::desc_advance(aq, off), ptx::desc_advance(be, off), idesc, acc);
          }
          ptx::umma_commit(bar_empty + 8 * s);  // stage reusable once everything issued so far retires
        }
        __syncwarp();
      };
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        const uint32_t st = base + s * C::STAGE_BYTES;
        ptx::mbar_wait(bar_full + 8 * s, ph);
        ptx::tcgen05_fence_after();
        const uint64_t a0 = ptx::make_kmajor_desc<kSwz>(st + C::OFF_A0);
        const uint64_t a1 = ptx::make_kmajor_desc<kSwz>(st + C::OFF_A1);
        const uint64_t b0 = ptx::make_kmajor_desc<kSwz>(st + C::OFF_B0);
        const uint64_t b1 = ptx::make_kmajor_desc<kSwz>(st + C::OFF_B1);
        if (elected) {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            const uint32_t off = k * 32;
            ptx::umma_ss<C::kBF16>(t_re, ptx::desc_advance(a0, off), ptx::desc_advance(b0, off), idesc, acc);
            if constexpr (kCplx) {
              ptx::umma_ss<C::kBF16>(t_re, ptx::desc_advance(a1, off), ptx::desc_advance(b1, off), idesc_na, 1u);
              ptx::umma_ss<C::kBF16>(t_im, ptx::desc_advance(a0, off), ptx::desc_advance(b1, off), idesc, acc);
              ptx::umma_ss<C::kBF16>(t_im, ptx::desc_advance(a1, off), ptx::desc_advance(b0, off), idesc, 1u);
            }
          }
        }
        __syncwarp();
        if constexpr (kVD && kXform) {
          if (kb > 0) issue_var(kb - 1);
        } else if constexpr (kVD) {
          issue_var(kb);  // operands landed with the same TMA transaction group
        } else {
          if (elected) ptx::umma_commit(bar_empty + 8 * s);
          __syncwarp();
        }
      }
      if constexpr (kVD && kXform) issue_var(num_kb - 1);
      if (elected) ptx::umma_commit(bar_accum);  // accumulators complete
      __syncwarp();
    }
  } else {
    // -------------------------------------------------- transform, then epilogue
    [[maybe_unused]] const int tt = threadIdx.x - 64;  // 0..127
    if constexpr (kVD && kXform) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        uint8_t* st = smem + s * C::STAGE_BYTES;
        ptx::mbar_wait(bar_full + 8 * s, ph);
        // elementwise on identically-swizzled tiles: walk raw 16-byte chunks
#pragma unroll 4
        for (int i = tt; i < C::TILE_BYTES / 16; i += 128) {
          Vec16<T> a, q;
          a.load_smem(reinterpret_cast<const T*>(st + C::OFF_A0 + 16 * i));
          if constexpr (kCplx) {
            Vec16<T> b;
            b.load_smem(reinterpret_cast<const T*>(st + C::OFF_A1 + 16 * i));
#pragma unroll
            for (int j = 0; j < Vec16<T>::N; ++j) q.v[j] = round_operand<T>(fmaf(a.v[j], a.v[j], b.v[j] * b.v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < Vec16<T>::N; ++j) q.v[j] = round_operand<T>(a.v[j] * a.v[j]);
          }
          q.store(reinterpret_cast<T*>(st + C::OFF_Q + 16 * i));
          Vec16<T> e;
          e.load_smem(reinterpret_cast<const T*>(st + C::OFF_E + 16 * i));
#pragma unroll
          for (int j = 0; j < Vec16<T>::N; ++j) e.v[j] = round_operand<T>(__expf(e.v[j]));
          e.store(reinterpret_cast<T*>(st + C::OFF_E + 16 * i));
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_xf + 8 * s);
      }
    }

    if constexpr (C::kPrefetchNoise) {
      const int quarter = warp & 3;          // TMEM lanes this warp may touch
      const int half = (warp - 2) >> 2;      // which 64 columns of the tile
      const int64_t m = static_cast<int64_t>(m0) + quarter * 32 + lane;
      const int64_t nb = static_cast<int64_t>(n0) + half * 64;
      float nre[64], nim[kCplx ? 64 : 1];
      noise_prefetch<T, kCplx, 64>(p.ep, m, nb, nre, nim);  // overlaps the MMA mainloop
      ptx::mbar_wait(bar_accum, 0);
      ptx::tcgen05_fence_after();
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {  // 8 columns at a time keeps the live set under 168 registers
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
        }
        epilogue_finish<T, kCplx, 8>(p.ep, m, nb + c * 8, f_re, f_im, f_s2, &nre[c * 8],
                                     &nim[kCplx ? c * 8 : 0]);
      }
      ptx::tcgen05_fence_before();
    } else {
    ptx::mbar_wait(bar_accum, 0);
    ptx::tcgen05_fence_after();
    const int quarter = warp & 3;
    const int64_t m = static_cast<int64_t>(m0) + quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < C::BN / 16; ++c) {
      uint32_t r_re[16], r_im[16], r_s2[16];
      ptx::tmem_ld_32x32b_x16(lane_base + c * 16, r_re);
      if constexpr (kCplx) ptx::tmem_ld_32x32b_x16(lane_base + C::BN + c * 16, r_im);
      if constexpr (kVD) ptx::tmem_ld_32x32b_x16(lane_base + C::NA * C::BN + c * 16, r_s2);
      ptx::tmem_ld_wait();
      float f_re[16], f_im[16], f_s2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f_re[j] = __uint_as_float(r_re[j]);
        f_im[j] = kCplx ? __uint_as_float(r_im[j]) : 0.f;
        f_s2[j] = kVD ? __uint_as_float(r_s2[j]) : 0.f;
      }
      epilogue_run<T, kCplx, kVD, 16>(p.ep, m, static_cast<int64_t>(n0) + c * 16, f_re, f_im, f_s2);
    }
    ptx::tcgen05_fence_before();
    }
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
// plane [rows, K] row-major -> box {kSwz bytes of K, 128 rows}
template <typename T, int kSwz>
static int make_plane_map(CUtensorMap* out, const void* ptr, int64_t rows, int64_t K,
                          bool mma_operand = true) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encode_fn();
  if (!enc) return CPLXK_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * sizeof(T)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kSwz / sizeof(T)), 128u};
  cuuint32_t estr[2] = {1u, 1u};
  // fp32 planes feed kind::tf32 MMAs: the TFLOAT32 tensor-map type makes TMA deliver tf32
  // values (round-to-nearest) instead of leaving the truncation to the tensor core, which
  // would bias every product by about -2^-10 relative.
  const bool raw_f32 = knobs().tma_raw_f32;
  CUtensorMapDataType dt = std::is_same<T, float>::value
                               ? ((raw_f32 || !mma_operand) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                            : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32)
                               : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMapSwizzle sw = kSwz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (kSwz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CPLXK_OK : CPLXK_ERR_CUDA;
}

// ---- operand preparation for the variational forward (workspace variant) ----------------
// q = |x|^2 (per input row) and E = exp(log_sigma2) (per weight row), rounded to the MMA
// operand precision, written once to a caller-provided workspace.  HBM-bound elementwise pass;
// it takes the square/exp work (and its shared-memory traffic) out of the GEMM mainloop.
template <typename T, bool kCplx>
__global__ void __launch_bounds__(256)
vd_prepare_kernel(const T* __restrict__ x_re, const T* __restrict__ x_im, int64_t nx,
                  T* __restrict__ q, const T* __restrict__ ls2, int64_t nw, T* __restrict__ e) {
  constexpr int V = Elem<T>::kVec;
  const int64_t vx = nx / V, vw = nw / V;  // K * sizeof(T) % 16 == 0 => both exact
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < vx + vw;
       i += stride) {
    Vec16<T> o;
    if (i < vx) {
      Vec16<T> a;
      a.load(x_re + i * V);
      if constexpr (kCplx) {
        Vec16<T> b;
        b.load(x_im + i * V);
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = round_operand<T>(fmaf(a.v[j], a.v[j], b.v[j] * b.v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = round_operand<T>(a.v[j] * a.v[j]);
      }
      o.store(q + i * V);
    } else {
      const int64_t k = i - vx;
      Vec16<T> a;
      a.load(ls2 + k * V);
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = round_operand<T>(__expf(a.v[j]));
      o.store(e + k * V);
    }
  }
}

// persistent CTA-pair kernel on 16-bit operands (fwd_tc3.cu)
size_t fwd_tc3_workspace_bytes(int dtype, int64_t M, int64_t N, int64_t K);
bool fwd_tc3_supported(int dtype, int64_t M, int64_t N, int64_t K);
int fwd_tc3_f32(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                const void* ls2, void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                cudaStream_t st, const KlFuse& kl);
int fwd_tc3_bf16(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                 const void* q, const void* e, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                 cudaStream_t st);

size_t fwd_tc_workspace_bytes(int dtype, int64_t M, int64_t N, int64_t K) {
  const size_t es = dtype == CPLXK_F32 ? 4 : 2;
  const size_t qb = (static_cast<size_t>(M) * K * es + 255) & ~static_cast<size_t>(255);
  const size_t eb = (static_cast<size_t>(N) * K * es + 255) & ~static_cast<size_t>(255);
  const size_t v3 = fwd_tc3_workspace_bytes(dtype, M, N, K);
  return qb + eb > v3 ? qb + eb : v3;
}

constexpr int64_t kShortK = 128;   // below: tf32 operands everywhere (see launch_tc)

template <typename T, bool kCplx, bool kVD, bool kXform, int kSwz>
static int launch_tc(bool f16_ok, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                     const void* ls2, void* workspace, int64_t M, int64_t N, int64_t K,
                     const EpiParams& ep, cudaStream_t st, const KlFuse& kl) {
  using C = TcCfg<T, kCplx, kVD, kXform, kSwz>;
  CUtensorMap tm_xr, tm_xi, tm_wr, tm_wi, tm_ls, tm_q;
  int rc;
  if ((rc = make_plane_map<T, kSwz>(&tm_xr, x_re, M, K))) return rc;
  if ((rc = make_plane_map<T, kSwz>(&tm_wr, w_re, N, K))) return rc;
  tm_xi = tm_xr, tm_wi = tm_wr, tm_ls = tm_wr, tm_q = tm_xr;
  if (kCplx) {
    if ((rc = make_plane_map<T, kSwz>(&tm_xi, x_im, M, K))) return rc;
    if ((rc = make_plane_map<T, kSwz>(&tm_wi, w_im, N, K))) return rc;
  }
  if (kVD && kXform) {
    // log_sigma2 is exponentiated in the kernel: it must arrive with all its fp32 bits
    if ((rc = make_plane_map<T, kSwz>(&tm_ls, ls2, N, K, false))) return rc;
  } else if (kVD) {
    T* q = static_cast<T*>(workspace);
    const size_t qb = (static_cast<size_t>(M) * K * sizeof(T) + 255) & ~static_cast<size_t>(255);
    T* e = reinterpret_cast<T*>(static_cast<uint8_t*>(workspace) + qb);
    // fp32 planes: per-row-scaled fp16 operands on kind::f16 in the persistent CTA-pair kernel
    // of fwd_tc3.cu (MATH_TENSOR_TF32: tf32 operands on the one-tile-per-CTA kernel below); bf16
    // planes: the same persistent kernel on the planes as they are.
    // bf16 variance operands: their rounding errors (2^-9 each) average out over K; for a
    // few dozen terms they do not (K = 64, log_sigma2 spread over 14 units: 1.0e-3 measured),
    // so short reductions keep tf32 everywhere
    const bool short_k = K < kShortK;
    if (!short_k && fwd_tc3_supported(std::is_same<T, float>::value ? CPLXK_F32 : CPLXK_BF16, M, N, K)) {
      if constexpr (std::is_same<T, float>::value) {
        if (f16_ok)
          return fwd_tc3_f32(kCplx, x_re, x_im, w_re, w_im, ls2, workspace, M, N, K, ep, st, kl);
      } else {
        const int64_t work3 = (M * K + N * K) / Elem<T>::kVec;
        int sms = 148;
        if ((rc = current_device_sm_count(&sms))) return rc;
        const int grid3 = static_cast<int>(work3 / 256 + 1 > sms * 16 ? sms * 16 : work3 / 256 + 1);
        vd_prepare_kernel<T, kCplx><<<grid3, 256, 0, st>>>(static_cast<const T*>(x_re),
                                                         static_cast<const T*>(x_im), M * K, q,
                                                         static_cast<const T*>(ls2), N * K, e);
        CPLXK_CUDA_TRY(cudaGetLastError());
        return fwd_tc3_bf16(kCplx, x_re, x_im, w_re, w_im, q, e, M, N, K, ep, st);
      }
    }
    const int64_t work = (M * K + N * K) / Elem<T>::kVec;
    int sms = 148;
    if ((rc = current_device_sm_count(&sms))) return rc;
    const int grid = static_cast<int>(work / 256 + 1 > sms * 16 ? sms * 16 : work / 256 + 1);
    vd_prepare_kernel<T, kCplx><<<grid, 256, 0, st>>>(static_cast<const T*>(x_re),
                                                    static_cast<const T*>(x_im), M * K, q,
                                                    static_cast<const T*>(ls2), N * K, e);
    CPLXK_CUDA_TRY(cudaGetLastError());
    if ((rc = make_plane_map<T, kSwz>(&tm_q, q, M, K, false))) return rc;   // already rounded
    if ((rc = make_plane_map<T, kSwz>(&tm_ls, e, N, K, false))) return rc;
  }

  TcParams p;
  p.M = M, p.N = N, p.K = K;
  p.tiles_m = static_cast<int>((M + C::BM - 1) / C::BM);
  p.tiles_n = static_cast<int>((N + C::BN - 1) / C::BN);
  p.ep = ep;
  const int64_t tiles = static_cast<int64_t>(p.tiles_m) * p.tiles_n;
  if (tiles > 0x7fffffff) return CPLXK_ERR_UNSUPPORTED;

  auto kern = fwd_tc_kernel<T, kCplx, kVD, kXform, kSwz>;
  CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  kern<<<static_cast<unsigned>(tiles), C::THREADS, C::SMEM_BYTES, st>>>(tm_xr, tm_xi, tm_wr, tm_wi, tm_ls, tm_q, p);
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

// true when fwd_tc_dispatch(vd, workspace) will also produce the layer's KL sum (pre-pass fusion)
bool fwd_tc_fuses_kl(int dtype, bool f16_ok, int64_t M, int64_t N, int64_t K) {
  return dtype == CPLXK_F32 && f16_ok && K >= kShortK && fwd_tc3_supported(dtype, M, N, K);
}

bool fwd_tc_supported(int dtype, bool cplx, const void* x_re, const void* x_im, const void* w_re,
                      const void* w_im, const void* ls2, int64_t M, int64_t N, int64_t K) {
  const int64_t es = dtype == CPLXK_F32 ? 4 : 2;
  if (M < 1 || N < 1 || K < 1) return false;
  if ((K * es) % 16 != 0) return false;
  if (M > 0x7fffff00 || N > 0x7fffff00 || K > 0x7fffff00) return false;
  if (!aligned16(x_re) || !aligned16(w_re)) return false;
  if (cplx && (!aligned16(x_im) || !aligned16(w_im))) return false;
  if (ls2 && !aligned16(ls2)) return false;
  return true;
}

int fwd_tc_dispatch(int dtype, bool cplx, bool vd, int swz, bool f16_ok, const void* x_re,
                    const void* x_im, const void* w_re, const void* w_im, const void* ls2,
                    void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                    cudaStream_t st, const KlFuse& kl) {
  const bool xform = vd && workspace == nullptr;
#define CPLXK_TC_ARGS f16_ok, x_re, x_im, w_re, w_im, ls2, workspace, M, N, K, ep, st, kl
#define CPLXK_TC_CASE(T, SW)                                                                   \
  if (cplx && vd && xform) return launch_tc<T, true, true, true, SW>(CPLXK_TC_ARGS);           \
  if (cplx && vd) return launch_tc<T, true, true, false, SW>(CPLXK_TC_ARGS);                   \
  if (cplx) return launch_tc<T, true, false, false, SW>(CPLXK_TC_ARGS);                        \
  if (vd && xform) return launch_tc<T, false, true, true, SW>(CPLXK_TC_ARGS);                  \
  if (vd) return launch_tc<T, false, true, false, SW>(CPLXK_TC_ARGS);                          \
  return launch_tc<T, false, false, false, SW>(CPLXK_TC_ARGS);
  if (dtype == CPLXK_F32) {
    if (swz == 64) { CPLXK_TC_CASE(float, 64) }
    CPLXK_TC_CASE(float, 128)
  }
  if (dtype == CPLXK_BF16) {
    if (swz == 64) { CPLXK_TC_CASE(__nv_bfloat16, 64) }
    CPLXK_TC_CASE(__nv_bfloat16, 128)
  }
#undef CPLXK_TC_CASE
#undef CPLXK_TC_ARGS
  return CPLXK_ERR_BADARG;
}

}  // namespace cplxk
