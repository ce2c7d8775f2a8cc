// Persistent CTA-pair (cta_group::2) fused variational forward on 16-bit operands.
//
//   re += a_re U^T - a_im V^T ,  im += a_re V^T + a_im U^T      (cplxmodule/cplx.py:641-642)
//   s2 += |x|^2 exp(log_sigma2)^T                               (nn/relevance/complex/base.py:50-54)
//   y   = mu * (sx[m] sw[n]) + b + eps sqrt(max(s2, 1e-8))      (complex/base.py:56, cplx.py:644-646)
//
// Differences to fwd_tc2.cu (one tile pair per cluster):
//  * fp32 planes do NOT run as kind::tf32.  A pre-pass (vd_prepare_f16_kernel below) rescales
//    every row of x and of W by a power of two so its largest entry sits at 2^13 and writes it
//    as fp16: same 11-bit significand as tf32 (round-to-nearest), no range problem because the
//    scale is per row, and kind::f16 runs at twice the tf32 rate.  The epilogue undoes the two
//    scales (exact: powers of two).  bf16 planes are consumed as they are.
//  * One cluster per SM pair loops over output tiles (static round-robin).  The TMA producers run
//    ahead into the ring while the epilogue warps drain TMEM, so the next tile's MMAs start the
//    moment the accumulators are released (tmem_empty barrier): no prologue, TMEM allocation or
//    pipeline refill per tile.
//  * The drain is short: each epilogue thread folds accumulator, scales, bias and its prefetched
//    noise IN PLACE into the registers that held the noise, releases TMEM after its last
//    tcgen05.ld, and only then issues the global stores, which trail into the next mainloop.
//
// Warps (320 threads / CTA): 0 = TMA producer, 1 = MMA issuer (leader CTA), 2..9 = noise
// prefetch (during the mainloop, registers) + drain.
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "epilogue.cuh"
#include "kl_math.cuh"
#include "knobs.cuh"
#include "prep_row.cuh"
#include "ptx.cuh"
#include "tc3_common.cuh"

namespace cplxk {

template <typename OutT, bool kCplx>
struct Tc3Cfg {
  static constexpr int BN = 128;
  static constexpr int BK = 64;                       // 16-bit elements: 128-byte rows (SW128)
  static constexpr int KSTEPS = 4;                    // K = 16 per MMA
  static constexpr int A_TILE = 128 * 128, B_HALF = 64 * 128;
  static constexpr int NA = kCplx ? 2 : 1;
  static constexpr int OFF_A0 = 0, OFF_A1 = A_TILE, OFF_Q = NA * A_TILE;
  static constexpr int OFF_B0 = OFF_Q + A_TILE, OFF_B1 = OFF_B0 + B_HALF;
  static constexpr int OFF_E = OFF_B0 + NA * B_HALF;
  static constexpr int STAGE_BYTES = OFF_E + B_HALF;  // 72 KB complex, 48 KB real
  static constexpr int EPI_WARPS = 8;
  static constexpr int AUX_BYTES = 4096;              // barriers, tmem slot, per-tile column vectors
  static constexpr int AVAIL = 227 * 1024 - 1024 - AUX_BYTES;
  static constexpr int STAGES = AVAIL / STAGE_BYTES > 8 ? 8 : AVAIL / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
  static constexpr int NACC = NA + 1;
  static constexpr int TMEM_COLS = NACC * BN <= 256 ? 256 : 512;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static_assert(STAGES >= 3, "ring too shallow");
};

struct Tc3Params {
  int64_t M, N, K;
  int tiles_m2, tiles_n;   // tiles of 256 rows, 128 columns
  int f16;                 // mean operands are fp16 (else bf16)
  int dbg;                 // CPLXK_DBG: 1 = MMAs without loads, 2 = loads without MMAs, 3 = no mainloop
  int group;               // 256-row tiles per raster group (CPLXK_RASTER, default 6)
  const float* sx;         // [M] inverse row scales of x (nullable)
  const float* sw;         // [N] inverse row scales of W (nullable)
  EpiParams ep;
#ifdef CPLXK_TRACE
  long long* trace;        // -DCPLXK_TRACE builds: per-tile SM clock stamps of cluster 0's leader CTA
#endif
};

#ifdef CPLXK_TRACE
// slots per tile: 0 MMA warp past tmem_empty, 1 last MMA committed, 2 noise prefetched (warp 2),
// 3 accumulators seen full, 4 TMEM handed back, 5 stores issued; 6 producer issued the tile's last load
#define CPLXK_STAMP(slot)                                                                  \
  do {                                                                                     \
    if (p.trace && cluster_id == 0 && leader && lane == 0) p.trace[(t / num_clusters) * 8 + (slot)] = clock64(); \
  } while (0)
static long long* g_tc3_trace = nullptr;
extern "C" void cplxk_debug_set_trace(void* dev_ptr) { g_tc3_trace = static_cast<long long*>(dev_ptr); }
#else
#define CPLXK_STAMP(slot) do { } while (0)
#endif

template <typename OutT, bool kCplx>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
fwd_tc3_kernel(const __grid_constant__ CUtensorMap tm_xr, const __grid_constant__ CUtensorMap tm_xi,
               const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_wr,
               const __grid_constant__ CUtensorMap tm_wi, const __grid_constant__ CUtensorMap tm_e,
               const Tc3Params p) {
  using C = Tc3Cfg<OutT, kCplx>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t aux = base + C::STAGES * C::STAGE_BYTES;
  // aux: full[8] empty[8] accum_full tmem_empty tmem_slot | colvec[2][3][128] floats at +1024
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_accum = aux + 128, bar_tfree = aux + 136;
  const uint32_t tmem_slot = aux + 144;
  uint8_t* aux_ptr = smem + C::STAGES * C::STAGE_BYTES;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(aux_ptr + 144);
  float* colvec = reinterpret_cast<float*>(aux_ptr + 1024);         // [2][3][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.tiles_m2 * p.tiles_n;
  const int num_kb = p.dbg == 3 ? 0 : static_cast<int>((p.K + C::BK - 1) / C::BK);

  auto decode_tile = [&](int t, int& tile_m, int& tile_n) {
    raster_tile(t, p.tiles_m2, p.tiles_n, p.group, tile_m, tile_n);
  };

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
    ptx::mbar_init(bar_tfree, 2 * C::EPI_WARPS);   // only the leader's is used
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
    // ------------------------------------------------------------------- TMA producer
    // (whole warp in the loop, one elected lane issues: see the MMA issuer below)
    if (p.dbg != 1) {
      const bool elected = ptx::elect_one();
      int s = 0;
      uint32_t ph = 0;
      ptx::grid_dep_wait();   // the operands are written by the pre-pass launch before this one
      // From here on everything launched before this kernel is complete: a kernel launched behind
      // it WITH the programmatic attribute (cplxk_kl_guard: it reads the parameters and the
      // pre-pass's KL sum, nothing of this kernel) may run under the mainloop.  Ordinary launches
      // still wait for this grid to finish.
      ptx::grid_dep_launch();
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
            ptx::tma_load_2d_pair(st + C::OFF_Q, &tm_q, fb, k0, m0);
            ptx::tma_load_2d_pair(st + C::OFF_B0, &tm_wr, fb, k0, nb0);
            if constexpr (kCplx) ptx::tma_load_2d_pair(st + C::OFF_B1, &tm_wi, fb, k0, nb0);
            ptx::tma_load_2d_pair(st + C::OFF_E, &tm_e, fb, k0, nb0);
          }
          __syncwarp();
          if (++s == C::STAGES) s = 0, ph ^= 1u;
        }
        CPLXK_STAMP(6);
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer: leader CTA only
    // The WHOLE warp walks the loop (waits included) and one elected lane issues: with
    // warp-uniform control flow the descriptors live in uniform registers and a tcgen05.mma costs
    // ~3 issue slots; inside an `if (lane == 0)` region the compiler re-elects and moves four
    // registers to the uniform file per MMA, and the single issuing thread -- which shares its
    // scheduler with two noise-generating warps -- cannot keep the tensor pipe fed.
    if (leader) {
      const bool elected = ptx::elect_one();
      const uint32_t fmt = p.f16 ? 0u : 1u;
      const uint32_t idesc = ptx::make_idesc_f16(fmt, 256, C::BN, false);
      const uint32_t idesc_na = ptx::make_idesc_f16(fmt, 256, C::BN, true);
      constexpr uint32_t idesc_var = ptx::make_idesc_f16(1u, 256, C::BN, false);
      const uint32_t t_re = tmem_base, t_im = tmem_base + C::BN, t_s2 = tmem_base + C::NA * C::BN;
      int s = 0;
      uint32_t ph = 0, tile_par = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, tile_par ^= 1u) {
        ptx::mbar_wait_cluster(bar_tfree, tile_par ^ 1u);   // previous tile drained by both CTAs
        ptx::tcgen05_fence_after();
        CPLXK_STAMP(0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t st = base + s * C::STAGE_BYTES;
          if (p.dbg != 1) ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tcgen05_fence_after();
          if (p.dbg != 2) {
            const uint64_t a0 = ptx::make_kmajor_desc<128>(st + C::OFF_A0);
            const uint64_t a1 = ptx::make_kmajor_desc<128>(st + C::OFF_A1);
            const uint64_t aq = ptx::make_kmajor_desc<128>(st + C::OFF_Q);
            const uint64_t b0 = ptx::make_kmajor_desc<128>(st + C::OFF_B0);
            const uint64_t b1 = ptx::make_kmajor_desc<128>(st + C::OFF_B1);
            const uint64_t be = ptx::make_kmajor_desc<128>(st + C::OFF_E);
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
                ptx::umma_ss_pair<true>(t_s2, ptx::desc_advance(aq, off), ptx::desc_advance(be, off), idesc_var, acc);
              }
            }
          }
          if (elected) ptx::umma_commit_pair(bar_empty + 8 * s);   // frees the stage in BOTH CTAs
          __syncwarp();
          if (++s == C::STAGES) s = 0, ph ^= 1u;
        }
        if (elected) ptx::umma_commit_pair(bar_accum);             // accumulators complete, both CTAs
        __syncwarp();
        CPLXK_STAMP(1);
      }
    }
  } else {
    // ------------------------------------ 8 epilogue warps: noise prefetch, then drain + row stores
    // (little is kept live across noise_prefetch: it needs ~all of the 168 registers)
    uint32_t tile_par = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, tile_par ^= 1u) {
      float nre[64], nim[kCplx ? 64 : 1];
      {
        int tile_m, tile_n;
        decode_tile(t, tile_m, tile_n);
        const int64_t m = static_cast<int64_t>(tile_m) * 256 + rank * 128 + (warp & 3) * 32 + lane;
        const int64_t nb = static_cast<int64_t>(tile_n) * C::BN + ((warp - 2) >> 2) * 64;
        noise_prefetch<OutT, kCplx, 64>(p.ep, m, nb, nre, nim);
      }
      if (warp == 2) CPLXK_STAMP(2);
      // the first tile's noise does not depend on the pre-pass: it is generated while that launch
      // drains (programmatic dependent launch); the row / column scales below do depend on it
      ptx::grid_dep_wait();
      const int quarter = warp & 3;
      const int half = (warp - 2) >> 2;
      const int te = threadIdx.x - 64;                       // 0..255
      const uint32_t tfree_remote = ptx::mapa_u32(bar_tfree, 0);
      int tile_m, tile_n;
      decode_tile(t, tile_m, tile_n);
      const int32_t m0 = tile_m * 256 + static_cast<int32_t>(rank) * 128;
      const int32_t n0 = tile_n * C::BN;
      const int64_t m = static_cast<int64_t>(m0) + quarter * 32 + lane;
      const int64_t nb = static_cast<int64_t>(n0) + half * 64;

      // per-tile column vectors (bias, inverse weight-row scale) -> smem, double buffered
      float* cv = colvec + tile_par * 3 * 128;
      if (te < 128) {
        const int64_t n = static_cast<int64_t>(n0) + te;
        const bool ok = n < p.N;
        const OutT* br = static_cast<const OutT*>(p.ep.b_re);
        const OutT* bi = static_cast<const OutT*>(p.ep.b_im);
        cv[te] = (ok && br) ? Elem<OutT>::to_f(__ldg(br + n)) : 0.f;
        cv[128 + te] = (kCplx && ok && bi) ? Elem<OutT>::to_f(__ldg(bi + n)) : 0.f;
        cv[256 + te] = (ok && p.sw) ? __ldg(p.sw + n) : 1.f;
      }
      const float sxm = (p.sx && m < p.M) ? __ldg(p.sx + m) : 1.f;
      ptx::named_bar_sync(1, 32 * C::EPI_WARPS);            // column vectors visible

      ptx::mbar_wait(bar_accum, tile_par);
      ptx::tcgen05_fence_after();
      if (warp == 2) CPLXK_STAMP(3);
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 64;
      const float* cvb = cv + half * 64;
      // The TMEM reads are software-pipelined: the loads of chunk c + 1 are issued right after the
      // wait for chunk c and fly while chunk c is folded into the noise registers.  (Serialised
      // load -> wait -> math rounds kept the accumulators busy for 2.5-5 us per tile and the tensor
      // pipe idle for as long: tools/tc3_trace.py, profiles/tc3_trace_r2.jsonl.)
      uint32_t r_re[2][8], r_im[2][8], r_s2[2][8];
      ptx::tmem_ld_32x32b_x8(lane_base, r_re[0]);
      if constexpr (kCplx) ptx::tmem_ld_32x32b_x8(lane_base + C::BN, r_im[0]);
      ptx::tmem_ld_32x32b_x8(lane_base + C::NA * C::BN, r_s2[0]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = c * 8;
        const int cur = c & 1, nxt = cur ^ 1;
        ptx::tmem_ld_wait();
        if (c < 7) {
          ptx::tmem_ld_32x32b_x8(lane_base + col + 8, r_re[nxt]);
          if constexpr (kCplx) ptx::tmem_ld_32x32b_x8(lane_base + C::BN + col + 8, r_im[nxt]);
          ptx::tmem_ld_32x32b_x8(lane_base + C::NA * C::BN + col + 8, r_s2[nxt]);
        } else {   // last TMEM read of this warp: hand the accumulators back to the MMA issuer
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tfree_remote);
          if (warp == 2) CPLXK_STAMP(4);
        }
        float f_s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sc = sxm * cvb[256 + col + j];
          f_s2[j] = __uint_as_float(r_s2[cur][j]);
          const float sd = sd_of(f_s2[j]);
          nre[col + j] = fmaf(nre[col + j], sd, fmaf(__uint_as_float(r_re[cur][j]), sc, cvb[col + j]));
          if constexpr (kCplx)
            nim[col + j] = fmaf(nim[col + j], sd, fmaf(__uint_as_float(r_im[cur][j]), sc, cvb[128 + col + j]));
        }
        if (p.ep.s2_out && m < p.M && nb + col < p.N) {
          const int64_t ncol = nb + col;
          const int nvalid = (p.N - ncol) < 8 ? static_cast<int>(p.N - ncol) : 8;
          store_s2_run<OutT, 8>(p.ep, m * p.N, ncol, nvalid, f_s2);
        }
      }
      // TMEM is released; the stores trail into the next tile's mainloop
      if (m < p.M && nb < p.N) {
        const int nvalid = (p.N - nb) < 64 ? static_cast<int>(p.N - nb) : 64;
        store_run64<OutT>(static_cast<OutT*>(p.ep.y_re) + m * p.N + nb, nre, nvalid);
        if constexpr (kCplx) store_run64<OutT>(static_cast<OutT*>(p.ep.y_im) + m * p.N + nb, nim, nvalid);
      }
      if (warp == 2) CPLXK_STAMP(5);
    }
    ptx::tcgen05_fence_before();
  }

  ptx::cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still use it
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ fp32 -> fp16 operand pre-pass
// One block per row of x or of W (interleaved); the row conversion itself is prep_row.cuh.
// resident blocks per SM / 8-element groups a thread keeps in registers between the two passes
// (side builds for A/B runs: -DCPLXK_PREP_BLOCKS=.. -DCPLXK_PREP_CACHE=..)
#ifndef CPLXK_PREP_BLOCKS
#define CPLXK_PREP_BLOCKS 4
#endif
#ifndef CPLXK_PREP_CACHE
#define CPLXK_PREP_CACHE 2
#endif
template <bool kCplx, bool kMask>
__global__ void __launch_bounds__(256, CPLXK_PREP_BLOCKS)
vd_prepare_f16_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im, int64_t M_,
                      const float* __restrict__ w_re, const float* __restrict__ w_im,
                      const float* __restrict__ ls2, const float* __restrict__ w_mask, int64_t N_,
                      int64_t K_, __half* __restrict__ xh_re, __half* __restrict__ xh_im,
                      __nv_bfloat16* __restrict__ q, __half* __restrict__ wh_re,
                      __half* __restrict__ wh_im, __nv_bfloat16* __restrict__ e,
                      float* __restrict__ isx, float* __restrict__ isw, int kl_kind,
                      float* __restrict__ kl_sum, KlWorkspace* __restrict__ kl_ws, int64_t kl_row0,
                      int64_t kl_row1, unsigned long long* __restrict__ kl_fp, int prefetch_next) {
  // (separate __restrict__ parameters, not the struct: the no-alias facts let the loads of the
  // write pass be hoisted above its stores)
  PrepArgs a;
  a.x_re = x_re, a.x_im = x_im, a.w_re = w_re, a.w_im = w_im, a.ls2 = ls2, a.w_mask = w_mask;
  a.M = M_, a.N = N_, a.K = K_;
  a.xh_re = xh_re, a.xh_im = xh_im, a.wh_re = wh_re, a.wh_im = wh_im, a.q = q, a.e = e;
  a.isx = isx, a.isw = isw, a.kl_kind = kl_kind, a.kl_row0 = kl_row0, a.kl_row1 = kl_row1;
  a.want_fp = (kl_kind >= 0 && kl_fp != nullptr) ? 1 : 0;
  unsigned long long fp_acc = 0ull;          // fingerprint share of the weight rows this block converts
  __shared__ float red[8];
  __shared__ double kl_sh[kKlThreads / 32];
  __shared__ bool kl_last;
  float kl_acc = 0.f;                        // KL penalty of the weight rows this thread converts
  const int tid = threadIdx.x;
  // let the GEMM launch behind this one start its prologue (barriers, TMEM, first tile's noise) as
  // soon as every block of this grid is running; it waits (griddepcontrol.wait) before it reads
  // anything written here
  ptx::grid_dep_launch();
  auto sync = [] { __syncthreads(); };
  // x rows (pure streaming) and W rows (streaming + ~50 instructions of KL math per element)
  // alternate in the global row order and the grid size is odd, so every block -- and every SM
  // at any time -- works on a mix of the two instead of all x rows first and all W rows last
  const int64_t M = a.M, N = a.N;
  const int64_t mn = M < N ? M : N;
  auto row_of = [&](int64_t g, bool& is_x, int64_t& r) {
    if (g < 2 * mn) {
      is_x = (g & 1) == 0;
      r = g >> 1;
    } else {
      is_x = M > N;
      r = mn + (g - 2 * mn);
    }
  };
  for (int64_t g = blockIdx.x; g < M + N; g += gridDim.x) {
    bool is_x;
    int64_t r;
    row_of(g, is_x, r);
    if (prefetch_next && g + gridDim.x < M + N) {     // the row this block converts next: towards L2 now
      bool nx;
      int64_t nr;
      row_of(g + gridDim.x, nx, nr);
      prep_prefetch_row<kCplx, 256>(a, nx, nr, tid);
    }
    kl_acc += prep_convert_row<kCplx, 256, CPLXK_PREP_CACHE, kMask>(a, is_x, r, tid, red, sync, fp_acc);
  }
  if (a.kl_kind >= 0) {
    __syncthreads();
    if (a.want_fp && fp_acc != 0ull) atomicAdd(&kl_ws->fp, fp_acc);   // thread 0 only; before its ticket
    const double bsum = block_sum(static_cast<double>(kl_acc), kl_sh);
    grid_sum_finish(bsum, kl_ws, kl_sum, 1.0, kl_sh, &kl_last, a.want_fp ? kl_fp : nullptr);
  }
}

// ------------------------------------------------------------------------ host side
struct Tc3Operands {
  const void *a_re, *a_im, *q, *b_re, *b_im, *e;   // 16-bit planes [M,K] / [N,K]
  const float *sx, *sw;
  bool f16;
  bool after_prepass;   // the launch right before this one in the stream is vd_prepare_f16_kernel
};

template <typename OutT, bool kCplx>
static int launch_tc3(const Tc3Operands& o, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                      cudaStream_t st) {
  using C = Tc3Cfg<OutT, kCplx>;
  CUtensorMap tm_xr, tm_xi, tm_q, tm_wr, tm_wi, tm_e;
  const CUtensorMapDataType dt_op = o.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapDataType dt_bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc;
  if ((rc = map2d(&tm_xr, dt_op, 2, o.a_re, M, K, C::BK, 128))) return rc;
  if ((rc = map2d(&tm_wr, dt_op, 2, o.b_re, N, K, C::BK, 64))) return rc;
  if ((rc = map2d(&tm_q, dt_bf, 2, o.q, M, K, C::BK, 128))) return rc;
  if ((rc = map2d(&tm_e, dt_bf, 2, o.e, N, K, C::BK, 64))) return rc;
  tm_xi = tm_xr, tm_wi = tm_wr;
  if (kCplx) {
    if ((rc = map2d(&tm_xi, dt_op, 2, o.a_im, M, K, C::BK, 128))) return rc;
    if ((rc = map2d(&tm_wi, dt_op, 2, o.b_im, N, K, C::BK, 64))) return rc;
  }
  Tc3Params p;
  p.M = M, p.N = N, p.K = K;
  p.tiles_m2 = static_cast<int>((M + 255) / 256);
  p.tiles_n = static_cast<int>((N + C::BN - 1) / C::BN);
  p.f16 = o.f16 ? 1 : 0;
  p.dbg = knobs().dbg;        // 0 unless built with -DCPLXK_DEBUG
  p.group = knobs().raster;
  p.sx = o.sx, p.sw = o.sw;
  p.ep = ep;
#ifdef CPLXK_TRACE
  p.trace = g_tc3_trace;
#endif
  const int64_t pairs = static_cast<int64_t>(p.tiles_m2) * p.tiles_n;
  if (pairs > 0x3fffffff) return CPLXK_ERR_UNSUPPORTED;
  int sm_count = 0;
  if ((rc = current_device_sm_count(&sm_count))) return rc;
  // cplxk_set_sm_reserve(n) leaves n SMs to kernels of other streams (a collective, the KL shard
  // kernel): a persistent grid that owns every SM would make them wait for its last tile
  int64_t clusters = (sm_count - sm_reserve()) / 2;
  if (clusters < 1) clusters = 1;
  if (clusters > pairs) clusters = pairs;
  auto kern = fwd_tc3_kernel<OutT, kCplx>;
  CPLXK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters)), cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = (o.after_prepass && knobs().pdl) ? 1 : 0;
  CPLXK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm_xr, tm_xi, tm_q, tm_wr, tm_wi, tm_e, p));
  return CPLXK_OK;
}


// workspace of the fp32-plane path: xh_re, xh_im, q [M,K]; wh_re, wh_im, E [N,K] (2 bytes each),
// isx [M], isw [N] floats
size_t fwd_tc3_workspace_bytes(int dtype, int64_t M, int64_t N, int64_t K) {
  if (dtype != CPLXK_F32) return 0;
  return 3 * align256(static_cast<size_t>(M) * K * 2) + 3 * align256(static_cast<size_t>(N) * K * 2) +
         align256(static_cast<size_t>(M) * 4) + align256(static_cast<size_t>(N) * 4);
}

bool fwd_tc3_supported(int dtype, int64_t M, int64_t N, int64_t K) {
  const int64_t es = dtype == CPLXK_F32 ? 4 : 2;
  (void)es;
  return M > 128 && K % 8 == 0;
}

int vd_prepare_f16_launch(bool cplx, const PrepArgs& a, const KlFuse& kl, cudaStream_t st) {
  const int64_t rows = a.M + a.N;
  int sms = 148;
  int rc0 = current_device_sm_count(&sms);
  if (rc0) return rc0;
  int64_t cap = static_cast<int64_t>(sms) * (2 * CPLXK_PREP_BLOCKS) - 1;         // odd (see the kernel)
  if (cap > kKlMaxBlocks) cap = kKlMaxBlocks - 1 + (kKlMaxBlocks & 1);           // <= kKlMaxBlocks partials, odd
  const int grid = static_cast<int>(rows > cap ? cap : (rows < 1 ? 1 : rows));
  auto kws = static_cast<KlWorkspace*>(kl.ws);
  const bool mask = a.w_mask != nullptr;
#define CPLXK_PREP(C, MK)                                                                          \
  vd_prepare_f16_kernel<C, MK><<<grid, 256, 0, st>>>(a.x_re, a.x_im, a.M, a.w_re, a.w_im, a.ls2, a.w_mask, \
                                                     a.N, a.K, a.xh_re, a.xh_im, a.q, a.wh_re, a.wh_im,   \
                                                     a.e, a.isx, a.isw, a.kl_kind, kl.sum, kws, a.kl_row0, \
                                                     a.kl_row1, static_cast<unsigned long long*>(kl.fp), \
                                                     knobs().prep_prefetch ? 1 : 0)
  if (cplx && mask) CPLXK_PREP(true, true);
  else if (cplx) CPLXK_PREP(true, false);
  else if (mask) CPLXK_PREP(false, true);
  else CPLXK_PREP(false, false);
#undef CPLXK_PREP
  CPLXK_CUDA_TRY(cudaGetLastError());
  return CPLXK_OK;
}

// fp32 planes: pre-pass to scaled fp16, then the persistent CTA-pair kernel above on kind::f16.
// (Converting part of the rows INSIDE the GEMM kernel, under its MMAs, was built and measured in
// round 2 -- profiles/tail_ab_r2_*.jsonl: two extra warps per CTA are too slow for the KL / exp
// math of the weight rows and even pure x rows cost the GEMM more than the shorter pre-pass saves.)
int fwd_tc3_f32(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                const void* ls2, void* workspace, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                cudaStream_t st, const KlFuse& kl) {
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t xb = align256(static_cast<size_t>(M) * K * 2), wb = align256(static_cast<size_t>(N) * K * 2);
  auto f = [](const void* p) { return static_cast<const float*>(p); };
  PrepArgs a{};
  a.x_re = f(x_re), a.x_im = f(x_im), a.w_re = f(w_re), a.w_im = f(w_im), a.ls2 = f(ls2), a.w_mask = nullptr;
  a.M = M, a.N = N, a.K = K;
  a.xh_re = reinterpret_cast<__half*>(ws);
  a.xh_im = reinterpret_cast<__half*>(ws + xb);
  a.q = reinterpret_cast<__nv_bfloat16*>(ws + 2 * xb);
  a.wh_re = reinterpret_cast<__half*>(ws + 3 * xb);
  a.wh_im = reinterpret_cast<__half*>(ws + 3 * xb + wb);
  a.e = reinterpret_cast<__nv_bfloat16*>(ws + 3 * xb + 2 * wb);
  a.isx = reinterpret_cast<float*>(ws + 3 * xb + 3 * wb);
  a.isw = reinterpret_cast<float*>(ws + 3 * xb + 3 * wb + align256(static_cast<size_t>(M) * 4));
  const bool want_kl = kl.sum && kl.ws && kl.kind >= 0;
  const int64_t row0 = kl.row_begin, row1 = kl.row_end < 0 ? N : kl.row_end;
  // The KL sum is a by-product of this pass (the weight rows and log_sigma2 are in registers
  // anyway).  Its ~50 instructions per weight cost 25-30 us of issue slots at 4096^2 wherever they
  // run serially; evaluating it inside the GEMM kernel instead (epilogue warps between two tiles,
  // or two extra warps) was built and measured in round 2 and is SLOWER (profiles/
  // kl_in_gemm_ab_r2.jsonl, tail_ab_r2_*.jsonl): those warps share their schedulers with the
  // Philox generation that already fills the mainloop's shadow.
  a.kl_kind = want_kl ? kl.kind : -1;
  a.kl_row0 = row0, a.kl_row1 = row1;
  int rc = vd_prepare_f16_launch(cplx, a, kl, st);
  if (rc) return rc;
  // the KL (partial) sum is final here: let a collective on another stream start under the GEMM
  if (kl.event) CPLXK_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(kl.event), st));
  if (!ep.y_re) return CPLXK_OK;   // cplxk_linear_vd_prepare: operands (and the KL sum) only
  Tc3Operands o{a.xh_re, a.xh_im, a.q, a.wh_re, a.wh_im, a.e, a.isx, a.isw, true, kl.event == nullptr};
  return cplx ? launch_tc3<float, true>(o, M, N, K, ep, st) : launch_tc3<float, false>(o, M, N, K, ep, st);
}

// bf16 planes: operands as they are; q / E (bf16) were written by the caller's pre-pass.
int fwd_tc3_bf16(bool cplx, const void* x_re, const void* x_im, const void* w_re, const void* w_im,
                 const void* q, const void* e, int64_t M, int64_t N, int64_t K, const EpiParams& ep,
                 cudaStream_t st) {
  Tc3Operands o{x_re, x_im, q, w_re, w_im, e, nullptr, nullptr, false, false};
  return cplx ? launch_tc3<__nv_bfloat16, true>(o, M, N, K, ep, st)
              : launch_tc3<__nv_bfloat16, false>(o, M, N, K, ep, st);
}

}  // namespace cplxk
