// Pieces shared by the persistent CTA-pair kernels (fwd_tc3.cu: variational forward,
// fwd_lin3.cu: plain affine map): cluster-scope mbarrier wrappers, the kind::f16
// instruction descriptor with an explicit operand format, row stores, tensor-map encoding.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace cplxk {

namespace ptx {
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of another CTA of the cluster.  Used to hand TMEM accumulators back to the
// MMA issuer: the tcgen05.ld's are complete (tcgen05.wait::ld) and fenced (tcgen05.fence::
// before_thread_sync) when it is called, and no ordinary memory is published through it, so the
// arrive carries no .release.cluster -- that form compiles to MEMBAR.ALL.GPU + ERRBAR and made
// every epilogue warp wait for its outstanding global stores once per tile (ncu, conv pair kernel:
// 20 % of all stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// kind::f16 instruction descriptor with explicit operand format (0 = fp16, 1 = bf16)
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t fmt, int M, int N, bool neg_a) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (neg_a ? (1u << 13) : 0u) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
}  // namespace ptx

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Grouped raster of the persistent kernels: tiles of 256 rows x 128 columns, `group` row tiles
// are walked against all column tiles before the next group (a wave of 74 tile pairs then
// covers about 6 x 12 tiles: the smallest operand footprint per wave).
__host__ __device__ __forceinline__ void raster_tile(int t, int tiles_m2, int tiles_n, int group, int& tile_m,
                                            int& tile_n) {
  const int per_group = group * tiles_n;
  const int g = t / per_group;
  const int first_m = g * group;
  const int gsize = (tiles_m2 - first_m) < group ? (tiles_m2 - first_m) : group;
  const int r = t - g * per_group;
  tile_m = first_m + r % gsize;
  tile_n = r / gsize;
}

// 64 consecutive outputs of one row: 32-byte (fp32: st.global.v8, one full sector per lane) or
// 16-byte vector stores when the row segment allows
template <typename T>
__device__ __forceinline__ void store_run64(T* dst, const float (&v)[64], int nvalid) {
  constexpr int V = Elem<T>::kVec;
  if constexpr (std::is_same<T, float>::value) {
    if (nvalid == 64 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * c),
                     "f"(v[8 * c]), "f"(v[8 * c + 1]), "f"(v[8 * c + 2]), "f"(v[8 * c + 3]),
                     "f"(v[8 * c + 4]), "f"(v[8 * c + 5]), "f"(v[8 * c + 6]), "f"(v[8 * c + 7])
                     : "memory");
      return;
    }
  }
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (nvalid == 64 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[16 * c + 2 * j], v[16 * c + 2 * j + 1]);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * c),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
      }
      return;
    }
  }
  if (nvalid == 64 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
#pragma unroll
    for (int c = 0; c < 64 / V; ++c) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = v[c * V + j];
      o.store(dst + c * V);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 64; ++j)
      if (j < nvalid) dst[j] = Elem<T>::from_f(v[j]);
  }
}

// sixteen consecutive outputs of one row (16-column drain chunk of fwd_lin3.cu); `full`: all
// sixteen are inside the row and dst is 32-byte (fp32) / 32-byte (bf16) aligned
template <typename T>
__device__ __forceinline__ void store_run16(T* dst, const float (&v)[16], bool full, int nvalid) {
  if (full) {
    if constexpr (std::is_same<T, float>::value) {
#pragma unroll
      for (int c = 0; c < 2; ++c)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * c),
                     "f"(v[8 * c]), "f"(v[8 * c + 1]), "f"(v[8 * c + 2]), "f"(v[8 * c + 3]),
                     "f"(v[8 * c + 4]), "f"(v[8 * c + 5]), "f"(v[8 * c + 6]), "f"(v[8 * c + 7])
                     : "memory");
    } else {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w[0]),
                   "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) dst[j] = Elem<T>::from_f(v[j]);
  }
}

// plane [rows, cols] row-major, element size es -> box {box_c, box_r}; swizzle = box_c * es bytes
static inline int map2d(CUtensorMap* out, CUtensorMapDataType dt, size_t es, const void* ptr, int64_t rows,
                 int64_t cols, int box_c, int box_r) {
  auto enc = tensor_map_encode_fn();
  if (!enc) return CPLXK_ERR_CUDA;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_r)};
  cuuint32_t estr[2] = {1u, 1u};
  const size_t inner = box_c * es;
  const CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CPLXK_OK : CPLXK_ERR_CUDA;
}

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace cplxk
