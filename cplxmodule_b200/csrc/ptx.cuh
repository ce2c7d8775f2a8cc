// Thin inline-PTX wrappers for the Blackwell (sm_100a) async machinery used by
// the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld) and the shared-memory matrix descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor"
// tables (same fields as cute::UMMA::SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace cplxk {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time);
// one lookup per process, shared by every tensor-core kernel's host side.
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tensorMapEncodeTiled tensor_map_driver_fn() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) != cudaSuccess ||
        r != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(q);
  }
  return fn;
}

// Descriptor cache: a layer called step after step passes the same (pointer, shape, box) tuples,
// so every host-side launcher encodes through a small per-thread direct-mapped table keyed by
// ALL arguments of the encode (and the current device) instead of calling the driver each time.
// Nothing but the 128-byte descriptors is kept; a colliding tuple simply overwrites its slot.
struct TensorMapKey {
  uint64_t ptr, gdim[5], gstr[4];
  uint32_t box[5], estr[5], dt, rank, interleave, swizzle, l2, oob;
  int32_t dev;
};
inline CUresult tensor_map_encode_cached(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank,
                                         void* ptr, const cuuint64_t* gdim, const cuuint64_t* gstr,
                                         const cuuint32_t* box, const cuuint32_t* estr,
                                         CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                                         CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
  PFN_tensorMapEncodeTiled enc = tensor_map_driver_fn();
  if (!enc) return CUDA_ERROR_NOT_INITIALIZED;
  if (rank < 1 || rank > 5) return enc(out, dt, rank, ptr, gdim, gstr, box, estr, il, sw, l2, oob);
  TensorMapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = reinterpret_cast<uint64_t>(ptr);
  for (cuuint32_t i = 0; i < rank; ++i) k.gdim[i] = gdim[i], k.box[i] = box[i], k.estr[i] = estr[i];
  for (cuuint32_t i = 0; i + 1 < rank; ++i) k.gstr[i] = gstr[i];
  k.dt = static_cast<uint32_t>(dt), k.rank = rank, k.interleave = static_cast<uint32_t>(il);
  k.swizzle = static_cast<uint32_t>(sw), k.l2 = static_cast<uint32_t>(l2), k.oob = static_cast<uint32_t>(oob);
  int dev = 0;
  cudaGetDevice(&dev);
  k.dev = dev;
  uint64_t h = 1469598103934665603ull;           // FNV-1a over the key's 8-byte words
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
  for (size_t i = 0; i < sizeof(k) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
  struct Slot { TensorMapKey key; CUtensorMap map; bool valid; };
  constexpr int kSlots = 256;
  static thread_local Slot table[kSlots];
  Slot& s = table[(h ^ (h >> 29)) & (kSlots - 1)];
  if (s.valid && memcmp(&s.key, &k, sizeof(k)) == 0) {
    *out = s.map;
    return CUDA_SUCCESS;
  }
  const CUresult r = enc(out, dt, rank, ptr, gdim, gstr, box, estr, il, sw, l2, oob);
  if (r == CUDA_SUCCESS) s.key = k, s.map = *out, s.valid = true;
  return r;
}
inline PFN_tensorMapEncodeTiled tensor_map_encode_fn() {
  return tensor_map_driver_fn() ? &tensor_map_encode_cached : nullptr;
}

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------- programmatic dependent launch (PDL)
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still running; grid_dep_wait() blocks until that predecessor has
// completed and its memory is visible (a no-op for an ordinary launch).  The predecessor calls
// grid_dep_launch() in every block to allow the early start.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-d tiled load: c0 = innermost (contiguous) coordinate, c1 = row coordinate
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; issued by ONE thread
template <bool kBF16>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  if constexpr (kBF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// all previously issued tcgen05.mma of this thread -> one arrive on `bar`
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t gets lane (base_lane + t)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------- CTA-pair (cta_group::2) forms
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are accounted on the
// mbarrier of the EVEN (leader) CTA: same smem offset, peer bit (24) of the address cleared.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap* m,
                                                 uint32_t bar, int32_t c0, int32_t c1) {
  const uint32_t leader_bar = bar & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                                 int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  const uint32_t leader_bar = bar & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// M = 256 (128 rows per CTA), B split along N across the pair; issued by the leader CTA only
template <bool kBF16>
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  if constexpr (kBF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// completion of all prior MMAs of the pair -> one arrive on `bar` in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}

// ----------------------------------------------------------- descriptors
// Shared-memory matrix descriptor for a K-major operand tile whose rows are
// `kSwizzleBytes` long (one swizzle atom wide), rows packed densely, 8-row
// groups `8*kSwizzleBytes` apart (exactly what a TMA box {kSwizzleBytes/elt, rows}
// with the matching CU_TENSOR_MAP_SWIZZLE_* mode writes).
//   bits [ 0,14) start address >> 4        bits [16,30) LBO >> 4 (unused for swizzled K-major: 1)
//   bits [32,46) SBO >> 4 (8-row group pitch) bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
template <int kSwizzleBytes>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  static_assert(kSwizzleBytes == 128 || kSwizzleBytes == 64 || kSwizzleBytes == 32, "swizzle");
  constexpr uint64_t layout = kSwizzleBytes == 128 ? 2 : (kSwizzleBytes == 64 ? 4 : 6);
  constexpr uint64_t sbo = (8 * kSwizzleBytes) >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= sbo << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= layout << 61;
  return d;
}
// advance a K-major descriptor by `bytes` along K inside the swizzle atom
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) {
  return d + static_cast<uint64_t>(bytes >> 4);
}

// Instruction descriptor, kind::f16 / kind::tf32, fp32 accumulate, K-major A and B.
//   [4,6) c_format (1 = f32)  [7,10) a_format  [10,13) b_format (0 f16, 1 bf16, 2 tf32)
//   [13] negate A  [14] negate B  [15] a_major  [16] b_major (0 = K-major)
//   [17,23) N >> 3   [24,29) M >> 4
template <bool kBF16>
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool neg_a, bool neg_b) {
  uint32_t fmt = kBF16 ? 1u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (neg_a ? (1u << 13) : 0u) |
         (neg_b ? (1u << 14) : 0u) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace ptx
}  // namespace cplxk
