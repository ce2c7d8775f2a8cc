// Process-level tuning switches and per-device facts of the host side.
//
// * A/B switches (kernel variants kept for same-box comparisons) are read from the environment
//   ONCE, when the library is first used -- never per call.  None of them changes results beyond
//   the documented tolerance of the selected arithmetic; the arithmetic itself (exact fp32 /
//   scaled-fp16 / tf32 operands) is chosen through the `math` ARGUMENT of the C ABI.
// * Measurement aids that drop work (CPLXK_DBG) only exist in a -DCPLXK_DEBUG build.
// * SM count / compute capability are cached per DEVICE (a process may drive several GPUs).
#pragma once
#include <cuda_runtime.h>

namespace cplxk {

struct Knobs {
  int tc_swizzle;          // CPLXK_TC_SWIZZLE = 64 | 128: smem swizzle of fwd_tc_kernel (0: per shape)
  int raster;              // CPLXK_RASTER: 256-row tiles per raster group of the persistent GEMMs
  bool lin3;               // CPLXK_LIN3=0: bf16 affine map on the one-tile-per-CTA kernel
  bool tma_raw_f32;        // CPLXK_TMA_RAW_F32=1: let the tensor core truncate fp32 -> tf32
  bool conv_pair;          // CPLXK_CONV_PAIR=0: conv on single-CTA tiles
  bool conv_persistent;    // CPLXK_CONV_NONPERSISTENT=1: one conv tile per CTA
  bool conv_row;           // CPLXK_CONV_ROW=0: CTA-pair conv kernel loads every tap separately (no row mode)
  bool combine_flat;       // CPLXK_COMBINE_FLAT=0: complex cplxk_vd_combine with the paired (three calls per four outputs) noise kernel
  bool conv_real_pair;     // CPLXK_CONV_REAL_PAIR=0: ungrouped real-plane conv on the one-tile-per-CTA kernel
  int conv_overlap;        // CPLXK_CONV_OVERLAP: image chunks of the fp32 NCHW conv whose pre-pass overlaps the previous chunk's GEMM (0 / 1: serial)
  bool conv_amax_pass;     // CPLXK_CONV_AMAX_PASS=1: fp32 NCHW conv input: separate amax pass before the fp16 conversion (default: optimistic single pass + fix-up)
  bool pdl;                // CPLXK_PDL=0: no programmatic dependent launch between pre-pass and GEMM
  bool prep_prefetch;      // CPLXK_PREP_PREFETCH=1: the operand pre-pass prefetches its next row to L2 (default off: slower)
  int dbg;                 // CPLXK_DBG (debug builds only; 0 otherwise)
};

const Knobs& knobs();             // api.cu
int sm_reserve();                 // SMs a persistent grid leaves free (cplxk_set_sm_reserve)
// SM count of the CURRENT device (cached per device); CPLXK_OK or CPLXK_ERR_CUDA
int current_device_sm_count(int* sms);

}  // namespace cplxk
