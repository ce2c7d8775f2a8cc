// Interface between the conv C-ABI entry point (conv.cu) and the tcgen05 implicit GEMM (conv_tc.cu).
#pragma once
#include "common.cuh"
#include "noise.cuh"

namespace cplxk {

struct ConvTcGeom {
  int64_t B, C, H, W, O, Ho, Wo;
  int Cp, Op;            // padded channel counts of the workspace planes
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int Wt, Ht;            // output patch of one tile (Wt * Ht == 128)
  int tiles_w, tiles_h, tiles_n;
  // grouped / real-plane convolutions (conv_tc_kernel only; groups == 1: Cg = C, Og = O, Cgp = Cp,
  // Ogp = Op, tiles_ng = tiles_n): channels per group, their padded counts in the weight planes
  // ([tap][group * Ogp + o][Cgp]) and the n-blocks of one group
  int groups, Cg, Og, Cgp, Ogp, tiles_ng;
  // Groups narrower than an n-block are PACKED: `groups` / Cg / Og above describe super-groups of
  // several true groups whose weights sit block-diagonally in the prepared planes (tCg / tOg =
  // channels per TRUE group) -- the MMAs multiply the off-diagonal zeros, the operand traffic
  // and the tile count are those of the dense layer.
  int tCg, tOg;
};

struct ConvTcEpi {
  const void* b_re;
  const void* b_im;
  const void* eps_re;
  const void* eps_im;
  void* y_re;
  void* y_im;
  int64_t plane_elems;
  int f16_ok = 1;        // 0: MATH_TENSOR_TF32 -- fp32 planes keep tf32 operands
  int nhwc;              // 1: x / y planes are channels-last (torch.channels_last): no transposing pre-pass
  // fp32 planes on fp16 operands (conv_tc_pair_kernel<float, true>): per-image maxima of |x|
  // (float bits, written by conv_amax_kernel) and inverse per-output-channel scales of W
  const unsigned int* amax = nullptr;
  const float* isw = nullptr;
  NoiseParams noise;
};

size_t conv_tc_workspace_bytes(int dtype, bool vd, int64_t B, int64_t C, int64_t H, int64_t W,
                               int64_t O, int64_t kh, int64_t kw, int groups = 1, bool real = false);
bool conv_tc_supported(int dtype, int64_t B, int64_t C, int64_t H, int64_t W, int64_t O, int64_t Ho,
                       int64_t Wo, int kh, int kw, int sh, int sw, int groups = 1, bool real = false);
// x_im == nullptr: real planes (w_im, y_im, b_im, eps_im unused); abs2: real conv of |x_re + i x_im|^2
// (the variance operand of a complex variational layer, squared inside the transposing pre-pass)
int conv_tc_dispatch(int dtype, bool vd, bool nhwc, const void* x_re, const void* x_im, const void* w_re,
                     const void* w_im, const void* ls2, void* workspace, int64_t B, int64_t C,
                     int64_t H, int64_t W, int64_t O, int64_t Ho, int64_t Wo, int kh, int kw, int sh,
                     int sw, int ph, int pw, int dh, int dw, const ConvTcEpi& ep, cudaStream_t st,
                     int groups = 1, bool abs2 = false);

}  // namespace cplxk
