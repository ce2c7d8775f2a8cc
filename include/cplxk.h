/*
 * cplxk.h -- C ABI of the B200-native (sm_100a) kernels behind cplxmodule's
 * complex linear / variational-dropout forward / KL `penalties()` hot path.
 *
 * The reference (ivannz/cplxmodule, pure Python) has no FFI of its own; the
 * seam a maintainer binds is its functional layer.  Each entry point below
 * names the reference function it replaces (paths relative to the reference
 * tree).  See INTEGRATION.md for the ctypes stub that goes into the
 * reference's `cplx.py` / `nn/relevance/*`.
 *
 * Conventions
 * -----------
 *  - plain pointers + sizes, no torch types; all pointers are DEVICE pointers
 *    on the device that is current on the calling thread, unless noted;
 *  - planes are row-major with unit inner stride ("split" complex: separate
 *    real and imaginary planes, as `Cplx.real` / `Cplx.imag`);
 *  - the library allocates nothing and keeps no state besides per-process
 *    function attributes; the caller owns every buffer (incl. workspace);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    no implicit synchronisation;
 *  - return value: 0 on success, negative cplxk_status otherwise -- never
 *    throws, never aborts.  `cplxk_strerror` maps a status to text;
 *  - there is NO CPU path: a machine without an sm_100 device gets
 *    CPLXK_ERR_ARCH / CPLXK_ERR_CUDA.
 */
#ifndef CPLXK_H_
#define CPLXK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPLXK_ABI_VERSION 1

typedef enum {
  CPLXK_OK = 0,
  CPLXK_ERR_BADARG = -1,   /* null pointer / bad enum / negative size          */
  CPLXK_ERR_ALIGN = -2,    /* pointer or pitch violates the alignment contract */
  CPLXK_ERR_ARCH = -3,     /* device is not compute capability 10.x            */
  CPLXK_ERR_CUDA = -4,     /* a CUDA runtime/driver call failed                */
  CPLXK_ERR_UNSUPPORTED = -5,
  CPLXK_ERR_WORKSPACE = -6 /* workspace too small                              */
} cplxk_status;

/* storage type of every plane in a call (inputs, parameters and outputs) */
typedef enum { CPLXK_F32 = 0, CPLXK_BF16 = 1 } cplxk_dtype;

/* arithmetic used by the GEMM-shaped part of a forward call */
typedef enum {
  CPLXK_MATH_AUTO = 0,   /* tensor cores when the shape/alignment allows, else SIMT */
  CPLXK_MATH_TENSOR = 1, /* tcgen05, fp32 accumulation in TMEM.  BF16 planes: bf16 operands.
                            F32 planes: 11-bit-significand operands -- per-row power-of-two
                            scaled fp16 (variational forward with workspace) or tf32
                            (round-to-nearest) otherwise.  Needs 16-byte aligned planes and
                            K*sizeof(elem) % 16 == 0.                          */
  CPLXK_MATH_SIMT = 2,   /* exact fp32 FMA on CUDA cores (any shape)          */
  CPLXK_MATH_TENSOR_TF32 = 3 /* as AUTO, but F32 planes always run on tf32 operands (rounded to
                            nearest by TMA): the arithmetic of torch's `allow_tf32`.  Same
                            11-bit significand as the scaled-fp16 form without its per-row
                            (linear) / per-image (conv) scale, at half the MMA rate.        */
} cplxk_math;

/* source of the local-reparameterisation noise */
typedef enum {
  CPLXK_NOISE_INJECT = 0,       /* read eps_re/eps_im planes [M,N] supplied by the caller */
  CPLXK_NOISE_PHILOX_TORCH = 1, /* regenerate, inside the epilogue, exactly the
                                   Philox4x32-10 stream `torch.randn(2,M,N,device='cuda')/sqrt(2)`
                                   (complex) or `torch.randn(M,N,device='cuda')` (real)
                                   would produce for (seed, offset, philox_threads)       */
  CPLXK_NOISE_PHILOX_FAST = 2   /* library-private counter layout: one Philox call
                                   feeds 4 normals of one thread (4x cheaper)             */
} cplxk_noise;

typedef enum {
  CPLXK_KL_REAL_VD = 0,  /* nn/relevance/real/vd.py:74-76     */
  CPLXK_KL_REAL_ARD = 1, /* nn/relevance/real/ard.py:39       */
  CPLXK_KL_CPLX_VD = 2,  /* nn/relevance/complex/vd.py:95-99  */
  CPLXK_KL_CPLX_ARD = 3, /* nn/relevance/complex/ard.py:39    */
  /* nn/relevance/extensions/complex.py */
  CPLXK_KL_CPLX_VD_APPROX = 4,    /* :77-100  softplus(-la) + 0.57810 sigmoid(-1.36526 la - 1.45926) */
  CPLXK_KL_CPLX_VD_SCALEFREE = 5  /* :18-44   log|w| - log_sigma2 - Ei(-1/alpha) / 2               */
} cplxk_kl_kind;

int cplxk_abi_version(void);
const char* cplxk_strerror(int status);

/* Persistent GEMM grids leave `n_sms` SMs free for kernels of other streams (a collective, a
 * KL shard kernel).  Process-wide, 0 by default (or CPLXK_SM_RESERVE at load). */
int cplxk_set_sm_reserve(int n_sms);

/* Number of SMs / compute capability of the current device (host ints). */
int cplxk_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * Complex (or real) affine map  y = x W^T + b.
 * Replaces cplx.linear == linear_naive (cplxmodule/cplx.py:634-648,698) and,
 * with x_im == w_im == NULL, torch.nn.functional.linear as used by
 * nn/relevance/real/base.py:44.
 *   x_re,x_im : [M,K]   w_re,w_im : [N,K]   b_re,b_im : [N] or NULL
 *   y_re,y_im : [M,N]   (y_im NULL for the real map)
 */
int cplxk_linear_fwd(const void* x_re, const void* x_im,
                     const void* w_re, const void* w_im,
                     const void* b_re, const void* b_im,
                     void* y_re, void* y_im,
                     int64_t M, int64_t N, int64_t K,
                     int dtype, int math, void* stream);

/*
 * Same map with a caller-provided scratch buffer (nullable; cplxk_linear_workspace_bytes()
 * bytes, 16-byte aligned, uninitialised).  With it, F32 planes (M > 128, K >= 64, K % 8 == 0)
 * are first rewritten as per-row power-of-two scaled fp16 copies (one pre-pass launch) and the
 * GEMM runs on kind::f16 -- tf32's 11-bit significand at twice the MMA rate -- in a persistent
 * CTA-pair kernel whose TMEM accumulators are double buffered.  Without it: cplxk_linear_fwd.
 */
size_t cplxk_linear_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype);
int cplxk_linear_fwd_ws(const void* x_re, const void* x_im,
                        const void* w_re, const void* w_im,
                        const void* b_re, const void* b_im,
                        void* y_re, void* y_im,
                        int64_t M, int64_t N, int64_t K,
                        int dtype, int math,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * Affine map of a fixed-sparsity ("masked") layer:  y = x (W * mask)^T + b.
 * Replaces CplxLinearMasked.forward / LinearMasked.forward, i.e.
 * cplx.linear(input, self.weight_masked, self.bias) with weight_masked = weight * mask
 * (nn/masked/complex.py:33-35, nn/masked/real.py:25-27, nn/masked/base.py:135-149).
 *   mask : [N,K] plane of `dtype` (0/1 or soft), applied to BOTH weight planes.
 * The mask is applied where the weights are staged for the GEMM: F32 planes on the tensor-core
 * path multiply inside the operand pre-pass (no masked copy of W is ever written); every other
 * path writes W * mask for both planes in ONE elementwise launch into the workspace and runs
 * the ordinary kernels on that.  workspace: cplxk_linear_masked_workspace_bytes() bytes.
 */
size_t cplxk_linear_masked_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype);
int cplxk_linear_masked_fwd(const void* x_re, const void* x_im,
                            const void* w_re, const void* w_im, const void* mask,
                            const void* b_re, const void* b_im,
                            void* y_re, void* y_im,
                            int64_t M, int64_t N, int64_t K,
                            int dtype, int math,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * Local-reparameterisation forward of a Gaussian (variational dropout) linear
 * layer, fused: mean GEMM(s), variance GEMM |x|^2 . exp(log_sigma2)^T, noise,
 * and  y = mu + eps * sqrt(max(s2, 1e-8)).
 * Replaces CplxLinearGaussian.forward (nn/relevance/complex/base.py:43-56)
 * and, with x_im == w_im == NULL, LinearGaussian.forward
 * (nn/relevance/real/base.py:43-49).
 *   log_sigma2 : [N,K]
 *   noise == INJECT        : eps_re (,eps_im) are [M,N] planes of dtype
 *   noise == PHILOX_*      : eps_* ignored; (seed, offset) as in torch's CUDA
 *                            generator state (offset % 4 == 0);
 *                            philox_threads = 256 * grid of torch's randn kernel
 *                            (only used by PHILOX_TORCH).
 *   workspace (nullable)   : cplxk_linear_vd_workspace_bytes() bytes, 16-byte aligned,
 *                            uninitialised scratch.  With it the call first writes the GEMM
 *                            operands there (one pre-pass launch) and the GEMM kernel streams
 *                            them by TMA: for F32 planes per-row power-of-two scaled fp16
 *                            copies of x and W, bf16 |x|^2 and exp(log_sigma2), and the
 *                            inverse row scales (M > 128, K >= 64, K % 8 == 0; otherwise, and
 *                            for BF16 planes, only |x|^2 [M,K] and exp(log_sigma2) [N,K]);
 *                            without it the derived operands are produced inside the GEMM
 *                            kernel's shared-memory pipeline (one launch, slower mainloop).
 */
size_t cplxk_linear_vd_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype);
int cplxk_linear_vd_fwd(const void* x_re, const void* x_im,
                        const void* w_re, const void* w_im,
                        const void* b_re, const void* b_im,
                        const void* log_sigma2,
                        const void* eps_re, const void* eps_im,
                        int noise, uint64_t seed, uint64_t offset,
                        uint32_t philox_threads,
                        void* y_re, void* y_im,
                        int64_t M, int64_t N, int64_t K,
                        int dtype, int math, void* s2_out /* nullable [M,N]: saved variance */,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * The operand pre-pass of the F32 tensor-core path on its own: fills `workspace` with the
 * row-scaled fp16 copies of x and W, bf16 |x|^2 and exp(log_sigma2) and the inverse row scales
 * (and, with kl_kind >= 0, writes the layer's KL sum to *kl_sum) WITHOUT launching the GEMM.
 * What cplxk_linear_vd_fwd[_kl] runs first; exposed so that the HBM-bound stage can be timed
 * and profiled by itself.  CPLXK_ERR_UNSUPPORTED unless F32, M > 128, K >= 64, K % 8 == 0.
 */
int cplxk_linear_vd_prepare(const void* x_re, const void* x_im,
                            const void* w_re, const void* w_im,
                            const void* log_sigma2,
                            int64_t M, int64_t N, int64_t K, int dtype,
                            void* workspace, size_t workspace_bytes,
                            int kl_kind, float* kl_sum,
                            void* kl_workspace, size_t kl_workspace_bytes, void* stream);

/*
 * Same forward, with the layer's KL penalty as a by-product.  The operand pre-pass of the
 * tensor-core path reads every weight row and log_sigma2 anyway; with kl_kind >= 0 it also
 * evaluates sum(penalty(kl_kind, log_alpha)) over all N*K parameters (see cplxk_kl below: the
 * same per-element device function, the same deterministic reduction) and writes it to
 * *kl_sum -- no second HBM pass over the parameters.  *kl_done (HOST int) is set to 1 when the
 * path taken produced the sum (fp32 planes, tensor cores, workspace given, M > 128, K % 8 == 0)
 * and to 0 otherwise, in which case the caller runs cplxk_kl.  kl_kind < 0: plain forward.
 *   kl_workspace : cplxk_kl_workspace_bytes() bytes, as for cplxk_kl.
 *   kl_row_begin, kl_row_end : only weight rows [begin, end) enter the sum (kl_row_end < 0: all
 *                  N rows) -- a rank's shard of the row-sharded multi-GPU KL, whose partial
 *                  sums are combined by ONE all-reduce;
 *   kl_event     : cudaEvent_t (nullable) recorded on `stream` right after the pre-pass launch:
 *                  *kl_sum is final there, so that all-reduce can run on another stream while
 *                  the GEMM kernel computes.
 *   kl_fingerprint : device uint64 (nullable): fingerprint of the parameters the sum was computed
 *                  from, written with the sum; cplxk_kl_guard compares it with the parameters
 *                  as they are when the sum is handed out.
 * Replaces the pair CplxLinearGaussian.forward + CplxVDMixin.penalty.sum()
 * (nn/relevance/complex/base.py:43-56, complex/vd.py:95-99, relevance/base.py:135-139).
 */
int cplxk_linear_vd_fwd_kl(const void* x_re, const void* x_im,
                           const void* w_re, const void* w_im,
                           const void* b_re, const void* b_im,
                           const void* log_sigma2,
                           const void* eps_re, const void* eps_im,
                           int noise, uint64_t seed, uint64_t offset,
                           uint32_t philox_threads,
                           void* y_re, void* y_im,
                           int64_t M, int64_t N, int64_t K,
                           int dtype, int math, void* s2_out,
                           void* workspace, size_t workspace_bytes,
                           int kl_kind, float* kl_sum,
                           void* kl_workspace, size_t kl_workspace_bytes,
                           int64_t kl_row_begin, int64_t kl_row_end, void* kl_event,
                           void* kl_fingerprint, int* kl_done, void* stream);
/* 1 if cplxk_linear_vd_fwd_kl would produce the KL by-product for this shape / dtype / math mode
 * (given a workspace), 0 if it runs a path without the fused pre-pass */
int cplxk_linear_vd_fuses_kl(int64_t M, int64_t N, int64_t K, int dtype, int math);


/*
 * Outer-product features of the bilinear layers: z[b, p * d2 + q] = conj?(x1[b, p]) * x2[b, q].
 * With them  cplx.bilinear(x1, x2, W, b, conjugate)  (cplxmodule/cplx.py:1062-1090; CplxBilinear,
 * nn/modules/linear.py:67-117)  ==  cplxk_linear_fwd(z, W viewed as [out, d1 * d2], b), and the
 * variance  F.bilinear(|x1|^2, |x2|^2, exp(log_sigma2))  of CplxBilinearGaussian.forward
 * (nn/relevance/complex/base.py:70-84) is the variance GEMM of cplxk_linear_vd_fwd on the same z.
 * Real layers (torch.nn.Bilinear, nn/relevance/real/base.py:52-80): x1_im == x2_im == z_im == NULL.
 *   x1 : [B,d1]   x2 : [B,d2]   z : [B, d1*d2]
 * cplxk_outer_bwd: gradients of a scalar loss through z given g = dL/dz (either output pair
 * may be NULL).
 */
int cplxk_outer_fwd(const void* x1_re, const void* x1_im, const void* x2_re, const void* x2_im,
                    void* z_re, void* z_im, int64_t B, int64_t d1, int64_t d2, int conjugate,
                    int dtype, void* stream);
int cplxk_outer_bwd(const void* g_re, const void* g_im, const void* x1_re, const void* x1_im,
                    const void* x2_re, const void* x2_im, void* d1_re, void* d1_im, void* d2_re,
                    void* d2_im, int64_t B, int64_t d1, int64_t d2, int conjugate, int dtype,
                    void* stream);

/*
 * KL penalty of a variational layer over n parameters, one HBM pass:
 *   log_alpha = log_sigma2 - 2 log(|w| + 1e-12)
 *     (nn/relevance/real/base.py:23-26, complex/base.py:27-31)
 *   penalty(kind, log_alpha)            (see cplxk_kl_kind)
 * out_elem (nullable): per-element penalty, dtype planes  (reduction=None)
 * out_sum  (nullable): one float, scale * sum(penalty)    (reduction="sum"/"mean",
 *                      nn/relevance/base.py:135-139)
 * w_im must be NULL for the REAL kinds (0, 1) and non-NULL for the complex ones (2..5).
 * workspace: cplxk_kl_workspace_bytes() bytes, 16-byte aligned, zero-filled
 *            once by the caller before first use (the kernel restores it).
 */
size_t cplxk_kl_workspace_bytes(void);
int cplxk_kl(int kind, const void* w_re, const void* w_im,
             const void* log_sigma2, int64_t n, int dtype,
             void* out_elem, float* out_sum, double scale,
             void* workspace, size_t workspace_bytes, void* stream);

/*
 * The KL sum AND the relevance mask (log_alpha <= threshold, as floats of `dtype`) of a layer
 * from ONE pass over its parameters: what a sparsification schedule asks for every time it
 * logs the penalty and re-derives the masks (named_penalties + compute_ard_masks,
 * nn/relevance/base.py:88-141,192-216; RelevanceMixin.relevance, complex/vd.py:50-53).
 * out_sum nullable (mask only); out_mask required.
 */
int cplxk_kl_mask(int kind, const void* w_re, const void* w_im,
                  const void* log_sigma2, int64_t n, int dtype,
                  float threshold, void* out_mask, float* out_sum, double scale,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Guard of the KL by-product of cplxk_linear_vd_fwd_kl (the reference recomputes the penalty from
 * the current parameters on every penalties() call, nn/relevance/base.py:88-141; handing out a
 * sum computed earlier is only valid while the parameters are unchanged).  A 64-bit fingerprint
 * of the first 8 entries of every row of w_re, w_im (nullable), log_sigma2 ([N, K] each):
 *   fp_ref == NULL : record   -> fp_out[0] (device uint64)
 *   fp_ref != NULL : compare  -> out_sum[0] = fused_sum[0] if the fingerprint still equals
 *                    fp_ref[0], NaN otherwise; *stale_flag (nullable, host-mapped int) = 1 then.
 * kl_workspace: cplxk_kl_workspace_bytes() bytes as for cplxk_kl (block sums + ticket).
 * One row head per thread over N / 128 blocks, asynchronous on `stream`.
 */
int cplxk_kl_guard(const void* w_re, const void* w_im, const void* log_sigma2,
                   int64_t N, int64_t K, int dtype, void* fp_out, const void* fp_ref,
                   const float* fused_sum, float* out_sum, int* stale_flag,
                   void* kl_workspace, size_t kl_workspace_bytes, void* stream);

/*
 * log_alpha itself and the relevance mask  (log_alpha <= threshold)
 * (complex/vd.py:47-60, real/vd.py:13-24).  Either output may be NULL.
 */
int cplxk_log_alpha(const void* w_re, const void* w_im, const void* log_sigma2,
                    int64_t n, int dtype, void* out_log_alpha,
                    float threshold, void* out_mask, void* stream);

/*
 * Complex (or real: x_im == w_im == y_im == NULL) 2-d cross-correlation, NCHW, zero padding.
 * Replaces cplx.conv2d -> convnd_quick (cplxmodule/cplx.py:729-742,822-838) and the F.conv2d
 * calls of the real layers (nn/relevance/real/base.py:149-163); with log_sigma2 != NULL also the
 * variational forward CplxConvNdGaussianMixin._forward_impl (nn/relevance/complex/base.py:120-135)
 * / ConvNdGaussianMixin._forward_impl (real/base.py:149-163).  conv1d is the H == kh == 1 case.
 *   x : [B,C,H,W]  w,log_sigma2 : [O,C/groups,kh,kw]  b : [O]  y,eps : [B,O,Ho,Wo]
 *   workspace (nullable): cplxk_conv2d_workspace_bytes[_g]() bytes of scratch, 16-byte aligned.
 *     With it the call runs the tcgen05 implicit GEMM: two elementwise launches write
 *     channels-last copies of the input (and |x|^2, exp(log_sigma2), tap-major weights) to the
 *     workspace, one kernel does the MMAs + epilogue.  Complex planes with groups == 1: CTA-pair
 *     / persistent kernels on the stacked [U;V] operand.  Real planes: one A tile and one
 *     accumulator per 128 real output channels.  groups > 1 (cplx.py:717-726 passes `groups` on
 *     to F.conv): ONE launch, the group is a factor of the n-block index.
 *     Without a workspace, or for geometries outside the TMA box limits, the exact-fp32
 *     CUDA-core kernel runs (groups == 1 only: CPLXK_ERR_UNSUPPORTED otherwise, the caller
 *     then issues one call per group).
 *   Variance operand of the complex variational layers (F.conv(abs(input)**2, exp(log_sigma2)),
 *   complex/base.py:100-117): x_im given with w_im == y_im == NULL is the REAL convolution of
 *   |x_re + i x_im|^2 with w_re -- the square is formed inside the transposing pre-pass, no
 *   |x|^2 plane is written.  Tensor-core path only (workspace of the real-plane size, NCHW planes
 *   with W % 8 == 0 (bf16) / W % 4 == 0 (fp32), not variational): CPLXK_ERR_UNSUPPORTED otherwise.
 */
size_t cplxk_conv2d_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t O,
                                    int64_t kh, int64_t kw, int dtype, int variational);
size_t cplxk_conv2d_workspace_bytes_g(int64_t B, int64_t C, int64_t H, int64_t W, int64_t O,
                                      int64_t kh, int64_t kw, int64_t groups, int is_complex,
                                      int dtype, int variational);
int cplxk_conv2d_fwd(const void* x_re, const void* x_im,
                     const void* w_re, const void* w_im,
                     const void* b_re, const void* b_im,
                     const void* log_sigma2,
                     const void* eps_re, const void* eps_im,
                     int noise, uint64_t seed, uint64_t offset,
                     uint32_t philox_threads,
                     void* y_re, void* y_im,
                     int64_t B, int64_t C, int64_t H, int64_t W,
                     int64_t O, int64_t kh, int64_t kw,
                     int64_t stride_h, int64_t stride_w,
                     int64_t pad_h, int64_t pad_w,
                     int64_t dil_h, int64_t dil_w,
                     int dtype, int math,
                     int channels_last /* 1: x and y planes are NHWC in memory (torch.channels_last);
                                          tensor-core path only, complex planes, groups == 1,
                                          C % 8 == 0 (fp32) / 16 (bf16); skips the transposing
                                          pre-pass.  eps stays NCHW. */,
                     void* workspace, size_t workspace_bytes, void* stream);
/* the same with `groups` (cplxk_conv2d_fwd is groups == 1) */
int cplxk_conv2d_fwd_g(const void* x_re, const void* x_im,
                       const void* w_re, const void* w_im,
                       const void* b_re, const void* b_im,
                       const void* log_sigma2,
                       const void* eps_re, const void* eps_im,
                       int noise, uint64_t seed, uint64_t offset,
                       uint32_t philox_threads,
                       void* y_re, void* y_im,
                       int64_t B, int64_t C, int64_t H, int64_t W,
                       int64_t O, int64_t kh, int64_t kw,
                       int64_t stride_h, int64_t stride_w,
                       int64_t pad_h, int64_t pad_w,
                       int64_t dil_h, int64_t dil_w,
                       int64_t groups,
                       int dtype, int math, int channels_last,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * ---- backward pass ---------------------------------------------------------------------
 * Every gradient GEMM runs through cplxk_linear_fwd on transposed / conjugated operands
 * (dx = g . conj(W), dW = g^T . conj(x), dq = g_s2 . E, dE = g_s2^T . q); the entry points
 * below produce those operands and the elementwise pieces.  The reference obtains all of
 * this from torch autograd over cplx.py:634-648 and nn/relevance/complex/base.py:43-56; its
 * one hand-written derivative is ExpiFunction.backward (complex/vd.py:38-41).
 */
/* out[cols, rows] = op(in[rows, cols]); op: 0 copy, 1 negate, 2 exp, 3 in^2 + in2^2, 4 in^2 */
int cplxk_transpose2d(const void* in, const void* in2, void* out, int64_t rows, int64_t cols,
                      int dtype, int op, void* stream);
/* out = op(a [, b]) elementwise, same op codes plus 5: a * b, no transposition (conv backward:
 * |x|^2, exp(log_sigma2), conj; masked conv layers: weight * mask) */
int cplxk_eltwise(int op, const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
/* out[N] = sum over rows of g[M, N]  (bias gradient) */
int cplxk_colsum(const void* g, void* out, int64_t M, int64_t N, int dtype, void* stream);
/* g_s2 = (g_re eps_re + g_im eps_im) / (2 sqrt(s2)) where s2 > 1e-8 (eps as in the forward) */
int cplxk_vd_grad_s2(const void* g_re, const void* g_im, const void* s2, const void* eps_re,
                     const void* eps_im, int noise, uint64_t seed, uint64_t offset,
                     uint32_t philox_threads, void* out, int64_t M, int64_t N, int dtype,
                     void* stream);
/* y += eps sqrt(max(s2, 1e-8)) in place (y_im, eps_im nullable: real planes); eps injected,
 * drawn in torch's layout (generated by Philox call: three calls per four complex outputs) or in
 * the conv kernels' private `fast` layout.  The variational conv forward
 * (nn/relevance/complex/base.py:120-135, real/base.py:149-163) as mean conv + variance conv of the
 * fast kernels + this. */
int cplxk_vd_combine(void* y_re, void* y_im, const void* s2, const void* eps_re,
                     const void* eps_im, int noise, uint64_t seed, uint64_t offset,
                     uint32_t philox_threads, int64_t numel, int dtype, void* stream);
/* dx_re += 2 x_re dq, dx_im += 2 x_im dq  (in place) */
int cplxk_vd_grad_input(void* dx_re, void* dx_im, const void* x_re, const void* x_im,
                        const void* dq, int64_t n, int dtype, void* stream);
/* out (+)= a * exp(b) */
int cplxk_mul_exp(const void* a, const void* b, void* out, int64_t n, int dtype, int accumulate,
                  void* stream);
/* gradient of scale * sum(grad * penalty) w.r.t. (w_re, w_im, log_sigma2); `grad` is a device
 * scalar (grad_is_tensor == 0) or an [n] plane, fp32 (grad_is_f32) or of `dtype` */
int cplxk_kl_bwd(int kind, const void* w_re, const void* w_im, const void* log_sigma2, int64_t n,
                 int dtype, const void* grad, int grad_is_tensor, int grad_is_f32, double scale,
                 void* d_w_re, void* d_w_im, void* d_log_sigma2, void* stream);

/*
 * Test hook: fill out[n] (float) with scale * N(0,1) using the very device
 * function the VD epilogue uses for CPLXK_NOISE_PHILOX_TORCH, i.e. element i
 * equals torch.randn(n, device='cuda')[i] * scale for the same generator state.
 */
int cplxk_randn_philox_torch(float* out, int64_t n, uint64_t seed,
                             uint64_t offset, uint32_t philox_threads,
                             float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CPLXK_H_ */
