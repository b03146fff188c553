/* dktb200 -- C ABI of the B200-native DKT per-episode GP inference path.
 *
 * The reference (BayesWatch/deep-kernel-transfer) is pure Python and has no FFI: its boundary for this
 * path is the Python plugin API of methods/DKT.py / methods/DKT_regression.py / backbone.py (kept by
 * deep_kernel_transfer_b200/methods + backbone).  This header is the boundary *underneath* that API:
 * every entry point replaces one implicit library call the reference makes through PyTorch/GPyTorch
 * (cited per function).  Conventions:
 *   - extern "C", plain pointers and sizes; every pointer is a DEVICE pointer owned by the caller;
 *   - explicit cudaStream_t; no hidden allocation, no global mutable state (re-entrant across streams);
 *   - caller-provided scratch (sizes stated per function);
 *   - return 0 = ok, <0 = bad argument, >0 = cudaError_t of the launch;
 *   - activations are NHWC fp32; "padded" tensors are [img][H+2][W+2][64] with a zero border that the
 *     caller zero-initialises once (kernels only ever write the interior);
 *   - E episodes are packed along the image axis, `ipe` images per episode; every BatchNorm batch is
 *     one episode (methods/DKT.py:140-141).
 */
#ifndef DKTB200_H
#define DKTB200_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

int dktb_version(void);

/* ---- backbone: ConvBlock = Conv2d(3x3,pad 1,bias) -> BatchNorm2d -> ReLU -> MaxPool2d(2)
 *      (backbone.py:105-132; nn.Conv2d / nn.BatchNorm2d / nn.MaxPool2d calls at 115-121) -------------- */
/* first layer, Cin = 3: x [B,3,H,W] NCHW (as the loader delivers it, train.py:142) -> y [B,H,W,64];
 * partials (nullable) [B*tiles][2][64] BatchNorm partial sums, tiles = dktb_conv1_tiles(H,W). */
int dktb_conv1_tiles(int H, int W);
int dktb_conv1_fwd(const float* x, const float* w, const float* bias, float* y, float* partials, int B, int H, int W,
                   cudaStream_t stream);
/* dw [64,3,3,3], db [64] (nullable); scratch: dktb_conv1_wgrad_nsplit()*28*64 floats. */
int dktb_conv1_wgrad_nsplit(void);
int dktb_conv1_wgrad(const float* x, const float* gy, float* dw, float* db, float* scratch, int B, int H, int W,
                     cudaStream_t stream);
/* Fused first-block backward: BatchNorm + ReLU + MaxPool2d(2) backward of the pooled gradient `gout`
 * [B,H/2+2*out_pad,W/2+2*out_pad,64] evaluated tile-wise and consumed directly by the conv1 weight / bias gradient
 * (the gradient of the pre-BN map never touches HBM).  Call dktb_bn_relu_pool_bwd(..., gy = NULL, ...) first: it
 * produces `sums` [B/ipe][2][64] and dgamma / dbeta.  Replaces loss.backward() through trunk[0] (backbone.py:93-102).
 * scratch: dktb_conv1_wgrad_nsplit()*28*64 floats. */
int dktb_conv1_bwd_fused(const float* x, const float* y, const float* gout, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const float* sums, float* dw, float* db, float* scratch,
                         int B, int H, int W, int ipe, int out_pad, cudaStream_t stream);
/* The same contract with the weight-gradient GEMM on warp-level tensor-core MMAs (mma.sync m16n8k8 tf32, 3xTF32). */
int dktb_conv1_bwd_fused_mma(const float* x, const float* y, const float* gout, const float* mean, const float* invstd,
                             const float* gamma, const float* beta, const float* sums, float* dw, float* db,
                             float* scratch, int B, int H, int W, int ipe, int out_pad, cudaStream_t stream);
/* Fixed-order reduction of nsplit per-CTA partials [28][64] (27 taps + bias row) into dw [64,3,3,3] / db [64]. */
int dktb_conv1_wgrad_reduce(const float* partial, int nsplit, float* dw, float* db, cudaStream_t stream);
/* weight re-layout for the 64->64 kernels: w [64,64,3,3] -> wt_fwd [9][ci][co], wt_dgrad [9][co][ci] (flipped). */
int dktb_prep_weights(const float* w, float* wt_fwd, float* wt_dgrad, cudaStream_t stream);
/* 64->64 3x3 conv over padded NHWC: out[q] = sum_tap A[q+off(tap)] * wt[tap]; forward (wt_fwd, bias, partials)
 * and dgrad (a = gy, wt = wt_dgrad, bias = partials = NULL).  partials [B*tiles][2][64], tiles =
 * dktb_conv3x3_tiles(H,W).  fp32 CUDA-core version; dktb_conv3x3_tc_* is the tcgen05 3xTF32 version. */
int dktb_conv3x3_tiles(int H, int W);
int dktb_conv3x3_fwd(const float* a, const float* wt, const float* bias, float* out, float* partials, int B, int H,
                     int W, cudaStream_t stream);
int dktb_conv3x3_wgrad_nsplit(void);
long dktb_conv3x3_wgrad_scratch_floats(void);
int dktb_conv3x3_wgrad(const float* a, const float* gy, float* dw, float* db, float* scratch, int B, int H, int W,
                       cudaStream_t stream);

/* tcgen05 + TMA version of the 64->64 convolution (forward and dgrad), fp32-class accuracy via an error-compensated
 * 3xTF32 split.  wb = [9][2][64][64] hi/lo weight tensor written by dktb_prep_weights_tc ([tap][hl][n][k]; wb_dgrad has
 * the taps flipped and n/k swapped).  err: device int, zero-initialised by the caller, set to 1 when a pipeline
 * barrier wait timed out (a bug guard: results are then invalid).  Same layouts/partials as dktb_conv3x3_fwd.
 * Persistent CTAs (one per SM): activation halo loaded once per tile by TMA, A operand split in registers and staged in
 * TMEM, weight ring streaming across tiles, double-buffered halo and accumulator, dedicated epilogue warps. */
long dktb_conv3x3_tc_weight_floats(void);      /* floats per prepared tensor: tf32 [9][hi|lo][64][64] + bf16 w_hi [9][64][64] */
int dktb_prep_weights_tc(const float* w, float* wb_fwd, float* wb_dgrad, cudaStream_t stream);
int dktb_conv3x3_tc_fwd(const float* a, const float* wb, const float* bias, float* out, float* partials, int* err,
                        int B, int H, int W, cudaStream_t stream);
/* ResNet layers on tcgen05 (reference backbone.py:135-247): any Cin, Cout that are multiples of 64, 3x3 / stride 1 /
 * pad 1 over padded-flat NHWC tensors [B][H+2][W+2][C] (zero border; interior of `out` written) or 1x1 / stride 1 as a
 * plain GEMM over the rows of dense [B*H*W][C] tensors; forward (wb_fwd + bias) and dgrad (wb_dgrad, Cin / Cout swapped
 * by the caller); 3xTF32 error-compensated like dktb_conv3x3_tc_fwd.  dktb_conv_tcg_ok tells whether a layer qualifies;
 * the weight tensors hold dktb_conv_tcg_weight_floats(Cin, Cout, R) floats each.  dktb_pad_copy moves a dense tensor
 * into the interior of a (zero-bordered) padded one (dir 0) or back (dir 1); dktb_zero_border clears the border. */
int dktb_conv_tcg_ok(int Cin, int Cout, int R, int stride, int pad, int dil, int W);
long dktb_conv_tcg_weight_floats(int Cin, int Cout, int R);
int dktb_prep_weights_tcg(const float* w, float* wb_fwd, float* wb_dgrad, int Cout, int Cin, int R, cudaStream_t stream);
int dktb_conv_tcg(const float* a, const float* wb, const float* bias, float* out, int* err, int B, int H, int W, int Cin,
                  int Cout, int R, cudaStream_t stream);
/* Stride-2 3x3 / pad 1 layers (backbone.py:150-152, 195-197) on the same tcgen05 kernel: a stride-1 convolution over the
 * space-to-depth input xs [B][H/2+2][W/2+2][4C] (padded-flat, channel = ((ih&1)*2 + (iw&1))*C + c), the 27 structurally
 * zero (parity plane, tap) pairs skipped.  dktb_s2d packs x (dense, or padded-flat when x_pad) into xs (dir 0) or unpacks
 * (dir 1); dktb_conv_tcg_s2: dgrad = 0: xs -> out [B][Ho+2][Wo+2][Cout]; dgrad = 1: dy [.., Cout] -> dxs [.., 4C]. */
int dktb_conv_tcg_s2_ok(int C, int Cout, int H, int W);
long dktb_conv_tcg_s2_weight_floats(int C, int Cout);
int dktb_prep_weights_tcg_s2(const float* w, float* wb_fwd, float* wb_dgrad, int Cout, int C, cudaStream_t stream);
int dktb_s2d(float* x, float* xs, int B, int H, int W, int C, int x_pad, int dir, cudaStream_t stream);
/* stride-2 1x1 shortcuts: gather x[:, ::2, ::2, :] into a dense tensor (dir 0; then dktb_conv_tcg with R = 1) or scatter a
 * gradient back, zeros at the skipped pixels (dir 1) */
int dktb_subsample2(float* x, float* xg, int B, int H, int W, int C, int x_pad, int dir, cudaStream_t stream);
int dktb_conv_tcg_s2(const float* a, const float* wb, const float* bias, float* out, int* err, int B, int Ho, int Wo, int C,
                     int Cout, int dgrad, cudaStream_t stream);
/* weight gradient of the same layers (W/2 <= 29: the four parity planes of 64 channels are staged together): xs, gy
 * padded-flat with zero borders -> dw [Cout][C][3][3], db [Cout] or NULL */
int dktb_wgrad_tcg_s2_ok(int C, int Cout, int H, int W);
long dktb_wgrad_tcg_s2_scratch_floats(int B, int Ho, int Wo, int C, int Cout);
int dktb_wgrad_tcg_s2(const float* xs, const float* gy, float* dw, float* db, float* scratch, int* err, int B, int Ho, int Wo,
                      int C, int Cout, cudaStream_t stream);

/* ResNet stem (backbone.py:336-340: Conv2d(3, 64, 7, stride 2, padding 3)) on tcgen05: x [B,3,H,W] NCHW -> y [B,H/2,W/2,64]
 * NHWC; implicit GEMM over k = ci*49 + r*7 + s (147 -> 160), 3xTF32.  wb: dktb_stem_tc_weight_floats() floats written by
 * dktb_prep_weights_stem_tc from w [64][3][7][7].  err: device int (zero-initialised), 1 = a pipeline wait timed out. */
int dktb_stem_tc_ok(int Cin, int Cout, int R, int stride, int pad, int dil, int H, int W);
long dktb_stem_tc_weight_floats(void);
int dktb_prep_weights_stem_tc(const float* w, float* wb, cudaStream_t stream);
int dktb_stem_tc(const float* x, const float* wb, const float* bias, float* y, int* err, int B, int H, int W,
                 cudaStream_t stream);
/* its weight gradient (K = output pixels; the im2col rows are gathered transposed into TMEM): x NCHW, gy [B,H/2,W/2,64] NHWC
 * -> dw [64][3][7][7]; scratch: dktb_stem_wgrad_tc_scratch_floats(B, H) floats; deterministic (fixed-order reduce) */
long dktb_stem_wgrad_tc_scratch_floats(int B, int H);
int dktb_stem_wgrad_tc(const float* x, const float* gy, float* dw, float* scratch, int* err, int B, int H, int W,
                       cudaStream_t stream);

/* weight gradient of the same layers on tcgen05: x [.., Cin], gy [.., Cout] in the layouts of dktb_conv_tcg (R = 3: both
 * padded-flat with zero borders; R = 1: dense rows) -> dw [Cout][Cin][R][R], db [Cout] or NULL; scratch holds
 * dktb_wgrad_tcg_scratch_floats(...) floats; per-CTA partials are reduced in a fixed order (deterministic). */
long dktb_wgrad_tcg_scratch_floats(int B, int H, int W, int Cin, int Cout, int R);
int dktb_wgrad_tcg(const float* x, const float* gy, float* dw, float* db, float* scratch, int* err, int B, int H, int W,
                   int Cin, int Cout, int R, cudaStream_t stream);
int dktb_zero_border(float* x, int B, int H, int W, int C, cudaStream_t stream);
int dktb_pad_copy(float* dense, float* padded, int B, int H, int W, int C, int dir, cudaStream_t stream);

/* First layer (3->64, K=27 padded to 32) on tcgen05: im2col rows gathered from an NCHW patch, split and staged in TMEM.
 * wb1 [2][64][32] from dktb_prep_weights_conv1_tc.  mode 0: y + BatchNorm partials (tile numbering of dktb_conv1_fwd);
 * mode 1: partials only; mode 2: fused BatchNorm(mean, invstd: [B/ipe][64], ipe == 0: one row) + ReLU + MaxPool2d(2)
 * writing the zero-bordered block output act [B,H/2+2,W/2+2,64] -- the pre-BN tensor never reaches HBM. */
int dktb_prep_weights_conv1_tc(const float* w, float* wb1, cudaStream_t stream);
int dktb_conv1_tc(const float* x, const float* wb1, const float* bias, float* y, float* partials, const float* mean,
                  const float* invstd, const float* gamma, const float* beta, float* act, int* err, int B, int H, int W,
                  int ipe, int mode, cudaStream_t stream);

/* tcgen05 wgrad (A = X^T staged in TMEM, B = gy re-laid-out K-major in smem, 3xTF32); same contract as
 * dktb_conv3x3_wgrad plus the err flag.  dktb_conv3x3_wgrad_reduce: fixed-order reduction of per-CTA partials
 * [nsplit][9*64*64 + 64] into dw [64,64,3,3] / db [64]. */
int dktb_conv3x3_wgrad_tc(const float* a, const float* gy, float* dw, float* db, float* scratch, int* err, int B, int H,
                          int W, cudaStream_t stream);
int dktb_conv3x3_wgrad_reduce(const float* partial, int nsplit, float* dw, float* db, cudaStream_t stream);

/* Generic NHWC fp32 convolution (any Cin/Cout, RxS, stride, padding, dilation) for the layers outside the 64->64 3x3
 * case: Conv3 of the regression path (backbone.py:379-402) and strided / 1x1 / 7x7 ResNet layers.  Weights in the
 * reference layout [Cout][Cin][R][S].  relu != 0: forward applies ReLU; dgrad / wgrad then take the forward output
 * `yout` and mask gy with it.  wgrad scratch: dktb_conv2d_wgrad_nsplit(N*Ho*Wo)*R*S*Cin*Cout floats. */
int dktb_conv2d_out_size(int H, int R, int stride, int pad, int dil);
int dktb_conv2d_fwd(const float* x, const float* w, const float* bias, float* out, int N, int H, int W, int Cin,
                    int Cout, int R, int S, int stride, int pad, int dil, int relu, cudaStream_t stream);
int dktb_conv2d_dgrad(const float* gy, const float* yout, const float* w, float* gx, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int dil, int relu, cudaStream_t stream);
int dktb_conv2d_wgrad_nsplit(long npix);
int dktb_conv2d_wgrad(const float* x, const float* gy, const float* yout, float* dw, float* db, float* scratch, int N,
                      int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int relu,
                      cudaStream_t stream);
/* Tensor-core variants (mma.sync 3xTF32 tile products) for layers whose reduction widths are multiples of 32 -- every
 * ResNet layer but the stem (backbone.py:135-247, 330-376); dktb_conv2d_mma_ok() says whether a layer qualifies
 * (DKTB_RESNET_CONV=fp32 turns them off).  Weights are read from two k-contiguous copies refreshed once per step by
 * dktb_conv2d_prep_mma: wf [R*S][Cout][Cin] (forward), wd [R*S][Cin][Cout] (dgrad).  No fused ReLU.  dktb_conv2d_wgrad
 * picks its tensor-core kernel by itself (same arguments, same scratch). */
int dktb_conv2d_mma_ok(int Cin, int Cout);
int dktb_conv2d_prep_mma(const float* w, float* wf, float* wd, int Cout, int Cin, int R, int S, cudaStream_t stream);
int dktb_conv2d_fwd_mma(const float* x, const float* wf, const float* bias, float* out, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, int dil, cudaStream_t stream);
int dktb_conv2d_dgrad_mma(const float* gy, const float* wd, float* gx, int N, int H, int W, int Cin, int Cout, int R,
                          int S, int stride, int pad, int dil, cudaStream_t stream);
/* Narrow inputs (the 3-channel 7x7 stem, backbone.py:343-347): the reduction runs over the flattened (r, s, ci) index,
 * weights from wflat [Cout][dktb_conv2d_flat_k(Cin, R, S)] written by dktb_conv2d_prep_flat_mma. */
int dktb_conv2d_flat_k(int Cin, int R, int S);
int dktb_conv2d_prep_flat_mma(const float* w, float* wflat, int Cout, int Cin, int R, int S, cudaStream_t stream);
int dktb_conv2d_fwd_flat_mma(const float* x, const float* wflat, const float* bias, float* out, int N, int H, int W,
                             int Cin, int Cout, int R, int S, int stride, int pad, int dil, cudaStream_t stream);
int dktb_nchw_to_nhwc(const float* x, float* out, int N, int C, int H, int W, cudaStream_t stream);

/* BatchNorm2d statistics (train: per-episode batch stats from the conv partial sums + running-stat EMA,
 * momentum 0.1, unbiased running variance; eval: running stats).  scratch_d: dktb_bn_scratch_doubles(B/ipe). */
long dktb_bn_scratch_doubles(int E);
int dktb_bn_finalize(const float* partials, int B, int T, int ipe, int hw, float* mean, float* invstd,
                     float* running_mean, float* running_var, double* scratch_d, float momentum, float eps,
                     cudaStream_t stream);
int dktb_bn_eval_prepare(const float* running_mean, const float* running_var, float* mean, float* invstd, int n,
                         float eps, cudaStream_t stream);
/* out = maxpool2(relu(bn(y))); y [B][H+2*in_pad][W+2*in_pad][64]; out [B][Ho+2*out_pad][Wo+2*out_pad][64];
 * mean/invstd [B/ipe][64] (ipe == 0: one row for all images = eval mode); pool = 0 keeps the resolution. */
int dktb_bn_relu_pool_fwd(const float* y, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, float* out, int B, int H, int W, int ipe, int in_pad, int out_pad,
                          int pool, cudaStream_t stream);
/* backward of the above (train mode): gy (same layout as y), dgamma/dbeta [64];
 * partial: B*dktb_bn_bwd_chunks(H,W,pool)*128 floats, sums: (B/ipe)*128 floats,
 * scratch_d: dktb_bn_scratch_doubles(B/ipe) doubles. */
int dktb_bn_bwd_chunks(int H, int W, int pool);
int dktb_bn_relu_pool_bwd(const float* y, const float* gout, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, float* gy, float* dgamma, float* dbeta,
                          float* partial, float* sums, double* scratch_d, int B, int H, int W, int ipe, int in_pad,
                          int out_pad, int pool, cudaStream_t stream);

/* ---- channel-generic NHWC blocks of the ResNet backbones (backbone.py:135-247, 330-376): BatchNorm2d with per-episode
 * batch statistics for any C (multiple of 4), optional fused residual add + ReLU; MaxPool2d(3,2,1); global AvgPool. */
long dktb_bn2d_partial_floats(int B, int HW, int C);   /* size of `partial` below (images x pixel splits x C x 2) */
int dktb_bn2d_stats(const float* x, float* mean, float* invstd, float* running_mean, float* running_var, float* partial,
                    int B, int HW, int C, int ipe, float momentum, float eps, cudaStream_t stream);
int dktb_bn2d_apply(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                    const float* res, float* y, int B, int HW, int C, int ipe, int relu, cudaStream_t stream);
int dktb_bn2d_bwd(const float* x, const float* y, const float* gy, const float* mean, const float* invstd,
                  const float* gamma, float* gx, float* gres, float* dgamma, float* dbeta, float* partial, float* sums,
                  int B, int HW, int C, int ipe, int relu, cudaStream_t stream);   /* sums: (B/ipe)*C*2 */
/* Layout-aware variants: every NHWC tensor argument is dense [B][H][W][C] or the interior of a padded-flat buffer
 * [B][H+2][W+2][C] (pointer = buffer base, border never touched) -- the layout dktb_conv_tcg reads and writes, so no copy
 * separates the tcgen05 convolutions from the BatchNorm kernels.  lay: bit 0 = x (add: a), bit 1 = y (add: b), bit 2 = gy,
 * bit 3 = gx, bit 4 = res / gres; W = row width (HW % W == 0). */
int dktb_bn2d_stats_l(const float* x, float* mean, float* invstd, float* running_mean, float* running_var, float* partial,
                      int B, int HW, int C, int ipe, float momentum, float eps, int W, int lay, cudaStream_t stream);
int dktb_bn2d_apply_l(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const float* res, float* y, int B, int HW, int C, int ipe, int relu, int W, int lay,
                      cudaStream_t stream);
int dktb_bn2d_bwd_l(const float* x, const float* y, const float* gy, const float* mean, const float* invstd,
                    const float* gamma, float* gx, float* gres, float* dgamma, float* dbeta, float* partial, float* sums,
                    int B, int HW, int C, int ipe, int relu, int W, int lay, cudaStream_t stream);
int dktb_add_inplace_l(float* a, const float* b, int B, int HW, int C, int W, int lay, cudaStream_t stream);
int dktb_maxpool3_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C, cudaStream_t stream);
int dktb_maxpool3_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C,
                      cudaStream_t stream);
int dktb_avgpool_fwd(const float* x, float* y, int B, int HW, int C, cudaStream_t stream);
int dktb_avgpool_bwd(const float* gy, float* gx, int B, int HW, int C, cudaStream_t stream);
int dktb_add_inplace(float* a, const float* b, long n, cudaStream_t stream);

/* ---- feature head: bn_out BatchNorm1d (methods/DKT.py:45-48) + F.normalize (methods/DKT.py:142) -------
 * features f [E][N][D] NHWC-flattened (j = p*Cch + c); parameters indexed in the reference's NCHW-flatten
 * order (c*P + p); P <= 1 means identity mapping. */
int dktb_bn1d_fwd(const float* f, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  float* z, float* mean, float* invstd, float* var, int E, int N, int D, int Cch, int P, int training,
                  int update_running, float momentum, float eps, cudaStream_t stream);
int dktb_bn1d_bwd(const float* f, const float* gz, const float* gamma, const float* mean, const float* invstd,
                  float* gf, float* dgamma, float* dbeta, float* pgrad /* E*2*D */, int E, int N, int D, int Cch,
                  int P, cudaStream_t stream);
int dktb_l2norm_fwd(const float* z, float* zhat, float* inv, long rows, int D, float eps, cudaStream_t stream);
int dktb_l2norm_bwd(const float* zhat, const float* gzhat, const float* inv, float* gz, long rows, int D,
                    cudaStream_t stream);

/* ---- exact GP (GPyTorch call sites methods/DKT.py:161-162,177,187,265; DKT_regression.py:52-54,90-93) -- */
/* out[e][m][n] = <x1[e][m], x2[e][n]>  (LinearKernel / cross-kernel) */
int dktb_gram(const float* x1, const float* x2, float* out, int E, int M, int N, int D, cudaStream_t stream);
/* the same product on tcgen05 / TMA (128 x 128 tiles, 3xTF32 error-compensated, fp32 accumulation in TMEM) for the
 * shapes dktb_gram_tc_ok accepts (D % 4 == 0, D >= 32, at least 128 rows per operand tensor); err as dktb_conv3x3_tc_fwd */
int dktb_gram_tc_ok(int E, int M, int N, int D);
long dktb_gram_tc_scratch_floats(int E, int M, int N, int D);      /* split-K partials, one per 64 features */
int dktb_gram_tc(const float* x1, const float* x2, float* out, float* scratch, int* err, int E, int M, int N, int D,
                 cudaStream_t stream);
/* per (episode, class): K~ = softplus(raw_outputscale_c)*Kb + (softplus(raw_noise_c)+1e-4) I -> Cholesky ->
 * alpha_c = K~^-1 (y_c - constant_c), loss_terms[e][c] = -log p_c/(N*C).
 * Cholesky follows GPyTorch's psd_safe_cholesky (utils/cholesky.py; the reference relies on it, README.md:27,
 * methods/DKT.py:162): a plain attempt, then -- when jitter > 0 -- retries with jitter, 10*jitter, 100*jitter added to the
 * diagonal (pass 1e-6 for the fp32 schedule).  info[e][c] = 0 ok | -k ok after the k-th retry | > 0 the 1-based failing
 * pivot of the last attempt; a failed system gets a NaN loss term and ZERO alpha / dkbase / dhyper.
 * Optional: linv [E][C][N][N]; dkbase [E][C][N][N] = grad_scale*dLoss/dKb_c; dhyper [E][C][3] =
 * grad_scale*dLoss/d(raw_outputscale, constant, raw_noise).  raw_outputscale == NULL: no ScaleKernel.
 * N <= dktb_gp_max_n(). */
int dktb_gp_max_n(void);
int dktb_gp_fit(const float* kbase, long kbase_class_stride, const float* y, long y_episode_stride,
                const float* raw_outputscale, const float* constant, const float* raw_noise, float* alpha,
                float* linv, float* loss_terms, int* info, float* dkbase, float* dhyper, float grad_scale,
                float jitter, int E, int C, int N, cudaStream_t stream);
/* The same contract for any N <= dktb_gp_large_max_n() = 512 (20-way training episodes, BASELINE configs[4] Gram-N
 * sweep; faster than dktb_gp_fit from N ~ 100 on): tiled blocked Cholesky (tensor-core tile products) on a
 * caller-provided global workspace of dktb_gp_large_work_floats(E, C, N) floats.  kbase must be symmetric (it is a
 * kernel matrix); the first pivot that is not positive is reported in info like dktb_gp_fit does. */
int dktb_gp_large_max_n(void);
long dktb_gp_large_work_floats(int E, int C, int N);
int dktb_gp_fit_large(const float* kbase, long kbase_class_stride, const float* y, long y_episode_stride,
                      const float* raw_outputscale, const float* constant, const float* raw_noise, float* alpha,
                      float* linv, float* loss_terms, int* info, float* dkbase, float* dhyper, float* work,
                      float grad_scale, float jitter, int E, int C, int N, cudaStream_t stream);
int dktb_gp_reduce(const float* loss_terms, const float* dhyper, float* loss, float* hyper, int E, int C,
                   cudaStream_t stream);

/* Cholesky status over many steps with one read-back: sticky[0] = max(sticky[0], max info), sticky[1] = min(sticky[1],
 * min info).  info > 0: not positive definite even after the jitter retries (GPyTorch psd_safe_cholesky raises there;
 * methods/DKT.py:162 relies on it, README.md:27); info = -k: factorised after the k-th retry (jitter * 10^(k-1)). */
int dktb_gp_info_accumulate(const int* info, int* sticky, int n, cudaStream_t stream);
/* dZ = scale * (S + S^T) Z,  S = sum_c w[e][c] */
int dktb_gram_bwd(const float* w, const float* z, float* dz, int E, int C, int N, int D, float scale,
                  cudaStream_t stream);
/* mean[e][c][m] = constant_c + s_c * kx[e][(c)][m][:] . alpha[e][c][:]; pred[e][m] = argmax_c sigmoid(mean) */
int dktb_gp_predict(const float* kx, long kx_class_stride, const float* alpha, const float* raw_outputscale,
                    const float* constant, float* mean, int* pred, int E, int C, int M, int N, cudaStream_t stream);

/* ---- base-kernel family (ExactGPLayer, methods/DKT.py:352-372; DKT_regression.py:117-124) as an epilogue on the Gram
 * matrix of mean-centred features.  kind: 0 linear (softplus(raw)=variance), 1 rbf, 2 matern-2.5 (lengthscale),
 * 3 poli1, 4 poli2 (offset).  kb [E][C][M][N]: one kernel matrix per one-vs-rest model. */
int dktb_center_rows(const float* x, const float* ref, float* out, int E, int N, int Nr, int D, cudaStream_t stream);
int dktb_row_sqnorm(const float* x, float* sq, long rows, int D, cudaStream_t stream);
int dktb_sqdist(const float* x1, const float* x2, float* out, int E, int M, int N, int D, cudaStream_t stream);
/* g = Gram matrix (linear / poli), d2 = squared distances from dktb_sqdist (rbf / matern); the unused one may be NULL */
int dktb_kernel_fwd(int kind, const float* g, const float* d2, const float* raw_param, float* kb, int E, int C, int M,
                    int N, cudaStream_t stream);
/* dkb [E][C][N][N] = dLoss/dKb_c -> dg [E][N][N] (summed over classes, incl. the diagonal terms through the norms),
 * dparam [C] = dLoss/d raw_param; scratch: E*C*N floats */
int dktb_kernel_bwd(int kind, const float* g, const float* d2, const float* raw_param, const float* dkb, float* dg,
                    float* dparam, float* scratch, int E, int C, int N, cudaStream_t stream);
/* predictive variance of likelihood(model(x*)): s*kss - ||L^-1 s kx||^2 + noise  (DKT_regression.py:90-93) */
int dktb_gp_predict_var(const float* kx, long kx_class_stride, const float* kss, long kss_class_stride,
                        const float* linv, const float* raw_outputscale, const float* raw_noise, float* var, int E,
                        int C, int M, int N, cudaStream_t stream);

/* ---- spectral-mixture kernel with ARD (DKT_regression.py:122, sines/train_DKT.py:132): k = sum_q w_q prod_d
 * exp(-2 pi^2 tau_d^2 v_qd^2) cos(2 pi tau_d mu_qd).  raw_w [Q], raw_mu / raw_v [Q][D] in the reference's feature order
 * (index (j % Cch)*P + j / Cch for NHWC feature j; P <= 1: identity).  ec [E][Q][M][N] keeps the per-mixture terms
 * for the backward pass.  bwd: dkb [E][N][N] -> dw [Q], dmu / dv [Q][D], dx [E][N][D]. */
int dktb_spectral_fwd(const float* x1, const float* x2, const float* raw_w, const float* raw_mu, const float* raw_v,
                      float* kb, float* ec, int E, int M, int N, int D, int Q, int Cch, int P, cudaStream_t stream);
int dktb_spectral_bwd(const float* x, const float* raw_w, const float* raw_mu, const float* raw_v, const float* dkb,
                      const float* ec, float* dw, float* dmu, float* dv, float* dx, int E, int N, int D, int Q, int Cch,
                      int P, cudaStream_t stream);

/* ---- optimiser (torch.optim.Adam, methods/DKT.py:114-115,164) ---------------------------------------- */
int dktb_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                   float eps, int step, float grad_scale, cudaStream_t stream);
/* The same update with the 1-based step count read from device memory (CUDA-graph replay of a whole meta-step: no
 * per-step scalar is baked into a launch); dktb_counter_add advances such a counter on the stream. */
int dktb_adam_step_dev(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                       float eps, const int* step_dev, float grad_scale, cudaStream_t stream);
int dktb_counter_add(int* counter, int delta, cudaStream_t stream);
int dktb_scale(float* x, long n, float a, cudaStream_t stream);

/* ---- episode feeder: the per-image transform chain of the reference's episode loader
 *      (data/datamgr.py:37-46 composed transforms, applied in data/dataset.py:66-70; ImageJitter =
 *      data/additional_transforms.py:24-34), bit-identical to PIL / torchvision for the same parameters ---------- */
/* store: HWC uint8 RGB images back to back; desc [n_images][3] long = (byte offset, height, width).
 * params [B][8] int = (image id, crop top, crop left, crop h, crop w, flip, jitter on/off, unused);
 * factors [B][3] float = ImageEnhance factors (Brightness, Contrast, Color), read when jitter is on.
 * Per output image: crop box -> Pillow bilinear resize to RH x RW (two-pass 22-bit fixed point) -> the S x S window
 * at (oy, ox) [aug: RH = RW = S, window 0,0 = RandomSizedCrop; plain: RH = RW = int(1.15 S) over the whole image,
 * window = CenterCrop(S)] -> ImageJitter -> horizontal flip -> ToTensor -> Normalize(mean, std) -> out [B,3,S,S] fp32.
 * kmax: upper bound of resample taps per output pixel (Pillow: 2 ceil(max(in/out, 1)) + 1); tmp_rows: rows of
 * horizontally resampled input kept in shared memory per band; dktb_episode_transform_smem() must be <= 227 KB.
 * err: device int, zero-initialised by the caller; 1 = more than kmax taps needed, 2 = one output row needs more than
 * tmp_rows input rows, 3 = bad image id / crop box outside its image. */
long dktb_episode_transform_smem(int S, int kmax, int tmp_rows);
int dktb_episode_transform(const unsigned char* store, const long* desc, long n_images, const int* params,
                           const float* factors, float* out, int B, int S, int RH, int RW, int oy, int ox, int kmax,
                           int tmp_rows, float mean_r, float mean_g, float mean_b, float std_r, float std_g,
                           float std_b, int* err, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DKTB200_H */
