/* sgb200.h -- C ABI of libsgb200.so: hand-written sm_100a CUDA kernels for the StackGAN
 * training step of anishbasnet969/ImageGenerator.
 *
 * The reference has no FFI of its own (it is pure Python on torch.nn); every
 * entry point below replaces the torch operator(s) the cited reference line executes.  A host
 * program binds these with ctypes / cffi / dlopen (INTEGRATION.md shows the ctypes stub).
 *
 * Conventions
 *   - every function returns 0 on success, a cudaError_t (> 0) or SG_ERR_* (< 0) otherwise;
 *     sg_last_error() returns a thread-local message for the last non-zero return.
 *   - all buffers are caller-owned DEVICE pointers; nothing is allocated or freed across the ABI (one exception:
 *     sg_check_device() allocates, once per device, 4 x 19 MB of scratch for the conv kernel's tail-wave K-split),
 *     no host synchronisation happens inside, every call is asynchronous on `stream`
 *     (a cudaStream_t passed as void*) and is CUDA-graph capturable.
 *   - "T" tensors are activations/packed weights in the storage type selected by `dtype`
 *     (SG_F32 or SG_BF16), laid out NHWC ([rows][C], channel fastest).  Parameters, gradients,
 *     scores and statistics are fp32; statistic sums are fp64.
 *   - a convolution *operator* is described in Conv2d orientation: weight [Co][Ci][k][k],
 *     x [N][H][W][Ci] -> y [N][Ho][Wo][Co], stride s, padding p.  A ConvTranspose2d layer whose
 *     weight is [Cin_t][Cout_t][k][k] is the SAME operator with Co=Cin_t, Ci=Cout_t run in its
 *     data-gradient direction (sg_conv_dgrad), and vice versa.
 *   - act codes: 0 none, 1 ReLU, 2 LeakyReLU(0.1), 3 Tanh.
 */
#ifndef SGB200_H
#define SGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_F32 0
#define SG_BF16 1

#define SG_ACT_NONE 0
#define SG_ACT_RELU 1
#define SG_ACT_LRELU 2
#define SG_ACT_TANH 3

#define SG_ERR_BAD_ARG (-1)
#define SG_ERR_UNSUPPORTED (-2)
#define SG_ERR_NO_DEVICE (-3)

/* ---- library / device probes ------------------------------------------------------------- */
int sg_version(void);                       /* ABI version, currently 1 */
const char* sg_last_error(void);
/* 0 if the current device is sm_100 (B200); SG_ERR_NO_DEVICE otherwise.  Kernels are compiled
 * for sm_100a only -- there is no other code path. */
int sg_check_device(void);
/* number of kernels launched through this library since load (all threads) */
int64_t sg_launch_count(void);
/* tuning switches (A/B measurements; defaults are the measured-best settings; SG_OPTS="name=value,..." sets them from the
 * environment through the Python binding):
 *  conv_tcp (conv_tc.cu):  "force_cg" / "force_bn" pin the CTA-group size / tile width (0 = cost model); "force_stages" caps the
 *    pipeline depth; "slab" = 1/0 shared activation slabs; "nsplit" = 1/0 column slices for a partly filled last round; "rotate" =
 *    rotate the slab order per cluster; "dyn_sched" = 1/0 dynamic tile schedule (default 0, measured slower); "epi_alt" = 1/0
 *    alternate-tile epilogue for tiles <= 64 columns (default 0); "bstats_min_k" = shallowest reduction that gets the fused
 *    BatchNorm-backward statistics (default 4000); "pdl" = 1/0 programmatic dependent launch for the launches that follow.
 *  wgrad (conv_tc.cu):  "wgrad_mc" = 1/0 TMA multicast between co tiles, "wgrad_mc_max" largest co-tile cluster (default 2),
 *    "wgrad_mc_odd" = 0 pairs only; "wgrad_smem_kb" shared memory the pipeline may take (default 200; ~150 leaves room for other
 *    kernels' CTAs on the SM -- co-residency experiment, measured slower).
 *  narrow layers (narrow_conv.cu / direct_tc.cu):  "narrow" = bit mask of the shapes routed to the direct kernels (1 forward
 *    16->32, 2 data gradient 16<-32, 4 forward 32->64, 8 data gradient 32<-64; default 7); "narrow_cfg" bit 4 = the mma.sync
 *    kernels instead of the tcgen05 ones, bits 1 / 2 = one CTA per SM for the mma.sync forward / data-gradient kernel;
 *    "dtc_wide" = 1/0 two / one M tiles per tile in direct_tc.cu (default 1); "dtc_diag" = timing experiments that produce WRONG
 *    results (1 contiguous boxes, 2 no stores, 4 no MMAs; default 0 -- never set it outside tools/).
 *  BatchNorm backward (bn_fast.cu):  "bn_fused" = 1/0 sg_bn_bwd as ONE launch where (da, y) fit the SMs' shared memory (default 0:
 *    faster alone -- 19 vs 26 us on the 12.6 MB critic layer -- but its 190 KB CTAs cannot share an SM with the side stream's
 *    wgrad CTAs and the captured steps measured 1-2 % slower; the engines switch it on for the gradient penalty's first-order
 *    pass, where the side streams are nearly idle); "gp_bn_fused" = 1/0 the same scheme for sg_gp_bn (default 1: the penalty's
 *    second-order pass runs alone on the GPU, Stage-I 4.82 -> 4.76 ms); "bn_act_bulk" = 1/0 sg_bn_finalize_act with bulk-copy
 *    staged ranges, one CTA per SM (default 0: measured neutral in the step, its input is L2-resident); "bn_fused_keep_pct" = least share of a range that must fit
 *    (default 50); "bn_fused_steal_ns" = patience at the rendezvous before resident CTAs take over ranges of CTAs that have not
 *    started (default 30000); "bn_fused_dbg" = globaltimer stamps of the first / last CTA in the work words.
 *  "dbg" = verbose launch decisions on stderr. */
int sg_set_option(const char* name, int value);
/* one-time device allocations of the library (the launch entry points never allocate): the counter pool of the dynamic conv
   schedule.  Once per process, before the first launch, outside stream capture. */
int sg_init_workspace(void);
/* further options: "dyn_sched" = 1/0 dynamic work distribution in the persistent conv kernel (item numbers drawn from a global
   counter instead of static per-cluster lists; env SG_DYN_SCHED); "bstats_min_k" = least reduction depth for which the
   BatchNorm-backward statistics ride in a conv epilogue (env SG_BSTATS_MIN_K) */


/* ---- memory helpers ------------------------------------------------------------------------ */
int sg_zero(void* ptr, int64_t bytes, void* stream);
/* n <= 32 small buffers (4-byte aligned, sizes multiples of 4) zeroed by ONE kernel node: the per-channel sums of every
   BatchNorm layer of a backward pass (torch zero-fills them inside native_batch_norm_backward).  A memset in front of every
   reduction is a graph node between two dependent kernels: 60 per Stage-I step, 0.2 ms of its 5.3.  ptrs / bytes: host arrays. */
int sg_zero_multi(void* const* ptrs, const int64_t* bytes, int n, void* stream);
int sg_fill_f32(float* ptr, float value, int64_t n, void* stream);
/* out = a*x + b on fp32 vectors (e.g. 1 - eps of utils.py:11) */
int sg_affine_f32(const float* x, float a, float b, float* out, int64_t n, void* stream);

/* ---- layout: the reference API is NCHW fp32 (e.g. generator_1.py:38-40 output) -------------- */
int sg_nchw_to_nhwc(const float* src, void* dst, int N, int C, int H, int W, int dtype, void* stream);
int sg_nhwc_to_nchw(const void* src, float* dst, int N, int C, int H, int W, int dtype, void* stream);
/* tanh output in [-1, 1] (NHWC, T) -> NCHW uint8 image round((x + 1) * 127.5): the de-normalised picture, a quarter of the
   fp32 read-back (sampling / serving path; the reference normalises with mean = std = 0.5, train.py:104-110) */
int sg_nhwc_to_nchw_u8(const void* src, unsigned char* dst, int N, int C, int H, int W, int dtype, void* stream);
/* w [Co][Ci][k*k] fp32 -> pf [Co][k*k][Ci] and pd [Ci][k*k][Co] in T (either may be NULL) */
int sg_pack_weight(const float* w, void* pf, void* pd, int Co, int Ci, int kk, int dtype, void* stream);
/* wt[(t, ci)][Kp] (T) = w[co][ci][t], columns co >= Co zero: forward operand of ConvTranspose2d(Co -> Ci, k, s1, p0) on a 1x1
   input run as the GEMM [B, Kp] x [Kp, kk*Ci] (generator_1.py:9-13; imagegenerator_b200.engine.Up0Gemm) */
int sg_pack_gemm_t(const float* w, void* wt, int Co, int Ci, int kk, int Kp, int dtype, void* stream);

/* P[n,oh,ow, ci*k*k + kh*k + kw] = x[n, oh*s-p+kh, ow*s-p+kw, ci] (0 outside): patch matrix of a thin
 * (3-channel) image in the PyTorch weight order, so that discrminator_1.py:10 / discriminator_2.py:9 run as
 * a 1x1 convolution on the tensor-core kernels */
int sg_patchify(const void* x, void* P, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p,
                int dtype, void* stream);

/* col2im for a thin (<= 4 channel) result: out[n,oh,ow,c] = act(bias[c] + sum of col[n,ih,iw, c*k*k + kh*k + kw] over
 * the taps with oh = ih*s-p+kh, ow = iw*s-p+kw); col is FP32 (sg_conv_fprop_f32out), out is T.
 * Behind the 1x1 GEMM col = x W^T on the tensor-core kernels this is
 * ConvTranspose2d(C -> 3)+Tanh (generator_1.py:20,38-40; generator_2.py:55) and d/d image of the critics' first
 * conv (discrminator_1.py:10, needed by utils.py:15 and by the generator step) */
int sg_unpatchify(const float* col, const float* bias, void* out, int N, int Hi, int Wi, int C, int Ho, int Wo,
                  int k, int s, int p, int act, int dtype, void* stream);
/* y (FP32) = conv(x, W), operands as in sg_conv_fprop: the accumulators are stored un-rounded (bf16 mode: tcgen05
 * kernel only) so that a following col2im sums exact partial products */
int sg_conv_fprop_f32out(const void* x, const void* pf, float* y, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                         int k, int s, int p, int dtype, void* stream);
int sg_conv_fprop_tc_f32out(const void* x, const void* pf, float* y, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                            int k, int s, int p, void* stream);

/* ---- the 3-channel image side as direct kernels (thin_conv.cu; bf16 only) --------------------------------------------
 * sg_conv_fprop / sg_conv_dgrad route here by themselves when Ci == 3, k4 s2 p1 and the output grid tiles (rows % 8,
 * columns % 32 == 0, Co % 8 == 0, Co <= 128): every activation byte is read once, no patch / col matrix in HBM.
 *   fprop: Conv2d(3 -> Co) + bias + act  -- critics' first layer (discrminator_1.py:17-18, discriminator_2.py:11-12),
 *          G2's first layer (generator_2.py:9-10), the input gradient of the generators' ConvT(C -> 3);
 *   dgrad: ConvTranspose2d(Co -> 3) + bias + act -- the generators' output layer + Tanh (generator_1.py:30-33,
 *          generator_2.py:55-57), d/d image of the critics (utils.py:15-21).
 * mode: 0 fprop, 1 dgrad; shapes in Conv2d orientation like every sg_conv_* call. */
int sg_conv_thin_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p);
int sg_conv_thin_fprop(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Co, int act,
                       void* stream);
int sg_conv_thin_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Ho, int Wo, int Co, int act,
                       void* stream);

/* ---- narrow-channel k4 s2 p1 convs as direct kernels (direct_tc.cu, narrow_conv.cu; bf16 only) -------------------------------
 * (Ci, Co) in {(16, 32), (32, 64)} on large maps -- the Stage-II critic's second and third layer (discriminator_2.py:13-18) and
 * their data gradients: HBM-bound shapes (150 MB for 13 GFLOP) whose 32-byte operand rows the implicit-GEMM kernel cannot feed.
 * direct_tc.cu: persistent CTAs stage a spatial tile with TMA (one strided box per filter column, zero padding = out-of-bounds
 * fill) and tcgen05.mma reads every tap's operand as a window of it; forward (both shapes, optional BatchNorm statistics) and the
 * 16 <- 32 data gradient; output grid rows % 16 == 0, columns % 8 == 0.  narrow_conv.cu: the same operators with warp-level
 * mma.sync (rows % 8, columns % 32), kept as the second implementation behind option "narrow_cfg" and for 32 <- 64.
 * sg_conv_fprop / sg_conv_fprop_stats / sg_conv_dgrad route a supported shape here when option "narrow" selects it (bit mask,
 * see sg_set_option; sg_conv_narrow_routed reports the decision). */
int sg_conv_narrow_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p);
int sg_conv_narrow_routed(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p);
int sg_conv_narrow_fprop(const void* x, const void* pf, const float* bias, void* y, double* stats, int groups, int N, int H, int W,
                         int Ci, int Co, int act, void* stream);
int sg_conv_narrow_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Ho, int Wo, int Ci, int Co, int act,
                         void* stream);

/* out[n,hw,:Cx] = x, out[n,hw,Cx:] = c[n]  (generator_2.py:61-63 reshape/repeat/cat);  backward:
 * dx = dout[..., :Cx], dc[n] (fp32) = sum_hw dout[n,hw,Cx:] */
int sg_concat_rep(const void* x, const float* c, void* out, int N, int HW, int Cx, int Cc, int dtype, void* stream);
int sg_split_rep_bwd(const void* dout, void* dx, float* dc, int N, int HW, int Cx, int Cc, int dtype, void* stream);

/* ---- convolution operator (replaces nn.Conv2d / nn.ConvTranspose2d and their autograd:
 *      generator_1.py:20,26  generator_2.py:30,46,55,71,87  discrminator_1.py:10,29
 *      discriminator_2.py:9,44) ------------------------------------------------------------- */
/* y = act(conv(x, W) + bias);  pf is the [Co][k][k][Ci] pack; bias may be NULL */
int sg_conv_fprop(const void* x, const void* pf, const float* bias, void* y,
                  int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p,
                  int act, int dtype, void* stream);
/* dx = act(conv_transpose(dy, W) + bias);  pd is the [Ci][k][k][Co] pack; bias ([Ci]) may be NULL */
int sg_conv_dgrad(const void* dy, const void* pd, const float* bias, void* dx,
                  int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p,
                  int act, int dtype, void* stream);
/* dw[Co][Ci][k][k] (fp32) += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh*s-p+kh,ow*s-p+kw,ci] */
int sg_conv_wgrad(const void* x, const void* dy, float* dw,
                  int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p,
                  int dtype, void* stream);
/* weight gradient accumulated in a channels-last buffer gw[Co][k][k][Ci] (fp32, += semantics): every 16 accumulator
 * columns of the tcgen05 kernel are then 64 contiguous bytes and leave as vector reductions whatever k is (the
 * PyTorch layout puts the k*k taps innermost, which for 3x3 layers -- generator_2.py:30 -- means 4-byte atomics).
 * sg_fold_grad_cl adds gw into dw[Co][Ci][k][k] and clears it; call it once before the optimizer step. */
int sg_conv_wgrad_cl_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int dtype);
int sg_conv_wgrad_cl(const void* x, const void* dy, float* gw, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                     int k, int s, int p, int dtype, void* stream);
int sg_fold_grad_cl(float* gw, float* dw, int Co, int Ci, int kk, void* stream);
/* the same three entry points pinned to one implementation: *_ffma = CUDA-core implicit GEMM (fp32 or bf16
 * storage, fp32 accumulate), *_tc = tcgen05/TMEM/TMA (bf16 only; sg_conv_tc_supported / sg_conv_wgrad_tc_supported
 * say whether a shape is eligible).  sg_conv_* above dispatch between them. */
int sg_conv_fprop_ffma(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho,
                       int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream);
int sg_conv_dgrad_ffma(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho,
                       int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream);
int sg_conv_wgrad_ffma(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                       int k, int s, int p, int dtype, void* stream);
int sg_conv_fprop_tc(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho,
                     int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream);
int sg_conv_dgrad_tc(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho,
                     int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream);
int sg_conv_wgrad_tc(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                     int k, int s, int p, int dtype, void* stream);
int sg_conv_tc_supported(int mode /*0 fprop, 1 dgrad*/, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p);
int sg_conv_wgrad_tc_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p);

/* conv + statistics of the BatchNorm2d that follows it (every BN layer of the reference sits behind a bias-free
 * conv: generator_1.py:26-34, generator_2.py:30-38, discrminator_1.py:29-37, discriminator_2.py:44-52):
 * y = conv(x, W) / dx = conv_transpose(dy, W), and stats[groups][C][2] (fp64) += (sum, sum of squares) of the
 * stored result per image group (N/groups consecutive images each).  On the tcgen05 path the sums are reduced in
 * the epilogue (warp shuffles -> shared memory -> one fp64 atomic per channel and tile); other shapes run the conv
 * followed by sg_col_stats. */
int sg_conv_fprop_stats(const void* x, const void* pf, void* y, double* stats, int groups,
                        int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream);
int sg_conv_dgrad_stats(const void* dy, const void* pd, void* dx, double* stats, int groups,
                        int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream);

/* conv whose RESULT is d loss / d a of the BatchNorm'ed layer below (a = act(bn(ybn)), generator_1.py:26-34, discrminator_1.py:29-37,
 * generator_2.py:30-38): besides y / dx (T) the launch reduces that layer's BatchNorm-BACKWARD statistics
 * sums[groups][C][2] = (sum dz, sum dz * xhat), dz = result * act'(gamma * xhat + beta), xhat = (ybn - mean) * rstd -- in the tcgen05
 * epilogue when the shape allows (bf16), otherwise conv + sg_bn_bwd_reduce_y.  Replaces one pass over (da, y) per layer of every
 * backward chain.  ybn has the result's layout; mr [groups][C][2]; act in {none, relu, lrelu}; C % 8 == 0. */
int sg_conv_fprop_bstats(const void* x, const void* pf, void* y, const void* ybn, const float* mr, const float* gamma,
                         const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                         int k, int s, int p, int dtype, void* stream);
int sg_conv_dgrad_bstats(const void* dy, const void* pd, void* dx, const void* ybn, const float* mr, const float* gamma,
                         const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                         int k, int s, int p, int dtype, void* stream);
/* sg_conv_dgrad_bstats with the MASKED gradient stored: dx = dz = conv result * act'(gamma * xhat + beta), sums[.][.][0] = the
 * column sums of dz -- data-gradient conv + activation backward in one kernel (tcgen05 epilogue; bf16, Ci % 32 == 0, Ci <= 256).
 * With the identity table (mean 0, rstd 1, gamma 1, beta 0) and ybn = the stored activation of a conv + bias + activation layer
 * (discrminator_1.py:17-18) it replaces the activation-backward pass over the critics' largest activation, and sums[.][.][0] is
 * that layer's bias gradient.  sums_zeroed: the caller zeroed sums (sg_zero_multi). */
int sg_conv_dgrad_tc_bstats_masked(const void* dy, const void* pd, void* dx, const void* ybn, const float* mr, const float* gamma,
                                   const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho,
                                   int Wo, int Co, int k, int s, int p, int sums_zeroed, void* stream);
int sg_conv_dgrad_tc_bstats_masked_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups);
/* 1: the two calls above reduce the statistics in the tcgen05 epilogue for this shape; 0: they run the conv followed by
 * sg_bn_bwd_reduce_y (a caller that follows up with sg_bn_bwd -- reduce + apply in one launch -- then prefers the plain conv) */
int sg_conv_bstats_in_epilogue(int dgrad, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups,
                               int dtype);
int sg_conv_fprop_tc_bstats(const void* x, const void* pf, void* y, const void* ybn, const float* mr, const float* gamma,
                            const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo,
                            int Co, int k, int s, int p, void* stream);
int sg_conv_dgrad_tc_bstats(const void* dy, const void* pd, void* dx, const void* ybn, const float* mr, const float* gamma,
                            const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo,
                            int Co, int k, int s, int p, void* stream);
int sg_conv_fprop_tc_stats(const void* x, const void* pf, void* y, double* stats, int groups, int N, int H, int W, int Ci,
                           int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream);
int sg_conv_dgrad_tc_stats(const void* dy, const void* pd, void* dx, double* stats, int groups, int N, int H, int W, int Ci,
                           int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream);
int sg_conv_tc_stats_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups);

/* ---- forward-only path (sampling, stage_2_train_fn.py:181-195): eval-mode BatchNorm folded into the conv ------ */
/* scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale */
int sg_bn_fold(const float* running_mean, const float* running_var, const float* gamma, const float* beta, float eps,
               float* scale, float* shift, int C, void* stream);
/* sg_pack_weight with every weight multiplied by scale[co] (axis 0: forward convs) or scale[ci] (axis 1: the
 * output channels of a ConvTranspose2d run as sg_conv_dgrad) */
int sg_pack_weight_scaled(const float* w, const float* scale, int axis, void* pf, void* pd, int Co, int Ci, int kk,
                          int dtype, void* stream);
/* y = act(conv(x, W) + bias + residual): closing layer of ResidualBlock.forward (generator_2.py:23-26) */
int sg_conv_fprop_res(const void* x, const void* pf, const float* bias, const void* residual, void* y,
                      int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream);
int sg_conv_fprop_tc_res(const void* x, const void* pf, const float* bias, const void* residual, void* y, int N, int H, int W,
                         int Ci, int Ho, int Wo, int Co, int k, int s, int p, int act, void* stream);
/* out = act(a + b) */
int sg_add_act(const void* a, const void* b, void* out, int64_t n, int act, int dtype, void* stream);

/* out[C] (fp32) += column sums of x[rows][C]  (bias gradients) */
int sg_colsum(const void* x, float* out, int64_t rows, int C, int dtype, void* stream);

/* ---- BatchNorm2d, train and eval mode (replaces nn.BatchNorm2d: generator_1.py:34,
 *      generator_2.py:38,79,95  discrminator_1.py:37  discriminator_2.py:52) ------------------ */
/* stats[G][C][2] (fp64) += (sum, sum of squares) of y, rows split evenly into G groups */
int sg_col_stats(const void* y, double* stats, int64_t rows_per_group, int C, int groups, int dtype, void* stream);
/* mr[G][C][2] = (mean, 1/sqrt(var_biased+eps)); running stats EMA-updated once per group in group
 * order, group 0 `dup_first` times; *nbt += dup_first + G - 1 */
int sg_bn_finalize(const double* stats, int64_t count, float* mr, float* running_mean, float* running_var,
                   int64_t* nbt, int dup_first, int update_running, float momentum, float eps,
                   int groups, int C, void* stream);
int sg_bn_eval_mr(const float* running_mean, const float* running_var, float* mr, float eps, int C, void* stream);
/* out = act(gamma*(y-mean)*rstd + beta [+ residual]) */
int sg_bn_act(const void* y, const float* mr, const float* gamma, const float* beta, const void* residual,
              void* out, int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream);
/* sg_bn_finalize followed by sg_bn_act in ONE launch (the training-mode forward of nn.BatchNorm2d + activation):
 * every CTA derives (mean, rstd) of the groups it touches from the raw sums, CTA 0 writes mr and the running statistics */
int sg_bn_finalize_act(const double* stats, int64_t count, float* mr, float* running_mean, float* running_var,
                       int64_t* nbt, int dup_first, int update_running, float momentum, float eps, const void* y,
                       const float* gamma, const float* beta, const void* residual, void* out,
                       int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream);
/* sums[G][C][2] (fp64) = (sum dz, sum dz*xhat),  dz = da * act'(a_out) */
int sg_bn_bwd_reduce(const void* da, const void* a_out, const void* y, const float* mr, double* sums,
                     int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream);
/* dy = gamma*rstd/N*(N dz - S1 - xhat S2) [+ inject on the rows of group inject_group] */
int sg_bn_bwd_apply(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma,
                    const double* sums, const void* inject, int inject_group, void* dy,
                    int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream);
/* sg_bn_bwd_reduce / sg_bn_bwd_apply without the activation tensor: act'(a) comes from the sign of gamma*xhat+beta,
 * recomputed from y -- one tensor less to stream (C % 8 == 0; ReLU / LeakyReLU / none directly behind the BN) */
int sg_bn_bwd_reduce_y(const void* da, const void* y, const float* mr, const float* gamma, const float* beta, double* sums,
                       int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream);
int sg_bn_bwd_apply_y(const void* da, const void* y, const float* mr, const float* gamma, const float* beta,
                      const double* sums, const void* inject, int inject_group, void* dy,
                      int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream);
/* BatchNorm backward in one call (sg_bn_bwd_reduce[_y] + sg_bn_bwd_apply[_y]; torch's native_batch_norm_backward behind
 * generator_1.py:26-34 / discrminator_1.py:29-37 / generator_2.py:30-38).  Tensors whose (da, y[, a_out]) fit the SMs' shared
 * memory run as ONE launch: per-channel sums, a grid-wide rendezvous that counts finished ranges (no co-residency assumed),
 * apply out of shared memory; larger tensors run the two kernels.  a_out == NULL: act' from the sign of gamma*xhat+beta (needs beta).
 * sums_zeroed != 0: the caller zeroed sums (sg_zero_multi at the start of its pass) -- no memset node in front of the reduction.
 * work: 1 KB of zero-initialised words owned by the call site (re-armed by the kernel), NULL = never fuse. */
int sg_bn_bwd(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma, const float* beta,
              double* sums, const void* inject, int inject_group, void* dy, int64_t rows_per_group, int C, int groups,
              int act, int dtype, int sums_zeroed, void* work, void* stream);
/* dgamma += sum_g S2, dbeta += sum_g S1 */
int sg_bn_param_grad(const double* sums, float* dgamma, float* dbeta, int groups, int C, void* stream);
/* the same for n_layers BatchNorm layers in one launch (host arrays of device pointers / sizes, <= 24 layers): a network's
   backward pass queues every layer's (sums, gamma.grad, beta.grad) and ends with ONE of these */
int sg_bn_param_grad_multi(const double* const* sums, float* const* dgamma, float* const* dbeta, const int* groups, const int* C,
                           int n_layers, void* stream);
/* out = da * act'(a_out)   (LeakyReLU / ReLU / Tanh backward without BN) */
int sg_act_bwd(const void* da, const void* a_out, void* out, int64_t n, int act, int dtype, void* stream);
/* the same on a [rows][C] tensor AND colsum[c] += sum_rows out[., c]: the activation backward of a conv + bias + activation layer
   together with that layer's bias gradient (discrminator_1.py:17-18) in one pass over the tensor */
int sg_act_bwd_colsum(const void* da, const void* a_out, void* out, float* colsum, int64_t rows, int C, int act, int dtype,
                      void* stream);

/* ---- WGAN-GP second order through a train-mode BN (replaces autograd's double backward of
 *      utils.py:15-21 under stage_1_train_fn.py:147) ------------------------------------------ */
/* tsums[C][3] (fp64) = (sum v, sum v*xhat, sum v*dz) */
int sg_gp_bn_reduce(const void* v, const void* da, const void* a_out, const void* y, const float* mr,
                    double* tsums, int64_t rows, int C, int act, int dtype, void* stream);
/* the same, ADDING to tsums which the caller zeroed (sg_zero_multi at the start of the pass): no memset node in the chain */
int sg_gp_bn_reduce_acc(const void* v, const void* da, const void* a_out, const void* y, const float* mr,
                        double* tsums, int64_t rows, int C, int act, int dtype, void* stream);
/* sg_gp_bn_reduce + sg_gp_bn_apply behind ONE call: one launch where (v, da, a_out, y) fit the SMs' shared memory (the scheme of
 * sg_bn_bwd's one-launch kernel; this pass runs while the side streams are idle, where it pays), else the two kernels.
 * tsums_zeroed: the caller zeroed tsums; work: 1 KB of zeroed words owned by the call site (NULL = two kernels); option "gp_bn_fused". */
int sg_gp_bn(const void* v, const void* da, const void* a_out, const void* y, const float* mr, const float* gamma,
             const double* sums, double* tsums, void* w_out, void* gy_out, float* dgamma, int64_t rows, int C, int act,
             int dtype, int tsums_zeroed, void* work, void* stream);
int sg_gp_bn_apply(const void* v, const void* da, const void* a_out, const void* y, const float* mr,
                   const float* gamma, const double* sums, const double* tsums, void* w_out, void* gy_out,
                   float* dgamma, int64_t rows, int C, int act, int dtype, void* stream);

/* ---- small dense layers, fp32 (replaces nn.Linear: con_augment.py:9-11, discrminator_1.py:16) - */
/* out[N][M] = x[N][K] w[M][K]^T + b  (optional ReLU) */
int sg_linear_fwd(const float* x, const float* w, const float* b, float* out, int N, int K, int M, int relu, void* stream);
/* dout masked by relu_out>0 if given; dw += dout^T x; db += colsum(dout); dx (+)= dout w.  dw/db/dx may be NULL */
int sg_linear_bwd(const float* x, const float* w, const float* dout, const float* relu_out,
                  float* dw, float* db, float* dx, int dx_acc, int N, int K, int M, void* stream);

/* ---- critic head: compress + replicate + concat + 1x1 conv + flatten + linear
 *      (discrminator_1.py:43-52, discriminator_2.py:29-38) collapsed to score = <A,a4>+<Bv,ce>+c0 - */
int sg_head_prepare(const float* wcr, const float* bcr, const float* wcs, const float* bcs,
                    float* A, float* Bv, float* c0, int K, int Cx, int Nd, void* stream);
int sg_head_fwd(const void* a4, const float* ce, const float* A, const float* Bv, const float* c0,
                float* score, int N, int M, int Nd, int dtype, void* stream);
/* up to four sg_head_fwd calls of one critic forward in one launch (discrminator_1.py:41-52 is evaluated on the real,
 * mismatched, fake and interpolated rows of the same activation buffer): job j scores rows a4[a_row0[j]+n] with text rows
 * ce[ce_row0[j]+n] into score[score_off[j]+n], n < N.  The three index arrays are HOST arrays of n_jobs ints. */
int sg_head_fwd_multi(const void* a4, const float* ce, const float* A, const float* Bv, const float* c0, float* score,
                      int n_jobs, const int* a_row0, const int* ce_row0, const int* score_off, int N, int M, int Nd,
                      int dtype, void* stream);
/* out[n][m] = coef[n] * vec[m]   (out in T when out_dtype says so, else fp32) */
int sg_outer(const float* coef, const float* vec, void* out, int N, int M, int out_dtype, void* stream);
/* out[m] (fp32) += sum_n coef[n] * x[n][m] */
int sg_wsum_rows(const float* coef, const void* x, float* out, int N, int M, int dtype, void* stream);
int sg_head_param_grads(const float* dA, const float* dBv, const float* dc0, const float* wcr, const float* bcr,
                        const float* wcs, float* dwcr, float* dbcr, float* dwcs, float* dbcs,
                        int K, int Cx, int Nd, void* stream);

/* ---- conditioning augmentation (con_augment.py:18-22; stage_1_train_fn.py:120-122,156-159) ---- */
/* c_hat = mu + sigma*eps; cg[N][ld] (T) = [c_hat, z, 0...] when cg != NULL (z may be NULL; ld >= C+nz, the columns past
   C+nz are zero padding so that the generator's first layer sees a K that is a multiple of 64) */
int sg_ca_reparam(const float* mu, const float* sigma, const float* eps, const float* z, float* c_hat,
                  void* cg, int N, int C, int nz, int ld, int dtype, void* stream);
/* dmu = dc + kl*(-2mu); dsigma = dc*eps + kl*(2/sigma-2sigma); dc = first C of each dcg row (ld = row length) */
int sg_ca_bwd_seed(const void* dcg, const float* eps, const float* mu, const float* sigma, float kl_scale,
                   float* dmu, float* dsigma, int N, int C, int ld, int dtype, void* stream);

/* The whole module in one launch each way (dense.cu): tem [N][Tm] -> h = relu(Wh tem + bh) [N][Hd] -> mu, sigma [N][C]
   -> c_hat = mu + sigma*eps -> cg row [c_hat, z, 0...] (con_augment.py:13-22, stage_1_train_fn.py:120-122).  eps NULL =
   encode only; z / cg NULL = no generator input row.  h, mu, sigma, c_hat are kept for the backward. */
int sg_ca_forward(const float* tem, const float* Wh, const float* bh, const float* Wmu, const float* bmu, const float* Wsg,
                  const float* bsg, const float* eps, const float* z, float* h, float* mu, float* sigma, float* c_hat, void* cg,
                  int N, int Tm, int Hd, int C, int nz, int ld, int dtype, void* stream);
/* backward of the above in two launches (per-sample data gradients, then the six parameter gradients together):
   dmu/dsigma as sg_ca_bwd_seed, dh = relu'(h)*(Wmu^T dmu + Wsg^T dsigma) [stored masked], dtem (+)= Wh^T dh (dtem NULL: text
   side frozen, stage_2_train_fn.py:52-57), gW*, gb* += the weight / bias gradients. */
int sg_ca_backward(const void* dcg, const float* eps, const float* mu, const float* sigma, float kl_scale, const float* h,
                   const float* tem, const float* Wmu, const float* Wsg, const float* Wh, float* dmu, float* dsigma, float* dh,
                   float* gWmu, float* gbmu, float* gWsg, float* gbsg, float* gWh, float* gbh, float* dtem, int dtem_acc, int N,
                   int Tm, int Hd, int C, int ld, int dtype, void* stream);

/* ---- losses (utils.py:8-26; stage_1_train_fn.py:134-144,154-159) ------------------------------ */
int sg_interp(const void* real, const void* fake, const float* eps, void* out, int N, int64_t per_sample, int dtype, void* stream);
int sg_sample_sqnorm(const void* g, float* out, int N, int64_t per_sample, int dtype, void* stream);
int sg_sample_sqnorm_acc(const void* g, float* out, int N, int64_t per_sample, int dtype, void* stream);   /* out zeroed by the caller */
int sg_gp_seed(const void* g, const float* sq, float coef, void* v, int N, int64_t per_sample, int dtype, void* stream);
int sg_critic_loss(const float* s_real, const float* s_mis, const float* s_fake, const float* sq, float lam,
                   float* out2, int N, void* stream);
int sg_gen_loss(const float* s_fake, const float* mu, const float* sigma, float* out2, int N, int C, void* stream);
/* out[n,...] (+)= scale[n] * x[n,...] */
int sg_scale_rows_add(const void* x, const float* scale, void* out, int accumulate, int N, int64_t per_sample, int dtype, void* stream);

/* ---- Adam (train.py:92-102; torch.optim.Adam semantics, bias-corrected, eps outside sqrt/bc2) -- */
/* hyper (device, 8 floats) = [lr, beta1, beta2, eps, step, step_lo, step_hi, -]; the kernel increments step first
   (step as a float for the bias correction; the exact count is step_hi * 2^23 + step_lo) */
int sg_adam_step(float* p, const float* g, float* m, float* v, float* hyper, int64_t n, void* stream);

/* ---- data-parallel optimizer step over NVLink peer memory (xm.optimizer_step: stage_1_train_fn.py:149,166-172) -------- */
/* One kernel = reduce-scatter of the replicas' gradients (P2P loads of the owned shard) + Adam on the shard (optimizer
   state sharded over ranks) + all-gather of the new parameters (P2P stores).  grad_ptrs / param_ptrs / flag_ptrs: host
   arrays of `world` device pointers to the peers' symmetric buffers; flag block = sg_dp_flag_ints() int32, zeroed once;
   sync = sg_dp_sync_ints() local int32, zeroed once (its last word is raised when a peer did not answer within ~2 s).
   Graph-capturable; replicas end bit-identical.  write_avg != 0 also stores the averaged gradient into every replica. */
int sg_dp_max_world(void);
int sg_dp_flag_ints(void);
int sg_dp_sync_ints(void);
int sg_dp_adam_step(const void* const* grad_ptrs, const void* const* param_ptrs, const void* const* flag_ptrs, float* m, float* v,
                    float* hyper, int* sync, int64_t n, int rank, int world, int slot, int write_avg, void* stream);

/* ---- profiling hook -------------------------------------------------------------------------- */
/* While buf (device, >= 296*16 uint64) is set, every persistent conv launch writes 16 %globaltimer stamps per CTA into
   it; NULL switches the hook off (tools/exp_conv_trace.py). */
int sg_debug_conv_trace(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H */
