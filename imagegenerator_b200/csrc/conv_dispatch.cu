// conv_dispatch.cu -- public convolution entry points.  bf16 storage = the tcgen05 kernels, and ONLY those: a shape they
// cannot take is an error (SG_ERR_UNSUPPORTED), not a silent switch to the CUDA-core kernels -- a second backend hides
// 10-50x performance cliffs (round 1: the generator's 228-channel first layer ran on FFMA unnoticed).  fp32 storage = the
// CUDA-core implicit GEMM of conv_ffma.cu, the validation mode.  The sg_conv_*_ffma entry points stay callable by name.
#include "common.cuh"
#include <stdlib.h>

namespace sg { extern int g_bstats_min_k; }

extern "C" {
int sg_conv_fprop_ffma(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int,
                       int, void*);
int sg_conv_dgrad_ffma(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int,
                       int, void*);
int sg_conv_wgrad_ffma(const void*, const void*, float*, int, int, int, int, int, int, int, int, int, int, int, void*);

int sg_conv_fprop_tc(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int, int,
                     void*);
int sg_conv_dgrad_tc(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int, int,
                     void*);
int sg_conv_tc_supported(int, int, int, int, int, int, int, int, int, int, int);
int sg_conv_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, int, int, int, int, int, int, void*);
int sg_conv_wgrad_tc_supported(int, int, int, int, int, int, int, int, int, int);

int sg_conv_tc_stats_supported(int, int, int, int, int, int, int, int, int, int, int, int);
int sg_conv_fprop_tc_stats(const void*, const void*, void*, double*, int, int, int, int, int, int, int, int, int, int, int, int,
                           void*);
int sg_conv_dgrad_tc_stats(const void*, const void*, void*, double*, int, int, int, int, int, int, int, int, int, int, int, int,
                           void*);
int sg_col_stats(const void*, double*, int64_t, int, int, int, void*);
int sg_conv_fprop_tc_f32out(const void*, const void*, float*, int, int, int, int, int, int, int, int, int, int, void*);
int sg_conv_fprop_tc_res(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int, int,
                         int, int, void*);
int sg_add_act(const void*, const void*, void*, int64_t, int, int, void*);
int sg_conv_fprop_tc_bstats(const void*, const void*, void*, const void*, const float*, const float*, const float*, double*, int, int,
                            int, int, int, int, int, int, int, int, int, int, void*);
int sg_conv_dgrad_tc_bstats(const void*, const void*, void*, const void*, const float*, const float*, const float*, double*, int, int,
                            int, int, int, int, int, int, int, int, int, int, void*);
int sg_bn_bwd_reduce_y(const void*, const void*, const float*, const float*, const float*, double*, int64_t, int, int, int, int, void*);
int sg_conv_narrow_supported(int, int, int, int, int, int, int, int, int, int, int);
int sg_conv_narrow_routed(int, int, int, int, int, int, int, int, int, int, int);
int sg_conv_narrow_fprop(const void*, const void*, const float*, void*, double*, int, int, int, int, int, int, int, void*);
int sg_conv_narrow_dgrad(const void*, const void*, const float*, void*, int, int, int, int, int, int, void*);
int sg_conv_thin_supported(int, int, int, int, int, int, int, int, int, int, int);
int sg_conv_thin_fprop(const void*, const void*, const float*, void*, int, int, int, int, int, void*);
int sg_conv_thin_dgrad(const void*, const void*, const float*, void*, int, int, int, int, int, void*);

static int unsupported(const char* what, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    sg::set_error("%s: the tensor-core kernels cannot take N=%d %dx%dx%d -> %dx%dx%d k%d s%d p%d in bf16 mode (reduction channels must "
                  "be a multiple of 8, grids powers of two); there is no CUDA-core fallback in bf16 mode -- pad the operand "
                  "(engine.Up0Gemm), go through a patch matrix (sg_patchify), or call the _ffma entry point explicitly",
                  what, N, H, W, Ci, Ho, Wo, Co, k, s, p);
    return SG_ERR_UNSUPPORTED;
}

int sg_conv_fprop(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho, int Wo,
                  int Co, int k, int s, int p, int act, int dtype, void* stream) {
    if (dtype == SG_BF16) {
        // the 3-channel image side: direct kernel (thin_conv.cu), the 128-row tcgen05 tiles have nothing to contract there
        if (sg_conv_thin_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p)) return sg_conv_thin_fprop(x, pf, bias, y, N, H, W, Co, act, stream);
        // 16 / 32 input channels on large maps: HBM-bound, direct kernel (narrow_conv.cu) where it is the faster one
        if (sg_conv_narrow_routed(0, N, H, W, Ci, Ho, Wo, Co, k, s, p))
            return sg_conv_narrow_fprop(x, pf, bias, y, nullptr, 1, N, H, W, Ci, Co, act, stream);
        if (!sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p)) return unsupported("conv_fprop", N, H, W, Ci, Ho, Wo, Co, k, s, p);
        return sg_conv_fprop_tc(x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, dtype, stream);
    }
    return sg_conv_fprop_ffma(x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, dtype, stream);
}
int sg_conv_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho, int Wo,
                  int Co, int k, int s, int p, int act, int dtype, void* stream) {
    if (dtype == SG_BF16) {
        if (sg_conv_thin_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p)) return sg_conv_thin_dgrad(dy, pd, bias, dx, N, Ho, Wo, Co, act, stream);
        if (sg_conv_narrow_routed(1, N, H, W, Ci, Ho, Wo, Co, k, s, p))
            return sg_conv_narrow_dgrad(dy, pd, bias, dx, N, Ho, Wo, Ci, Co, act, stream);
        if (!sg_conv_tc_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p)) return unsupported("conv_dgrad", N, H, W, Ci, Ho, Wo, Co, k, s, p);
        return sg_conv_dgrad_tc(dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, dtype, stream);
    }
    return sg_conv_dgrad_ffma(dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, dtype, stream);
}
int sg_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k,
                  int s, int p, int dtype, void* stream) {
    if (dtype == SG_BF16) {
        if (!sg_conv_wgrad_tc_supported(N, H, W, Ci, Ho, Wo, Co, k, s, p)) return unsupported("conv_wgrad", N, H, W, Ci, Ho, Wo, Co, k, s, p);
        return sg_conv_wgrad_tc(x, dy, dw, N, H, W, Ci, Ho, Wo, Co, k, s, p, dtype, stream);
    }
    return sg_conv_wgrad_ffma(x, dy, dw, N, H, W, Ci, Ho, Wo, Co, k, s, p, dtype, stream);
}

// y = conv(x, W) and stats[groups][Co][2] += (sum, sum^2) of y per image group -- the batch statistics of the
// BatchNorm2d that follows every bias-free conv of the reference (e.g. discrminator_1.py:29-37).  Fused into the
// tensor-core epilogue when the shape allows, otherwise conv + sg_col_stats.
int sg_conv_fprop_stats(const void* x, const void* pf, void* y, double* stats, int groups, int N, int H, int W, int Ci,
                        int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream) {
    if (dtype == SG_BF16 && groups >= 1 && N % groups == 0 && sg_conv_narrow_routed(0, N, H, W, Ci, Ho, Wo, Co, k, s, p))
        return sg_conv_narrow_fprop(x, pf, nullptr, y, stats, groups, N, H, W, Ci, Co, SG_ACT_NONE, stream);
    if (dtype == SG_BF16 && sg_conv_tc_stats_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups))
        return sg_conv_fprop_tc_stats(x, pf, y, stats, groups, N, H, W, Ci, Ho, Wo, Co, k, s, p, dtype, stream);
    int e = sg_conv_fprop(x, pf, nullptr, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, dtype, stream);
    if (e) return e;
    return sg_col_stats(y, stats, (int64_t)(N / groups) * Ho * Wo, Co, groups, dtype, stream);
}
int sg_conv_dgrad_stats(const void* dy, const void* pd, void* dx, double* stats, int groups, int N, int H, int W, int Ci,
                        int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream) {
    if (dtype == SG_BF16 && sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups))
        return sg_conv_dgrad_tc_stats(dy, pd, dx, stats, groups, N, H, W, Ci, Ho, Wo, Co, k, s, p, dtype, stream);
    int e = sg_conv_dgrad(dy, pd, nullptr, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, dtype, stream);
    if (e) return e;
    return sg_col_stats(dx, stats, (int64_t)(N / groups) * H * W, Ci, groups, dtype, stream);
}

// y = conv(x, W) is d loss / d a of the BatchNorm'ed layer below (a = act(bn(ybn))): also reduce that layer's backward
// statistics sums[groups][Co][2] = (sum dz, sum dz * xhat) -- in the tensor-core epilogue when the shape allows (bf16), otherwise
// conv + sg_bn_bwd_reduce_y.  act in {none, relu, lrelu}; Co % 8 == 0.
// 1 when sg_conv_{fprop,dgrad}_bstats reduces the statistics in the tcgen05 epilogue for this shape, 0 when it would run the
// conv followed by sg_bn_bwd_reduce_y -- a caller that follows up with sg_bn_bwd anyway (reduce + apply in one launch) then
// prefers the plain conv.
int sg_conv_bstats_in_epilogue(int dgrad, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups,
                               int dtype) {
    if (dtype != SG_BF16) return 0;
    if (dgrad) return Co * k * k / (s * s) >= sg::g_bstats_min_k && sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups);
    return Co <= 256 * 8 && Ci * k * k >= sg::g_bstats_min_k && sg_conv_tc_stats_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups);
}
int sg_conv_fprop_bstats(const void* x, const void* pf, void* y, const void* ybn, const float* mr, const float* gamma,
                         const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                         int k, int s, int p, int dtype, void* stream) {
    if (dtype == SG_BF16 && Co <= 256 * 8 && Ci * k * k >= sg::g_bstats_min_k &&
        sg_conv_tc_stats_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups))
        return sg_conv_fprop_tc_bstats(x, pf, y, ybn, mr, gamma, beta, sums, groups, act, N, H, W, Ci, Ho, Wo, Co, k, s, p, stream);
    int e = sg_conv_fprop(x, pf, nullptr, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, dtype, stream);
    if (e) return e;
    return sg_bn_bwd_reduce_y(y, ybn, mr, gamma, beta, sums, (int64_t)(N / groups) * Ho * Wo, Co, groups, act, dtype, stream);
}
int sg_conv_dgrad_bstats(const void* dy, const void* pd, void* dx, const void* ybn, const float* mr, const float* gamma,
                         const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                         int k, int s, int p, int dtype, void* stream) {
    if (dtype == SG_BF16 && Co * k * k / (s * s) >= sg::g_bstats_min_k &&
        sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups))
        return sg_conv_dgrad_tc_bstats(dy, pd, dx, ybn, mr, gamma, beta, sums, groups, act, N, H, W, Ci, Ho, Wo, Co, k, s, p, stream);
    int e = sg_conv_dgrad(dy, pd, nullptr, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, dtype, stream);
    if (e) return e;
    return sg_bn_bwd_reduce_y(dx, ybn, mr, gamma, beta, sums, (int64_t)(N / groups) * H * W, Ci, groups, act, dtype, stream);
}

// y (fp32) = conv(x, W): same operands as sg_conv_fprop, result kept in fp32 (no activation / bias).  In fp32 storage
// mode this is sg_conv_fprop itself; in bf16 mode the tensor-core kernel stores its accumulators un-rounded.
int sg_conv_fprop_f32out(const void* x, const void* pf, float* y, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k,
                         int s, int p, int dtype, void* stream) {
    if (dtype == SG_F32) return sg_conv_fprop(x, pf, nullptr, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, dtype, stream);
    if (sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p))
        return sg_conv_fprop_tc_f32out(x, pf, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, stream);
    sg::set_error("conv_fprop_f32out: shape not eligible for the tensor-core kernel (Ci %% 8 != 0?)");
    return SG_ERR_UNSUPPORTED;
}

// y = act(conv(x, W) + bias + residual).  Tensor-core shapes: fused in the epilogue; otherwise conv then sg_add_act.
int sg_conv_fprop_res(const void* x, const void* pf, const float* bias, const void* residual, void* y, int N, int H, int W,
                      int Ci, int Ho, int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream) {
    if (dtype == SG_BF16 && sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p))
        return sg_conv_fprop_tc_res(x, pf, bias, residual, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, stream);
    int e = sg_conv_fprop(x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, dtype, stream);
    if (e) return e;
    return sg_add_act(y, residual, y, (int64_t)N * Ho * Wo * Co, act, dtype, stream);
}
}
