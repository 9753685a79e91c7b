// conv_dispatch.cu -- public convolution entry points: route each call to the tcgen05 kernels
// (bf16, tensor-core-shaped layers) or to the CUDA-core implicit GEMM (fp32 mode, thin layers).
#include "common.cuh"

extern "C" {
int sg_conv_fprop_ffma(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int,
                       int, void*);
int sg_conv_dgrad_ffma(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int,
                       int, void*);
int sg_conv_wgrad_ffma(const void*, const void*, float*, int, int, int, int, int, int, int, int, int, int, int, void*);

int sg_conv_fprop(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho, int Wo,
                  int Co, int k, int s, int p, int act, int dtype, void* stream) {
    return sg_conv_fprop_ffma(x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, dtype, stream);
}
int sg_conv_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho, int Wo,
                  int Co, int k, int s, int p, int act, int dtype, void* stream) {
    return sg_conv_dgrad_ffma(dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, dtype, stream);
}
int sg_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k,
                  int s, int p, int dtype, void* stream) {
    return sg_conv_wgrad_ffma(x, dy, dw, N, H, W, Ci, Ho, Wo, Co, k, s, p, dtype, stream);
}
}
