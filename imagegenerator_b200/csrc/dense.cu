// dense.cu -- the small fp32 dense pieces of the StackGAN step: conditioning-augmentation /
// text-compression linears and the critic's collapsed affine head.  These are latency-bound
// (<= 1.5 MB of traffic per call); they are written for full coalescing
// and one launch per logical operation rather than for FLOP/s.
#include "common.cuh"

namespace sg {

// out[n][m] = sum_k x[n][k] w[m][k] + b[m]; one warp per output element
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ out, int N,
                                                         int K, int M, int relu) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N * M) return;
    int n = warp / M, m = warp - n * M;
    const float* xr = x + (int64_t)n * K;
    const float* wr = w + (int64_t)m * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += xr[k] * wr[k];
    acc = warp_sum(acc);
    if (lane == 0) {
        acc += b ? b[m] : 0.f;
        out[warp] = relu ? fmaxf(acc, 0.f) : acc;
    }
}

// dw[m][k] += sum_n dout[n][m] x[n][k]  (thread per (m,k), k fastest)
__global__ void __launch_bounds__(256) linear_dw_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                        const float* __restrict__ relu_out, float* __restrict__ dw,
                                                        int N, int K, int M) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * K) return;
    int m = (int)(i / K), k = (int)(i - (int64_t)m * K);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
        float d = dout[(int64_t)n * M + m];
        if (relu_out && !(relu_out[(int64_t)n * M + m] > 0.f)) d = 0.f;
        acc += d * x[(int64_t)n * K + k];
    }
    dw[i] += acc;
}
__global__ void linear_db_kernel(const float* __restrict__ dout, const float* __restrict__ relu_out,
                                 float* __restrict__ db, int N, int M) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
        float d = dout[(int64_t)n * M + m];
        if (relu_out && !(relu_out[(int64_t)n * M + m] > 0.f)) d = 0.f;
        acc += d;
    }
    db[m] += acc;
}
// dx[n][k] (+)= sum_m dout[n][m] w[m][k]
__global__ void __launch_bounds__(256) linear_dx_kernel(const float* __restrict__ w, const float* __restrict__ dout,
                                                        const float* __restrict__ relu_out, float* __restrict__ dx,
                                                        int acc_flag, int N, int K, int M) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    int n = (int)(i / K), k = (int)(i - (int64_t)n * K);
    float acc = 0.f;
    for (int m = 0; m < M; ++m) {
        float d = dout[(int64_t)n * M + m];
        if (relu_out && !(relu_out[(int64_t)n * M + m] > 0.f)) d = 0.f;
        acc += d * w[(int64_t)m * K + k];
    }
    dx[i] = acc_flag ? dx[i] + acc : acc;
}

// ---- head: A[hw][c] = sum_k wcs[k][hw] wcr[k][c]; Bv[j] = sum_k sw[k] wcr[k][Cx+j]; c0
__global__ void head_prepare_kernel(const float* __restrict__ wcr, const float* __restrict__ bcr,
                                    const float* __restrict__ wcs, const float* __restrict__ bcs, float* A, float* Bv,
                                    float* c0, int K, int Cx, int Nd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int W = Cx + Nd;
    if (i < 16 * Cx) {
        int hw = i / Cx, c = i - hw * Cx;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += wcs[k * 16 + hw] * wcr[(int64_t)k * W + c];
        A[i] = acc;
    } else if (i < 16 * Cx + Nd) {
        int j = i - 16 * Cx;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            float sw = 0.f;
            for (int h = 0; h < 16; ++h) sw += wcs[k * 16 + h];
            acc += sw * wcr[(int64_t)k * W + Cx + j];
        }
        Bv[j] = acc;
    } else if (i == 16 * Cx + Nd) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            float sw = 0.f;
            for (int h = 0; h < 16; ++h) sw += wcs[k * 16 + h];
            acc += sw * bcr[k];
        }
        c0[0] = acc + bcs[0];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ a4, const float* __restrict__ ce,
                                                       const float* __restrict__ A, const float* __restrict__ Bv,
                                                       const float* __restrict__ c0, float* __restrict__ score, int M,
                                                       int Nd) {
    int n = blockIdx.x;
    const T* a = a4 + (int64_t)n * M;
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < M; i += 256 * 4) {
        F4 v = ld4(a + i);
        float4 w = *reinterpret_cast<const float4*>(A + i);
        acc += v.v[0] * w.x + v.v[1] * w.y + v.v[2] * w.z + v.v[3] * w.w;
    }
    for (int j = threadIdx.x; j < Nd; j += 256) acc += ce[(int64_t)n * Nd + j] * Bv[j];
    __shared__ float sh[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = c0[0];
        for (int i = 0; i < 8; ++i) s += sh[i];
        score[n] = s;
    }
}

// out[m] += sum_n coef[n] x[n][m]; grid.y splits n
template <typename T>
__global__ void __launch_bounds__(256) wsum_rows_kernel(const float* __restrict__ coef, const T* __restrict__ x,
                                                        float* __restrict__ out, int N, int M, int n_per_block) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int n0 = blockIdx.y * n_per_block, n1 = min(N, n0 + n_per_block);
    float acc = 0.f;
    for (int n = n0; n < n1; ++n) {
        float c = coef[n];
        if (c != 0.f) acc += c * ldf(x + (int64_t)n * M + m);
    }
    atomicAdd(out + m, acc);
}

__global__ void __launch_bounds__(256) head_param_grads_kernel(const float* __restrict__ dA, const float* __restrict__ dBv,
                                                               const float* __restrict__ dc0, const float* __restrict__ wcr,
                                                               const float* __restrict__ bcr, const float* __restrict__ wcs,
                                                               float* dwcr, float* dbcr, float* dwcs, float* dbcs, int K,
                                                               int Cx, int Nd) {
    int W = Cx + Nd;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n1 = (int64_t)K * W, n2 = n1 + (int64_t)K * 16;
    float d0 = dc0[0];
    if (i < n1) {                      // dwcr[k][c]
        int k = (int)(i / W), c = (int)(i - (int64_t)k * W);
        float acc = 0.f;
        if (c < Cx) {
            for (int h = 0; h < 16; ++h) acc += wcs[k * 16 + h] * dA[h * Cx + c];
        } else {
            float sw = 0.f;
            for (int h = 0; h < 16; ++h) sw += wcs[k * 16 + h];
            acc = sw * dBv[c - Cx];
        }
        dwcr[i] += acc;
    } else if (i < n2) {               // dwcs[k][hw]
        int64_t r = i - n1;
        int k = (int)(r / 16), h = (int)(r - (int64_t)k * 16);
        float acc = bcr[k] * d0;
        for (int c = 0; c < Cx; ++c) acc += wcr[(int64_t)k * W + c] * dA[h * Cx + c];
        for (int j = 0; j < Nd; ++j) acc += wcr[(int64_t)k * W + Cx + j] * dBv[j];
        dwcs[r] += acc;
    } else if (i < n2 + K) {           // dbcr[k]
        int k = (int)(i - n2);
        float sw = 0.f;
        for (int h = 0; h < 16; ++h) sw += wcs[k * 16 + h];
        dbcr[k] += sw * d0;
    } else if (i == n2 + K) {
        dbcs[0] += d0;
    }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_linear_fwd(const float* x, const float* w, const float* b, float* out, int N, int K, int M, int relu, void* stream) {
    int64_t threads = (int64_t)N * M * 32;
    linear_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, SG_STREAM(stream)>>>(x, w, b, out, N, K, M, relu);
    SG_LAUNCHED("linear_fwd");
    return 0;
}

int sg_linear_bwd(const float* x, const float* w, const float* dout, const float* relu_out, float* dw, float* db,
                  float* dx, int dx_acc, int N, int K, int M, void* stream) {
    cudaStream_t st = SG_STREAM(stream);
    if (dw) {
        linear_dw_kernel<<<(unsigned)(((int64_t)M * K + 255) / 256), 256, 0, st>>>(x, dout, relu_out, dw, N, K, M);
        SG_LAUNCHED("linear_dw");
    }
    if (db) {
        linear_db_kernel<<<(M + 127) / 128, 128, 0, st>>>(dout, relu_out, db, N, M);
        SG_LAUNCHED("linear_db");
    }
    if (dx) {
        linear_dx_kernel<<<(unsigned)(((int64_t)N * K + 255) / 256), 256, 0, st>>>(w, dout, relu_out, dx, dx_acc, N, K, M);
        SG_LAUNCHED("linear_dx");
    }
    return 0;
}

int sg_head_prepare(const float* wcr, const float* bcr, const float* wcs, const float* bcs, float* A, float* Bv,
                    float* c0, int K, int Cx, int Nd, void* stream) {
    int n = 16 * Cx + Nd + 1;
    head_prepare_kernel<<<(n + 255) / 256, 256, 0, SG_STREAM(stream)>>>(wcr, bcr, wcs, bcs, A, Bv, c0, K, Cx, Nd);
    SG_LAUNCHED("head_prepare");
    return 0;
}

int sg_head_fwd(const void* a4, const float* ce, const float* A, const float* Bv, const float* c0, float* score, int N,
                int M, int Nd, int dtype, void* stream) {
    SG_REQUIRE(M % 4 == 0, "head_fwd: M %% 4 != 0");
    SG_DISPATCH_T(dtype, (head_fwd_kernel<T><<<N, 256, 0, SG_STREAM(stream)>>>((const T*)a4, ce, A, Bv, c0, score, M, Nd)));
    SG_LAUNCHED("head_fwd");
    return 0;
}

int sg_wsum_rows(const float* coef, const void* x, float* out, int N, int M, int dtype, void* stream) {
    int splits = 16;
    if (splits > N) splits = N;
    int npb = (N + splits - 1) / splits;
    splits = (N + npb - 1) / npb;
    dim3 grid((M + 255) / 256, splits);
    SG_DISPATCH_T(dtype, (wsum_rows_kernel<T><<<grid, 256, 0, SG_STREAM(stream)>>>(coef, (const T*)x, out, N, M, npb)));
    SG_LAUNCHED("wsum_rows");
    return 0;
}

int sg_head_param_grads(const float* dA, const float* dBv, const float* dc0, const float* wcr, const float* bcr,
                        const float* wcs, float* dwcr, float* dbcr, float* dwcs, float* dbcs, int K, int Cx, int Nd,
                        void* stream) {
    int64_t n = (int64_t)K * (Cx + Nd) + (int64_t)K * 16 + K + 1;
    head_param_grads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, SG_STREAM(stream)>>>(dA, dBv, dc0, wcr, bcr, wcs, dwcr,
                                                                                       dbcr, dwcs, dbcs, K, Cx, Nd);
    SG_LAUNCHED("head_param_grads");
    return 0;
}

}  // extern "C"
