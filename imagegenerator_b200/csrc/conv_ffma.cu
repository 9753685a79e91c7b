// conv_ffma.cu -- fp32-accumulate CUDA-core implicit-GEMM convolution (fprop / dgrad / wgrad).
//
// Role: (1) the fp32 validation mode of the train step (tolerance 1e-4 needs real fp32 products;
// kind::tf32 MMAs would not meet it, SURVEY.md section 7 hard part 5); (2) the thin layers whose channel
// count is not a tensor-core shape (Cin=3 / Cout=3 / 24 / 228); (3) bring-up reference for the
// tcgen05 kernels in conv_tc.cu.  No im2col buffer is materialised: the gather happens while the
// 64x16 operand tiles are staged in shared memory.
//
// GEMM views (NHWC, Conv2d orientation, K index = (tap, channel) with channel fastest):
//   fprop : C[m=(n,oh,ow)][co]  = sum_{tap,ci} x[n, oh*s-p+kh, ow*s-p+kw, ci] * pf[co][tap][ci]
//   dgrad : C[m=(n,ih,iw)][ci]  = sum_{tap,co} dy[n,(ih+p-kh)/s,(iw+p-kw)/s,co] * pd[ci][tap][co]
//           split into s*s output-parity phases so that only the taps that hit are visited
//   wgrad : C[co][(tap,ci)]    += sum_{pix}    dy[pix][co] * x[pix@tap][ci]      (split over pixels)
#include "common.cuh"

namespace sg {

constexpr int BM = 64, BN = 64, BK = 16, PADW = 4;

template <typename T>
__device__ __forceinline__ void load4_or_zero(const T* base, bool ok, float* o) {
    if (ok) {
        F4 v = ld4(base);
        o[0] = v.v[0]; o[1] = v.v[1]; o[2] = v.v[2]; o[3] = v.v[3];
    } else {
        o[0] = o[1] = o[2] = o[3] = 0.f;
    }
}

__device__ __forceinline__ void mma_tile(const float (*As)[BM + PADW], const float (*Bs)[BN + PADW], int ty, int tx,
                                         float acc[4][4]) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
    }
}

// ------------------------------------------------------------------------------------------------ fprop
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) conv_fprop_ffma_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                              const float* __restrict__ bias, T* __restrict__ y, int N,
                                                              int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s,
                                                              int p, int act) {
    __shared__ __align__(16) float As[BK][BM + PADW];
    __shared__ __align__(16) float Bs[BK][BN + PADW];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t M = (int64_t)N * Ho * Wo;
    const int Kt = k * k * Ci;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;

    int64_t m = m0 + lrow;
    bool mvalid = m < M;
    int n_img = 0, ih0 = 0, iw0 = 0;
    if (mvalid) {
        n_img = (int)(m / (Ho * Wo));
        int r = (int)(m - (int64_t)n_img * Ho * Wo);
        int oh = r / Wo, ow = r - oh * Wo;
        ih0 = oh * s - p; iw0 = ow * s - p;
    }
    const T* xn = x + (int64_t)n_img * H * W * Ci;
    const int co_l = n0 + lrow;
    const bool covalid = co_l < Co;
    const T* wr = w + (int64_t)co_l * Kt;

    float acc[4][4] = {};
    for (int k0 = 0; k0 < Kt; k0 += BK) {
        float av[4], bv[4];
        int kk = k0 + lk;
        if (VEC) {
            bool ok = mvalid && kk < Kt;
            const T* src = xn;
            if (ok) {
                int tap = kk / Ci, ci = kk - tap * Ci;
                int kh = tap / k, kw = tap - kh * k;
                int ih = ih0 + kh, iw = iw0 + kw;
                ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
                src = xn + ((int64_t)ih * W + iw) * Ci + ci;
            }
            load4_or_zero(src, ok, av);
            load4_or_zero(wr + kk, covalid && kk < Kt, bv);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int q = kk + j;
                float a = 0.f, b = 0.f;
                if (q < Kt) {
                    if (mvalid) {
                        int tap = q / Ci, ci = q - tap * Ci;
                        int kh = tap / k, kw = tap - kh * k;
                        int ih = ih0 + kh, iw = iw0 + kw;
                        if (ih >= 0 && ih < H && iw >= 0 && iw < W) a = ldf(xn + ((int64_t)ih * W + iw) * Ci + ci);
                    }
                    if (covalid) b = ldf(wr + q);
                }
                av[j] = a; bv[j] = b;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { As[lk + j][lrow] = av[j]; Bs[lk + j][lrow] = bv[j]; }
        __syncthreads();
        mma_tile(As, Bs, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t mm = m0 + ty * 4 + i;
        if (mm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co < Co) {
                float v = acc[i][j] + (bias ? bias[co] : 0.f);
                stf(y + mm * Co + co, act_fwd(v, act));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ dgrad
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) conv_dgrad_ffma_kernel(const T* __restrict__ dy, const T* __restrict__ pd,
                                                              const float* __restrict__ bias, T* __restrict__ dx, int N,
                                                              int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s,
                                                              int p, int act, int tiles_per_phase) {
    __shared__ __align__(16) float As[BK][BM + PADW];
    __shared__ __align__(16) float Bs[BK][BN + PADW];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int phase = blockIdx.x / tiles_per_phase, tile = blockIdx.x - phase * tiles_per_phase;
    const int ph = phase / s, pw = phase - ph * s;
    const int Hq = (H - ph + s - 1) / s, Wq = (W - pw + s - 1) / s;
    const int64_t Mq = (int64_t)N * Hq * Wq;
    const int64_t m0 = (int64_t)tile * BM;
    if (m0 >= Mq) return;
    const int rh = (ph + p) % s, rw = (pw + p) % s;
    const int njh = rh < k ? (k - rh + s - 1) / s : 0, njw = rw < k ? (k - rw + s - 1) / s : 0;
    const int base_h = (ph + p - rh) / s, base_w = (pw + p - rw) / s;
    const int Kt = njh * njw * Co;
    const int n0 = blockIdx.y * BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;

    int64_t m = m0 + lrow;
    bool mvalid = m < Mq;
    int n_img = 0, q = 0, pp = 0;
    if (mvalid) {
        n_img = (int)(m / (Hq * Wq));
        int r = (int)(m - (int64_t)n_img * Hq * Wq);
        q = r / Wq; pp = r - q * Wq;
    }
    const T* dyn = dy + (int64_t)n_img * Ho * Wo * Co;
    const int ci_l = n0 + lrow;
    const bool civalid = ci_l < Ci;
    const T* wr = pd + (int64_t)ci_l * k * k * Co;

    float acc[4][4] = {};
    for (int k0 = 0; k0 < Kt; k0 += BK) {
        float av[4], bv[4];
        int kk = k0 + lk;
        if (VEC) {
            bool okk = kk < Kt;
            int jt = 0, co = 0, jh = 0, jw = 0;
            if (okk) { jt = kk / Co; co = kk - jt * Co; jh = jt / njw; jw = jt - jh * njw; }
            int oh = q + base_h - jh, ow = pp + base_w - jw;
            bool oka = okk && mvalid && oh >= 0 && oh < Ho && ow >= 0 && ow < Wo;
            load4_or_zero(dyn + ((int64_t)oh * Wo + ow) * Co + co, oka, av);
            int kh = rh + s * jh, kw = rw + s * jw;
            load4_or_zero(wr + (int64_t)(kh * k + kw) * Co + co, okk && civalid, bv);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int qq = kk + j;
                float a = 0.f, b = 0.f;
                if (qq < Kt) {
                    int jt = qq / Co, co = qq - jt * Co;
                    int jh = jt / njw, jw = jt - jh * njw;
                    int oh = q + base_h - jh, ow = pp + base_w - jw;
                    if (mvalid && oh >= 0 && oh < Ho && ow >= 0 && ow < Wo) a = ldf(dyn + ((int64_t)oh * Wo + ow) * Co + co);
                    int kh = rh + s * jh, kw = rw + s * jw;
                    if (civalid) b = ldf(wr + (int64_t)(kh * k + kw) * Co + co);
                }
                av[j] = a; bv[j] = b;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { As[lk + j][lrow] = av[j]; Bs[lk + j][lrow] = bv[j]; }
        __syncthreads();
        mma_tile(As, Bs, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t mm = m0 + ty * 4 + i;
        if (mm >= Mq) continue;
        int ni = (int)(mm / (Hq * Wq));
        int r = (int)(mm - (int64_t)ni * Hq * Wq);
        int qq = r / Wq, pq = r - qq * Wq;
        int ih = qq * s + ph, iw = pq * s + pw;
        T* o = dx + (((int64_t)ni * H + ih) * W + iw) * Ci;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = n0 + tx * 4 + j;
            if (ci < Ci) stf(o + ci, act_fwd(acc[i][j] + (bias ? bias[ci] : 0.f), act));
        }
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_ffma_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                              float* __restrict__ dw, int N, int H, int W, int Ci, int Ho,
                                                              int Wo, int Co, int k, int s, int p, int64_t pix_per_split) {
    __shared__ __align__(16) float As[BK][BM + PADW];   // [pixel][co]
    __shared__ __align__(16) float Bs[BK][BN + PADW];   // [pixel][(tap,ci)]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t Mp = (int64_t)N * Ho * Wo;
    const int kk2 = k * k, Nt = kk2 * Ci;
    const int co0 = blockIdx.x * BM, c0 = blockIdx.y * BN;
    int64_t p0 = (int64_t)blockIdx.z * pix_per_split, p1 = p0 + pix_per_split;
    if (p1 > Mp) p1 = Mp;
    const int lcol = tid & 63, lkb = (tid >> 6) * 4;
    const int co_l = co0 + lcol;
    const int col = c0 + lcol;
    const bool covalid = co_l < Co, colvalid = col < Nt;
    int tap = 0, ci = 0, kh = 0, kw = 0;
    if (colvalid) { tap = col / Ci; ci = col - tap * Ci; kh = tap / k; kw = tap - kh * k; }

    float acc[4][4] = {};
    for (int64_t pk = p0; pk < p1; pk += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t pix = pk + lkb + j;
            float a = 0.f, b = 0.f;
            if (pix < p1) {
                if (covalid) a = ldf(dy + pix * Co + co_l);
                if (colvalid) {
                    int n_img = (int)(pix / (Ho * Wo));
                    int r = (int)(pix - (int64_t)n_img * Ho * Wo);
                    int oh = r / Wo, ow = r - oh * Wo;
                    int ih = oh * s - p + kh, iw = ow * s - p + kw;
                    if (ih >= 0 && ih < H && iw >= 0 && iw < W) b = ldf(x + (((int64_t)n_img * H + ih) * W + iw) * Ci + ci);
                }
            }
            As[lkb + j][lcol] = a;
            Bs[lkb + j][lcol] = b;
        }
        __syncthreads();
        mma_tile(As, Bs, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= Co) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int cc = c0 + tx * 4 + j;
            if (cc < Nt) {
                int t = cc / Ci, c = cc - t * Ci;
                atomicAdd(dw + ((int64_t)co * Ci + c) * kk2 + t, acc[i][j]);
            }
        }
    }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_conv_fprop_ffma(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho,
                       int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(Ho == (H + 2 * p - k) / s + 1 && Wo == (W + 2 * p - k) / s + 1, "conv_fprop: inconsistent output size");
    int64_t M = (int64_t)N * Ho * Wo;
    dim3 grid((unsigned)((M + BM - 1) / BM), (Co + BN - 1) / BN);
    bool vec = (Ci % 4 == 0);
    SG_DISPATCH_T(dtype, {
        if (vec)
            conv_fprop_ffma_kernel<T, true><<<grid, 256, 0, SG_STREAM(stream)>>>((const T*)x, (const T*)pf, bias, (T*)y, N, H,
                                                                                 W, Ci, Ho, Wo, Co, k, s, p, act);
        else
            conv_fprop_ffma_kernel<T, false><<<grid, 256, 0, SG_STREAM(stream)>>>((const T*)x, (const T*)pf, bias, (T*)y, N,
                                                                                  H, W, Ci, Ho, Wo, Co, k, s, p, act);
    });
    SG_LAUNCHED("conv_fprop_ffma");
    return 0;
}

int sg_conv_dgrad_ffma(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho,
                       int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k,
               "conv_dgrad: dx size must equal the transposed-conv output size");
    // 1x1 input (Stage-I generator's first layer): the stride is immaterial; using s = k turns the k*k
    // output pixels into k*k parity phases with exactly ONE live tap each instead of k*k mostly-empty taps
    if (Ho == 1 && Wo == 1 && p == 0 && H == k && W == k) s = k;
    SG_REQUIRE(H % s == 0 && W % s == 0, "conv_dgrad: H, W must be multiples of the stride");
    int Hq = H / s, Wq = W / s;
    int64_t Mq = (int64_t)N * Hq * Wq;
    int tiles = (int)((Mq + BM - 1) / BM);
    dim3 grid((unsigned)(tiles * s * s), (Ci + BN - 1) / BN);
    bool vec = (Co % 4 == 0);
    SG_DISPATCH_T(dtype, {
        if (vec)
            conv_dgrad_ffma_kernel<T, true><<<grid, 256, 0, SG_STREAM(stream)>>>((const T*)dy, (const T*)pd, bias, (T*)dx, N,
                                                                                 H, W, Ci, Ho, Wo, Co, k, s, p, act, tiles);
        else
            conv_dgrad_ffma_kernel<T, false><<<grid, 256, 0, SG_STREAM(stream)>>>((const T*)dy, (const T*)pd, bias, (T*)dx,
                                                                                  N, H, W, Ci, Ho, Wo, Co, k, s, p, act, tiles);
    });
    SG_LAUNCHED("conv_dgrad_ffma");
    return 0;
}

int sg_conv_wgrad_ffma(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k,
                       int s, int p, int dtype, void* stream) {
    int64_t Mp = (int64_t)N * Ho * Wo;
    int gx = (Co + BM - 1) / BM, gy = (k * k * Ci + BN - 1) / BN;
    int64_t want = (3 * SG_NUM_SMS + (int64_t)gx * gy - 1) / ((int64_t)gx * gy);
    int64_t max_splits = (Mp + 4 * BK - 1) / (4 * BK);
    int64_t splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
    int64_t pps = ((Mp + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (Mp + pps - 1) / pps;
    dim3 grid(gx, gy, (unsigned)splits);
    SG_DISPATCH_T(dtype, (conv_wgrad_ffma_kernel<T><<<grid, 256, 0, SG_STREAM(stream)>>>((const T*)x, (const T*)dy, dw, N, H,
                                                                                        W, Ci, Ho, Wo, Co, k, s, p, pps)));
    SG_LAUNCHED("conv_wgrad_ffma");
    return 0;
}

}  // extern "C"
