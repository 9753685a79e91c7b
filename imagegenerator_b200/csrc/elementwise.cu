// elementwise.cu -- HBM/L2-bound kernels of the StackGAN step: layout, BatchNorm (statistics,
// apply, backward, the gradient-penalty double backward), activations' backward, losses, Adam.
// All are streaming kernels: 4-wide vector accesses, channel-fastest (NHWC) indexing, per-channel
// reductions finished with one fp64 atomic per channel per CTA.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace sg {

std::atomic<long long> g_launches{0};
int g_use_pdl = 1;
int g_dbg_skip_memset = 0;     // option "dbg_skip_memset": timing experiment, WRONG results (what do the memset nodes cost?)
static thread_local char g_err[512] = "";


void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// generic 4-wide elementwise driver: F::apply(i4, lanes...) handles elements [4*i4, 4*i4+4)
// ------------------------------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) ew4_kernel(F f, int64_t n4) {
    SG_PDL_SYNC();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) f(i);
}

template <typename F>
static int launch_ew4(F f, int64_t n, cudaStream_t st, const char* what) {
    if (n % 4 != 0) {
        set_error("%s: element count %lld not a multiple of 4", what, (long long)n);
        return SG_ERR_BAD_ARG;
    }
    int64_t n4 = n / 4;
    if (n4 == 0) return 0;
    launch_pdl(ew4_kernel<F>, dim3(grid_for(n4, 256, 16)), dim3(256), 0, st, f, n4);
    g_launches.fetch_add(1);
    return check_launch(what);
}

// ---- BN apply + activation (+ residual)
template <typename T>
struct BnActF {
    const T* y; const float* mr; const float* gamma; const float* beta; const T* res; T* out;
    int64_t rows_per_group; int C; int act;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        F4 v = ld4(y + e), r;
        F4 rs;
        if (res) rs = ld4(res + e);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t row = (e + j) / C;
            int c = (int)((e + j) - row * C);
            int g = (int)(row / rows_per_group);
            const float* m = mr + ((int64_t)g * C + c) * 2;
            float z = (v.v[j] - m[0]) * m[1] * gamma[c] + beta[c];
            if (res) z += rs.v[j];
            r.v[j] = act_fwd(z, act);
        }
        st4(out + e, r);
    }
};

// ---- BN backward apply
template <typename T>
struct BnBwdApplyF {
    const T* da; const T* a_out; const T* y; const float* mr; const float* gamma; const double* sums;
    const T* inject; int inject_group; T* dy; int64_t rows_per_group; int C; int act;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        F4 d = ld4(da + e), a = ld4(a_out + e), yy = ld4(y + e), r, inj;
        int64_t row0 = e / C;
        int g0 = (int)(row0 / rows_per_group);
        bool has_inj = inject != nullptr && g0 == inject_group;   // 4 elements never straddle groups (group size % 4 == 0)
        if (has_inj) inj = ld4(inject + (e - (int64_t)inject_group * rows_per_group * C));
        float n = (float)rows_per_group;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t row = (e + j) / C;
            int c = (int)((e + j) - row * C);
            int g = (int)(row / rows_per_group);
            const float* m = mr + ((int64_t)g * C + c) * 2;
            const double* s = sums + ((int64_t)g * C + c) * 2;
            float xh = (yy.v[j] - m[0]) * m[1];
            float dz = d.v[j] * act_mask(a.v[j], act);
            float coef = gamma[c] * m[1] / n;
            float o = coef * (n * dz - (float)s[0] - xh * (float)s[1]);
            if (has_inj) o += inj.v[j];
            r.v[j] = o;
        }
        st4(dy + e, r);
    }
};

template <typename T>
struct ActBwdF {
    const T* da; const T* a_out; T* out; int act;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        F4 d = ld4(da + e), a = ld4(a_out + e), r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = d.v[j] * act_mask(a.v[j], act);
        st4(out + e, r);
    }
};

// ---- gradient-penalty BN double backward, apply part
template <typename T>
struct GpBnApplyF {
    const T* v; const T* da; const T* a_out; const T* y; const float* mr; const float* gamma;
    const double* sums; const double* tsums; T* w_out; T* gy_out; int64_t rows; int C; int act;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        F4 vv = ld4(v + e), d = ld4(da + e), a = ld4(a_out + e), yy = ld4(y + e), w, gy;
        float n = (float)rows;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = (int)((e + j) % C);
            float mean = mr[c * 2], r = mr[c * 2 + 1];
            float S1 = (float)sums[c * 2], S2 = (float)sums[c * 2 + 1];
            float T1 = (float)tsums[c * 3], T2 = (float)tsums[c * 3 + 1], T3 = (float)tsums[c * 3 + 2];
            float al = gamma[c] * r / n;
            float xh = (yy.v[j] - mean) * r;
            float mk = act_mask(a.v[j], act);
            float dz = d.v[j] * mk;
            float u = al * (n * vv.v[j] - T1 - xh * T2);
            w.v[j] = u * mk;
            float P = al * (n * T3 - S1 * T1 - S2 * T2);
            float G = -al * (vv.v[j] * S2 + dz * T2);
            float sG = -al * (S2 * T1 + S1 * T2);
            float sGx = -2.f * al * S2 * T2;
            gy.v[j] = r * (G - sG / n - xh * sGx / n) - P * r * xh / n;
        }
        st4(w_out + e, w);
        st4(gy_out + e, gy);
    }
};

__global__ void gp_bn_dgamma_kernel(const float* mr, const double* sums, const double* tsums, float* dgamma,
                                    double n, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double r = mr[c * 2 + 1];
    double S1 = sums[c * 2], S2 = sums[c * 2 + 1];
    double T1 = tsums[c * 3], T2 = tsums[c * 3 + 1], T3 = tsums[c * 3 + 2];
    dgamma[c] += (float)(r / n * (n * T3 - S1 * T1 - S2 * T2));
}

// ---- per-sample scale kernels
template <typename T>
struct InterpF {
    const T* real; const T* fake; const float* eps; T* out; int64_t per_sample;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        float ep = eps[e / per_sample];
        F4 a = ld4(real + e), b = ld4(fake + e), r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = a.v[j] * ep + b.v[j] * (1.f - ep);
        st4(out + e, r);
    }
};

template <typename T>
struct GpSeedF {
    const T* g; const float* sq; float coef; T* v; int64_t per_sample;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        // d||g||/dg at g = 0 is taken as 0, like torch's norm backward (1 - 1/0 would seed NaNs through g * -inf)
        const float q = sq[e / per_sample];
        float f = q > 0.f ? coef * (1.f - rsqrtf(q)) : 0.f;
        F4 a = ld4(g + e), r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = a.v[j] * f;
        st4(v + e, r);
    }
};

template <typename T>
struct ScaleRowsAddF {
    const T* x; const float* scale; T* out; int accumulate; int64_t per_sample;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        float s = scale[e / per_sample];
        F4 a = ld4(x + e), r;
        if (accumulate) r = ld4(out + e);
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = (accumulate ? r.v[j] : 0.f) + a.v[j] * s;
        st4(out + e, r);
    }
};

template <typename T>
struct OuterF {
    const float* coef; const float* vec; T* out; int M;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        int64_t n = e / M;
        int m = (int)(e - n * M);
        float c = coef[n];
        F4 r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = c * vec[m + j];
        st4(out + e, r);
    }
};

struct AdamF {
    float* p; const float* g; float* m; float* v; const float* hyper;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], t = hyper[4];
        float bc1 = 1.f - powf(b1, t), bc2s = sqrtf(1.f - powf(b2, t));
        F4 pp = ld4(p + e), gg = ld4(g + e), mm = ld4(m + e), vv = ld4(v + e);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mm.v[j] = b1 * mm.v[j] + (1.f - b1) * gg.v[j];
            vv.v[j] = b2 * vv.v[j] + (1.f - b2) * gg.v[j] * gg.v[j];
            float denom = sqrtf(vv.v[j]) / bc2s + eps;
            pp.v[j] -= (lr / bc1) * (mm.v[j] / denom);
        }
        st4(p + e, pp); st4(m + e, mm); st4(v + e, vv);
    }
};

// hyper[4]: the step as a float (bias correction; saturates harmlessly at 2^24, where beta^t == 0 anyway);
// hyper[5], hyper[6]: the exact count as lo + 2^23 * hi for checkpoints
__global__ void adam_tick_kernel(float* hyper) {
    hyper[4] += 1.f;
    float lo = hyper[5] + 1.f;
    if (lo >= 8388608.f) { lo = 0.f; hyper[6] += 1.f; }
    hyper[5] = lo;
}

struct FillF {
    float* p; float val;
    __device__ void operator()(int64_t i4) const {
        F4 r;
        r.v[0] = r.v[1] = r.v[2] = r.v[3] = val;
        st4(p + i4 * 4, r);
    }
};
__global__ void affine_f32_kernel(const float* x, float a, float b, float* out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + b;
}
__global__ void fill_tail_kernel(float* p, float v, int64_t from, int64_t n) {
    int64_t i = from + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// per-channel reductions over rows of a [rows][C] tensor
//   block = (TX channel lanes) x (256/TX row lanes); grid = (row chunks, groups, channel chunks)
// ------------------------------------------------------------------------------------------------
template <int NV, typename F, typename OutT>
__global__ void __launch_bounds__(256) rowreduce_kernel(F f, OutT* out, int64_t rows_per_group, int C,
                                                         int64_t rows_per_block, int out_group_stride) {
    SG_PDL_SYNC();
    __shared__ double sh[NV][256];
    int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    int c = blockIdx.z * TX + tx;
    int g = blockIdx.y;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > rows_per_group) r1 = rows_per_group;
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
    if (c < C) {
        for (int64_t r = r0 + ty; r < r1; r += TY) {
            float vals[NV];
            f((int64_t)g * rows_per_group + r, c, g, vals);
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] += vals[k];
        }
    }
    int tid = ty * TX + tx;
#pragma unroll
    for (int k = 0; k < NV; ++k) sh[k][tid] = (double)acc[k];
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (int j = 0; j < TY; ++j) s += sh[k][j * TX + tx];
            atomicAdd(out + (int64_t)g * out_group_stride + (int64_t)c * NV + k, (OutT)s);
        }
    }
}

template <int NV, typename F, typename OutT>
static int launch_rowreduce(F f, OutT* out, int64_t rows_per_group, int C, int groups, cudaStream_t st,
                            const char* what) {
    int TX = 4;
    while (TX < C && TX < 64) TX *= 2;
    int TY = 256 / TX;
    int cchunks = (C + TX - 1) / TX;
    int64_t target_blocks = (4 * SG_NUM_SMS + groups * cchunks - 1) / (groups * cchunks);
    if (target_blocks < 1) target_blocks = 1;
    int64_t rpb = (rows_per_group + target_blocks - 1) / target_blocks;
    int64_t min_rpb = (int64_t)TY * 4;
    if (rpb < min_rpb) rpb = min_rpb;
    if (rpb > (int64_t)TY * 256) rpb = (int64_t)TY * 256;   // bound the fp32 partial length per thread
    int64_t rblocks = (rows_per_group + rpb - 1) / rpb;
    dim3 grid((unsigned)rblocks, groups, cchunks), block(TX, TY);
    launch_pdl(rowreduce_kernel<NV, F, OutT>, grid, block, 0, st, f, out, rows_per_group, C, rpb, C * NV);
    g_launches.fetch_add(1);
    return check_launch(what);
}

template <typename T>
struct ColStatsF {
    const T* y; int C;
    __device__ void operator()(int64_t row, int c, int g, float* o) const {
        float v = ldf(y + row * C + c);
        o[0] = v; o[1] = v * v;
    }
};
template <typename T>
struct ColSumF {
    const T* x; int C;
    __device__ void operator()(int64_t row, int c, int g, float* o) const { o[0] = ldf(x + row * C + c); }
};
template <typename T>
struct BnBwdReduceF {
    const T* da; const T* a_out; const T* y; const float* mr; int C; int act;
    __device__ void operator()(int64_t row, int c, int g, float* o) const {
        int64_t i = row * C + c;
        const float* m = mr + ((int64_t)g * C + c) * 2;
        float dz = ldf(da + i) * act_mask(ldf(a_out + i), act);
        o[0] = dz;
        o[1] = dz * (ldf(y + i) - m[0]) * m[1];
    }
};
template <typename T>
struct GpBnReduceF {
    const T* v; const T* da; const T* a_out; const T* y; const float* mr; int C; int act;
    __device__ void operator()(int64_t row, int c, int g, float* o) const {
        int64_t i = row * C + c;
        float vv = ldf(v + i);
        float dz = ldf(da + i) * act_mask(ldf(a_out + i), act);
        o[0] = vv;
        o[1] = vv * (ldf(y + i) - mr[c * 2]) * mr[c * 2 + 1];
        o[2] = vv * dz;
    }
};

__global__ void bn_finalize_kernel(const double* stats, double count, float* mr, float* rm, float* rv,
                                   long long* nbt, int dup_first, int update_running, float momentum, float eps,
                                   int G, int C) {
    SG_PDL_SYNC();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && update_running && nbt) *nbt += dup_first + G - 1;
    if (c >= C) return;
    float m_run = 0.f, v_run = 0.f;
    if (update_running) { m_run = rm[c]; v_run = rv[c]; }
    for (int g = 0; g < G; ++g) {
        double s = stats[((int64_t)g * C + c) * 2], q = stats[((int64_t)g * C + c) * 2 + 1];
        double mean = s / count;
        double var = q / count - mean * mean;
        if (var < 0) var = 0;
        mr[((int64_t)g * C + c) * 2] = (float)mean;
        mr[((int64_t)g * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
        if (update_running) {
            float unb = (float)(var * count / (count > 1 ? count - 1 : 1));
            int reps = g == 0 ? dup_first : 1;
            for (int r = 0; r < reps; ++r) {
                m_run = (1.f - momentum) * m_run + momentum * (float)mean;
                v_run = (1.f - momentum) * v_run + momentum * unb;
            }
        }
    }
    if (update_running) { rm[c] = m_run; rv[c] = v_run; }
}

__global__ void bn_eval_mr_kernel(const float* rm, const float* rv, float* mr, float eps, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        mr[c * 2] = rm[c];
        mr[c * 2 + 1] = (float)(1.0 / sqrt((double)rv[c] + (double)eps));
    }
}

__global__ void bn_param_grad_kernel(const double* sums, float* dgamma, float* dbeta, int G, int C) {
    SG_PDL_SYNC();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s1 = 0, s2 = 0;
    for (int g = 0; g < G; ++g) {
        s1 += sums[((int64_t)g * C + c) * 2];
        s2 += sums[((int64_t)g * C + c) * 2 + 1];
    }
    dgamma[c] += (float)s2;
    dbeta[c] += (float)s1;
}

// the same for every BatchNorm layer of a network in ONE launch (blockIdx.y = layer): a backward pass used to end in one tiny
// launch per layer (20 per Stage-I step, 121 per Stage-II step)
constexpr int BN_PG_MAX = 24;
struct BnParamGradArgs {
    const double* sums[BN_PG_MAX];
    float* dgamma[BN_PG_MAX];
    float* dbeta[BN_PG_MAX];
    int G[BN_PG_MAX], C[BN_PG_MAX];
};
__global__ void bn_param_grad_multi_kernel(const BnParamGradArgs A) {
    SG_PDL_SYNC();
    const int l = blockIdx.y, C = A.C[l], G = A.G[l];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double* sums = A.sums[l];
    double s1 = 0, s2 = 0;
    for (int g = 0; g < G; ++g) {
        s1 += sums[((int64_t)g * C + c) * 2];
        s2 += sums[((int64_t)g * C + c) * 2 + 1];
    }
    A.dgamma[l][c] += (float)s2;
    A.dbeta[l][c] += (float)s1;
}

// ---- layout
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* src, T* dst, int N, int C, int64_t HW) {
    int64_t total = (int64_t)N * HW;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int64_t n = i / HW, hw = i - n * HW;
        const float* s = src + n * C * HW + hw;
        T* d = dst + i * C;
        for (int c = 0; c < C; ++c) stf(d + c, s[(int64_t)c * HW]);
    }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* src, float* dst, int N, int C, int64_t HW) {
    int64_t total = (int64_t)N * HW;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int64_t n = i / HW, hw = i - n * HW;
        const T* s = src + i * C;
        float* d = dst + n * C * HW + hw;
        for (int c = 0; c < C; ++c) d[(int64_t)c * HW] = ldf(s + c);
    }
}
// NHWC tanh output in [-1, 1] -> NCHW uint8 image: round((x + 1) * 127.5), clamped (the usual de-normalisation of the
// transforms.Normalize((0.5,)*3, (0.5,)*3) the reference trains with, train.py:104-110) -- a quarter of the fp32 read-back
template <typename T>
__global__ void nhwc_to_nchw_u8_kernel(const T* src, unsigned char* dst, int N, int C, int64_t HW) {
    int64_t total = (int64_t)N * HW;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int64_t n = i / HW, hw = i - n * HW;
        const T* s = src + i * C;
        unsigned char* d = dst + n * C * HW + hw;
        for (int c = 0; c < C; ++c) {
            const float v = fminf(fmaxf((ldf(s + c) + 1.f) * 127.5f, 0.f), 255.f);
            d[(int64_t)c * HW] = (unsigned char)__float2int_rn(v);
        }
    }
}
// w[co][ci][t] (fp32 master) -> wt[(t, ci)][Kp] in the storage type, columns co >= Co zero: the forward operand of a
// ConvTranspose2d on a 1x1 input run as a GEMM (engine.Up0Gemm).  Block = one ci; a thread owns one co, reads its kk <= 16
// contiguous taps and stores them kk rows apart -- consecutive threads write consecutive columns.
template <typename T>
__global__ void __launch_bounds__(256) pack_gemm_t_kernel(const float* __restrict__ w, T* __restrict__ wt, int Co, int Ci, int kk,
                                                          int Kp) {
    const int ci = blockIdx.x;
    for (int co = threadIdx.x; co < Kp; co += blockDim.x) {
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = 0.f;
        if (co < Co) {
            const float* src = w + ((int64_t)co * Ci + ci) * kk;
            for (int t = 0; t < kk; ++t) v[t] = src[t];
        }
        for (int t = 0; t < kk; ++t) stf(wt + ((int64_t)t * Ci + ci) * Kp + co, v[t]);
    }
}

// w[co][ci][t] (fp32 master) -> pf[co][t][ci] and pd[ci][t][co] in the storage type.  A thread owns one (co, ci) pair
// and reads its kk contiguous taps; blockIdx.y = 0 runs with ci fastest across threads (pf stores coalesced), = 1 with
// co fastest (pd stores coalesced) -- element-order stores were 2-byte scatters at stride Ci / Co (32 us for the 2 M
// element critic layer, on the path between the optimizer step and the next forward).
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, T* __restrict__ pf,
                                                          T* __restrict__ pd, int Co, int Ci, int kk, int first_dir) {
    const int dir = first_dir + blockIdx.y;
    const int64_t pairs = (int64_t)Co * Ci;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
        int co, ci;
        if (dir == 0) { co = (int)(i / Ci); ci = (int)(i - (int64_t)co * Ci); }
        else          { ci = (int)(i / Co); co = (int)(i - (int64_t)ci * Co); }
        const float* src = w + ((int64_t)co * Ci + ci) * kk;
        T* dst = dir == 0 ? pf + (int64_t)co * kk * Ci + ci : pd + (int64_t)ci * kk * Co + co;
        const int64_t ds = dir == 0 ? Ci : Co;
        for (int t = 0; t < kk; ++t) stf(dst + t * ds, src[t]);
    }
}

// dw[co][ci][t] += gw[co][t][ci]; gw = 0   (fold a channels-last accumulation buffer into the PyTorch-layout gradient)
__global__ void __launch_bounds__(256) fold_grad_cl_kernel(float* __restrict__ gw, float* __restrict__ dw, int Ci, int kk,
                                                           int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int ci = (int)(i % Ci);
        const int64_t r = i / Ci;
        const int t = (int)(r % kk);
        const int64_t co = r / kk;
        dw[(co * Ci + ci) * kk + t] += gw[i];
        gw[i] = 0.f;
    }
}

// ---- inference: eval-mode BatchNorm folded into the conv that precedes it
__global__ void bn_fold_kernel(const float* rm, const float* rv, const float* gamma, const float* beta, float eps, float* scale,
                               float* shift, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        float sc = (float)((double)gamma[c] / sqrt((double)rv[c] + (double)eps));
        scale[c] = sc;
        shift[c] = beta[c] - rm[c] * sc;
    }
}
template <typename T>
__global__ void pack_weight_scaled_kernel(const float* w, const float* scale, int axis, T* pf, T* pd, int Co, int Ci, int kk) {
    int64_t total = (int64_t)Co * Ci * kk;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int t = (int)(i % kk);
        int64_t r = i / kk;
        int ci = (int)(r % Ci);
        int co = (int)(r / Ci);
        float v = w[i] * scale[axis == 0 ? co : ci];
        if (pf) stf(pf + ((int64_t)co * kk + t) * Ci + ci, v);
        if (pd) stf(pd + ((int64_t)ci * kk + t) * Co + co, v);
    }
}
template <typename T>
struct AddActF {
    const T* a; const T* b; T* out; int act;
    __device__ void operator()(int64_t i4) const {
        int64_t e = i4 * 4;
        F4 x = ld4(a + e), y = ld4(b + e), r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = act_fwd(x.v[j] + y.v[j], act);
        st4(out + e, r);
    }
};

// ---- patch matrix of a thin (<= 4 channel) image: P[n,oh,ow, ci*k*k + kh*k + kw] = x[n, oh*s-p+kh, ow*s-p+kw, ci]
// (the PyTorch weight order, so the layer becomes a 1x1 convolution over P for the tensor-core kernels)
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* __restrict__ x, T* __restrict__ P, int N, int H, int W,
                                                       int C, int Ho, int Wo, int k, int s, int p) {
    const int kk = k * k, K = C * kk;
    int64_t total = (int64_t)N * Ho * Wo * kk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i % kk);
        int64_t pix = i / kk;
        int ow = (int)(pix % Wo);
        int64_t r = pix / Wo;
        int oh = (int)(r % Ho), n = (int)(r / Ho);
        int kh = t / k, kw = t - kh * k;
        int ih = oh * s - p + kh, iw = ow * s - p + kw;
        bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
        const T* src = x + (((int64_t)n * H + ih) * W + iw) * C;
        T* dst = P + pix * K + t;
        for (int c = 0; c < C; ++c) dst[c * kk] = ok ? src[c] : T(0.f);
    }
}

// k=4, s=2, p=1 (every thin layer of the StackGAN networks): one thread per (output pixel, kernel row) reads the
// 4 input pixels of that row once and writes the 4 taps of each channel as one vector (8 B in bf16) -- the four
// threads of a pixel together fill whole 32-byte sectors of its 96-byte record.
template <typename T, int C>
__global__ void __launch_bounds__(256) patchify_k4s2_kernel(const T* __restrict__ x, T* __restrict__ P, int H, int W, int Ho,
                                                            int Wo, unsigned total) {
    SG_PDL_SYNC();
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned kh = i & 3u, pix = i >> 2;
        const unsigned ow = pix % (unsigned)Wo, r = pix / (unsigned)Wo;
        const unsigned oh = r % (unsigned)Ho, n = r / (unsigned)Ho;
        const int ih = (int)oh * 2 - 1 + (int)kh, iw0 = (int)ow * 2 - 1;
        const bool row_ok = ih >= 0 && ih < H;
        const T* src = x + (((int64_t)n * H + ih) * W + iw0) * C;
        float v[4][C];
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
            const bool ok = row_ok && iw0 + kw >= 0 && iw0 + kw < W;
#pragma unroll
            for (int c = 0; c < C; ++c) v[kw][c] = ok ? ldf(src + kw * C + c) : 0.f;
        }
        T* dst = P + (int64_t)pix * (C * 16) + kh * 4;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            F4 o;
            o.v[0] = v[0][c]; o.v[1] = v[1][c]; o.v[2] = v[2][c]; o.v[3] = v[3][c];
            st4(dst + c * 16, o);
        }
    }
}

// ---- col2im for a thin output (inverse of the patch matrix): out[n,oh,ow,c] = act(bias[c] + sum over the taps
// (kh,kw) with oh = ih*s-p+kh, ow = iw*s-p+kw of col[n,ih,iw, c*k*k + kh*k + kw]).  With the 1x1 tensor-core GEMM
// col = x * W^T in front, this is ConvTranspose2d(Cin -> 3) (generator_1.py:20, generator_2.py:55) and the data
// gradient of the critics' first conv (discrminator_1.py:10) without 3-wide MMA tiles.
template <typename T>
__global__ void __launch_bounds__(256) unpatchify_kernel(const float* __restrict__ col, const float* __restrict__ bias,
                                                         T* __restrict__ out, int Hi, int Wi, int Ho, int Wo, int C, int k,
                                                         int s, int p, int act, unsigned total) {
    SG_PDL_SYNC();
    const int kk = k * k, K = C * kk;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned ow = i % (unsigned)Wo, r = i / (unsigned)Wo;
        const unsigned oh = r % (unsigned)Ho, n = r / (unsigned)Ho;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kh = ((int)oh + p) % s; kh < k; kh += s) {
            const int ih = ((int)oh + p - kh) / s;
            if (ih < 0 || ih >= Hi) continue;
            for (int kw = ((int)ow + p) % s; kw < k; kw += s) {
                const int iw = ((int)ow + p - kw) / s;
                if (iw < 0 || iw >= Wi) continue;
                const float* src = col + (((int64_t)n * Hi + ih) * Wi + iw) * K + kh * k + kw;
                for (int c = 0; c < C; ++c) acc[c] += __ldg(src + c * kk);
            }
        }
        T* dst = out + (int64_t)i * C;
        for (int c = 0; c < C; ++c) stf(dst + c, act_fwd(acc[c] + (bias ? bias[c] : 0.f), act));
    }
}

// The k4 s2 p1 case (every thin layer of StackGAN) tiled through shared memory: a CTA stages the col rows of an 8x16
// block of input pixels plus a one-pixel halo with coalesced 16-byte loads (the per-pixel gather above touches every
// 32-byte sector about four times: 1.4 TB/s on the 201 MB col matrix of G2's last layer), then each thread gathers two
// output pixels of the 16x32 output block from shared memory.  Output (oh, ow) takes taps kh = (oh+1)%2 + {0,2} from
// input rows ih = (oh+1-kh)/2, likewise in w.
constexpr int UP_TH = 8, UP_TW = 16, UP_RS = 52;          // row stride 52 floats: 16-byte aligned, 2-way bank conflicts
template <typename T, int C>
__global__ void __launch_bounds__(256) unpatchify_k4s2p1_kernel(const float* __restrict__ col, const float* __restrict__ bias,
                                                                T* __restrict__ out, int Hi, int Wi, int act, int tiles_w,
                                                                int tiles_h) {
    SG_PDL_SYNC();
    constexpr int K = C * 16, K4 = K / 4, HR = UP_TH + 2, HC = UP_TW + 2;
    __shared__ __align__(16) float tile[HR * HC * UP_RS];
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int n = b / tiles_h;
    const int ih0 = th * UP_TH - 1, iw0 = tw * UP_TW - 1;
    for (int q = threadIdx.x; q < HR * HC * K4; q += 256) {
        const int pix = q / K4, k4 = q - pix * K4;
        const int r = pix / HC, c = pix - r * HC;
        const int ih = ih0 + r, iw = iw0 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi)
            v = __ldg(reinterpret_cast<const float4*>(col + (((int64_t)n * Hi + ih) * Wi + iw) * K) + k4);
        *reinterpret_cast<float4*>(tile + pix * UP_RS + k4 * 4) = v;
    }
    __syncthreads();
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    const int owl = threadIdx.x & 31, ow = tw * (2 * UP_TW) + owl;
    const int kw0 = (owl + 1) & 1;                               // tile origin is even, so parity of ow = parity of owl
    const int c0 = (owl + 1 - kw0) / 2 + 1;                      // halo column of tap kw0; tap kw0+2 sits one column left
    float bs[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) bs[ch] = bias ? bias[ch] : 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ohl = (threadIdx.x >> 5) + half * 8, oh = th * (2 * UP_TH) + ohl;
        if (oh >= Ho || ow >= Wo) continue;
        const int kh0 = (ohl + 1) & 1;
        const int r0 = (ohl + 1 - kh0) / 2 + 1;
        float acc[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int bq = 0; bq < 2; ++bq) {
                const float* src = tile + ((r0 - a) * HC + (c0 - bq)) * UP_RS + (kh0 + 2 * a) * 4 + kw0 + 2 * bq;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) acc[ch] += src[ch * 16];
            }
        }
        T* dst = out + (((int64_t)n * Ho + oh) * Wo + ow) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) stf(dst + ch, act_fwd(acc[ch] + bs[ch], act));
    }
}

// ---- text replicate + channel concat (generator_2.py:61-63) and its backward
template <typename T>
__global__ void __launch_bounds__(256) concat_rep_kernel(const T* __restrict__ x, const float* __restrict__ c,
                                                         T* __restrict__ out, int64_t rows, int HW, int Cx, int Cc) {
    const int C = Cx + Cc;
    int64_t total = rows * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = i / C;
        int ch = (int)(i - row * C);
        if (ch < Cx) out[i] = x[row * Cx + ch];
        else stf(out + i, c[(row / HW) * Cc + (ch - Cx)]);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) split_copy_kernel(const T* __restrict__ dout, T* __restrict__ dx, int64_t rows, int Cx,
                                                         int C) {
    int64_t total = rows * Cx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = i / Cx;
        int ch = (int)(i - row * Cx);
        dx[i] = dout[row * C + ch];
    }
}
// dc[n][j] = sum_hw dout[n,hw,Cx+j]; one CTA per (n, 32-channel slab)
template <typename T>
__global__ void __launch_bounds__(256) rep_bwd_kernel(const T* __restrict__ dout, float* __restrict__ dc, int HW, int Cx,
                                                      int Cc) {
    __shared__ float sh[8][33];
    const int n = blockIdx.y, j = blockIdx.x * 32 + (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
    const int C = Cx + Cc;
    float acc = 0.f;
    if (j < Cc)
        for (int r = lane_r; r < HW; r += 8) acc += ldf(dout + ((int64_t)n * HW + r) * C + Cx + j);
    sh[lane_r][threadIdx.x & 31] = acc;
    __syncthreads();
    if (lane_r == 0 && j < Cc) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x & 31];
        dc[(int64_t)n * Cc + j] = s;
    }
}

// ---- losses
template <typename T>
__global__ void __launch_bounds__(256) sample_sqnorm_kernel(const T* g, float* out, int64_t per_sample,
                                                            int64_t per_block) {
    int n = blockIdx.y;
    int64_t b0 = (int64_t)blockIdx.x * per_block, b1 = b0 + per_block;
    if (b1 > per_sample) b1 = per_sample;
    const T* p = g + (int64_t)n * per_sample;
    float acc = 0.f;
    for (int64_t i = b0 + threadIdx.x * 4; i < b1; i += 256 * 4) {
        F4 v = ld4(p + i);
        acc += v.v[0] * v.v[0] + v.v[1] * v.v[1] + v.v[2] * v.v[2] + v.v[3] * v.v[3];
    }
    __shared__ float sh[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += sh[i];
        atomicAdd(out + n, s);
    }
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    return s;
}

__global__ void __launch_bounds__(256) critic_loss_kernel(const float* s_real, const float* s_mis, const float* s_fake,
                                                          const float* sq, float lam, float* out, int N) {
    __shared__ double sh[8];
    double a = 0, b = 0, c = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        a += (double)s_mis[i] + (double)s_fake[i];
        b += s_real[i];
        double d = sqrt((double)sq[i]) - 1.0;
        c += d * d;
    }
    a = block_sum_d(a, sh); b = block_sum_d(b, sh); c = block_sum_d(c, sh);
    if (threadIdx.x == 0) {
        double gp = c / N;
        out[0] = (float)(a / (2.0 * N) - b / N + lam * gp);
        out[1] = (float)gp;
    }
}

__global__ void __launch_bounds__(256) gen_loss_kernel(const float* s, const float* mu, const float* sigma, float* out,
                                                       int N, int C) {
    __shared__ double sh[8];
    double a = 0, k = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) a += s[i];
    for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
        double m = mu[i], sg_ = sigma[i];
        k += 1.0 + log(sg_ * sg_) - m * m - sg_ * sg_;
    }
    a = block_sum_d(a, sh); k = block_sum_d(k, sh);
    if (threadIdx.x == 0) {
        out[0] = (float)(-a / N + k);
        out[1] = (float)k;
    }
}

// ---- conditioning augmentation elementwise parts
template <typename T>
__global__ void ca_reparam_kernel(const float* mu, const float* sigma, const float* eps, const float* z, float* c_hat,
                                  T* cg, int N, int C, int nz, int W) {
    // W = row length of cg >= C + nz: columns past C + nz are zero padding (the generator's first layer runs as a GEMM
    // over K = W channels, a multiple of 64)
    int64_t total = (int64_t)N * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(i / W), j = (int)(i - (int64_t)n * W);
        if (j < C) {
            int64_t k = (int64_t)n * C + j;
            float c = mu[k] + sigma[k] * eps[k];
            c_hat[k] = c;
            if (cg) stf(cg + i, c);
        } else if (cg && z && j < C + nz) {
            stf(cg + i, z[(int64_t)n * nz + (j - C)]);
        } else if (cg && j >= C + nz) {
            stf(cg + i, 0.f);
        }
    }
}
template <typename T>
__global__ void ca_bwd_seed_kernel(const T* dcg, const float* eps, const float* mu, const float* sigma, float kl,
                                   float* dmu, float* dsigma, int N, int C, int ld) {
    int64_t total = (int64_t)N * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(i / C), j = (int)(i - (int64_t)n * C);
        float dc = dcg ? ldf(dcg + (int64_t)n * ld + j) : 0.f;
        float s = sigma[i];
        dmu[i] = dc + kl * (-2.f * mu[i]);
        dsigma[i] = dc * eps[i] + kl * (2.f / s - 2.f * s);
    }
}

// 8-wide kernels of bn_fast.cu (C % 8 == 0)
template <typename T> int bn_act8(const void*, const float*, const float*, const float*, const void*, void*, int64_t, int, int, int, cudaStream_t);
template <typename T> int bn_finalize_act8(const double*, double, float*, float*, float*, long long*, int, int, float, float, const void*, const float*, const float*, const void*, void*, int64_t, int, int, int, cudaStream_t);
template <typename T> int bn_bwd_reduce8(const void*, const void*, const void*, const float*, const float*, const float*, double*, int64_t, int, int, int, cudaStream_t);
template <typename T> int bn_bwd_apply8(const void*, const void*, const void*, const float*, const float*, const float*, const double*, const void*, int, void*, int64_t, int, int, int, cudaStream_t);
template <typename T> int bn_bwd_fused8(const void*, const void*, const void*, const float*, const float*, const float*, double*, const void*, int, void*, int64_t, int, int, int, unsigned*, cudaStream_t);
template <typename T> int act_bwd8(const void*, const void*, void*, int64_t, int, cudaStream_t);
template <typename T> int gp_bn_reduce8(const void*, const void*, const void*, const void*, const float*, double*, int64_t, int, int, cudaStream_t);
template <typename T> int act_bwd8_colsum(const void*, const void*, void*, float*, int64_t, int, int, cudaStream_t);
template <typename T> int bn_finalize_act8_bulk(const double*, double, float*, float*, float*, long long*, int, int, float, float, const void*, const float*, const float*, void*, int64_t, int, int, int, cudaStream_t);
template <typename T> int gp_bn_fused8(const void*, const void*, const void*, const void*, const float*, const float*, const double*, double*, void*, void*, float*, int64_t, int, int, unsigned*, cudaStream_t);
template <typename T> int gp_bn_apply8(const void*, const void*, const void*, const void*, const float*, const float*, const double*, const double*, void*, void*, int64_t, int, int, float*, cudaStream_t);

// Several small buffers zeroed by ONE kernel node (the per-channel sums of every BatchNorm layer of a backward pass): a
// cudaMemsetAsync in front of every reduction is a graph node of its own between two dependent kernels -- ~3 us each on the
// latency-bound main chain, and it cuts the programmatic (PDL) edge between its neighbours (Stage-I: 60 per step,
// 5.30 -> 5.10 ms without them).
struct ZeroMultiArgs {
    uint32_t* p[32];
    int words[32];
};
__global__ void __launch_bounds__(256) zero_multi_kernel(ZeroMultiArgs A) {
    SG_PDL_SYNC();
    uint32_t* p = A.p[blockIdx.x];
    const int n = A.words[blockIdx.x];
    for (int i = threadIdx.x; i < n; i += 256) p[i] = 0u;
}

}  // namespace sg

using namespace sg;

// ================================================================================================ C ABI
extern "C" {

int sg_version(void) { return 1; }
const char* sg_last_error(void) { return sg::g_err; }
int64_t sg_launch_count(void) { return (int64_t)g_launches.load(); }

int sg_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return SG_ERR_NO_DEVICE; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return SG_ERR_NO_DEVICE; }
    if (p.major != 10) {
        set_error("libsgb200 is built for sm_100a (B200) only; device is sm_%d%d", p.major, p.minor);
        return SG_ERR_NO_DEVICE;
    }
    if (const char* e = getenv("SG_PDL")) sg::g_use_pdl = atoi(e);      // 0 disables programmatic dependent launch
    return 0;
}

int sg_zero_multi(void* const* ptrs, const int64_t* bytes, int n, void* stream) {
    SG_REQUIRE(n >= 1 && n <= 32, "zero_multi: 1..32 buffers per launch");
    ZeroMultiArgs A;
    for (int i = 0; i < n; ++i) {
        SG_REQUIRE(bytes[i] % 4 == 0 && bytes[i] >= 0 && bytes[i] < (1ll << 31) && ((uintptr_t)ptrs[i] & 3) == 0,
                   "zero_multi: 4-byte aligned buffers below 2 GB");
        A.p[i] = (uint32_t*)ptrs[i];
        A.words[i] = (int)(bytes[i] / 4);
    }
    launch_pdl(zero_multi_kernel, dim3(n), dim3(256), 0, SG_STREAM(stream), A);
    SG_LAUNCHED("zero_multi");
    return 0;
}

int sg_zero(void* ptr, int64_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, SG_STREAM(stream));
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int sg_fill_f32(float* ptr, float value, int64_t n, void* stream) {
    int64_t n4 = n / 4 * 4;
    if (n4) {
        FillF f{ptr, value};
        int e = launch_ew4(f, n4, SG_STREAM(stream), "fill");
        if (e) return e;
    }
    if (n4 < n) {
        fill_tail_kernel<<<1, 4, 0, SG_STREAM(stream)>>>(ptr, value, n4, n);
        SG_LAUNCHED("fill_tail");
    }
    return 0;
}

int sg_affine_f32(const float* x, float a, float b, float* out, int64_t n, void* stream) {
    affine_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, SG_STREAM(stream)>>>(x, a, b, out, n);
    SG_LAUNCHED("affine_f32");
    return 0;
}

int sg_nchw_to_nhwc(const float* src, void* dst, int N, int C, int H, int W, int dtype, void* stream) {
    int64_t HW = (int64_t)H * W;
    SG_DISPATCH_T(dtype, (nchw_to_nhwc_kernel<T><<<grid_for((int64_t)N * HW, 256), 256, 0, SG_STREAM(stream)>>>(
                             src, (T*)dst, N, C, HW)));
    SG_LAUNCHED("nchw_to_nhwc");
    return 0;
}
int sg_pack_gemm_t(const float* w, void* wt, int Co, int Ci, int kk, int Kp, int dtype, void* stream) {
    SG_REQUIRE(kk >= 1 && kk <= 16 && Kp >= Co, "pack_gemm_t: kk %d (<= 16), Kp %d >= Co %d", kk, Kp, Co);
    SG_DISPATCH_T(dtype, (pack_gemm_t_kernel<T><<<Ci, 256, 0, SG_STREAM(stream)>>>(w, (T*)wt, Co, Ci, kk, Kp)));
    SG_LAUNCHED("pack_gemm_t");
    return 0;
}
int sg_nhwc_to_nchw(const void* src, float* dst, int N, int C, int H, int W, int dtype, void* stream) {
    int64_t HW = (int64_t)H * W;
    SG_DISPATCH_T(dtype, (nhwc_to_nchw_kernel<T><<<grid_for((int64_t)N * HW, 256), 256, 0, SG_STREAM(stream)>>>(
                             (const T*)src, dst, N, C, HW)));
    SG_LAUNCHED("nhwc_to_nchw");
    return 0;
}
int sg_nhwc_to_nchw_u8(const void* src, unsigned char* dst, int N, int C, int H, int W, int dtype, void* stream) {
    int64_t HW = (int64_t)H * W;
    SG_DISPATCH_T(dtype, (nhwc_to_nchw_u8_kernel<T><<<grid_for((int64_t)N * HW, 256), 256, 0, SG_STREAM(stream)>>>(
                             (const T*)src, dst, N, C, HW)));
    SG_LAUNCHED("nhwc_to_nchw_u8");
    return 0;
}
int sg_concat_rep(const void* x, const float* c, void* out, int N, int HW, int Cx, int Cc, int dtype, void* stream) {
    int64_t rows = (int64_t)N * HW;
    SG_DISPATCH_T(dtype, (concat_rep_kernel<T><<<grid_for(rows * (Cx + Cc), 256, 16), 256, 0, SG_STREAM(stream)>>>(
                             (const T*)x, c, (T*)out, rows, HW, Cx, Cc)));
    SG_LAUNCHED("concat_rep");
    return 0;
}

int sg_split_rep_bwd(const void* dout, void* dx, float* dc, int N, int HW, int Cx, int Cc, int dtype, void* stream) {
    int64_t rows = (int64_t)N * HW;
    cudaStream_t st = SG_STREAM(stream);
    SG_DISPATCH_T(dtype, (split_copy_kernel<T><<<grid_for(rows * Cx, 256, 16), 256, 0, st>>>((const T*)dout, (T*)dx, rows, Cx,
                                                                                            Cx + Cc)));
    SG_LAUNCHED("split_copy");
    dim3 grid((Cc + 31) / 32, N);
    SG_DISPATCH_T(dtype, (rep_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)dout, dc, HW, Cx, Cc)));
    SG_LAUNCHED("rep_bwd");
    return 0;
}

int sg_patchify(const void* x, void* P, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p, int dtype,
                void* stream) {
    int64_t n = (int64_t)N * Ho * Wo * k * k;
    if (k == 4 && s == 2 && p == 1 && C == 3 && Ho * 2 == H && Wo * 2 == W && (int64_t)N * Ho * Wo * 4 < (1ll << 31)) {
        unsigned total = (unsigned)((int64_t)N * Ho * Wo * 4);
        SG_DISPATCH_T(dtype, (launch_pdl(patchify_k4s2_kernel<T, 3>, dim3(grid_for(total, 256, 16)), dim3(256), 0, SG_STREAM(stream),
                                         (const T*)x, (T*)P, H, W, Ho, Wo, total)));
        SG_LAUNCHED("patchify_k4s2");
        return 0;
    }
    SG_DISPATCH_T(dtype, (patchify_kernel<T><<<grid_for(n, 256, 16), 256, 0, SG_STREAM(stream)>>>((const T*)x, (T*)P, N, H, W, C,
                                                                                                Ho, Wo, k, s, p)));
    SG_LAUNCHED("patchify");
    return 0;
}

int sg_unpatchify(const float* col, const float* bias, void* out, int N, int Hi, int Wi, int C, int Ho, int Wo, int k, int s,
                  int p, int act, int dtype, void* stream) {
    SG_REQUIRE(C >= 1 && C <= 4, "unpatchify: 1..4 output channels");
    SG_REQUIRE(Ho == (Hi - 1) * s - 2 * p + k && Wo == (Wi - 1) * s - 2 * p + k, "unpatchify: inconsistent sizes");
    SG_REQUIRE((int64_t)N * Ho * Wo < (1ll << 31), "unpatchify: too many pixels");
    unsigned total = (unsigned)((int64_t)N * Ho * Wo);
    if (k == 4 && s == 2 && p == 1 && C == 3 && (reinterpret_cast<uintptr_t>(col) & 15) == 0) {
        const int tiles_w = (Wi + UP_TW - 1) / UP_TW, tiles_h = (Hi + UP_TH - 1) / UP_TH;
        SG_REQUIRE((int64_t)N * tiles_w * tiles_h < (1ll << 31), "unpatchify: too many tiles");
        SG_DISPATCH_T(dtype, (launch_pdl(unpatchify_k4s2p1_kernel<T, 3>, dim3((unsigned)(N * tiles_w * tiles_h)), dim3(256), 0,
                                         SG_STREAM(stream), col, bias, (T*)out, Hi, Wi, act, tiles_w, tiles_h)));
        SG_LAUNCHED("unpatchify");
        return 0;
    }
    SG_DISPATCH_T(dtype, (launch_pdl(unpatchify_kernel<T>, dim3(grid_for(total, 256, 16)), dim3(256), 0, SG_STREAM(stream), col, bias,
                                     (T*)out, Hi, Wi, Ho, Wo, C, k, s, p, act, total)));
    SG_LAUNCHED("unpatchify");
    return 0;
}

int sg_bn_fold(const float* running_mean, const float* running_var, const float* gamma, const float* beta, float eps,
               float* scale, float* shift, int C, void* stream) {
    bn_fold_kernel<<<(C + 127) / 128, 128, 0, SG_STREAM(stream)>>>(running_mean, running_var, gamma, beta, eps, scale, shift, C);
    SG_LAUNCHED("bn_fold");
    return 0;
}
int sg_pack_weight_scaled(const float* w, const float* scale, int axis, void* pf, void* pd, int Co, int Ci, int kk, int dtype,
                          void* stream) {
    SG_REQUIRE(axis == 0 || axis == 1, "pack_weight_scaled: axis 0 (Co) or 1 (Ci)");
    int64_t n = (int64_t)Co * Ci * kk;
    SG_DISPATCH_T(dtype, (pack_weight_scaled_kernel<T><<<grid_for(n, 256), 256, 0, SG_STREAM(stream)>>>(w, scale, axis, (T*)pf,
                                                                                                          (T*)pd, Co, Ci, kk)));
    SG_LAUNCHED("pack_weight_scaled");
    return 0;
}
int sg_add_act(const void* a, const void* b, void* out, int64_t n, int act, int dtype, void* stream) {
    int e = 0;
    SG_DISPATCH_T(dtype, {
        AddActF<T> f{(const T*)a, (const T*)b, (T*)out, act};
        e = launch_ew4(f, n, SG_STREAM(stream), "add_act");
    });
    return e;
}

int sg_fold_grad_cl(float* gw, float* dw, int Co, int Ci, int kk, void* stream) {
    int64_t total = (int64_t)Co * Ci * kk;
    fold_grad_cl_kernel<<<grid_for(total, 256, 8), 256, 0, SG_STREAM(stream)>>>(gw, dw, Ci, kk, total);
    SG_LAUNCHED("fold_grad_cl");
    return 0;
}

int sg_pack_weight(const float* w, void* pf, void* pd, int Co, int Ci, int kk, int dtype, void* stream) {
    if (!pf && !pd) return 0;
    int64_t n = (int64_t)Co * Ci;
    dim3 grid(grid_for(n, 256), (pf && pd) ? 2 : 1);
    SG_DISPATCH_T(dtype, (pack_weight_kernel<T><<<grid, 256, 0, SG_STREAM(stream)>>>(w, (T*)pf, (T*)pd, Co, Ci, kk,
                                                                                      pf ? 0 : 1)));
    SG_LAUNCHED("pack_weight");
    return 0;
}

int sg_colsum(const void* x, float* out, int64_t rows, int C, int dtype, void* stream) {
    int e = 0;
    SG_DISPATCH_T(dtype, {
        ColSumF<T> f{(const T*)x, C};
        e = launch_rowreduce<1>(f, out, rows, C, 1, SG_STREAM(stream), "colsum");
    });
    return e;
}

int sg_col_stats(const void* y, double* stats, int64_t rows_per_group, int C, int groups, int dtype, void* stream) {
    int e = 0;
    SG_DISPATCH_T(dtype, {
        ColStatsF<T> f{(const T*)y, C};
        e = launch_rowreduce<2>(f, stats, rows_per_group, C, groups, SG_STREAM(stream), "col_stats");
    });
    return e;
}

int sg_bn_finalize(const double* stats, int64_t count, float* mr, float* running_mean, float* running_var,
                   int64_t* nbt, int dup_first, int update_running, float momentum, float eps, int groups, int C,
                   void* stream) {
    launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, SG_STREAM(stream), stats, (double)count, mr, running_mean,
               running_var, (long long*)nbt, dup_first, update_running, momentum, eps, groups, C);
    SG_LAUNCHED("bn_finalize");
    return 0;
}
int sg_bn_eval_mr(const float* running_mean, const float* running_var, float* mr, float eps, int C, void* stream) {
    bn_eval_mr_kernel<<<(C + 127) / 128, 128, 0, SG_STREAM(stream)>>>(running_mean, running_var, mr, eps, C);
    SG_LAUNCHED("bn_eval_mr");
    return 0;
}

int sg_bn_act(const void* y, const float* mr, const float* gamma, const float* beta, const void* residual, void* out,
              int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream) {
    int64_t n = rows_per_group * C * groups;
    int e = 0;
    if (C % 8 == 0 && C <= 2048) {
        SG_DISPATCH_T(dtype, e = bn_act8<T>(y, mr, gamma, beta, residual, out, rows_per_group, C, groups, act, SG_STREAM(stream)));
        return e;
    }
    SG_DISPATCH_T(dtype, {
        BnActF<T> f{(const T*)y, mr, gamma, beta, (const T*)residual, (T*)out, rows_per_group, C, act};
        e = launch_ew4(f, n, SG_STREAM(stream), "bn_act");
    });
    return e;
}

int sg_bn_finalize_act(const double* stats, int64_t count, float* mr, float* running_mean, float* running_var, int64_t* nbt,
                       int dup_first, int update_running, float momentum, float eps, const void* y, const float* gamma,
                       const float* beta, const void* residual, void* out, int64_t rows_per_group, int C, int groups, int act,
                       int dtype, void* stream) {
    if (C % 8 == 0 && C <= 2048) {
        int e = -1;
        if (residual == nullptr)                                      // option "bn_act_bulk": the range-parked variant
            SG_DISPATCH_T(dtype, e = bn_finalize_act8_bulk<T>(stats, (double)count, mr, running_mean, running_var, (long long*)nbt,
                                                              dup_first, update_running, momentum, eps, y, gamma, beta, out,
                                                              rows_per_group, C, groups, act, SG_STREAM(stream)));
        if (e >= 0) return e;
        SG_DISPATCH_T(dtype, e = bn_finalize_act8<T>(stats, (double)count, mr, running_mean, running_var, (long long*)nbt,
                                                     dup_first, update_running, momentum, eps, y, gamma, beta, residual, out,
                                                     rows_per_group, C, groups, act, SG_STREAM(stream)));
        if (e >= 0) return e;
    }
    int e = sg_bn_finalize(stats, count, mr, running_mean, running_var, nbt, dup_first, update_running, momentum, eps, groups,
                           C, stream);
    if (e) return e;
    return sg_bn_act(y, mr, gamma, beta, residual, out, rows_per_group, C, groups, act, dtype, stream);
}

int sg_bn_bwd_reduce(const void* da, const void* a_out, const void* y, const float* mr, double* sums,
                     int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream) {
    cudaStream_t st = SG_STREAM(stream);
    if (!sg::g_dbg_skip_memset) cudaMemsetAsync(sums, 0, (size_t)groups * C * 2 * sizeof(double), st);
    int e = 0;
    if (C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH) {
        SG_DISPATCH_T(dtype, e = bn_bwd_reduce8<T>(da, a_out, y, mr, nullptr, nullptr, sums, rows_per_group, C, groups, act, st));
        return e;
    }
    SG_DISPATCH_T(dtype, {
        BnBwdReduceF<T> f{(const T*)da, (const T*)a_out, (const T*)y, mr, C, act};
        e = launch_rowreduce<2>(f, sums, rows_per_group, C, groups, st, "bn_bwd_reduce");
    });
    return e;
}

int sg_bn_bwd_apply(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma,
                    const double* sums, const void* inject, int inject_group, void* dy, int64_t rows_per_group, int C,
                    int groups, int act, int dtype, void* stream) {
    int64_t n = rows_per_group * C * groups;
    SG_REQUIRE((rows_per_group * C) % 4 == 0, "bn_bwd_apply: group size must be a multiple of 4 elements");
    int e = 0;
    if (C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH) {
        SG_DISPATCH_T(dtype, e = bn_bwd_apply8<T>(da, a_out, y, mr, gamma, nullptr, sums, inject, inject_group, dy, rows_per_group,
                                                  C, groups, act, SG_STREAM(stream)));
        return e;
    }
    SG_DISPATCH_T(dtype, {
        BnBwdApplyF<T> f{(const T*)da, (const T*)a_out, (const T*)y, mr, gamma, sums, (const T*)inject, inject_group,
                         (T*)dy, rows_per_group, C, act};
        e = launch_ew4(f, n, SG_STREAM(stream), "bn_bwd_apply");
    });
    return e;
}

// the same two kernels without the activation tensor: act'(a) is taken from the sign of gamma*xhat+beta recomputed from
// y (valid for ReLU / LeakyReLU / none directly behind the BN, i.e. every BN layer except the one closing a residual block)
int sg_bn_bwd_reduce_y(const void* da, const void* y, const float* mr, const float* gamma, const float* beta, double* sums,
                       int64_t rows_per_group, int C, int groups, int act, int dtype, void* stream) {
    SG_REQUIRE(C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH, "bn_bwd_reduce_y: C %% 8 == 0, act in {none, relu, lrelu}");
    cudaStream_t st = SG_STREAM(stream);
    if (!sg::g_dbg_skip_memset) cudaMemsetAsync(sums, 0, (size_t)groups * C * 2 * sizeof(double), st);
    int e = 0;
    SG_DISPATCH_T(dtype, e = bn_bwd_reduce8<T>(da, nullptr, y, mr, gamma, beta, sums, rows_per_group, C, groups, act, st));
    return e;
}
int sg_bn_bwd_apply_y(const void* da, const void* y, const float* mr, const float* gamma, const float* beta, const double* sums,
                      const void* inject, int inject_group, void* dy, int64_t rows_per_group, int C, int groups, int act,
                      int dtype, void* stream) {
    SG_REQUIRE(C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH, "bn_bwd_apply_y: C %% 8 == 0, act in {none, relu, lrelu}");
    int e = 0;
    SG_DISPATCH_T(dtype, e = bn_bwd_apply8<T>(da, nullptr, y, mr, gamma, beta, sums, inject, inject_group, dy, rows_per_group, C,
                                              groups, act, SG_STREAM(stream)));
    return e;
}

// BatchNorm backward as ONE call: reduce + apply.  Tensors that fit the SMs' shared memory run as a single launch
// (bn_bwd_fused8_kernel: sums, grid-wide rendezvous, apply out of shared memory); larger ones as the two kernels above.
// a_out == NULL: the activation's sign is recomputed from y (needs beta).  work: 1 KB of zeroed words owned by the call site.
int sg_bn_bwd(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma, const float* beta, double* sums,
              const void* inject, int inject_group, void* dy, int64_t rows_per_group, int C, int groups, int act, int dtype,
              int sums_zeroed, void* work, void* stream) {
    SG_REQUIRE(a_out != nullptr || beta != nullptr, "bn_bwd: a_out or beta (sign from y) is needed");
    if (C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH) {
        cudaStream_t st = SG_STREAM(stream);
        if (!sums_zeroed && !sg::g_dbg_skip_memset) cudaMemsetAsync(sums, 0, (size_t)groups * C * 2 * sizeof(double), st);
        int e = -1;
        SG_DISPATCH_T(dtype, e = bn_bwd_fused8<T>(da, a_out, y, mr, gamma, beta, sums, inject, inject_group, dy, rows_per_group, C,
                                                  groups, act, (unsigned*)work, st));
        if (e >= 0) return e;
        SG_DISPATCH_T(dtype, e = bn_bwd_reduce8<T>(da, a_out, y, mr, gamma, beta, sums, rows_per_group, C, groups, act, st));
        if (e) return e;
        SG_DISPATCH_T(dtype, e = bn_bwd_apply8<T>(da, a_out, y, mr, gamma, beta, sums, inject, inject_group, dy, rows_per_group, C,
                                                  groups, act, st));
        return e;
    }
    SG_REQUIRE(a_out != nullptr, "bn_bwd: this shape needs the activation tensor");
    int e = sg_bn_bwd_reduce(da, a_out, y, mr, sums, rows_per_group, C, groups, act, dtype, stream);
    if (e) return e;
    return sg_bn_bwd_apply(da, a_out, y, mr, gamma, sums, inject, inject_group, dy, rows_per_group, C, groups, act, dtype, stream);
}

int sg_bn_param_grad(const double* sums, float* dgamma, float* dbeta, int groups, int C, void* stream) {
    launch_pdl(bn_param_grad_kernel, dim3((C + 127) / 128), dim3(128), 0, SG_STREAM(stream), sums, dgamma, dbeta, groups, C);
    SG_LAUNCHED("bn_param_grad");
    return 0;
}

int sg_bn_param_grad_multi(const double* const* sums, float* const* dgamma, float* const* dbeta, const int* groups, const int* C,
                           int n_layers, void* stream) {
    SG_REQUIRE(n_layers >= 1 && n_layers <= BN_PG_MAX, "bn_param_grad_multi: 1..%d layers per launch", BN_PG_MAX);
    BnParamGradArgs A;
    int cmax = 0;
    for (int l = 0; l < n_layers; ++l) {
        A.sums[l] = sums[l]; A.dgamma[l] = dgamma[l]; A.dbeta[l] = dbeta[l]; A.G[l] = groups[l]; A.C[l] = C[l];
        if (C[l] > cmax) cmax = C[l];
    }
    launch_pdl(bn_param_grad_multi_kernel, dim3((cmax + 127) / 128, n_layers), dim3(128), 0, SG_STREAM(stream), A);
    SG_LAUNCHED("bn_param_grad_multi");
    return 0;
}

int sg_act_bwd(const void* da, const void* a_out, void* out, int64_t n, int act, int dtype, void* stream) {
    int e = 0;
    if (n % 8 == 0) {
        SG_DISPATCH_T(dtype, e = act_bwd8<T>(da, a_out, out, n, act, SG_STREAM(stream)));
        return e;
    }
    SG_DISPATCH_T(dtype, {
        ActBwdF<T> f{(const T*)da, (const T*)a_out, (T*)out, act};
        e = launch_ew4(f, n, SG_STREAM(stream), "act_bwd");
    });
    return e;
}

// out = da * act'(a_out) and colsum[c] += column sums of out ([rows][C] tensors; the bias gradient of a conv + bias + activation
// layer, discrminator_1.py:17-18) in ONE pass.  C % 8 == 0, act in {none, relu, lrelu}; otherwise sg_act_bwd + sg_colsum.
int sg_act_bwd_colsum(const void* da, const void* a_out, void* out, float* colsum, int64_t rows, int C, int act, int dtype,
                      void* stream) {
    if (C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH) {
        int e = 0;
        SG_DISPATCH_T(dtype, e = act_bwd8_colsum<T>(da, a_out, out, colsum, rows, C, act, SG_STREAM(stream)));
        return e;
    }
    int e = sg_act_bwd(da, a_out, out, rows * C, act, dtype, stream);
    if (e) return e;
    return sg_colsum(out, colsum, rows, C, dtype, stream);
}

static int gp_bn_reduce_impl(const void* v, const void* da, const void* a_out, const void* y, const float* mr, double* tsums,
                             int64_t rows, int C, int act, int dtype, void* stream, bool zero) {
    cudaStream_t st = SG_STREAM(stream);
    if (zero && !sg::g_dbg_skip_memset) cudaMemsetAsync(tsums, 0, (size_t)C * 3 * sizeof(double), st);
    int e = 0;
    if (C % 8 == 0 && C <= 2048 && act != SG_ACT_TANH) {
        SG_DISPATCH_T(dtype, e = gp_bn_reduce8<T>(v, da, a_out, y, mr, tsums, rows, C, act, st));
        return e;
    }
    SG_DISPATCH_T(dtype, {
        GpBnReduceF<T> f{(const T*)v, (const T*)da, (const T*)a_out, (const T*)y, mr, C, act};
        e = launch_rowreduce<3>(f, tsums, rows, C, 1, st, "gp_bn_reduce");
    });
    return e;
}

int sg_gp_bn_reduce(const void* v, const void* da, const void* a_out, const void* y, const float* mr, double* tsums,
                    int64_t rows, int C, int act, int dtype, void* stream) {
    return gp_bn_reduce_impl(v, da, a_out, y, mr, tsums, rows, C, act, dtype, stream, true);
}
// the same, ADDING to tsums: the caller zeroed them (sg_zero_multi at the start of the pass) -- no memset node in the chain
int sg_gp_bn_reduce_acc(const void* v, const void* da, const void* a_out, const void* y, const float* mr, double* tsums,
                        int64_t rows, int C, int act, int dtype, void* stream) {
    return gp_bn_reduce_impl(v, da, a_out, y, mr, tsums, rows, C, act, dtype, stream, false);
}

int sg_gp_bn_apply(const void* v, const void* da, const void* a_out, const void* y, const float* mr,
                   const float* gamma, const double* sums, const double* tsums, void* w_out, void* gy_out,
                   float* dgamma, int64_t rows, int C, int act, int dtype, void* stream) {
    int e = 0;
    if (C % 8 == 0 && C <= 1024 && act != SG_ACT_TANH) {
        // (the gamma gradient of the penalty term rides in CTA 0's prologue: it needs the same per-channel sums)
        SG_DISPATCH_T(dtype, e = gp_bn_apply8<T>(v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, rows, C, act, dgamma,
                                                 SG_STREAM(stream)));
        return e;
    }
    SG_DISPATCH_T(dtype, {
        GpBnApplyF<T> f{(const T*)v, (const T*)da, (const T*)a_out, (const T*)y, mr, gamma, sums, tsums,
                        (T*)w_out, (T*)gy_out, rows, C, act};
        e = launch_ew4(f, rows * C, SG_STREAM(stream), "gp_bn_apply");
    });
    if (e) return e;
    gp_bn_dgamma_kernel<<<(C + 127) / 128, 128, 0, SG_STREAM(stream)>>>(mr, sums, tsums, dgamma, (double)rows, C);
    SG_LAUNCHED("gp_bn_dgamma");
    return 0;
}

// sg_gp_bn_reduce + sg_gp_bn_apply behind ONE call; one launch (gp_bn_fused8_kernel) where the four tensors fit the SMs' shared
// memory.  tsums_zeroed: the caller zeroed tsums (sg_zero_multi).  work: 1 KB of zeroed words owned by the call site, NULL = two kernels.
int sg_gp_bn(const void* v, const void* da, const void* a_out, const void* y, const float* mr, const float* gamma, const double* sums,
             double* tsums, void* w_out, void* gy_out, float* dgamma, int64_t rows, int C, int act, int dtype, int tsums_zeroed,
             void* work, void* stream) {
    if (C % 8 == 0 && C <= 1024 && act != SG_ACT_TANH && work != nullptr) {
        cudaStream_t st = SG_STREAM(stream);
        if (!tsums_zeroed && !sg::g_dbg_skip_memset) cudaMemsetAsync(tsums, 0, (size_t)C * 3 * sizeof(double), st);
        int e = -1;
        SG_DISPATCH_T(dtype, e = gp_bn_fused8<T>(v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, rows, C, act,
                                                 (unsigned*)work, st));
        if (e >= 0) return e;
        tsums_zeroed = 1;                               // (zeroed just above)
    }
    int e = gp_bn_reduce_impl(v, da, a_out, y, mr, tsums, rows, C, act, dtype, stream, !tsums_zeroed);
    if (e) return e;
    return sg_gp_bn_apply(v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, rows, C, act, dtype, stream);
}

int sg_outer(const float* coef, const float* vec, void* out, int N, int M, int out_dtype, void* stream) {
    SG_REQUIRE(M % 4 == 0, "outer: M %% 4 != 0");
    int e = 0;
    SG_DISPATCH_T(out_dtype, {
        OuterF<T> f{coef, vec, (T*)out, M};
        e = launch_ew4(f, (int64_t)N * M, SG_STREAM(stream), "outer");
    });
    return e;
}

int sg_ca_reparam(const float* mu, const float* sigma, const float* eps, const float* z, float* c_hat, void* cg, int N,
                  int C, int nz, int ld, int dtype, void* stream) {
    SG_REQUIRE(ld >= C + nz, "ca_reparam: row length %d < %d + %d", ld, C, nz);
    int64_t n = (int64_t)N * ld;
    SG_DISPATCH_T(dtype, (ca_reparam_kernel<T><<<grid_for(n, 256), 256, 0, SG_STREAM(stream)>>>(mu, sigma, eps, z, c_hat,
                                                                                               (T*)cg, N, C, nz, ld)));
    SG_LAUNCHED("ca_reparam");
    return 0;
}
int sg_ca_bwd_seed(const void* dcg, const float* eps, const float* mu, const float* sigma, float kl_scale, float* dmu,
                   float* dsigma, int N, int C, int ld, int dtype, void* stream) {
    int64_t n = (int64_t)N * C;
    SG_DISPATCH_T(dtype, (ca_bwd_seed_kernel<T><<<grid_for(n, 256), 256, 0, SG_STREAM(stream)>>>(
                             (const T*)dcg, eps, mu, sigma, kl_scale, dmu, dsigma, N, C, ld)));
    SG_LAUNCHED("ca_bwd_seed");
    return 0;
}

int sg_interp(const void* real, const void* fake, const float* eps, void* out, int N, int64_t per_sample, int dtype,
              void* stream) {
    SG_REQUIRE(per_sample % 4 == 0, "interp: per_sample %% 4 != 0");
    int e = 0;
    SG_DISPATCH_T(dtype, {
        InterpF<T> f{(const T*)real, (const T*)fake, eps, (T*)out, per_sample};
        e = launch_ew4(f, (int64_t)N * per_sample, SG_STREAM(stream), "interp");
    });
    return e;
}

static int sample_sqnorm_impl(const void* g, float* out, int N, int64_t per_sample, int dtype, void* stream, bool zero);
int sg_sample_sqnorm(const void* g, float* out, int N, int64_t per_sample, int dtype, void* stream) {
    return sample_sqnorm_impl(g, out, N, per_sample, dtype, stream, true);
}
// the same, ADDING to out: the caller zeroed it (sg_zero_multi)
int sg_sample_sqnorm_acc(const void* g, float* out, int N, int64_t per_sample, int dtype, void* stream) {
    return sample_sqnorm_impl(g, out, N, per_sample, dtype, stream, false);
}
static int sample_sqnorm_impl(const void* g, float* out, int N, int64_t per_sample, int dtype, void* stream, bool zero) {
    SG_REQUIRE(per_sample % 4 == 0, "sample_sqnorm: per_sample %% 4 != 0");
    cudaStream_t st = SG_STREAM(stream);
    if (zero && !sg::g_dbg_skip_memset) cudaMemsetAsync(out, 0, N * sizeof(float), st);
    int64_t per_block = 256 * 4 * 4;
    int bx = (int)((per_sample + per_block - 1) / per_block);
    dim3 grid(bx, N);
    SG_DISPATCH_T(dtype, (sample_sqnorm_kernel<T><<<grid, 256, 0, st>>>((const T*)g, out, per_sample, per_block)));
    SG_LAUNCHED("sample_sqnorm");
    return 0;
}

int sg_gp_seed(const void* g, const float* sq, float coef, void* v, int N, int64_t per_sample, int dtype, void* stream) {
    SG_REQUIRE(per_sample % 4 == 0, "gp_seed: per_sample %% 4 != 0");
    int e = 0;
    SG_DISPATCH_T(dtype, {
        GpSeedF<T> f{(const T*)g, sq, coef, (T*)v, per_sample};
        e = launch_ew4(f, (int64_t)N * per_sample, SG_STREAM(stream), "gp_seed");
    });
    return e;
}

int sg_scale_rows_add(const void* x, const float* scale, void* out, int accumulate, int N, int64_t per_sample, int dtype,
                      void* stream) {
    SG_REQUIRE(per_sample % 4 == 0, "scale_rows_add: per_sample %% 4 != 0");
    int e = 0;
    SG_DISPATCH_T(dtype, {
        ScaleRowsAddF<T> f{(const T*)x, scale, (T*)out, accumulate, per_sample};
        e = launch_ew4(f, (int64_t)N * per_sample, SG_STREAM(stream), "scale_rows_add");
    });
    return e;
}

int sg_critic_loss(const float* s_real, const float* s_mis, const float* s_fake, const float* sq, float lam, float* out2,
                   int N, void* stream) {
    critic_loss_kernel<<<1, 256, 0, SG_STREAM(stream)>>>(s_real, s_mis, s_fake, sq, lam, out2, N);
    SG_LAUNCHED("critic_loss");
    return 0;
}
int sg_gen_loss(const float* s_fake, const float* mu, const float* sigma, float* out2, int N, int C, void* stream) {
    gen_loss_kernel<<<1, 256, 0, SG_STREAM(stream)>>>(s_fake, mu, sigma, out2, N, C);
    SG_LAUNCHED("gen_loss");
    return 0;
}

int sg_adam_step(float* p, const float* g, float* m, float* v, float* hyper, int64_t n, void* stream) {
    SG_REQUIRE(n % 4 == 0, "adam: n %% 4 != 0 (pad the flat buffer)");
    adam_tick_kernel<<<1, 1, 0, SG_STREAM(stream)>>>(hyper);
    SG_LAUNCHED("adam_tick");
    AdamF f{p, g, m, v, hyper};
    return launch_ew4(f, n, SG_STREAM(stream), "adam");
}

}  // extern "C"
