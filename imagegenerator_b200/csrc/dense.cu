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

// dw[m][k] += sum_n d[n][m] x[n][k], db[m] += sum_n d[n][m]   (d = dout masked by relu_out > 0)
// grid (ceil(K/256), M, splits): a block owns one output row m and a slice of the batch; the loop is unrolled so that
// the loads of 8 samples are in flight together (one dependent L2 round trip per sample made this kernel 35 us for
// 17 MFLOP); slices combine with atomics (the result is += anyway).
template <bool RELU>
__global__ void __launch_bounds__(256) linear_dw_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                        const float* __restrict__ relu_out, float* __restrict__ dw,
                                                        float* __restrict__ db, int N, int K, int M, int n_per_split) {
    const int k = blockIdx.x * 256 + threadIdx.x, m = blockIdx.y;
    const int n0 = blockIdx.z * n_per_split, n1 = min(N, n0 + n_per_split);
    const bool live = k < K;
    const float* xp = x + (live ? k : 0);
    float acc = 0.f, dsum = 0.f;
#pragma unroll 8
    for (int n = n0; n < n1; ++n) {
        float d = dout[(int64_t)n * M + m];
        if (RELU) d = relu_out[(int64_t)n * M + m] > 0.f ? d : 0.f;
        acc += d * xp[(int64_t)n * K];
        dsum += d;
    }
    if (live && dw) atomicAdd(dw + (int64_t)m * K + k, acc);
    if (db && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(db + m, dsum);
}
// dx[n][k] (+)= sum_m d[n][m] w[m][k]; grid (ceil(K/256), N): the masked row d[n][:] is staged in shared memory once
template <bool RELU>
__global__ void __launch_bounds__(256) linear_dx_kernel(const float* __restrict__ w, const float* __restrict__ dout,
                                                        const float* __restrict__ relu_out, float* __restrict__ dx,
                                                        int acc_flag, int N, int K, int M) {
    extern __shared__ float drow[];
    const int n = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    for (int m = threadIdx.x; m < M; m += 256) {
        float d = dout[(int64_t)n * M + m];
        if (RELU) d = relu_out[(int64_t)n * M + m] > 0.f ? d : 0.f;
        drow[m] = d;
    }
    __syncthreads();
    if (k >= K) return;
    const float* wp = w + k;
    float acc = 0.f;
#pragma unroll 8
    for (int m = 0; m < M; ++m) acc += drow[m] * wp[(int64_t)m * K];
    float* o = dx + (int64_t)n * K + k;
    *o = acc_flag ? *o + acc : acc;
}

// ---- head: A[hw][c] = sum_k wcs[k][hw] wcr[k][c]; Bv[j] = sum_k sw[k] wcr[k][Cx+j]; c0 = sum_k sw[k] bcr[k] + bcs,
// sw[k] = sum_hw wcs[k][hw].  Blocks [0, gridDim.x-1) compute A (thread per element, K independent coalesced loads);
// the last block computes sw in shared memory, then Bv and c0.
constexpr int HEAD_MAX_K = 1024;
__global__ void __launch_bounds__(256) head_prepare_kernel(const float* __restrict__ wcr, const float* __restrict__ bcr,
                                                           const float* __restrict__ wcs, const float* __restrict__ bcs,
                                                           float* __restrict__ A, float* __restrict__ Bv,
                                                           float* __restrict__ c0, int K, int Cx, int Nd) {
    const int W = Cx + Nd;
    if (blockIdx.x + 1 < gridDim.x) {
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i >= 16 * Cx) return;
        const int hw = i / Cx, c = i - hw * Cx;
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < K; ++k) acc += wcs[k * 16 + hw] * wcr[(int64_t)k * W + c];
        A[i] = acc;
        return;
    }
    __shared__ float sw[HEAD_MAX_K];
    __shared__ float red[8];
    float part = 0.f;
    for (int k = threadIdx.x; k < K; k += 256) {
        float v = 0.f;          // (the parameter may sit at any 4-byte offset of the flat buffer: no vector loads)
#pragma unroll
        for (int h = 0; h < 16; ++h) v += wcs[k * 16 + h];
        sw[k] = v;
        part += v * bcr[k];
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = bcs[0];
        for (int i = 0; i < 8; ++i) t += red[i];
        c0[0] = t;
    }
    for (int j = threadIdx.x; j < Nd; j += 256) {
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < K; ++k) acc += sw[k] * wcr[(int64_t)k * W + Cx + j];
        Bv[j] = acc;
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

// several head_fwd calls of one critic forward (real / mismatched / fake / interpolated rows) in one launch:
// job j scores rows a4[a_row0[j] + n] with text rows ce[ce_row0[j] + n] into score[s_off[j] + n]
struct HeadJobs {
    int a_row0[4], ce_row0[4], s_off[4];
};
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_multi_kernel(const T* __restrict__ a4, const float* __restrict__ ce,
                                                             const float* __restrict__ A, const float* __restrict__ Bv,
                                                             const float* __restrict__ c0, float* __restrict__ score,
                                                             HeadJobs jobs, int M, int Nd) {
    const int n = blockIdx.x, j = blockIdx.y;
    const T* a = a4 + (int64_t)(jobs.a_row0[j] + n) * M;
    const float* cr = ce + (int64_t)(jobs.ce_row0[j] + n) * Nd;
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < M; i += 256 * 4) {
        F4 v = ld4(a + i);
        float4 w = *reinterpret_cast<const float4*>(A + i);
        acc += v.v[0] * w.x + v.v[1] * w.y + v.v[2] * w.z + v.v[3] * w.w;
    }
    for (int q = threadIdx.x; q < Nd; q += 256) acc += cr[q] * Bv[q];
    __shared__ float sh[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = c0[0];
        for (int i = 0; i < 8; ++i) t += sh[i];
        score[jobs.s_off[j] + n] = t;
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

// Parameter gradients of the collapsed head.  Block ranges: [0, nb1) dwcr (thread per element, 16 terms);
// [nb1, nb1+nb2) dwcs (a warp per (k, hw): the 640-term dot products are read coalesced and reduced by shuffles -- a
// thread per element walked them serially, 49 us); the last block dbcr / dbcs.
__global__ void __launch_bounds__(256) head_param_grads_kernel(const float* __restrict__ dA, const float* __restrict__ dBv,
                                                               const float* __restrict__ dc0, const float* __restrict__ wcr,
                                                               const float* __restrict__ bcr, const float* __restrict__ wcs,
                                                               float* dwcr, float* dbcr, float* dwcs, float* dbcs, int K,
                                                               int Cx, int Nd, int nb1, int nb2) {
    const int W = Cx + Nd;
    const float d0 = dc0[0];
    int blk = blockIdx.x;
    if (blk < nb1) {                   // dwcr[k][c]
        const int64_t i = (int64_t)blk * 256 + threadIdx.x;
        if (i >= (int64_t)K * W) return;
        const int k = (int)(i / W), c = (int)(i - (int64_t)k * W);
        float acc = 0.f;
        if (c < Cx) {
#pragma unroll
            for (int h = 0; h < 16; ++h) acc += wcs[k * 16 + h] * dA[h * Cx + c];
        } else {
            float sw = 0.f;
#pragma unroll
            for (int h = 0; h < 16; ++h) sw += wcs[k * 16 + h];
            acc = sw * dBv[c - Cx];
        }
        dwcr[i] += acc;
        return;
    }
    blk -= nb1;
    if (blk < nb2) {                   // dwcs[k][hw]
        const int r = blk * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        if (r >= K * 16) return;
        const int k = r >> 4, h = r & 15;
        const float* wr = wcr + (int64_t)k * W;
        float acc = 0.f;
        for (int c = lane; c < Cx; c += 32) acc += wr[c] * dA[h * Cx + c];
        for (int j = lane; j < Nd; j += 32) acc += wr[Cx + j] * dBv[j];
        acc = warp_sum(acc);
        if (lane == 0) dwcs[r] += acc + bcr[k] * d0;
        return;
    }
    for (int k = threadIdx.x; k < K; k += 256) {   // dbcr[k]
        float sw = 0.f;
#pragma unroll
        for (int h = 0; h < 16; ++h) sw += wcs[k * 16 + h];
        dbcr[k] += sw * d0;
    }
    if (threadIdx.x == 0) dbcs[0] += d0;
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
    if (dw || db) {
        int splits = (N + 31) / 32;
        if (splits > 8) splits = 8;
        if (splits < 1) splits = 1;
        int nps = (N + splits - 1) / splits;
        splits = (N + nps - 1) / nps;
        dim3 grid(dw ? (K + 255) / 256 : 1, M, splits);
        if (relu_out)
            linear_dw_kernel<true><<<grid, 256, 0, st>>>(x, dout, relu_out, dw, db, N, dw ? K : 0, M, nps);
        else
            linear_dw_kernel<false><<<grid, 256, 0, st>>>(x, dout, relu_out, dw, db, N, dw ? K : 0, M, nps);
        SG_LAUNCHED("linear_dw");
    }
    if (dx) {
        dim3 grid((K + 255) / 256, N);
        size_t smem = (size_t)M * sizeof(float);
        SG_REQUIRE(smem <= 48 * 1024, "linear_bwd: M too large for the shared-memory row");
        if (relu_out)
            linear_dx_kernel<true><<<grid, 256, smem, st>>>(w, dout, relu_out, dx, dx_acc, N, K, M);
        else
            linear_dx_kernel<false><<<grid, 256, smem, st>>>(w, dout, relu_out, dx, dx_acc, N, K, M);
        SG_LAUNCHED("linear_dx");
    }
    return 0;
}

int sg_head_prepare(const float* wcr, const float* bcr, const float* wcs, const float* bcs, float* A, float* Bv,
                    float* c0, int K, int Cx, int Nd, void* stream) {
    SG_REQUIRE(K <= HEAD_MAX_K, "head_prepare: K > %d", HEAD_MAX_K);
    int nbA = (16 * Cx + 255) / 256;
    head_prepare_kernel<<<nbA + 1, 256, 0, SG_STREAM(stream)>>>(wcr, bcr, wcs, bcs, A, Bv, c0, K, Cx, Nd);
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

int sg_head_fwd_multi(const void* a4, const float* ce, const float* A, const float* Bv, const float* c0, float* score,
                      int n_jobs, const int* a_row0, const int* ce_row0, const int* score_off, int N, int M, int Nd, int dtype,
                      void* stream) {
    SG_REQUIRE(M % 4 == 0, "head_fwd_multi: M %% 4 != 0");
    SG_REQUIRE(n_jobs >= 1 && n_jobs <= 4, "head_fwd_multi: 1..4 jobs");
    HeadJobs jobs{};
    for (int j = 0; j < n_jobs; ++j) { jobs.a_row0[j] = a_row0[j]; jobs.ce_row0[j] = ce_row0[j]; jobs.s_off[j] = score_off[j]; }
    SG_DISPATCH_T(dtype, (head_fwd_multi_kernel<T><<<dim3(N, n_jobs), 256, 0, SG_STREAM(stream)>>>((const T*)a4, ce, A, Bv, c0,
                                                                                                   score, jobs, M, Nd)));
    SG_LAUNCHED("head_fwd_multi");
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
    int nb1 = (int)(((int64_t)K * (Cx + Nd) + 255) / 256), nb2 = (K * 16 + 7) / 8;
    head_param_grads_kernel<<<nb1 + nb2 + 1, 256, 0, SG_STREAM(stream)>>>(dA, dBv, dc0, wcr, bcr, wcs, dwcr, dbcr, dwcs, dbcs,
                                                                         K, Cx, Nd, nb1, nb2);
    SG_LAUNCHED("head_param_grads");
    return 0;
}

}  // extern "C"
