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

// ---- conditioning augmentation, fused (reference con_augment.py:13-22 + the concat of stage_1_train_fn.py:120-122)
// forward: tem -> h = relu(Wh tem + bh) -> (mu, sigma) -> c_hat = mu + sigma * eps -> row [c_hat, z, 0...] of the generator's
// input, ONE launch, one CTA per sample: the 197 k parameters stream through each CTA once (L2 hits after the first CTA),
// nothing but the tensors the backward needs (h, mu, sigma, c_hat) goes back to memory.  A warp computes four outputs at a
// time so that 16 independent 16-byte weight loads per lane are in flight (the chain is latency-bound, not FLOP-bound).
template <typename T>
__global__ void __launch_bounds__(256) ca_forward_kernel(const float* __restrict__ tem, const float* __restrict__ Wh,
                                                         const float* __restrict__ bh, const float* __restrict__ Wmu,
                                                         const float* __restrict__ bmu, const float* __restrict__ Wsg,
                                                         const float* __restrict__ bsg, const float* __restrict__ eps,
                                                         const float* __restrict__ z, float* __restrict__ h,
                                                         float* __restrict__ mu, float* __restrict__ sigma,
                                                         float* __restrict__ c_hat, T* __restrict__ cg, int Tm, int Hd, int C,
                                                         int nz, int ld) {
    extern __shared__ float ca_sm[];
    float* xs = ca_sm;                 // [Tm]
    float* hs = xs + Tm;               // [Hd]
    float* ms = hs + Hd;               // [2C]: mu, sigma
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < Tm; i += 256) xs[i] = tem[(int64_t)n * Tm + i];
    __syncthreads();
    auto dot4 = [&](const float* w0, const float* w1, const float* w2, const float* w3, const float* xv, int K, float (&acc)[4]) {
        const float4* x4 = reinterpret_cast<const float4*>(xv);
        const float4 *a = reinterpret_cast<const float4*>(w0), *b = reinterpret_cast<const float4*>(w1),
                     *c = reinterpret_cast<const float4*>(w2), *d = reinterpret_cast<const float4*>(w3);
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        for (int k = lane; k < (K >> 2); k += 32) {
            const float4 xx = x4[k], wa = __ldg(a + k), wb = __ldg(b + k), wc = __ldg(c + k), wd = __ldg(d + k);
            acc[0] += wa.x * xx.x + wa.y * xx.y + wa.z * xx.z + wa.w * xx.w;
            acc[1] += wb.x * xx.x + wb.y * xx.y + wb.z * xx.z + wb.w * xx.w;
            acc[2] += wc.x * xx.x + wc.y * xx.y + wc.z * xx.z + wc.w * xx.w;
            acc[3] += wd.x * xx.x + wd.y * xx.y + wd.z * xx.z + wd.w * xx.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
    };
    for (int m0 = warp * 4; m0 < Hd; m0 += 32) {
        float acc[4];
        dot4(Wh + (int64_t)m0 * Tm, Wh + (int64_t)(m0 + 1) * Tm, Wh + (int64_t)(m0 + 2) * Tm, Wh + (int64_t)(m0 + 3) * Tm, xs, Tm, acc);
        if (lane < 4) {
            const float sel = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            const float v = fmaxf(sel + bh[m0 + lane], 0.f);
            hs[m0 + lane] = v;
            h[(int64_t)n * Hd + m0 + lane] = v;
        }
    }
    __syncthreads();
    for (int m0 = warp * 4; m0 < 2 * C; m0 += 32) {
        const bool sg = m0 >= C;
        const float* W = sg ? Wsg : Wmu;
        const int r = sg ? m0 - C : m0;
        float acc[4];
        dot4(W + (int64_t)r * Hd, W + (int64_t)(r + 1) * Hd, W + (int64_t)(r + 2) * Hd, W + (int64_t)(r + 3) * Hd, hs, Hd, acc);
        if (lane < 4) {
            const float sel = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            const float v = sel + (sg ? bsg : bmu)[r + lane];
            ms[m0 + lane] = v;
            (sg ? sigma : mu)[(int64_t)n * C + r + lane] = v;
        }
    }
    if (eps == nullptr) return;
    __syncthreads();
    for (int j = tid; j < ld; j += 256) {
        if (j < C) {
            const float c = ms[j] + ms[C + j] * eps[(int64_t)n * C + j];
            c_hat[(int64_t)n * C + j] = c;
            if (cg) stf(cg + (int64_t)n * ld + j, c);
        } else if (cg) {
            stf(cg + (int64_t)n * ld + j, (z != nullptr && j < C + nz) ? z[(int64_t)n * nz + (j - C)] : 0.f);
        }
    }
}

// backward, data part, one CTA per sample:
//   dmu = dc + kl * (-2 mu), dsigma = dc * eps + kl * (2/sigma - 2 sigma)         (dc = d loss / d c_hat, row n of dcg)
//   dh  = relu'(h) * (Wmu^T dmu + Wsigma^T dsigma)                                  [stored MASKED]
//   dtem (+)= Wh^T dh                                                               (optional)
template <typename T>
__global__ void __launch_bounds__(256) ca_backward_data_kernel(const T* __restrict__ dcg, const float* __restrict__ eps,
                                                               const float* __restrict__ mu, const float* __restrict__ sigma,
                                                               float kl, const float* __restrict__ h,
                                                               const float* __restrict__ Wmu, const float* __restrict__ Wsg,
                                                               const float* __restrict__ Wh, float* __restrict__ dmu,
                                                               float* __restrict__ dsigma, float* __restrict__ dh,
                                                               float* __restrict__ dtem, int dtem_acc, int Tm, int Hd, int C,
                                                               int ld) {
    extern __shared__ float ca_sm[];
    float* dm = ca_sm;                 // [2C]: dmu, dsigma
    float* dhs = dm + 2 * C;           // [Hd]
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int j = tid; j < C; j += 256) {
        const int64_t i = (int64_t)n * C + j;
        const float dc = dcg ? ldf(dcg + (int64_t)n * ld + j) : 0.f;
        const float s = sigma[i];
        const float a = dc + kl * (-2.f * mu[i]), b = dc * eps[i] + kl * (2.f / s - 2.f * s);
        dm[j] = a; dm[C + j] = b;
        dmu[i] = a; dsigma[i] = b;
    }
    __syncthreads();
    for (int k = tid; k < Hd; k += 256) {
        float acc = 0.f;
#pragma unroll 8
        for (int j = 0; j < C; ++j) acc += dm[j] * __ldg(Wmu + (int64_t)j * Hd + k) + dm[C + j] * __ldg(Wsg + (int64_t)j * Hd + k);
        acc = h[(int64_t)n * Hd + k] > 0.f ? acc : 0.f;
        dhs[k] = acc;
        dh[(int64_t)n * Hd + k] = acc;
    }
    if (dtem == nullptr) return;
    __syncthreads();
    for (int k = tid; k < Tm; k += 256) {
        float acc = 0.f;
#pragma unroll 8
        for (int m = 0; m < Hd; ++m) acc += dhs[m] * __ldg(Wh + (int64_t)m * Tm + k);
        float* o = dtem + (int64_t)n * Tm + k;
        *o = dtem_acc ? *o + acc : acc;
    }
}

// backward, parameter part: the three weight / bias gradients in ONE launch.  grid (ceil(Tm/256), 2C + Hd, splits): row
// r < C: mu layer (x = h, d = dmu); r < 2C: sigma layer; else the hidden layer (x = tem, d = masked dh).
__global__ void __launch_bounds__(256) ca_backward_params_kernel(const float* __restrict__ tem, const float* __restrict__ h,
                                                                 const float* __restrict__ dmu, const float* __restrict__ dsigma,
                                                                 const float* __restrict__ dh, float* __restrict__ gWmu,
                                                                 float* __restrict__ gbmu, float* __restrict__ gWsg,
                                                                 float* __restrict__ gbsg, float* __restrict__ gWh,
                                                                 float* __restrict__ gbh, int N, int Tm, int Hd, int C,
                                                                 int n_per_split) {
    const int r = blockIdx.y;
    const float *x, *d;
    float *gw, *gb;
    int K, M, m;
    if (r < C) { x = h; d = dmu; gw = gWmu; gb = gbmu; K = Hd; M = C; m = r; }
    else if (r < 2 * C) { x = h; d = dsigma; gw = gWsg; gb = gbsg; K = Hd; M = C; m = r - C; }
    else { x = tem; d = dh; gw = gWh; gb = gbh; K = Tm; M = Hd; m = r - 2 * C; }
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x * 256 >= K) return;
    const int n0 = blockIdx.z * n_per_split, n1 = min(N, n0 + n_per_split);
    const bool live = k < K;
    const float* xp = x + (live ? k : 0);
    float acc = 0.f, dsum = 0.f;
#pragma unroll 8
    for (int n = n0; n < n1; ++n) {
        const float dv = d[(int64_t)n * M + m];
        acc += dv * xp[(int64_t)n * K];
        dsum += dv;
    }
    if (live) atomicAdd(gw + (int64_t)m * K + k, acc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(gb + m, dsum);
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

// Conditioning augmentation, fused (con_augment.py:13-22 + the [c_hat, z] concat of stage_1_train_fn.py:120-122): see the
// kernels above.  eps NULL = encode only (h, mu, sigma); z / cg NULL = no generator input row.
int sg_ca_forward(const float* tem, const float* Wh, const float* bh, const float* Wmu, const float* bmu, const float* Wsg,
                  const float* bsg, const float* eps, const float* z, float* h, float* mu, float* sigma, float* c_hat, void* cg,
                  int N, int Tm, int Hd, int C, int nz, int ld, int dtype, void* stream) {
    SG_REQUIRE(Tm % 4 == 0 && Hd % 4 == 0 && C % 4 == 0, "ca_forward: sizes must be multiples of 4");
    SG_REQUIRE(cg == nullptr || ld >= C + nz, "ca_forward: row length %d < %d + %d", ld, C, nz);
    const size_t smem = (size_t)(Tm + Hd + 2 * C) * sizeof(float);
    SG_REQUIRE(smem <= 48 * 1024, "ca_forward: sizes exceed the shared-memory staging");
    SG_DISPATCH_T(dtype, (ca_forward_kernel<T><<<N, 256, smem, SG_STREAM(stream)>>>(tem, Wh, bh, Wmu, bmu, Wsg, bsg, eps, z, h, mu,
                                                                                    sigma, c_hat, (T*)cg, Tm, Hd, C, nz,
                                                                                    cg ? ld : C)));
    SG_LAUNCHED("ca_forward");
    return 0;
}

// d loss / d(mu, sigma, h, tem) and the six parameter gradients (accumulated) in two launches.  dcg: gradient of the
// generator input rows (T, row length ld; only its first C columns are read) or NULL; kl_scale multiplies the gradient of
// sum(1 + log sigma^2 - mu^2 - sigma^2) (stage_1_train_fn.py:156-159); dtem NULL = the text side is frozen.
int sg_ca_backward(const void* dcg, const float* eps, const float* mu, const float* sigma, float kl_scale, const float* h,
                   const float* tem, const float* Wmu, const float* Wsg, const float* Wh, float* dmu, float* dsigma, float* dh,
                   float* gWmu, float* gbmu, float* gWsg, float* gbsg, float* gWh, float* gbh, float* dtem, int dtem_acc, int N,
                   int Tm, int Hd, int C, int ld, int dtype, void* stream) {
    cudaStream_t st = SG_STREAM(stream);
    const size_t smem = (size_t)(2 * C + Hd) * sizeof(float);
    SG_REQUIRE(smem <= 48 * 1024, "ca_backward: sizes exceed the shared-memory staging");
    SG_DISPATCH_T(dtype, (ca_backward_data_kernel<T><<<N, 256, smem, st>>>((const T*)dcg, eps, mu, sigma, kl_scale, h, Wmu, Wsg, Wh,
                                                                            dmu, dsigma, dh, dtem, dtem_acc, Tm, Hd, C, ld)));
    SG_LAUNCHED("ca_backward_data");
    int splits = (N + 31) / 32;
    if (splits > 8) splits = 8;
    if (splits < 1) splits = 1;
    const int nps = (N + splits - 1) / splits;
    splits = (N + nps - 1) / nps;
    const int kmax = Tm > Hd ? Tm : Hd;
    ca_backward_params_kernel<<<dim3((kmax + 255) / 256, 2 * C + Hd, splits), 256, 0, st>>>(tem, h, dmu, dsigma, dh, gWmu, gbmu, gWsg,
                                                                                            gbsg, gWh, gbh, N, Tm, Hd, C, nps);
    SG_LAUNCHED("ca_backward_params");
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
