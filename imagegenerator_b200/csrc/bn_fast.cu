// bn_fast.cu -- HBM-bound BatchNorm kernels for channel counts that are a multiple of 8 (every BN layer of
// the StackGAN networks: 16 ... 640 channels), 16-byte accesses.
//
// Work decomposition shared by all kernels here: the [rows][C] tensor is a stream of 8-channel vectors;
// CTA b owns the contiguous range [b*chunk, (b+1)*chunk) (chunk a multiple of CV = C/8) and walks it with
// `active` = the largest multiple of CV <= 256 threads, so that thread t always meets channel vector
// t % CV: the per-channel constants (mean, rstd, gamma, the backward sums ...) are folded ONCE into
// registers and the inner loop is loads -> a few FMAs per element -> store, no integer division.  A CTA's
// range touches at most two image groups, so per-channel reductions are combined in shared memory and
// leave the CTA as one fp64 atomic per (group, channel, quantity).
//
// Second family (round 2, final session): RANGE-PARKING kernels -- bn_bwd_fused8_kernel (BatchNorm backward in one launch),
// gp_bn_fused8_kernel (the gradient penalty's second-order pair in one launch), bn_act8_bulk_kernel (forward apply).  One CTA per
// SM owns a range that never crosses an image group, brings ALL of it into shared memory with cp.async.bulk + mbarrier
// transaction counts (the whole range in flight at once), and -- for the two backward kernels -- meets the other CTAs at a
// rendezvous that counts finished ranges, so that nothing depends on the grid being co-resident.  They take every SM they run
// on: the engines use them in the passes that have the GPU to themselves and keep the register-staged kernels (which share SMs
// with the weight-gradient stream) everywhere else.
#include <cstdlib>
#include "common.cuh"

// Register budget of the HBM-bound BatchNorm kernels.  Default: two CTAs of 256 threads per SM at up to 128 registers (the whole
// register file).  -DSG_BN_MAXREG=88 caps them so that two CTAs (45 k registers) leave room for one conv_wgrad2 CTA (192 threads x
// 92 registers) on the same SM: the side stream's tensor-bound weight-gradient kernels can then run UNDER the main chain's
// bandwidth-bound passes instead of alternating with them.
#ifdef SG_BN_MAXREG
#define SG_BN_BOUNDS __maxnreg__(SG_BN_MAXREG)
#else
#define SG_BN_BOUNDS __launch_bounds__(256, 2)
#endif

namespace sg {

struct V8 {
    float v[8];
};
__device__ __forceinline__ V8 ld8(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    V8 r;
    r.v[0] = __uint_as_float(t.x << 16); r.v[1] = __uint_as_float(t.x & 0xffff0000u);
    r.v[2] = __uint_as_float(t.y << 16); r.v[3] = __uint_as_float(t.y & 0xffff0000u);
    r.v[4] = __uint_as_float(t.z << 16); r.v[5] = __uint_as_float(t.z & 0xffff0000u);
    r.v[6] = __uint_as_float(t.w << 16); r.v[7] = __uint_as_float(t.w & 0xffff0000u);
    return r;
}
__device__ __forceinline__ V8 ld8(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    V8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st8(bf16* p, const V8& r) {
    uint4 t;
    t.x = pack_bf16x2(r.v[0], r.v[1]); t.y = pack_bf16x2(r.v[2], r.v[3]);
    t.z = pack_bf16x2(r.v[4], r.v[5]); t.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ void st8(float* p, const V8& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// packed 8-element loads: issue now, unpack at use (keeps 4 vectors per tensor in flight per thread)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> { uint4 q; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ Raw8<bf16> ldraw(const bf16* p) {
    Raw8<bf16> r;
    r.q = __ldg(reinterpret_cast<const uint4*>(p));
    return r;
}
__device__ __forceinline__ Raw8<float> ldraw(const float* p) {
    Raw8<float> r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p + 4));
    return r;
}
__device__ __forceinline__ V8 unpack(const Raw8<bf16>& t) {
    V8 r;
    r.v[0] = __uint_as_float(t.q.x << 16); r.v[1] = __uint_as_float(t.q.x & 0xffff0000u);
    r.v[2] = __uint_as_float(t.q.y << 16); r.v[3] = __uint_as_float(t.q.y & 0xffff0000u);
    r.v[4] = __uint_as_float(t.q.z << 16); r.v[5] = __uint_as_float(t.q.z & 0xffff0000u);
    r.v[6] = __uint_as_float(t.q.w << 16); r.v[7] = __uint_as_float(t.q.w & 0xffff0000u);
    return r;
}
__device__ __forceinline__ V8 unpack(const Raw8<float>& t) {
    V8 r;
    r.v[0] = t.a.x; r.v[1] = t.a.y; r.v[2] = t.a.z; r.v[3] = t.a.w;
    r.v[4] = t.b.x; r.v[5] = t.b.y; r.v[6] = t.b.z; r.v[7] = t.b.w;
    return r;
}
template <typename T> struct Unroll { static constexpr int U = 4; };
template <> struct Unroll<float> { static constexpr int U = 2; };

struct Chunking {
    int64_t nvec, gvec, chunk;   // vectors in the tensor, per image group, per CTA
    int CV, active, blocks;
};

static Chunking make_chunking(int64_t rows_per_group, int C, int groups, int waves) {
    Chunking k;
    k.CV = C / 8;
    k.gvec = rows_per_group * k.CV;
    k.nvec = k.gvec * groups;
    k.active = 256 / k.CV * k.CV;
    int64_t want = (int64_t)SG_NUM_SMS * waves;
    int64_t per = (k.nvec + want - 1) / want;
    // Short CTAs are latency-bound (per-channel constants, first loads, final atomics: ~3 us each whatever the chunk),
    // so mid-size tensors want about one round of CTAs (measured optimum ~192 for 6-25 MB tensors: 1.5-2x faster than
    // 4-8 waves of 4-vector threads); large tensors keep `waves` rounds for load balance.
    static const int minvec = getenv("SG_BN_MINVEC") ? atoi(getenv("SG_BN_MINVEC")) : 0;
    int64_t min_per = (int64_t)k.active * minvec;
    if (minvec == 0) {
        min_per = k.nvec / 192;                                     // <= ~50 MB: one round of ~192 CTAs
        if (min_per < (int64_t)k.active * 4) min_per = (int64_t)k.active * 4;
        if (min_per > (int64_t)k.active * 64) min_per = (int64_t)k.active * 32;
    }
    if (per < min_per) per = min_per;
    k.chunk = (per + k.active - 1) / k.active * k.active;
    k.blocks = (int)((k.nvec + k.chunk - 1) / k.chunk);
    return k;
}

// activation derivative from the stored output for none / ReLU / LeakyReLU: 1 where a > 0, else `slope`
static inline float act_slope(int act) { return act == SG_ACT_RELU ? 0.f : (act == SG_ACT_LRELU ? 0.1f : 1.f); }

// ---- out = act((y - mean) * rstd*gamma + beta [+ residual])
// FIN = true: the batch statistics are still raw sums (BnFinalize::stats, straight out of the conv epilogue).  Every CTA
// turns the sums of the image groups it touches into (mean, rstd*gamma, beta) in shared memory; CTA 0 also does what the
// separate sg_bn_finalize launch did (writes mr for the backward pass, updates the running statistics) -- one launch and
// one kernel-to-kernel dependency less per BatchNorm layer.
struct BnFinalize {
    const double* stats;
    double count;
    float* mr;
    float* rm;
    float* rv;
    long long* nbt;
    int dup_first, update_running, G;
    float momentum, eps;
};
__device__ __forceinline__ void bn_mean_var(const double* stats, int64_t gc, double count, double& mean, double& var) {
    const double s = stats[gc * 2], q = stats[gc * 2 + 1];
    mean = s / count;
    var = q / count - mean * mean;
    if (var < 0) var = 0;
}
template <typename T, bool FIN>
__global__ void SG_BN_BOUNDS
bn_act8_kernel(const T* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
               const float* __restrict__ beta, const T* __restrict__ res, T* __restrict__ out, Chunking k, int act,
               BnFinalize f) {
    SG_PDL_SYNC();
    extern __shared__ float fin_sm[];                // FIN: [groups of this CTA][C][3]
    const int C = k.CV * 8;
    int64_t i = (int64_t)blockIdx.x * k.chunk + threadIdx.x;
    int64_t end = (int64_t)(blockIdx.x + 1) * k.chunk;
    if (end > k.nvec) end = k.nvec;
    int gbase = 0;
    if (FIN) {
        int g_lo = (int)(((int64_t)blockIdx.x * k.chunk) / k.gvec), g_hi = (int)((end - 1) / k.gvec);
        if (blockIdx.x == 0) { g_lo = 0; g_hi = f.G - 1; }
        gbase = g_lo;
        for (int idx = threadIdx.x; idx < (g_hi - g_lo + 1) * C; idx += 256) {
            const int c = idx % C;
            const int64_t gc = (int64_t)g_lo * C + idx;
            double mean, var;
            bn_mean_var(f.stats, gc, f.count, mean, var);
            const float mf = (float)mean, rf = (float)(1.0 / sqrt(var + (double)f.eps));
            fin_sm[idx * 3] = mf; fin_sm[idx * 3 + 1] = rf * gamma[c]; fin_sm[idx * 3 + 2] = beta[c];
            if (blockIdx.x == 0) { f.mr[gc * 2] = mf; f.mr[gc * 2 + 1] = rf; }
        }
        if (blockIdx.x == 0 && f.update_running) {
            if (threadIdx.x == 0 && f.nbt) *f.nbt += f.dup_first + f.G - 1;
            for (int c = threadIdx.x; c < C; c += 256) {
                float m_run = f.rm[c], v_run = f.rv[c];
                for (int g = 0; g < f.G; ++g) {
                    double mean, var;
                    bn_mean_var(f.stats, (int64_t)g * C + c, f.count, mean, var);
                    const float unb = (float)(var * f.count / (f.count > 1 ? f.count - 1 : 1));
                    const int reps = g == 0 ? f.dup_first : 1;
                    for (int r = 0; r < reps; ++r) {
                        m_run = (1.f - f.momentum) * m_run + f.momentum * (float)mean;
                        v_run = (1.f - f.momentum) * v_run + f.momentum * unb;
                    }
                }
                f.rm[c] = m_run; f.rv[c] = v_run;
            }
        }
        __syncthreads();
    }
    if ((int)threadIdx.x >= k.active) return;
    if (i >= end) return;
    const int c0 = ((int)threadIdx.x % k.CV) * 8;
    int g = (int)(i / k.gvec);
    int64_t next = (int64_t)(g + 1) * k.gvec;
    float m[8], rg[8], b[8];
    auto load = [&](int gg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FIN) {
                const float* q = fin_sm + ((int64_t)(gg - gbase) * C + c0 + j) * 3;
                m[j] = q[0]; rg[j] = q[1]; b[j] = q[2];
            } else {
                const float* q = mr + ((int64_t)gg * C + c0 + j) * 2;
                m[j] = q[0]; rg[j] = q[1] * gamma[c0 + j]; b[j] = beta[c0 + j];
            }
        }
    };
    load(g);
    const float slope = act == SG_ACT_RELU ? 0.f : (act == SG_ACT_LRELU ? 0.1f : 1.f);
    constexpr int U = Unroll<T>::U;
    auto one = [&](const V8& v, const V8* r, int64_t at) {
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = (v.v[j] - m[j]) * rg[j] + b[j] + (r ? r->v[j] : 0.f);
        if (act == SG_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = tanhf(o.v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = o.v[j] > 0.f ? o.v[j] : slope * o.v[j];
        }
        st8(out + at * 8, o);
    };
    while (i < end) {
        if (i >= next) {
            do { ++g; next += k.gvec; } while (i >= next);
            load(g);
        }
        const int64_t lim = end < next ? end : next;
        if (i + (int64_t)(U - 1) * k.active < lim) {
            Raw8<T> ry[U], rr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) ry[u] = ldraw(y + (i + (int64_t)u * k.active) * 8);
            if (res != nullptr) {
#pragma unroll
                for (int u = 0; u < U; ++u) rr[u] = ldraw(res + (i + (int64_t)u * k.active) * 8);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                V8 v = unpack(ry[u]);
                if (res != nullptr) { V8 r = unpack(rr[u]); one(v, &r, i + (int64_t)u * k.active); }
                else one(v, nullptr, i + (int64_t)u * k.active);
            }
            i += (int64_t)U * k.active;
        } else {
            V8 v = unpack(ldraw(y + i * 8));
            if (res != nullptr) { V8 r = unpack(ldraw(res + i * 8)); one(v, &r, i); }
            else one(v, nullptr, i);
            i += k.active;
        }
    }
}

// ---- sums[g][c] += (sum dz, sum dz*xhat), dz = da * act'(a_out).  HAS_A = false: the activation output is not read;
// its sign is recomputed from y (a = act(gamma*xhat + beta) has the sign of its argument), one tensor less to stream
template <typename T, bool HAS_A>
__global__ void SG_BN_BOUNDS
bn_bwd_reduce8_kernel(const T* __restrict__ da, const T* __restrict__ a_out, const T* __restrict__ y,
                      const float* __restrict__ mr, const float* __restrict__ gamma, const float* __restrict__ beta,
                      double* __restrict__ sums, Chunking k, float slope) {
    SG_PDL_SYNC();
    extern __shared__ float sacc[];                 // [touched groups][C][2]
    const int C = k.CV * 8;
    const int64_t begin = (int64_t)blockIdx.x * k.chunk;
    int64_t end = begin + k.chunk;
    if (end > k.nvec) end = k.nvec;
    const int g_first = (int)(begin / k.gvec), g_last = (int)((end - 1) / k.gvec);
    const int nacc = (g_last - g_first + 1) * C * 2;
    for (int t = threadIdx.x; t < nacc; t += 256) sacc[t] = 0.f;
    __syncthreads();
    int64_t i = begin + threadIdx.x;
    // The threads' partial sums meet in shared memory: part[16][256(+1)] -> one owner per (channel, sum) adds the column
    // entries of its channel vector.  (Shared-memory atomics serialise C/8-fold: 85 threads per address for the
    // 24-channel generator layer, 36 us for a 6 MB tensor.)
    __shared__ float part[16][257];
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    int g_mine = -1;
    if ((int)threadIdx.x < k.active && i < end) {
        const int c0 = ((int)threadIdx.x % k.CV) * 8;
        int g = (int)(i / k.gvec);
        int64_t next = (int64_t)(g + 1) * k.gvec;
        float m[8], r[8], rg[HAS_A ? 1 : 8], bt[HAS_A ? 1 : 8];
        auto load = [&](int gg) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* q = mr + ((int64_t)gg * C + c0 + j) * 2;
                m[j] = q[0]; r[j] = q[1]; s1[j] = 0.f; s2[j] = 0.f;
                if (!HAS_A) { rg[j] = q[1] * gamma[c0 + j]; bt[j] = beta[c0 + j]; }
            }
        };
        auto flush = [&](int gg) {
            float* dst = sacc + ((gg - g_first) * C + c0) * 2;
#pragma unroll
            for (int j = 0; j < 8; ++j) { atomicAdd(dst + 2 * j, s1[j]); atomicAdd(dst + 2 * j + 1, s2[j]); }
        };
        load(g);
        constexpr int U = Unroll<T>::U;
        auto one = [&](const V8& d, const V8& a, const V8& yy) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float sgn = HAS_A ? a.v[j] : (yy.v[j] - m[j]) * rg[HAS_A ? 0 : j] + bt[HAS_A ? 0 : j];
                float dz = sgn > 0.f ? d.v[j] : slope * d.v[j];
                s1[j] += dz;
                s2[j] += dz * ((yy.v[j] - m[j]) * r[j]);
            }
        };
        while (i < end) {
            if (i >= next) {
                flush(g);
                do { ++g; next += k.gvec; } while (i >= next);
                load(g);
            }
            const int64_t lim = end < next ? end : next;
            if (i + (int64_t)(U - 1) * k.active < lim) {
                Raw8<T> rd[U], ra[HAS_A ? U : 1], ry[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t at = (i + (int64_t)u * k.active) * 8;
                    rd[u] = ldraw(da + at); ry[u] = ldraw(y + at);
                    if (HAS_A) ra[u] = ldraw(a_out + at);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const V8 yy = unpack(ry[u]);
                    one(unpack(rd[u]), HAS_A ? unpack(ra[HAS_A ? u : 0]) : yy, yy);
                }
                i += (int64_t)U * k.active;
            } else {
                const V8 yy = unpack(ldraw(y + i * 8));
                one(unpack(ldraw(da + i * 8)), HAS_A ? unpack(ldraw(a_out + i * 8)) : yy, yy);
                i += k.active;
            }
        }
        g_mine = g;
        if (g != g_last) {                           // (only when the group boundary falls into the CTA's last stride)
            flush(g);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        }
    }
    (void)g_mine;
#pragma unroll
    for (int j = 0; j < 8; ++j) { part[2 * j][threadIdx.x] = s1[j]; part[2 * j + 1][threadIdx.x] = s2[j]; }
    __syncthreads();
    for (int o = threadIdx.x; o < C * 2; o += 256) {
        const int c = o >> 1, row = 2 * (c & 7) + (o & 1);
        float acc = 0.f;
        for (int t = c >> 3; t < k.active; t += k.CV) acc += part[row][t];
        sacc[(g_last - g_first) * C * 2 + o] += acc;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nacc; t += 256) atomicAdd(sums + (int64_t)g_first * C * 2 + t, (double)sacc[t]);
}

// ---- dy = gamma*rstd/n * (n dz - S1 - xhat S2) [+ inject on one group]
template <typename T, bool HAS_A>
__global__ void SG_BN_BOUNDS
bn_bwd_apply8_kernel(const T* __restrict__ da, const T* __restrict__ a_out, const T* __restrict__ y,
                     const float* __restrict__ mr, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const double* __restrict__ sums, const T* __restrict__ inject, int inject_group, T* __restrict__ dy,
                     Chunking k, float slope, float n) {
    SG_PDL_SYNC();
    if ((int)threadIdx.x >= k.active) return;
    int64_t i = (int64_t)blockIdx.x * k.chunk + threadIdx.x;
    int64_t end = (int64_t)(blockIdx.x + 1) * k.chunk;
    if (end > k.nvec) end = k.nvec;
    if (i >= end) return;
    const int c0 = ((int)threadIdx.x % k.CV) * 8, C = k.CV * 8;
    int g = (int)(i / k.gvec);
    int64_t next = (int64_t)(g + 1) * k.gvec;
    float m[8], r[8], gr[8], c1[8], c2[8], rg[HAS_A ? 1 : 8], bt[HAS_A ? 1 : 8];
    auto load = [&](int gg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float* q = mr + ((int64_t)gg * C + c0 + j) * 2;
            const double* s = sums + ((int64_t)gg * C + c0 + j) * 2;
            m[j] = q[0]; r[j] = q[1];
            float coef = gamma[c0 + j] * q[1] / n;
            gr[j] = coef * n; c1[j] = coef * (float)s[0]; c2[j] = coef * (float)s[1];
            if (!HAS_A) { rg[j] = q[1] * gamma[c0 + j]; bt[j] = beta[c0 + j]; }
        }
    };
    load(g);
    constexpr int U = Unroll<T>::U;
    auto one = [&](const V8& d, const V8& a, const V8& yy, int64_t at) {
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float sgn = HAS_A ? a.v[j] : (yy.v[j] - m[j]) * rg[HAS_A ? 0 : j] + bt[HAS_A ? 0 : j];
            float dz = sgn > 0.f ? d.v[j] : slope * d.v[j];
            float xh = (yy.v[j] - m[j]) * r[j];
            o.v[j] = gr[j] * dz - c1[j] - xh * c2[j];
        }
        if (inject != nullptr && g == inject_group) {
            V8 q = unpack(ldraw(inject + (at - (int64_t)inject_group * k.gvec) * 8));
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] += q.v[j];
        }
        st8(dy + at * 8, o);
    };
    while (i < end) {
        if (i >= next) {
            do { ++g; next += k.gvec; } while (i >= next);
            load(g);
        }
        const int64_t lim = end < next ? end : next;
        if (i + (int64_t)(U - 1) * k.active < lim) {
            Raw8<T> rd[U], ra[HAS_A ? U : 1], ry[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t at = (i + (int64_t)u * k.active) * 8;
                rd[u] = ldraw(da + at); ry[u] = ldraw(y + at);
                if (HAS_A) ra[u] = ldraw(a_out + at);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const V8 yy = unpack(ry[u]);
                one(unpack(rd[u]), HAS_A ? unpack(ra[HAS_A ? u : 0]) : yy, yy, i + (int64_t)u * k.active);
            }
            i += (int64_t)U * k.active;
        } else {
            const V8 yy = unpack(ldraw(y + i * 8));
            one(unpack(ldraw(da + i * 8)), HAS_A ? unpack(ldraw(a_out + i * 8)) : yy, yy, i);
            i += k.active;
        }
    }
}

// ---- BatchNorm backward in ONE launch for tensors whose (da, y[, a]) fit the SMs' shared memory (every BN layer of Stage-I at
// batch 128; the <= 40 MB layers of Stage-II).  CTA b owns the contiguous RANGE b of the vector stream:
//   * one thread brings the range into shared memory with bulk async copies (cp.async.bulk + mbarrier transaction counts, one
//     barrier per piece of U*active vectors) BEFORE anything else happens -- the whole range is in flight at once (up to
//     ~190 KB per SM; the register-staged kernels above keep ~40 KB per SM in flight and are latency-bound on 6-25 MB
//     tensors), and the per-channel constants are fetched under it;
//   * phase 1 = bn_bwd_reduce8 out of shared memory, per-channel sums into global fp64 atomics;
//   * a grid-wide rendezvous on the sums;
//   * phase 2 = bn_bwd_apply8 out of shared memory: (da, y) cross the L2 -> SM path once instead of twice, and one launch +
//     one dependency edge per layer disappear.  What does not fit (`keep` < range) takes the register path in both phases.
//
// The rendezvous does NOT assume co-residency (the captured steps run up to three streams): it counts finished
// RANGES, not CTAs.  A CTA claims its own range with an atomic exchange; one that has waited `steal_ns` at the rendezvous
// starts claiming ranges nobody has started (CTAs that are not resident yet) and reduces them through the register path, so
// whatever subset of the grid is resident finishes by itself; CTAs that start late find their range taken and leave.
// `work` = {-, ranges done, CTAs exited, error, claim[nranges]}: zero before the first launch, re-armed by the last CTA to
// leave; owned by the caller (one per call site), so concurrent launches never share it.
constexpr int FNT = 512;           // threads per CTA (one CTA per SM: the parked range is the SM's shared memory)
constexpr int FMAXR = 160;         // ranges per launch (<= one per SM)
constexpr int FPIECES = 16;        // bulk-copy pieces (mbarriers) per range
struct FusedPlan {
    int64_t gvec;                  // vectors per image group
    int range, keep;               // vectors per range (multiple of active), parked per range (multiple of active)
    int CV, active, rpg, nranges, dbg;     // rpg: ranges per image group -- a range never crosses a group boundary
    int steal_ns;                          // patience at the rendezvous before taking over ranges nobody has started
};
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t fsaddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fbar_wait(uint64_t* bar, uint32_t parity) {     // bounded: a protocol bug traps, never hangs
    const uint32_t addr = fsaddr(bar);
    uint32_t done;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fbulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(fsaddr(dst)), "l"(src), "r"(bytes), "r"(fsaddr(bar)) : "memory");
}
// thread 0 at the rendezvous: -1 = every range is reduced; r >= 0 = range r was not started by anybody, reduce it as well
__device__ int fused_wait_or_steal(unsigned* work, int nranges, int steal_ns) {
    unsigned* claim = work + 4;
    long long t0 = globaltimer_ns();
    const long long t_begin = t0;
    for (;;) {
        if (ld_acquire_u32(&work[1]) >= (unsigned)nranges) { __threadfence(); return -1; }
        __nanosleep(40);
        const long long now = globaltimer_ns();
        if (now - t0 > steal_ns) {
            for (int r = 0; r < nranges; ++r)
                if (ld_acquire_u32(&claim[r]) == 0u && atomicExch(&claim[r], 1u) == 0u) return r;
            t0 = now;
            if (now - t_begin > 2000000000ll) { work[3] = 0xdeadu; __trap(); }      // 2 s: never hang the device
        }
    }
}

template <typename T, bool HAS_A>
__global__ void __launch_bounds__(FNT, 1)
bn_bwd_fused8_kernel(const T* __restrict__ da, const T* __restrict__ a_out, const T* __restrict__ y,
                     const float* __restrict__ mr, const float* __restrict__ gamma, const float* __restrict__ beta,
                     double* __restrict__ sums, const T* __restrict__ inject, int inject_group, T* __restrict__ dy,
                     FusedPlan k, float slope, float n, unsigned* __restrict__ work) {
    SG_PDL_SYNC();
    extern __shared__ __align__(128) unsigned char fsm[];
    typedef Raw8<T> R8;
    R8* p_da = reinterpret_cast<R8*>(fsm);
    R8* p_y = p_da + k.keep;
    R8* p_a = p_y + k.keep;
    const int C = k.CV * 8, tid = (int)threadIdx.x;
    float* cst = reinterpret_cast<float*>(p_a + (HAS_A ? k.keep : 0));        // [4][C]: mean, rstd, gamma, beta of the range's group
    float* sacc = cst + 4 * C;                                                // [C][2]: the CTA's sums, later the grid's
    __shared__ float part[16][FNT + 1];
    __shared__ uint64_t bars[FPIECES];
    __shared__ int s_cmd;
    __shared__ int mine[FMAXR];
    const int c0 = (tid % k.CV) * 8;
    const bool worker = tid < k.active;
    constexpr int U = Unroll<T>::U;
    const int piece = k.active * U;
    int nmine = 0;
    // option bn_fused_dbg: globaltimer stamps of the first and the last CTA at work + 800 bytes (tools/bench_bn_fused.py)
    long long* stamps = reinterpret_cast<long long*>(work + 200) + (blockIdx.x == 0 ? 0 : 8);
    const bool stamp_on = k.dbg && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
#define FSTAMP(slot) do { if (stamp_on) stamps[slot] = globaltimer_ns(); } while (0)
    FSTAMP(0);
    auto range_of = [&](int r, int& g, int64_t& begin, int& len) {
        g = r / k.rpg;
        const int64_t off = (int64_t)(r % k.rpg) * k.range;
        begin = (int64_t)g * k.gvec + off;
        const int64_t left = k.gvec - off;
        len = (int)(left < k.range ? left : k.range);
    };

    if (tid == 0) {
        for (int p = 0; p < FPIECES; ++p)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fsaddr(&bars[p])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int b = (int)blockIdx.x;
        const bool own = atomicExch(&work[4 + b], 1u) == 0u;
        s_cmd = own ? b : -1;
        if (own) {                                   // the whole parked part of the range goes in flight right now
            int g, len;
            int64_t begin;
            range_of(b, g, begin, len);
            const int left = len < k.keep ? len : k.keep;
            for (int p = 0, v0 = 0; v0 < left; ++p, v0 += piece) {
                const int cnt = left - v0 < piece ? left - v0 : piece;
                const uint32_t bytes = (uint32_t)cnt * (uint32_t)sizeof(R8);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                             ::"r"(fsaddr(&bars[p])), "r"(bytes * (HAS_A ? 3u : 2u)) : "memory");
                fbulk_g2s(p_da + v0, da + (begin + v0) * 8, bytes, &bars[p]);
                fbulk_g2s(p_y + v0, y + (begin + v0) * 8, bytes, &bars[p]);
                if (HAS_A) fbulk_g2s(p_a + v0, a_out + (begin + v0) * 8, bytes, &bars[p]);
            }
        }
    }
    __syncthreads();
    int cmd = s_cmd;
    FSTAMP(1);

    // ---------------- phase 1: per-channel sums of this CTA's range (and of ranges nobody else started)
    while (cmd >= 0) {
        const int r = cmd;
        const int keep = nmine == 0 ? k.keep : 0;
        if (tid == 0) mine[nmine] = r;
        ++nmine;
        int g, len;
        int64_t begin;
        range_of(r, g, begin, len);
        // the group's per-channel constants: one coalesced pass into shared memory (under the copies in flight)
        for (int c = tid; c < C; c += FNT) {
            const float2 q = *reinterpret_cast<const float2*>(mr + ((int64_t)g * C + c) * 2);
            cst[c] = q.x; cst[C + c] = q.y; cst[2 * C + c] = gamma[c]; cst[3 * C + c] = HAS_A ? 0.f : beta[c];
            sacc[2 * c] = 0.f; sacc[2 * c + 1] = 0.f;
        }
        __syncthreads();
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        int v = tid;
        if (worker && v < len) {
            int waited = 0;                          // pieces of the parked range known to have landed
            float m[8], rs[8], rg[HAS_A ? 1 : 8], bt[HAS_A ? 1 : 8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                m[j] = cst[c0 + j]; rs[j] = cst[C + c0 + j];
                if (!HAS_A) { rg[j] = rs[j] * cst[2 * C + c0 + j]; bt[j] = cst[3 * C + c0 + j]; }
            }
            auto one = [&](const V8& d, const V8& a, const V8& yy) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float sgn = HAS_A ? a.v[j] : (yy.v[j] - m[j]) * rg[HAS_A ? 0 : j] + bt[HAS_A ? 0 : j];
                    const float dz = sgn > 0.f ? d.v[j] : slope * d.v[j];
                    s1[j] += dz;
                    s2[j] += dz * ((yy.v[j] - m[j]) * rs[j]);
                }
            };
            auto landed = [&](int vv) {              // vector vv of the parked range is in shared memory
                const int need = vv / piece;
                while (waited <= need) { fbar_wait(&bars[waited], 0); ++waited; }
            };
            while (v < len) {
                if (v + (U - 1) * k.active < len) {
                    R8 rd[U], ra[HAS_A ? U : 1], ry[U];
                    if (v + (U - 1) * k.active < keep) {
                        landed(v + (U - 1) * k.active);
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            rd[u] = p_da[v + u * k.active]; ry[u] = p_y[v + u * k.active];
                            if (HAS_A) ra[HAS_A ? u : 0] = p_a[v + u * k.active];
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int vv = v + u * k.active;
                            const int64_t at = (begin + vv) * 8;
                            if (vv < keep) {
                                landed(vv);
                                rd[u] = p_da[vv]; ry[u] = p_y[vv];
                                if (HAS_A) ra[HAS_A ? u : 0] = p_a[vv];
                            } else {
                                rd[u] = ldraw(da + at); ry[u] = ldraw(y + at);
                                if (HAS_A) ra[HAS_A ? u : 0] = ldraw(a_out + at);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const V8 yy = unpack(ry[u]);
                        one(unpack(rd[u]), HAS_A ? unpack(ra[HAS_A ? u : 0]) : yy, yy);
                    }
                    v += U * k.active;
                } else {
                    R8 rd, ry, ra;
                    const int64_t at = (begin + v) * 8;
                    if (v < keep) { landed(v); rd = p_da[v]; ry = p_y[v]; if (HAS_A) ra = p_a[v]; }
                    else { rd = ldraw(da + at); ry = ldraw(y + at); if (HAS_A) ra = ldraw(a_out + at); }
                    const V8 yy = unpack(ry);
                    one(unpack(rd), HAS_A ? unpack(ra) : yy, yy);
                    v += k.active;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { part[2 * j][tid] = s1[j]; part[2 * j + 1][tid] = s2[j]; }
        __syncthreads();
        FSTAMP(2);
        {   // one owner per (channel, sum, segment): the column entries of its channel vector, `nseg` threads share a column
            const int nout = C * 2, nseg = FNT / nout > 0 ? FNT / nout : 1;
            for (int w = tid; w < nout * nseg; w += FNT) {
                const int o = w % nout, seg = w / nout;
                const int c = o >> 1, row = 2 * (c & 7) + (o & 1);
                float acc = 0.f;
                for (int t = (c >> 3) + seg * k.CV; t < k.active; t += k.CV * nseg) acc += part[row][t];
                if (nseg == 1) atomicAdd(sums + (int64_t)g * C * 2 + o, (double)acc);
                else atomicAdd(sacc + o, acc);
            }
            if (nseg > 1) {
                __syncthreads();
                for (int o = tid; o < nout; o += FNT) atomicAdd(sums + (int64_t)g * C * 2 + o, (double)sacc[o]);
            }
        }
        __syncthreads();
        FSTAMP(3);
        if (tid == 0) {
            __threadfence();
            atomicAdd(&work[1], 1u);
            FSTAMP(4);
            s_cmd = fused_wait_or_steal(work, k.nranges, k.steal_ns);        // ---------------- the rendezvous
        }
        __syncthreads();
        cmd = s_cmd;
        FSTAMP(5);
    }

    // ---------------- phase 2: dy = gamma*rstd/n * (n dz - S1 - xhat S2) [+ inject]
    for (int q = 0; q < nmine; ++q) {
        const int r = mine[q];
        const int keep = q == 0 ? k.keep : 0;
        int g, len;
        int64_t begin;
        range_of(r, g, begin, len);
        // the grid's sums of this group: ONE coalesced L2 read per CTA (every thread fetching its 16 values itself makes
        // all SMs hammer the same few L2 lines: 15-20 us)
        if (nmine > 1) {
            __syncthreads();
            for (int c = tid; c < C; c += FNT) {
                const float2 t = *reinterpret_cast<const float2*>(mr + ((int64_t)g * C + c) * 2);
                cst[c] = t.x; cst[C + c] = t.y; cst[2 * C + c] = gamma[c]; cst[3 * C + c] = HAS_A ? 0.f : beta[c];
            }
        }
        for (int o = tid; o < C * 2; o += FNT) sacc[o] = (float)__ldcg(sums + (int64_t)g * C * 2 + o);
        __syncthreads();
        int v = tid;
        if (!worker || v >= len) continue;
        float m[8], rs[8], gr[8], c1[8], c2[8], rg[HAS_A ? 1 : 8], bt[HAS_A ? 1 : 8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            m[j] = cst[c0 + j]; rs[j] = cst[C + c0 + j];
            const float gm = cst[2 * C + c0 + j];
            const float coef = gm * rs[j] / n;
            gr[j] = coef * n; c1[j] = coef * sacc[(c0 + j) * 2]; c2[j] = coef * sacc[(c0 + j) * 2 + 1];
            if (!HAS_A) { rg[j] = rs[j] * gm; bt[j] = cst[3 * C + c0 + j]; }
        }
        const bool inj = inject != nullptr && g == inject_group;
        auto one = [&](const V8& d, const V8& a, const V8& yy, int vv) {
            V8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float sgn = HAS_A ? a.v[j] : (yy.v[j] - m[j]) * rg[HAS_A ? 0 : j] + bt[HAS_A ? 0 : j];
                const float dz = sgn > 0.f ? d.v[j] : slope * d.v[j];
                const float xh = (yy.v[j] - m[j]) * rs[j];
                o.v[j] = gr[j] * dz - c1[j] - xh * c2[j];
            }
            const int64_t at = begin + vv;
            if (inj) {
                V8 w = unpack(ldraw(inject + (at - (int64_t)inject_group * k.gvec) * 8));
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] += w.v[j];
            }
            st8(dy + at * 8, o);
        };
        while (v < len) {
            if (v + (U - 1) * k.active < len) {
                R8 rd[U], ra[HAS_A ? U : 1], ry[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int vv = v + u * k.active;
                    const int64_t at = (begin + vv) * 8;
                    if (vv < keep) {
                        rd[u] = p_da[vv]; ry[u] = p_y[vv];
                        if (HAS_A) ra[HAS_A ? u : 0] = p_a[vv];
                    } else {
                        rd[u] = ldraw(da + at); ry[u] = ldraw(y + at);
                        if (HAS_A) ra[HAS_A ? u : 0] = ldraw(a_out + at);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const V8 yy = unpack(ry[u]);
                    one(unpack(rd[u]), HAS_A ? unpack(ra[HAS_A ? u : 0]) : yy, yy, v + u * k.active);
                }
                v += U * k.active;
            } else {
                R8 rd, ry, ra;
                const int64_t at = (begin + v) * 8;
                if (v < keep) { rd = p_da[v]; ry = p_y[v]; if (HAS_A) ra = p_a[v]; }
                else { rd = ldraw(da + at); ry = ldraw(y + at); if (HAS_A) ra = ldraw(a_out + at); }
                const V8 yy = unpack(ry);
                one(unpack(rd), HAS_A ? unpack(ra) : yy, yy, v);
                v += k.active;
            }
        }
    }
    // ---------------- the last CTA to leave re-arms the work words for the next launch of this call site
    __syncthreads();
    FSTAMP(6);
#undef FSTAMP
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&work[2], 1u) == gridDim.x - 1) {
            for (int r = 0; r < k.nranges; ++r) work[4 + r] = 0u;
            work[1] = 0u; work[2] = 0u;
            __threadfence();
        }
    }
}

int g_bn_fused = 0;                // option "bn_fused": 1 = one launch where the tensor fits; default 0 (two kernels): inside the
                                   // captured steps the one-launch kernel cannot share an SM with the side stream's wgrad CTAs
                                   // (190 KB of shared memory each) and measured 1-2 % slower per step, see DESIGN.md section 5;
                                   // the engines switch it on around the gradient penalty's first-order pass (side streams idle)
int g_bn_fused_keep_pct = 50;      // fuse when at least this share of a range can be parked in shared memory
int g_bn_fused_steal_ns = 30000;   // option "bn_fused_steal_ns": taking over is a safety net, not a schedule (4 us: Stage-I 5.25 -> 5.64 ms)
int g_gp_bn_fused = 1;             // option "gp_bn_fused": the penalty's second-order BatchNorm pair as one launch
int g_bn_fused_dbg = 0;            // option "bn_fused_dbg": the first / last CTA leave globaltimer stamps in the work words

// plan + launch; returns -1 (nothing launched) when the tensor is too large to profit -- the caller then runs reduce + apply
template <typename T>
int bn_bwd_fused8(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma, const float* beta,
                  double* sums, const void* inject, int inject_group, void* dy, int64_t rows_per_group, int C, int groups,
                  int act, unsigned* work, cudaStream_t st) {
    if (!g_bn_fused || work == nullptr) return -1;
    constexpr int U = Unroll<T>::U;
    FusedPlan k;
    k.CV = C / 8;
    if (k.CV > FNT || groups > SG_NUM_SMS) return -1;
    k.gvec = rows_per_group * k.CV;
    k.active = FNT / k.CV * k.CV;
    const int slots = SG_NUM_SMS / groups;          // ranges per image group: one CTA per SM in total
    const int64_t per = (k.gvec + slots - 1) / slots;
    const int64_t range = (per + k.active - 1) / k.active * k.active;
    if (range > (1 << 24)) return -1;
    k.range = (int)range;
    k.rpg = (int)((k.gvec + range - 1) / range);
    k.nranges = k.rpg * groups;
    if (k.nranges > FMAXR) return -1;
    const size_t acc_bytes = (size_t)C * 6 * sizeof(float);       // per-channel constants [4][C] + sums [C][2]
    const size_t vb = sizeof(Raw8<T>) * (a_out != nullptr ? 3 : 2);
    const size_t budget = 188 * 1024;               // next to the 34 KB of static reduction scratch
    if (acc_bytes + vb * k.active > budget) return -1;
    int64_t keep = (int64_t)((budget - acc_bytes) / vb) / k.active * k.active;
    if (keep > range) keep = range;
    if (keep * 100 < range * g_bn_fused_keep_pct) return -1;
    if ((keep + (int64_t)k.active * U - 1) / ((int64_t)k.active * U) > FPIECES) return -1;
    k.keep = (int)keep;
    k.dbg = g_bn_fused_dbg;
    k.steal_ns = g_bn_fused_steal_ns;
    const size_t smem = acc_bytes + vb * (size_t)keep;
    if (a_out != nullptr) {
        static bool set = false;
        if (!set) { cudaFuncSetAttribute(bn_bwd_fused8_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget); set = true; }
        launch_pdl(bn_bwd_fused8_kernel<T, true>, dim3(k.nranges), dim3(FNT), smem, st, (const T*)da, (const T*)a_out, (const T*)y, mr,
                   gamma, beta, sums, (const T*)inject, inject_group, (T*)dy, k, act_slope(act), (float)rows_per_group, work);
    } else {
        static bool set = false;
        if (!set) { cudaFuncSetAttribute(bn_bwd_fused8_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget); set = true; }
        launch_pdl(bn_bwd_fused8_kernel<T, false>, dim3(k.nranges), dim3(FNT), smem, st, (const T*)da, (const T*)nullptr, (const T*)y, mr,
                   gamma, beta, sums, (const T*)inject, inject_group, (T*)dy, k, act_slope(act), (float)rows_per_group, work);
    }
    g_launches.fetch_add(1);
    return check_launch("bn_bwd_fused8");
}
template int bn_bwd_fused8<float>(const void*, const void*, const void*, const float*, const float*, const float*, double*, const void*, int, void*, int64_t, int, int, int, unsigned*, cudaStream_t);
template int bn_bwd_fused8<bf16>(const void*, const void*, const void*, const float*, const float*, const float*, double*, const void*, int, void*, int64_t, int, int, int, unsigned*, cudaStream_t);

// ---- out = da * act'(a_out), 8-wide (ReLU / LeakyReLU / Tanh / none)
template <typename T>
__global__ void __launch_bounds__(256, 4)
act_bwd8_kernel(const T* __restrict__ da, const T* __restrict__ a_out, T* __restrict__ out, int64_t nvec, int act) {
    SG_PDL_SYNC();
    const float slope = act == SG_ACT_RELU ? 0.f : (act == SG_ACT_LRELU ? 0.1f : 1.f);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * 256) {
        V8 d = ld8(da + i * 8), a = ld8(a_out + i * 8), o;
        if (act == SG_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = d.v[j] * (1.f - a.v[j] * a.v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = a.v[j] > 0.f ? d.v[j] : slope * d.v[j];
        }
        st8(out + i * 8, o);
    }
}

// ---- out = da * act'(a_out) AND colsum[c] += sum_rows out[., c] (ReLU / LeakyReLU / none): the activation backward of a
// conv + bias + activation layer and that layer's bias gradient in one pass -- the separate column-sum pass (25 us on the
// critic's 50 MB first activation) sat on the tail of every critic iteration between the last data gradient and the optimizer.
__device__ __forceinline__ float as_stored(float x, const bf16*) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float as_stored(float x, const float*) { return x; }
template <typename T>
__global__ void SG_BN_BOUNDS
act_bwd8_colsum_kernel(const T* __restrict__ da, const T* __restrict__ a_out, T* __restrict__ out, float* __restrict__ colsum,
                       Chunking k, float slope) {
    SG_PDL_SYNC();
    __shared__ float part[8][257];
    const int C = k.CV * 8;
    int64_t i = (int64_t)blockIdx.x * k.chunk + threadIdx.x;
    int64_t end = (int64_t)(blockIdx.x + 1) * k.chunk;
    if (end > k.nvec) end = k.nvec;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if ((int)threadIdx.x < k.active) {
        constexpr int U = Unroll<T>::U;
        auto one = [&](const V8& d, const V8& a, int64_t at) {
            V8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = a.v[j] > 0.f ? d.v[j] : slope * d.v[j];
            st8(out + at * 8, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += as_stored(o.v[j], out);      // sums of the STORED (rounded) gradient, as the separate pass saw it
        };
        while (i < end) {
            if (i + (int64_t)(U - 1) * k.active < end) {
                Raw8<T> rd[U], ra[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t at = (i + (int64_t)u * k.active) * 8;
                    rd[u] = ldraw(da + at); ra[u] = ldraw(a_out + at);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) one(unpack(rd[u]), unpack(ra[u]), i + (int64_t)u * k.active);
                i += (int64_t)U * k.active;
            } else {
                one(unpack(ldraw(da + i * 8)), unpack(ldraw(a_out + i * 8)), i);
                i += k.active;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[j][threadIdx.x] = s[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float acc = 0.f;
        for (int t = c >> 3; t < k.active; t += k.CV) acc += part[c & 7][t];
        atomicAdd(colsum + c, acc);
    }
}
template <typename T>
int act_bwd8_colsum(const void* da, const void* a_out, void* out, float* colsum, int64_t rows, int C, int act, cudaStream_t st) {
    Chunking k = make_chunking(rows, C, 1, 8);
    launch_pdl(act_bwd8_colsum_kernel<T>, dim3(k.blocks), dim3(256), 0, st, (const T*)da, (const T*)a_out, (T*)out, colsum, k,
               act_slope(act));
    g_launches.fetch_add(1);
    return check_launch("act_bwd8_colsum");
}
template int act_bwd8_colsum<float>(const void*, const void*, void*, float*, int64_t, int, int, cudaStream_t);
template int act_bwd8_colsum<bf16>(const void*, const void*, void*, float*, int64_t, int, int, cudaStream_t);

// ---- gradient-penalty double backward through a train-mode BN (one image group), 8-wide.
// reduce: tsums[c] += (sum v, sum v*xhat, sum v*dz)
template <typename T>
__global__ void SG_BN_BOUNDS
gp_bn_reduce8_kernel(const T* __restrict__ v, const T* __restrict__ da, const T* __restrict__ a_out, const T* __restrict__ y,
                     const float* __restrict__ mr, double* __restrict__ tsums, Chunking k, float slope) {
    SG_PDL_SYNC();
    extern __shared__ float sacc[];                 // [C][3]
    const int C = k.CV * 8;
    for (int t = threadIdx.x; t < C * 3; t += 256) sacc[t] = 0.f;
    __syncthreads();
    const int64_t begin = (int64_t)blockIdx.x * k.chunk;
    int64_t end = begin + k.chunk;
    if (end > k.nvec) end = k.nvec;
    int64_t i = begin + threadIdx.x;
    if ((int)threadIdx.x < k.active && i < end) {
        const int c0 = ((int)threadIdx.x % k.CV) * 8;
        float m[8], r[8], t1[8], t2[8], t3[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m[j] = mr[(c0 + j) * 2]; r[j] = mr[(c0 + j) * 2 + 1]; t1[j] = t2[j] = t3[j] = 0.f; }
        for (; i < end; i += k.active) {
            const V8 vv = unpack(ldraw(v + i * 8)), d = unpack(ldraw(da + i * 8)), a = unpack(ldraw(a_out + i * 8)),
                     yy = unpack(ldraw(y + i * 8));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dz = a.v[j] > 0.f ? d.v[j] : slope * d.v[j];
                t1[j] += vv.v[j];
                t2[j] += vv.v[j] * ((yy.v[j] - m[j]) * r[j]);
                t3[j] += vv.v[j] * dz;
            }
        }
        float* dst = sacc + c0 * 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(dst + 3 * j, t1[j]); atomicAdd(dst + 3 * j + 1, t2[j]); atomicAdd(dst + 3 * j + 2, t3[j]); }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < C * 3; t += 256) atomicAdd(tsums + t, (double)sacc[t]);
}

// apply: w = u*act', gy = d/dy of the penalty term; the 8 per-channel constants live in shared memory
//   u = A v - B1 - xhat B2,  G = -(E v + B2 dz),  gy = r (G - K1 - xhat K2)
template <typename T>
__global__ void SG_BN_BOUNDS
gp_bn_apply8_kernel(const T* __restrict__ v, const T* __restrict__ da, const T* __restrict__ a_out, const T* __restrict__ y,
                    const float* __restrict__ mr, const float* __restrict__ gamma, const double* __restrict__ sums,
                    const double* __restrict__ tsums, T* __restrict__ w_out, T* __restrict__ gy_out, Chunking k, float slope,
                    float n, float* __restrict__ dgamma) {
    SG_PDL_SYNC();
    extern __shared__ float cst[];                  // [8][C]: mean, r, A, B1, B2, E, K1, K2
    const int C = k.CV * 8;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float mean = mr[c * 2], r = mr[c * 2 + 1];
        if (blockIdx.x == 0 && dgamma != nullptr) {  // the penalty's gamma gradient (was a launch of its own on the main chain)
            const double S1d = sums[c * 2], S2d = sums[c * 2 + 1], nd = (double)n;
            dgamma[c] += (float)((double)r / nd * (nd * tsums[c * 3 + 2] - S1d * tsums[c * 3] - S2d * tsums[c * 3 + 1]));
        }
        const float S1 = (float)sums[c * 2], S2 = (float)sums[c * 2 + 1];
        const float T1 = (float)tsums[c * 3], T2 = (float)tsums[c * 3 + 1], T3 = (float)tsums[c * 3 + 2];
        const float al = gamma[c] * r / n;
        const float P = al * (n * T3 - S1 * T1 - S2 * T2);
        const float sG = -al * (S2 * T1 + S1 * T2), sGx = -2.f * al * S2 * T2;
        cst[c] = mean; cst[C + c] = r; cst[2 * C + c] = al * n; cst[3 * C + c] = al * T1; cst[4 * C + c] = al * T2;
        cst[5 * C + c] = al * S2; cst[6 * C + c] = sG / n; cst[7 * C + c] = sGx / n + P / n;
    }
    __syncthreads();
    if ((int)threadIdx.x >= k.active) return;
    int64_t i = (int64_t)blockIdx.x * k.chunk + threadIdx.x;
    int64_t end = (int64_t)(blockIdx.x + 1) * k.chunk;
    if (end > k.nvec) end = k.nvec;
    const int c0 = ((int)threadIdx.x % k.CV) * 8;
    for (; i < end; i += k.active) {
        const V8 vv = unpack(ldraw(v + i * 8)), d = unpack(ldraw(da + i * 8)), a = unpack(ldraw(a_out + i * 8)),
                 yy = unpack(ldraw(y + i * 8));
        V8 w, gy;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const float mk = a.v[j] > 0.f ? 1.f : slope;
            const float dz = d.v[j] * mk;
            const float r = cst[C + c];
            const float xh = (yy.v[j] - cst[c]) * r;
            const float u = cst[2 * C + c] * vv.v[j] - cst[3 * C + c] - xh * cst[4 * C + c];
            w.v[j] = u * mk;
            const float G = -(cst[5 * C + c] * vv.v[j] + cst[4 * C + c] * dz);
            gy.v[j] = r * (G - cst[6 * C + c] - xh * cst[7 * C + c]);
        }
        st8(w_out + i * 8, w);
        st8(gy_out + i * 8, gy);
    }
}

// ---- gradient-penalty double backward through a train-mode BN as ONE launch (gp_bn_reduce8 + rendezvous + gp_bn_apply8): the
// same scheme as bn_bwd_fused8_kernel -- range b of the vector stream is parked in CTA b's shared memory by bulk async copies
// (FOUR tensors: v, da, a, y), phase 1 reduces (sum v, sum v*xhat, sum v*dz), a rendezvous that counts finished ranges, phase 2
// writes w and gy out of the parked copy; CTA 0 adds the penalty's gamma gradient.  This pass runs while the side streams are
// idle (the second-order chain of a critic update), where the one-launch scheme pays.  tsums must be ZERO on entry.
template <typename T>
__global__ void __launch_bounds__(FNT, 1)
gp_bn_fused8_kernel(const T* __restrict__ v, const T* __restrict__ da, const T* __restrict__ a_out, const T* __restrict__ y,
                    const float* __restrict__ mr, const float* __restrict__ gamma, const double* __restrict__ sums,
                    double* __restrict__ tsums, T* __restrict__ w_out, T* __restrict__ gy_out, float* __restrict__ dgamma,
                    FusedPlan k, float slope, float n, unsigned* __restrict__ work) {
    SG_PDL_SYNC();
    extern __shared__ __align__(128) unsigned char fsm[];
    typedef Raw8<T> R8;
    R8* p_v = reinterpret_cast<R8*>(fsm);
    R8* p_da = p_v + k.keep;
    R8* p_a = p_da + k.keep;
    R8* p_y = p_a + k.keep;
    const int C = k.CV * 8, tid = (int)threadIdx.x;
    float* cst = reinterpret_cast<float*>(p_y + k.keep);                      // [8][C]: mean, r, A, B1, B2, E, K1, K2 (phase 1: rows 0-1)
    float* sacc = cst + 8 * C;                                                // [C][3] segment sums
    __shared__ float part[16][FNT + 1];
    __shared__ uint64_t bars[FPIECES];
    __shared__ int s_cmd;
    __shared__ int mine[FMAXR];
    const int c0 = (tid % k.CV) * 8;
    const bool worker = tid < k.active;
    constexpr int U = Unroll<T>::U > 2 ? 2 : Unroll<T>::U;                    // four tensors per vector: two vectors in flight
    const int piece = k.active * U;
    int nmine = 0;
    bool has0 = false;                               // this CTA reduced range 0: it adds the gamma gradient (exactly one CTA does)
    auto range_of = [&](int r, int64_t& begin, int& len) {
        begin = (int64_t)r * k.range;
        const int64_t left = k.gvec - begin;
        len = (int)(left < k.range ? left : k.range);
    };
    if (tid == 0) {
        for (int p = 0; p < FPIECES; ++p)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fsaddr(&bars[p])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int b = (int)blockIdx.x;
        const bool own = atomicExch(&work[4 + b], 1u) == 0u;
        s_cmd = own ? b : -1;
        if (own) {
            int len;
            int64_t begin;
            range_of(b, begin, len);
            const int left = len < k.keep ? len : k.keep;
            for (int p = 0, v0 = 0; v0 < left; ++p, v0 += piece) {
                const int cnt = left - v0 < piece ? left - v0 : piece;
                const uint32_t bytes = (uint32_t)cnt * (uint32_t)sizeof(R8);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fsaddr(&bars[p])), "r"(bytes * 4u) : "memory");
                fbulk_g2s(p_v + v0, v + (begin + v0) * 8, bytes, &bars[p]);
                fbulk_g2s(p_da + v0, da + (begin + v0) * 8, bytes, &bars[p]);
                fbulk_g2s(p_a + v0, a_out + (begin + v0) * 8, bytes, &bars[p]);
                fbulk_g2s(p_y + v0, y + (begin + v0) * 8, bytes, &bars[p]);
            }
        }
    }
    for (int c = tid; c < C; c += FNT) {             // (mean, rstd): one coalesced pass under the copies in flight
        const float2 q = *reinterpret_cast<const float2*>(mr + (int64_t)c * 2);
        cst[c] = q.x; cst[C + c] = q.y;
    }
    __syncthreads();
    int cmd = s_cmd;

    // ---------------- phase 1: (sum v, sum v*xhat, sum v*dz) of this CTA's range (and of ranges nobody else started)
    while (cmd >= 0) {
        const int r = cmd;
        const int keep = nmine == 0 ? k.keep : 0;
        if (tid == 0) mine[nmine] = r;
        ++nmine;
        if (r == 0) has0 = true;
        int len;
        int64_t begin;
        range_of(r, begin, len);
        for (int t = tid; t < C * 3; t += FNT) sacc[t] = 0.f;
        float t1[8], t2[8], t3[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { t1[j] = 0.f; t2[j] = 0.f; t3[j] = 0.f; }
        int vi = tid;
        if (worker && vi < len) {
            int waited = 0;
            float m[8], rs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { m[j] = cst[c0 + j]; rs[j] = cst[C + c0 + j]; }
            auto landed = [&](int vv) {
                const int need = vv / piece;
                while (waited <= need) { fbar_wait(&bars[waited], 0); ++waited; }
            };
            for (; vi < len; vi += k.active) {
                R8 rv, rd, ra, ry;
                const int64_t at = (begin + vi) * 8;
                if (vi < keep) { landed(vi); rv = p_v[vi]; rd = p_da[vi]; ra = p_a[vi]; ry = p_y[vi]; }
                else { rv = ldraw(v + at); rd = ldraw(da + at); ra = ldraw(a_out + at); ry = ldraw(y + at); }
                const V8 vv = unpack(rv), d = unpack(rd), a = unpack(ra), yy = unpack(ry);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dz = a.v[j] > 0.f ? d.v[j] : slope * d.v[j];
                    t1[j] += vv.v[j];
                    t2[j] += vv.v[j] * ((yy.v[j] - m[j]) * rs[j]);
                    t3[j] += vv.v[j] * dz;
                }
            }
        }
        // cross-thread sums in two rounds through the 16-row scratch: (t1, t2), then t3
#pragma unroll
        for (int j = 0; j < 8; ++j) { part[2 * j][tid] = t1[j]; part[2 * j + 1][tid] = t2[j]; }
        __syncthreads();
        for (int o = tid; o < C * 2; o += FNT) {
            const int c = o >> 1, row = 2 * (c & 7) + (o & 1);
            float acc = 0.f;
            for (int t = c >> 3; t < k.active; t += k.CV) acc += part[row][t];
            sacc[c * 3 + (o & 1)] = acc;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) part[j][tid] = t3[j];
        __syncthreads();
        for (int c = tid; c < C; c += FNT) {
            float acc = 0.f;
            for (int t = c >> 3; t < k.active; t += k.CV) acc += part[c & 7][t];
            sacc[c * 3 + 2] = acc;
        }
        __syncthreads();
        for (int t = tid; t < C * 3; t += FNT) atomicAdd(tsums + t, (double)sacc[t]);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&work[1], 1u);
            s_cmd = fused_wait_or_steal(work, k.nranges, k.steal_ns);        // ---------------- the rendezvous
        }
        __syncthreads();
        cmd = s_cmd;
    }

    // ---------------- phase 2: w = u*act', gy = d/dy of the penalty term (gp_bn_apply8)
    if (nmine > 0) {
        for (int c = tid; c < C; c += FNT) {
            const float mean = cst[c], r = cst[C + c];
            const double S1d = __ldcg(sums + c * 2), S2d = __ldcg(sums + c * 2 + 1);
            const double T1d = __ldcg(tsums + c * 3), T2d = __ldcg(tsums + c * 3 + 1), T3d = __ldcg(tsums + c * 3 + 2);
            if (has0 && dgamma != nullptr)
                dgamma[c] += (float)((double)r / (double)n * ((double)n * T3d - S1d * T1d - S2d * T2d));
            const float S1 = (float)S1d, S2 = (float)S2d, T1 = (float)T1d, T2 = (float)T2d, T3 = (float)T3d;
            const float al = gamma[c] * r / n;
            const float P = al * (n * T3 - S1 * T1 - S2 * T2);
            const float sG = -al * (S2 * T1 + S1 * T2), sGx = -2.f * al * S2 * T2;
            (void)mean;
            cst[2 * C + c] = al * n; cst[3 * C + c] = al * T1; cst[4 * C + c] = al * T2;
            cst[5 * C + c] = al * S2; cst[6 * C + c] = sG / n; cst[7 * C + c] = sGx / n + P / n;
        }
        __syncthreads();
        for (int q = 0; q < nmine; ++q) {
            const int r = mine[q];
            const int keep = q == 0 ? k.keep : 0;
            int len;
            int64_t begin;
            range_of(r, begin, len);
            if (!worker) continue;
            for (int vi = tid; vi < len; vi += k.active) {
                R8 rv, rd, ra, ry;
                const int64_t at = (begin + vi) * 8;
                if (vi < keep) { rv = p_v[vi]; rd = p_da[vi]; ra = p_a[vi]; ry = p_y[vi]; }
                else { rv = ldraw(v + at); rd = ldraw(da + at); ra = ldraw(a_out + at); ry = ldraw(y + at); }
                const V8 vv = unpack(rv), d = unpack(rd), a = unpack(ra), yy = unpack(ry);
                V8 w, gy;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j;
                    const float mk = a.v[j] > 0.f ? 1.f : slope;
                    const float dz = d.v[j] * mk;
                    const float rr = cst[C + c];
                    const float xh = (yy.v[j] - cst[c]) * rr;
                    const float u = cst[2 * C + c] * vv.v[j] - cst[3 * C + c] - xh * cst[4 * C + c];
                    w.v[j] = u * mk;
                    const float G = -(cst[5 * C + c] * vv.v[j] + cst[4 * C + c] * dz);
                    gy.v[j] = rr * (G - cst[6 * C + c] - xh * cst[7 * C + c]);
                }
                st8(w_out + at, w);
                st8(gy_out + at, gy);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&work[2], 1u) == gridDim.x - 1) {
            for (int r = 0; r < k.nranges; ++r) work[4 + r] = 0u;
            work[1] = 0u; work[2] = 0u;
            __threadfence();
        }
    }
}

// plan + launch; -1 = nothing launched (too large to profit / option off): the caller runs gp_bn_reduce8 + gp_bn_apply8
template <typename T>
int gp_bn_fused8(const void* v, const void* da, const void* a_out, const void* y, const float* mr, const float* gamma,
                 const double* sums, double* tsums, void* w_out, void* gy_out, float* dgamma, int64_t rows, int C, int act,
                 unsigned* work, cudaStream_t st) {
    if (!g_gp_bn_fused || work == nullptr) return -1;
    FusedPlan k;
    k.CV = C / 8;
    if (k.CV > FNT || C > 1024) return -1;
    k.gvec = rows * k.CV;
    k.active = FNT / k.CV * k.CV;
    const int64_t per = (k.gvec + SG_NUM_SMS - 1) / SG_NUM_SMS;
    const int64_t range = (per + k.active - 1) / k.active * k.active;
    if (range > (1 << 24)) return -1;
    k.range = (int)range;
    k.rpg = (int)((k.gvec + range - 1) / range);
    k.nranges = k.rpg;
    if (k.nranges > FMAXR) return -1;
    const size_t acc_bytes = (size_t)C * 11 * sizeof(float);      // constants [8][C] + segment sums [C][3]
    const size_t vb = sizeof(Raw8<T>) * 4;
    const size_t budget = 188 * 1024;
    if (acc_bytes + vb * k.active > budget) return -1;
    int64_t keep = (int64_t)((budget - acc_bytes) / vb) / k.active * k.active;
    if (keep > range) keep = range;
    if (keep * 100 < range * g_bn_fused_keep_pct) return -1;
    constexpr int U = Unroll<T>::U > 2 ? 2 : Unroll<T>::U;
    if ((keep + (int64_t)k.active * U - 1) / ((int64_t)k.active * U) > FPIECES) return -1;
    k.keep = (int)keep;
    k.dbg = 0;
    k.steal_ns = g_bn_fused_steal_ns;
    const size_t smem = acc_bytes + vb * (size_t)keep;
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(gp_bn_fused8_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget); set = true; }
    launch_pdl(gp_bn_fused8_kernel<T>, dim3(k.nranges), dim3(FNT), smem, st, (const T*)v, (const T*)da, (const T*)a_out, (const T*)y, mr,
               gamma, sums, tsums, (T*)w_out, (T*)gy_out, dgamma, k, act_slope(act), (float)rows, work);
    g_launches.fetch_add(1);
    return check_launch("gp_bn_fused8");
}
template int gp_bn_fused8<float>(const void*, const void*, const void*, const void*, const float*, const float*, const double*, double*, void*, void*, float*, int64_t, int, int, unsigned*, cudaStream_t);
template int gp_bn_fused8<bf16>(const void*, const void*, const void*, const void*, const float*, const float*, const double*, double*, void*, void*, float*, int64_t, int, int, unsigned*, cudaStream_t);

template <typename T>
int gp_bn_reduce8(const void* v, const void* da, const void* a_out, const void* y, const float* mr, double* tsums, int64_t rows,
                  int C, int act, cudaStream_t st) {
    Chunking k = make_chunking(rows, C, 1, 4);
    launch_pdl(gp_bn_reduce8_kernel<T>, dim3(k.blocks), dim3(256), (size_t)C * 3 * sizeof(float), st, (const T*)v, (const T*)da, (const T*)a_out,
                                                                                  (const T*)y, mr, tsums, k, act_slope(act));
    g_launches.fetch_add(1);
    return check_launch("gp_bn_reduce8");
}
template <typename T>
int gp_bn_apply8(const void* v, const void* da, const void* a_out, const void* y, const float* mr, const float* gamma,
                 const double* sums, const double* tsums, void* w_out, void* gy_out, int64_t rows, int C, int act,
                 float* dgamma, cudaStream_t st) {
    Chunking k = make_chunking(rows, C, 1, 8);
    launch_pdl(gp_bn_apply8_kernel<T>, dim3(k.blocks), dim3(256), (size_t)C * 8 * sizeof(float), st, 
        (const T*)v, (const T*)da, (const T*)a_out, (const T*)y, mr, gamma, sums, tsums, (T*)w_out, (T*)gy_out, k,
        act_slope(act), (float)rows, dgamma);
    g_launches.fetch_add(1);
    return check_launch("gp_bn_apply8");
}
template int gp_bn_reduce8<float>(const void*, const void*, const void*, const void*, const float*, double*, int64_t, int, int, cudaStream_t);
template int gp_bn_reduce8<bf16>(const void*, const void*, const void*, const void*, const float*, double*, int64_t, int, int, cudaStream_t);
template int gp_bn_apply8<float>(const void*, const void*, const void*, const void*, const float*, const float*, const double*, const double*, void*, void*, int64_t, int, int, float*, cudaStream_t);
template int gp_bn_apply8<bf16>(const void*, const void*, const void*, const void*, const float*, const float*, const double*, const double*, void*, void*, int64_t, int, int, float*, cudaStream_t);

// ---- bn_act8_kernel<T, true> (finalize + apply) with the CTA's range brought in by bulk async copies: the whole range (<= 188 KB
// per SM) is in flight at once while the CTA turns the raw sums of its image group into (mean, rstd*gamma, beta) -- the
// register-staged kernel keeps ~40 KB per SM in flight and is latency-bound on the 6-25 MB tensors of the Stage-I critic forward
// (16 / 12 / 12 us for 50 / 25 / 13 MB of traffic).  One CTA per SM (it does not share an SM with other kernels' CTAs, so the
// engines use it only where the side streams are idle: option "bn_act_bulk").  Ranges never cross an image group.
template <typename T>
__global__ void __launch_bounds__(FNT, 1)
bn_act8_bulk_kernel(const T* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ out,
                    FusedPlan k, int act, BnFinalize f) {
    SG_PDL_SYNC();
    extern __shared__ __align__(128) unsigned char fsm[];
    typedef Raw8<T> R8;
    R8* p_y = reinterpret_cast<R8*>(fsm);
    const int C = k.CV * 8, tid = (int)threadIdx.x;
    float* fin = reinterpret_cast<float*>(p_y + k.keep);                      // [C][3]: mean, rstd*gamma, beta of this CTA's group
    __shared__ uint64_t bars[FPIECES];
    constexpr int U = Unroll<T>::U;
    const int piece = k.active * U;
    const int r = (int)blockIdx.x;
    const int g = r / k.rpg;
    const int64_t off = (int64_t)(r % k.rpg) * k.range;
    const int64_t begin = (int64_t)g * k.gvec + off;
    const int len = (int)(k.gvec - off < k.range ? k.gvec - off : k.range);
    if (tid == 0) {
        for (int p = 0; p < FPIECES; ++p)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fsaddr(&bars[p])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int p = 0, v0 = 0; v0 < len; ++p, v0 += piece) {
            const int cnt = len - v0 < piece ? len - v0 : piece;
            const uint32_t bytes = (uint32_t)cnt * (uint32_t)sizeof(R8);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fsaddr(&bars[p])), "r"(bytes) : "memory");
            fbulk_g2s(p_y + v0, y + (begin + v0) * 8, bytes, &bars[p]);
        }
    }
    for (int c = tid; c < C; c += FNT) {
        const int64_t gc = (int64_t)g * C + c;
        double mean, var;
        bn_mean_var(f.stats, gc, f.count, mean, var);
        const float mf = (float)mean, rf = (float)(1.0 / sqrt(var + (double)f.eps));
        fin[c * 3] = mf; fin[c * 3 + 1] = rf * gamma[c]; fin[c * 3 + 2] = beta[c];
    }
    if (blockIdx.x == 0) {                           // what the separate sg_bn_finalize launch did
        for (int idx = tid; idx < f.G * C; idx += FNT) {
            double mean, var;
            bn_mean_var(f.stats, idx, f.count, mean, var);
            f.mr[(int64_t)idx * 2] = (float)mean; f.mr[(int64_t)idx * 2 + 1] = (float)(1.0 / sqrt(var + (double)f.eps));
        }
        if (f.update_running) {
            if (tid == 0 && f.nbt) *f.nbt += f.dup_first + f.G - 1;
            for (int c = tid; c < C; c += FNT) {
                float m_run = f.rm[c], v_run = f.rv[c];
                for (int gg = 0; gg < f.G; ++gg) {
                    double mean, var;
                    bn_mean_var(f.stats, (int64_t)gg * C + c, f.count, mean, var);
                    const float unb = (float)(var * f.count / (f.count > 1 ? f.count - 1 : 1));
                    const int reps = gg == 0 ? f.dup_first : 1;
                    for (int q = 0; q < reps; ++q) {
                        m_run = (1.f - f.momentum) * m_run + f.momentum * (float)mean;
                        v_run = (1.f - f.momentum) * v_run + f.momentum * unb;
                    }
                }
                f.rm[c] = m_run; f.rv[c] = v_run;
            }
        }
    }
    __syncthreads();
    if (tid >= k.active) return;
    const int c0 = (tid % k.CV) * 8;
    float m[8], rg[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = fin[(c0 + j) * 3]; rg[j] = fin[(c0 + j) * 3 + 1]; b[j] = fin[(c0 + j) * 3 + 2]; }
    const float slope = act == SG_ACT_RELU ? 0.f : (act == SG_ACT_LRELU ? 0.1f : 1.f);
    int waited = 0;
    for (int v = tid; v < len; v += k.active) {
        const int need = v / piece;
        while (waited <= need) { fbar_wait(&bars[waited], 0); ++waited; }
        const V8 yy = unpack(p_y[v]);
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = (yy.v[j] - m[j]) * rg[j] + b[j];
            o.v[j] = t > 0.f ? t : slope * t;
        }
        st8(out + (begin + v) * 8, o);
    }
}
int g_bn_act_bulk = 0;             // option "bn_act_bulk": set by the engines around passes that run alone on the GPU

// -1 = nothing launched (option off, residual / tanh, or the tensor does not fit the SMs' shared memory)
template <typename T>
int bn_finalize_act8_bulk(const double* stats, double count, float* mr, float* rm, float* rv, long long* nbt, int dup_first,
                          int update_running, float momentum, float eps, const void* y, const float* gamma, const float* beta,
                          void* out, int64_t rows_per_group, int C, int groups, int act, cudaStream_t st) {
    if (!g_bn_act_bulk || act == SG_ACT_TANH) return -1;
    constexpr int U = Unroll<T>::U;
    FusedPlan k;
    k.CV = C / 8;
    if (k.CV > FNT || groups > SG_NUM_SMS) return -1;
    k.gvec = rows_per_group * k.CV;
    k.active = FNT / k.CV * k.CV;
    const int slots = SG_NUM_SMS / groups;
    const int64_t per = (k.gvec + slots - 1) / slots;
    const int64_t range = (per + k.active - 1) / k.active * k.active;
    k.range = (int)range;
    k.rpg = (int)((k.gvec + range - 1) / range);
    k.nranges = k.rpg * groups;
    const size_t fin_bytes = (size_t)C * 3 * sizeof(float);
    const size_t budget = 200 * 1024;
    if (range > (1 << 24) || fin_bytes + sizeof(Raw8<T>) * (size_t)range > budget) return -1;
    if ((range + (int64_t)k.active * U - 1) / ((int64_t)k.active * U) > FPIECES) return -1;
    if (k.gvec * groups * (int64_t)sizeof(Raw8<T>) < (2 << 20)) return -1;          // tiny tensors: the plain kernel is as fast
    k.keep = (int)range;
    k.dbg = 0; k.steal_ns = 0;
    BnFinalize f{stats, count, mr, rm, rv, nbt, dup_first, update_running, groups, momentum, eps};
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(bn_act8_bulk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget); set = true; }
    launch_pdl(bn_act8_bulk_kernel<T>, dim3(k.nranges), dim3(FNT), fin_bytes + sizeof(Raw8<T>) * (size_t)range, st, (const T*)y, gamma,
               beta, (T*)out, k, act, f);
    g_launches.fetch_add(1);
    return check_launch("bn_finalize_act8_bulk");
}
template int bn_finalize_act8_bulk<float>(const double*, double, float*, float*, float*, long long*, int, int, float, float, const void*, const float*, const float*, void*, int64_t, int, int, int, cudaStream_t);
template int bn_finalize_act8_bulk<bf16>(const double*, double, float*, float*, float*, long long*, int, int, float, float, const void*, const float*, const float*, void*, int64_t, int, int, int, cudaStream_t);

template <typename T>
int bn_act8(const void* y, const float* mr, const float* gamma, const float* beta, const void* res, void* out,
            int64_t rows_per_group, int C, int groups, int act, cudaStream_t st) {
    Chunking k = make_chunking(rows_per_group, C, groups, 8);
    launch_pdl(bn_act8_kernel<T, false>, dim3(k.blocks), dim3(256), 0, st, (const T*)y, mr, gamma, beta, (const T*)res, (T*)out, k, act,
               BnFinalize{});
    g_launches.fetch_add(1);
    return check_launch("bn_act8");
}
// finalize + apply in one launch; returns -1 (nothing launched) when the per-CTA table would not fit
template <typename T>
int bn_finalize_act8(const double* stats, double count, float* mr, float* rm, float* rv, long long* nbt, int dup_first,
                     int update_running, float momentum, float eps, const void* y, const float* gamma, const float* beta,
                     const void* res, void* out, int64_t rows_per_group, int C, int groups, int act, cudaStream_t st) {
    Chunking k = make_chunking(rows_per_group, C, groups, 8);
    size_t smem = (size_t)groups * C * 3 * sizeof(float);
    if (smem > 40 * 1024) return -1;
    BnFinalize f{stats, count, mr, rm, rv, nbt, dup_first, update_running, groups, momentum, eps};
    launch_pdl(bn_act8_kernel<T, true>, dim3(k.blocks), dim3(256), smem, st, (const T*)y, (const float*)nullptr, gamma, beta, (const T*)res,
               (T*)out, k, act, f);
    g_launches.fetch_add(1);
    return check_launch("bn_finalize_act8");
}
template <typename T>
int bn_bwd_reduce8(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma, const float* beta,
                   double* sums, int64_t rows_per_group, int C, int groups, int act, cudaStream_t st) {
    Chunking k = make_chunking(rows_per_group, C, groups, 4);
    size_t smem = (size_t)(groups < 2 ? 1 : 2) * C * 2 * sizeof(float);
    if (k.chunk >= k.gvec) smem = (size_t)groups * C * 2 * sizeof(float);   // tiny tensors: a CTA may span every group
    if (a_out != nullptr)
        launch_pdl(bn_bwd_reduce8_kernel<T, true>, dim3(k.blocks), dim3(256), smem, st, (const T*)da, (const T*)a_out, (const T*)y, mr, gamma, beta,
                                                                    sums, k, act_slope(act));
    else
        launch_pdl(bn_bwd_reduce8_kernel<T, false>, dim3(k.blocks), dim3(256), smem, st, (const T*)da, nullptr, (const T*)y, mr, gamma, beta, sums, k,
                                                                     act_slope(act));
    g_launches.fetch_add(1);
    return check_launch("bn_bwd_reduce8");
}
template <typename T>
int bn_bwd_apply8(const void* da, const void* a_out, const void* y, const float* mr, const float* gamma, const float* beta,
                  const double* sums, const void* inject, int inject_group, void* dy, int64_t rows_per_group, int C, int groups,
                  int act, cudaStream_t st) {
    Chunking k = make_chunking(rows_per_group, C, groups, 8);
    if (a_out != nullptr)
        launch_pdl(bn_bwd_apply8_kernel<T, true>, dim3(k.blocks), dim3(256), 0, st, (const T*)da, (const T*)a_out, (const T*)y, mr, gamma, beta, sums,
                                                                (const T*)inject, inject_group, (T*)dy, k, act_slope(act),
                                                                (float)rows_per_group);
    else
        launch_pdl(bn_bwd_apply8_kernel<T, false>, dim3(k.blocks), dim3(256), 0, st, (const T*)da, nullptr, (const T*)y, mr, gamma, beta, sums,
                                                                 (const T*)inject, inject_group, (T*)dy, k, act_slope(act),
                                                                 (float)rows_per_group);
    g_launches.fetch_add(1);
    return check_launch("bn_bwd_apply8");
}
template <typename T>
int act_bwd8(const void* da, const void* a_out, void* out, int64_t n, int act, cudaStream_t st) {
    int64_t nvec = n / 8;
    launch_pdl(act_bwd8_kernel<T>, dim3(grid_for(nvec, 256, 8)), dim3(256), 0, st, (const T*)da, (const T*)a_out, (T*)out, nvec, act);
    g_launches.fetch_add(1);
    return check_launch("act_bwd8");
}

// explicit instantiations used by elementwise.cu
template int bn_act8<float>(const void*, const float*, const float*, const float*, const void*, void*, int64_t, int, int, int, cudaStream_t);
template int bn_act8<bf16>(const void*, const float*, const float*, const float*, const void*, void*, int64_t, int, int, int, cudaStream_t);
template int bn_bwd_reduce8<float>(const void*, const void*, const void*, const float*, const float*, const float*, double*, int64_t, int, int, int, cudaStream_t);
template int bn_bwd_reduce8<bf16>(const void*, const void*, const void*, const float*, const float*, const float*, double*, int64_t, int, int, int, cudaStream_t);
template int bn_bwd_apply8<float>(const void*, const void*, const void*, const float*, const float*, const float*, const double*, const void*, int, void*, int64_t, int, int, int, cudaStream_t);
template int bn_bwd_apply8<bf16>(const void*, const void*, const void*, const float*, const float*, const float*, const double*, const void*, int, void*, int64_t, int, int, int, cudaStream_t);
template int bn_finalize_act8<float>(const double*, double, float*, float*, float*, long long*, int, int, float, float, const void*, const float*, const float*, const void*, void*, int64_t, int, int, int, cudaStream_t);
template int bn_finalize_act8<bf16>(const double*, double, float*, float*, float*, long long*, int, int, float, float, const void*, const float*, const float*, const void*, void*, int64_t, int, int, int, cudaStream_t);
template int act_bwd8<float>(const void*, const void*, void*, int64_t, int, cudaStream_t);
template int act_bwd8<bf16>(const void*, const void*, void*, int64_t, int, cudaStream_t);

}  // namespace sg
