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
                    float n) {
    SG_PDL_SYNC();
    extern __shared__ float cst[];                  // [8][C]: mean, r, A, B1, B2, E, K1, K2
    const int C = k.CV * 8;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float mean = mr[c * 2], r = mr[c * 2 + 1];
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
                 cudaStream_t st) {
    Chunking k = make_chunking(rows, C, 1, 8);
    launch_pdl(gp_bn_apply8_kernel<T>, dim3(k.blocks), dim3(256), (size_t)C * 8 * sizeof(float), st, 
        (const T*)v, (const T*)da, (const T*)a_out, (const T*)y, mr, gamma, sums, tsums, (T*)w_out, (T*)gy_out, k,
        act_slope(act), (float)rows);
    g_launches.fetch_add(1);
    return check_launch("gp_bn_apply8");
}
template int gp_bn_reduce8<float>(const void*, const void*, const void*, const void*, const float*, double*, int64_t, int, int, cudaStream_t);
template int gp_bn_reduce8<bf16>(const void*, const void*, const void*, const void*, const float*, double*, int64_t, int, int, cudaStream_t);
template int gp_bn_apply8<float>(const void*, const void*, const void*, const void*, const float*, const float*, const double*, const double*, void*, void*, int64_t, int, int, cudaStream_t);
template int gp_bn_apply8<bf16>(const void*, const void*, const void*, const void*, const float*, const float*, const double*, const double*, void*, void*, int64_t, int, int, cudaStream_t);

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
