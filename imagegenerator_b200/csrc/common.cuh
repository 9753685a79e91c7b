// common.cuh -- shared device helpers for libsgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/sgb200.h"

#define SG_NUM_SMS 148

namespace sg {

typedef __nv_bfloat16 bf16;

extern std::atomic<long long> g_launches;
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SG_STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define SG_REQUIRE(cond, ...)                                   \
    do {                                                        \
        if (!(cond)) {                                          \
            sg::set_error(__VA_ARGS__);                         \
            return SG_ERR_BAD_ARG;                              \
        }                                                       \
    } while (0)
// count + post-launch error check
#define SG_LAUNCHED(what)                  \
    do {                                   \
        sg::g_launches.fetch_add(1);       \
        int _e = sg::check_launch(what);   \
        if (_e) return _e;                 \
    } while (0)

// dispatch on storage type
#define SG_DISPATCH_T(dtype, ...)                                         \
    do {                                                                  \
        if ((dtype) == SG_F32) {                                          \
            typedef float T;                                              \
            __VA_ARGS__;                                                  \
        } else if ((dtype) == SG_BF16) {                                  \
            typedef sg::bf16 T;                                           \
            __VA_ARGS__;                                                  \
        } else {                                                          \
            sg::set_error("bad dtype %d", (int)(dtype));                  \
            return SG_ERR_BAD_ARG;                                        \
        }                                                                 \
    } while (0)

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4-element vectors
struct F4 {
    float v[4];
};
__device__ __forceinline__ F4 ld4(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    F4 r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ F4 ld4(const bf16* p) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    F4 r;
    r.v[0] = __uint_as_float(t.x << 16);
    r.v[1] = __uint_as_float(t.x & 0xffff0000u);
    r.v[2] = __uint_as_float(t.y << 16);
    r.v[3] = __uint_as_float(t.y & 0xffff0000u);
    return r;
}
__device__ __forceinline__ void st4(float* p, const F4& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st4(bf16* p, const F4& r) {
    uint2 t;
    t.x = pack_bf16x2(r.v[0], r.v[1]);
    t.y = pack_bf16x2(r.v[2], r.v[3]);
    *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float act_fwd(float x, int act) {
    switch (act) {
        case SG_ACT_RELU: return x > 0.f ? x : 0.f;
        case SG_ACT_LRELU: return x > 0.f ? x : 0.1f * x;
        case SG_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
// derivative of the activation w.r.t. its input, from the stored OUTPUT
__device__ __forceinline__ float act_mask(float a_out, int act) {
    switch (act) {
        case SG_ACT_RELU: return a_out > 0.f ? 1.f : 0.f;
        case SG_ACT_LRELU: return a_out > 0.f ? 1.f : 0.1f;
        case SG_ACT_TANH: return 1.f - a_out * a_out;
        default: return 1.f;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- programmatic dependent launch (PDL): kernels launched through launch_pdl may be scheduled while the previous
// kernel of their stream is still draining its last wave (the hardware releases them once every CTA of that kernel has
// executed launch_dependents or exited); SG_PDL_SYNC() is the point up to which a kernel touches no global memory --
// it lets ITS successor go and then blocks until everything it depends on has completed and is visible.  In a captured
// step this hides the ~2-3 us launch latency between dependent kernels (600 launches per Stage-I step).
#define SG_PDL_SYNC()                                                   \
    do {                                                                \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
        asm volatile("griddepcontrol.wait;" ::: "memory");              \
    } while (0)

extern int g_use_pdl;
extern int g_dbg_skip_memset;
extern int g_bn_act_bulk;
extern int g_bn_fused, g_bn_fused_keep_pct, g_bn_fused_dbg, g_bn_fused_steal_ns, g_gp_bn_fused;      // bn_fast.cu: BatchNorm backward as one launch

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
    int64_t blocks = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)SG_NUM_SMS * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace sg
