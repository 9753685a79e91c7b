// dp_adam.cu -- the data-parallel optimizer step as ONE kernel over NVLink peer memory.
//
// The reference's xm.optimizer_step (stage_1_train_fn.py:149,166-172; stage_2_train_fn.py:155,164,167) is "average this
// optimizer's gradients over the replicas, then step".  Round 1 did that as ncclAllReduce + a local Adam kernel, with the
// CUDA graph of the step cut in two at every collective.  Here the W replicas' flat gradient / parameter buffers live in
// symmetric memory (every rank can address every peer's copy through NVLink / NVSwitch), and rank r owns elements
// [r*chunk, (r+1)*chunk) of every optimizer:
//
//   reduce-scatter   g = (1/W) * sum over peers of grad_peer[i]            -- P2P loads of the owned shard
//   Adam             m, v (owned shard only: the optimizer state is sharded), bias-corrected like torch.optim.Adam
//   all-gather       param_peer[i] = p_new for every peer                  -- P2P stores
//
// in one pass, launched on the compute stream like any other kernel and therefore CUDA-graph capturable: the whole outer
// step stays ONE graph at any world size.  Because every parameter element is computed exactly once (by its owner) and
// copied, replicas stay bit-identical by construction.
//
// Cross-GPU ordering uses two flag words per (optimizer slot, peer) in symmetric memory, written with st.release.sys and
// polled with ld.acquire.sys: "my gradients are complete" before the first peer load, and "my shard is in your memory"
// after the last peer store (which also means: I am done reading your gradients).  Epochs count up, so nothing is ever
// reset across launches or graph replays.  The polls are bounded (a peer that died must not hang this GPU): on time-out
// the kernel raises sync[SYNC_ERR] and carries on; the host checks it (sg_dp_check).
#include "common.cuh"

namespace sg {

constexpr int DP_MAX_WORLD = 16;
constexpr int DP_SLOTS = 8;
// local int32 sync block per slot: [0] epoch of the last completed step, [1] blocks finished in the current launch
constexpr int SYNC_PER_SLOT = 4;
constexpr int SYNC_ERR = DP_SLOTS * SYNC_PER_SLOT;          // one error word after the slots

struct DpArgs {
    float* grads[DP_MAX_WORLD];      // peer pointers to the flat gradient buffer of this optimizer
    float* params[DP_MAX_WORLD];     // peer pointers to the flat parameter buffer
    int* flags[DP_MAX_WORLD];        // peer pointers to the flag block: [slot][phase 2][DP_MAX_WORLD]
    float* m;
    float* v;
    float* hyper;                    // [lr, b1, b2, eps, step, step_lo, step_hi, -]
    int* sync;                       // local: DP_SLOTS * SYNC_PER_SLOT + 1 ints
    long long n, chunk;              // elements of the buffer; elements per rank (multiple of 4)
    int rank, world, slot, write_avg;
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// wait until flag >= e; ~2 s bound
__device__ __forceinline__ bool wait_flag(const int* p, int e) {
    for (int it = 0; it < (1 << 24); ++it) {
        if (ld_acquire_sys(p) - e >= 0) return true;
        __nanosleep(100);
    }
    return false;
}

__global__ void __launch_bounds__(256) dp_adam_kernel(const DpArgs A) {
    __shared__ int s_last;
    int* const epoch_p = A.sync + A.slot * SYNC_PER_SLOT;
    const int e = *reinterpret_cast<volatile int*>(epoch_p) + 1;          // written only after every block of a launch is done
    int* const my_flags = A.flags[A.rank] + (A.slot * 2) * DP_MAX_WORLD;   // [phase][peer]
    // ---- phase 0: my gradients are complete (everything before this kernel on the stream has finished) -> tell every peer,
    //               then wait until every peer has told me
    if (blockIdx.x == 0 && threadIdx.x < A.world)
        st_release_sys(A.flags[threadIdx.x] + (A.slot * 2 + 0) * DP_MAX_WORLD + A.rank, e);
    if (threadIdx.x < A.world) {
        if (!wait_flag(my_flags + threadIdx.x, e)) atomicExch(A.sync + SYNC_ERR, 1);
    }
    __syncthreads();

    const float lr = A.hyper[0], b1 = A.hyper[1], b2 = A.hyper[2], eps = A.hyper[3], t = A.hyper[4] + 1.f;
    const float bc1 = 1.f - powf(b1, t), bc2s = sqrtf(1.f - powf(b2, t));
    const float inv_w = 1.f / (float)A.world;
    const long long lo = (long long)A.rank * A.chunk;
    long long hi = lo + A.chunk;
    if (hi > A.n) hi = A.n;
    for (long long i = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += (long long)gridDim.x * blockDim.x * 4) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < A.world; ++r) {             // fixed order: every rank would compute the same sum
            const float4 x = ld_peer4(A.grads[r] + i);
            g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
        }
        g.x *= inv_w; g.y *= inv_w; g.z *= inv_w; g.w *= inv_w;
        float4 p = *reinterpret_cast<const float4*>(A.params[A.rank] + i);
        float4 mm = *reinterpret_cast<const float4*>(A.m + i), vv = *reinterpret_cast<const float4*>(A.v + i);
        float* pg = &g.x; float* pp = &p.x; float* pm = &mm.x; float* pv = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pm[j] = b1 * pm[j] + (1.f - b1) * pg[j];
            pv[j] = b2 * pv[j] + (1.f - b2) * pg[j] * pg[j];
            const float denom = sqrtf(pv[j]) / bc2s + eps;
            pp[j] -= (lr / bc1) * (pm[j] / denom);
        }
        *reinterpret_cast<float4*>(A.m + i) = mm;
        *reinterpret_cast<float4*>(A.v + i) = vv;
        for (int r = 0; r < A.world; ++r) {
            *reinterpret_cast<float4*>(A.params[r] + i) = p;
            if (A.write_avg) *reinterpret_cast<float4*>(A.grads[r] + i) = g;      // tests: the gradient the optimizer saw
        }
    }
    // ---- phase 1: my shard is in every peer's memory -> tell them (last block of the grid), wait for theirs
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(epoch_p + 1, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < A.world) {
        st_release_sys(A.flags[threadIdx.x] + (A.slot * 2 + 1) * DP_MAX_WORLD + A.rank, e);
        if (!wait_flag(my_flags + DP_MAX_WORLD + threadIdx.x, e)) atomicExch(A.sync + SYNC_ERR, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch_p[1] = 0;
        epoch_p[0] = e;
        // the step counter (adam_tick_kernel's job in the single-GPU path)
        A.hyper[4] = t;
        float s_lo = A.hyper[5] + 1.f;
        if (s_lo >= 8388608.f) { s_lo = 0.f; A.hyper[6] += 1.f; }
        A.hyper[5] = s_lo;
    }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_dp_max_world(void) { return DP_MAX_WORLD; }
int sg_dp_flag_ints(void) { return DP_SLOTS * 2 * DP_MAX_WORLD; }
int sg_dp_sync_ints(void) { return DP_SLOTS * SYNC_PER_SLOT + 1; }

// grad_ptrs / param_ptrs / flag_ptrs: HOST arrays of `world` device pointers (the peers' symmetric buffers, same offset on
// every rank); m, v, hyper, sync: local device memory.  n % 4 == 0.
int sg_dp_adam_step(const void* const* grad_ptrs, const void* const* param_ptrs, const void* const* flag_ptrs, float* m, float* v,
                    float* hyper, int* sync, int64_t n, int rank, int world, int slot, int write_avg, void* stream) {
    SG_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "dp_adam: rank %d / world %d", rank, world);
    SG_REQUIRE(slot >= 0 && slot < DP_SLOTS && n % 4 == 0, "dp_adam: slot %d, n %lld", slot, (long long)n);
    DpArgs A;
    for (int r = 0; r < world; ++r) {
        A.grads[r] = (float*)grad_ptrs[r]; A.params[r] = (float*)param_ptrs[r]; A.flags[r] = (int*)flag_ptrs[r];
    }
    A.m = m; A.v = v; A.hyper = hyper; A.sync = sync; A.n = n;
    A.chunk = ((n + world - 1) / world + 3) / 4 * 4;
    A.rank = rank; A.world = world; A.slot = slot; A.write_avg = write_avg;
    const long long mine = A.chunk / 4;
    int blocks = (int)((mine + 255) / 256);
    if (blocks > 2 * SG_NUM_SMS) blocks = 2 * SG_NUM_SMS;     // every block must be resident while it polls in phase 0
    if (blocks < 1) blocks = 1;
    dp_adam_kernel<<<blocks, 256, 0, SG_STREAM(stream)>>>(A);
    SG_LAUNCHED("dp_adam");
    return 0;
}

}  // extern "C"
