// conv_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate).
//
// One warp-specialised kernel serves the forward direction (fprop, any k/s/p) and the data-gradient
// direction (dgrad == ConvTranspose2d forward, decomposed into s*s output-parity phases so only the
// taps that hit are visited).  No im2col buffer exists anywhere: for every (tap, 64-channel block) the
// TMA producer loads one strided box  {64 ch, bw, bh, bn}  of the NHWC activation tensor straight
// into a 128B-swizzled K-major shared-memory tile -- padding and image borders are the TMA's
// out-of-bounds zero fill, stride-2 sampling is the tensor map's elementStrides -- and the matching
// [BN x 64] slab of the pre-packed weight matrix.  A single elected thread issues tcgen05.mma
// (M=128, N=BN, K=16) into a TMEM accumulator; four epilogue warps read it back with tcgen05.ld,
// apply bias/activation and store bf16 NHWC rows.
//
//   warp 0    : TMA producer            full[s]/empty[s] mbarrier ring, 3-4 stages
//   warp 1    : TMEM alloc + MMA issuer  tcgen05.commit -> empty[s], -> accum_full
//   warps 2-5 : epilogue                 TMEM lane quarter (warp % 4), 16 columns per tcgen05.ld
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <map>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <tuple>
#include <vector>

namespace sg {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused (=1)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

struct TcParams {
    int M, Hq, Wq;        // output grid handled by this launch (per phase): N*Hq*Wq rows
    int bw, bh, bn;       // pixel box of one 128-row tile (bw*bh*bn == 128)
    int n_total, BN;      // output channels, tile width
    int Ck, cblocks;      // reduction channels per tap, ceil(Ck/64)
    int mode;             // 0 fprop, 1 dgrad
    int k, s, p;
    int outH, outW;       // spatial dims of the output tensor
    int act;
    int stages;
    int tmem_cols;
    const float* bias;
    bf16* out;
};

constexpr int TC_THREADS = 192;
constexpr int A_STAGE_BYTES = 128 * 128;

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[4], empty_bar[4], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int b_stage_bytes = P.BN * 128;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;

    // ---- tile / phase bookkeeping (warp-uniform)
    const int phase = blockIdx.z;
    const int ph = phase / P.s, pw = phase - ph * P.s;
    int nth, ntw, rh = 0, rw = 0, base_h = 0, base_w = 0;
    if (P.mode == 0) {
        nth = ntw = P.k;
    } else {
        rh = (ph + P.p) % P.s; rw = (pw + P.p) % P.s;
        nth = (P.k - rh + P.s - 1) / P.s; ntw = (P.k - rw + P.s - 1) / P.s;
        base_h = (ph + P.p - rh) / P.s; base_w = (pw + P.p - rw) / P.s;
    }
    const int nkb = nth * ntw * P.cblocks;
    const int m0 = blockIdx.x * 128;
    const int w0 = m0 % P.Wq, h0 = (m0 / P.Wq) % P.Hq, n0 = m0 / (P.Wq * P.Hq);
    const int nt0 = blockIdx.y * P.BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"((uint32_t)P.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % P.stages, it = kb / P.stages;
                mbar_wait(&empty_bar[st], (it & 1) ^ 1);
                mbar_expect_tx(&full_bar[st], (uint32_t)stage_bytes);
                const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
                const int th = tap / ntw, tw = tap - th * ntw;
                int ca_w, ca_h, bk;
                if (P.mode == 0) {
                    ca_w = w0 * P.s - P.p + tw; ca_h = h0 * P.s - P.p + th;
                    bk = (th * P.k + tw) * P.Ck + cb * 64;
                } else {
                    ca_w = w0 + base_w - tw; ca_h = h0 + base_h - th;
                    bk = ((rh + P.s * th) * P.k + (rw + P.s * tw)) * P.Ck + cb * 64;
                }
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                tma_load_4d(&tmA, &full_bar[st], sa, cb * 64, ca_w, ca_h, n0);
                tma_load_2d(&tmB, &full_bar[st], sa + A_STAGE_BYTES, bk, nt0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.BN >> 3) << 17) | ((128u >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % P.stages, it = kb / P.stages;
                mbar_wait(&full_bar[st], it & 1);
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)st * stage_bytes;
                const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    uint64_t ad = make_kmajor_sw128_desc(sa + k4 * 32);
                    uint64_t bd = make_kmajor_sw128_desc(sb + k4 * 32);
                    tc_mma_bf16(tmem_base, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
                }
                tc_commit(&empty_bar[st]);      // frees the smem stage when these MMAs retire
            }
            tc_commit(&accum_bar);              // accumulator complete
        }
    } else {
        // ---- epilogue: TMEM -> registers -> bias/act -> bf16 NHWC rows
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;            // row of the tile
        const int dn = r / (P.bw * P.bh), rem = r - dn * (P.bw * P.bh);
        const int dh = rem / P.bw, dw = rem - dh * P.bw;
        const int n_img = n0 + dn, hh = h0 + dh, ww = w0 + dw;
        const bool row_ok = n_img < P.M / (P.Hq * P.Wq);   // tiles divide the [Hq][Wq] grid; only the image index can run out
        int oh = hh, ow = ww;
        if (P.mode == 1) { oh = hh * P.s + ph; ow = ww * P.s + pw; }
        bf16* orow = P.out + (((int64_t)n_img * P.outH + oh) * P.outW + ow) * P.n_total;
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        const bool vec_ok = (P.n_total % 8) == 0;
        for (int c0 = 0; c0 < P.BN; c0 += 16) {
            uint32_t v[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            const int n_base = nt0 + c0;
            if (!row_ok || n_base >= P.n_total) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(v[j]);
                int n = n_base + j;
                if (P.bias && n < P.n_total) x += P.bias[n];
                f[j] = act_fwd(x, P.act);
            }
            if (vec_ok && n_base + 16 <= P.n_total) {
                uint4 o0, o1;
                o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]);
                o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
                o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]);
                o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
                *reinterpret_cast<uint4*>(orow + n_base) = o0;
                *reinterpret_cast<uint4*>(orow + n_base + 8) = o1;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n_base + j < P.n_total) orow[n_base + j] = __float2bfloat16_rn(f[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ 2-CTA variant
// Same dataflow with a CTA PAIR (cluster 2x1x1, tcgen05 cta_group::2): the pair owns a 256-row x BN tile,
// each CTA stages its own 128 rows of A and HALF of the weight slab (BN/2 rows); one tcgen05.mma
// (M=256, N=BN, K=16) issued by the leader reads both CTAs' shared memory and writes each CTA's half of
// the accumulator into its own TMEM.  Operand traffic per FLOP halves for B: 32 KB per 128x256x64 block
// per CTA = 128 FLOP/B instead of 64.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0
// relaxed: the producer publishes nothing of its own through this arrival (the TMA's complete_tx carries the data), and the
// default release.cluster form costs a MEMBAR + ERRBAR per k-block on the producer's critical path (ncu source page)
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(bar) & PEER_BIT_MASK),
                 "r"(bytes)
                 : "memory");
}
// wait with back-off: the epilogue warps wait a whole mainloop for their accumulator; polling flat out they steal issue
// slots from the MMA-issuing warp that shares their scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(128);
    }
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc2_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[4], empty_bar[4], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int half_n = P.BN / 2;
    const int b_stage_bytes = half_n * 128;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;

    const int phase = blockIdx.z;
    const int ph = phase / P.s, pw = phase - ph * P.s;
    int nth, ntw, rh = 0, rw = 0, base_h = 0, base_w = 0;
    if (P.mode == 0) {
        nth = ntw = P.k;
    } else {
        rh = (ph + P.p) % P.s; rw = (pw + P.p) % P.s;
        nth = (P.k - rh + P.s - 1) / P.s; ntw = (P.k - rw + P.s - 1) / P.s;
        base_h = (ph + P.p - rh) / P.s; base_w = (pw + P.p - rw) / P.s;
    }
    const int nkb = nth * ntw * P.cblocks;
    const int m0 = blockIdx.x * 128;                 // this CTA's 128 rows (blockIdx.x = 2*pair + rank)
    const int w0 = m0 % P.Wq, h0 = (m0 / P.Wq) % P.Hq, n0 = m0 / (P.Wq * P.Hq);
    const int nt0 = blockIdx.y * P.BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"((uint32_t)P.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % P.stages, it = kb / P.stages;
                mbar_wait(&empty_bar[st], (it & 1) ^ 1);
                mbar_expect_tx_leader(&full_bar[st], (uint32_t)stage_bytes);
                const int tap = kb / P.cblocks, cb = kb - tap * P.cblocks;
                const int th = tap / ntw, tw = tap - th * ntw;
                int ca_w, ca_h, bk;
                if (P.mode == 0) {
                    ca_w = w0 * P.s - P.p + tw; ca_h = h0 * P.s - P.p + th;
                    bk = (th * P.k + tw) * P.Ck + cb * 64;
                } else {
                    ca_w = w0 + base_w - tw; ca_h = h0 + base_h - th;
                    bk = ((rh + P.s * th) * P.k + (rw + P.s * tw)) * P.Ck + cb * 64;
                }
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                tma2_load_4d(&tmA, &full_bar[st], sa, cb * 64, ca_w, ca_h, n0);
                tma2_load_2d(&tmB, &full_bar[st], sa + A_STAGE_BYTES, bk, nt0 + (int)rank * half_n);
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.BN >> 3) << 17) | ((256u >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % P.stages, it = kb / P.stages;
                mbar_wait(&full_bar[st], it & 1);
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)st * stage_bytes;
                const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    uint64_t ad = make_kmajor_sw128_desc(sa + k4 * 32);
                    uint64_t bd = make_kmajor_sw128_desc(sb + k4 * 32);
                    tc2_mma_bf16(tmem_base, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
                }
                tc2_commit_mc(&empty_bar[st]);
            }
            tc2_commit_mc(&accum_bar);
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int dn = r / (P.bw * P.bh), rem = r - dn * (P.bw * P.bh);
        const int dh = rem / P.bw, dw = rem - dh * P.bw;
        const int n_img = n0 + dn, hh = h0 + dh, ww = w0 + dw;
        const bool row_ok = n_img < P.M / (P.Hq * P.Wq);
        int oh = hh, ow = ww;
        if (P.mode == 1) { oh = hh * P.s + ph; ow = ww * P.s + pw; }
        bf16* orow = P.out + (((int64_t)n_img * P.outH + oh) * P.outW + ow) * P.n_total;
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        const bool vec_ok = (P.n_total % 8) == 0;
        for (int c0 = 0; c0 < P.BN; c0 += 16) {
            uint32_t v[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            const int n_base = nt0 + c0;
            if (!row_ok || n_base >= P.n_total) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(v[j]);
                int n = n_base + j;
                if (P.bias && n < P.n_total) x += P.bias[n];
                f[j] = act_fwd(x, P.act);
            }
            if (vec_ok && n_base + 16 <= P.n_total) {
                uint4 o0, o1;
                o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]);
                o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
                o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]);
                o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
                *reinterpret_cast<uint4*>(orow + n_base) = o0;
                *reinterpret_cast<uint4*>(orow + n_base + 8) = o1;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n_base + j < P.n_total) orow[n_base + j] = __float2bfloat16_rn(f[j]);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ persistent kernel
// conv_tcp_kernel<CG>: the same dataflow as a PERSISTENT kernel -- one CTA (CG=1) or CTA pair (CG=2,
// cta_group::2) per SM / SM pair walks a static list of output tiles.  The accumulator is double
// buffered in TMEM (2 x BN columns), so the epilogue warps drain tile i (tcgen05.ld -> bias/act -> bf16
// NHWC rows, plus the per-channel (sum, sum^2) of the BatchNorm that follows, reduced across the 32 rows
// of a warp by a shuffle butterfly) while the MMA warp already accumulates tile i+1; barrier set-up, TMEM
// allocation and the pipeline fill are paid once per SM instead of once per tile.
//
//   warp 0    : TMA producer             full[s] / empty[s] ring (up to 8 stages)
//   warp 1    : MMA issuer (leader CTA)  tcgen05.commit -> empty[s], -> tfull[buf]
//   warps 2-5 : epilogue                 wait tfull[buf] ... arrive tempty[buf] (on the leader)
struct TcpParams {
    int Hq, Wq, n_img;    // output grid handled per phase, images
    int bw, bh, bn;       // pixel box of one 128-row tile
    int n_total, BN;      // output channels, tile width
    int Ck, cblocks;
    int mode, k, s, p;
    int outH, outW;
    int act;
    int stages;
    int tmem_cols, acc_stride, nbuf;      // TMEM columns allocated, columns per accumulator buffer, buffers in the ring
    int m_tiles, n_tiles, total_tiles;   // m_tiles counts CG*128-row cluster tiles
    int imgs_per_group;
    const float* bias;
    bf16* out;
    double* stats;        // [groups][n_total][2] or NULL
    float* out32;         // fp32 result instead of bf16 `out` (col2im input of the thin layers) or NULL
    const bf16* residual; // added before the activation (same layout as out) or NULL
    int lgW, lgHW;        // log2(Wq), log2(Hq*Wq)
    int full_tiles, split, kb_slice;   // tail-wave K-split: tiles below full_tiles are whole; see next_work()
    float* ws;            // fp32 partial tiles of the split tail wave
    int* flags;           // one per (leftover tile, non-owner slice, CTA of the pair)
    int epi_alt;          // narrow tiles: the two epilogue warp groups drain ALTERNATE tiles (see the epilogue)
};

__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// default (.release.cta) semantics on the cluster address, as CUTLASS's ClusterBarrier::arrive does: the explicit
// .release.cluster form costs a cluster-scope MEMBAR (~2600 cycles per tile, measured), and nothing but the TMEM reads --
// already complete after tcgen05.wait::ld and ordered by tcgen05.fence::before_thread_sync -- is handed over here
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// per-column sums over the 32 lanes of a warp for 16 columns held as s[0..15]: recursive halving, 16 shuffles.
// On return s[0] of lane l is the sum of column  8*b4 + 4*b3 + 2*b2 + b1  (b_i = bit i of l); lanes l and l^1 agree.
__device__ __forceinline__ void warp_colsum16(float (&s)[16], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float send = hi ? s[j] : s[j + 8], keep = hi ? s[j + 8] : s[j];
            s[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float send = hi ? s[j] : s[j + 4], keep = hi ? s[j + 4] : s[j];
            s[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float send = hi ? s[j] : s[j + 2], keep = hi ? s[j + 2] : s[j];
            s[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        float send = hi ? s[0] : s[1], keep = hi ? s[1] : s[0];
        s[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
}

struct PhaseGeo {
    int ph, pw, nth, ntw, rh, rw, base_h, base_w;
};
__device__ __forceinline__ PhaseGeo phase_geo(const TcpParams& P, int phase) {
    PhaseGeo g;
    g.ph = phase / P.s; g.pw = phase - g.ph * P.s;
    g.rh = g.rw = g.base_h = g.base_w = 0;
    if (P.mode == 0) {
        g.nth = g.ntw = P.k;
    } else {
        g.rh = (g.ph + P.p) % P.s; g.rw = (g.pw + P.p) % P.s;
        g.nth = (P.k - g.rh + P.s - 1) / P.s; g.ntw = (P.k - g.rw + P.s - 1) / P.s;
        g.base_h = (g.ph + P.p - g.rh) / P.s; g.base_w = (g.pw + P.p - g.rw) / P.s;
    }
    return g;
}

constexpr int TCP_EPI_WARPS = 8;
constexpr int TCP_THREADS = 64 + 32 * TCP_EPI_WARPS;

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Work list of one cluster: whole tiles c, c+C, c+2C, ... below P.full_tiles, then -- when the last round would be
// only partly full -- ONE K-slice of a leftover tile: the R = total - full leftover tiles are each cut into P.split
// slices of P.kb_slice k-blocks, slice s of leftover tile t goes to cluster t*split + s.  Slice 0 is the tile's owner:
// the other slices store their fp32 partial accumulators in a workspace and raise a flag, the owner adds them in its
// epilogue.  Non-owners never wait, every cluster holds at most one slice, all clusters are co-resident: no deadlock.
struct Work {
    int tile, kb0, kb1, slice;
};

__device__ __forceinline__ bool next_work(const TcpParams& P, int cluster_id, int num_clusters, int nkb, int i, Work* w) {
    const int t = cluster_id + i * num_clusters;
    if (t < P.full_tiles) { w->tile = t; w->kb0 = 0; w->kb1 = nkb; w->slice = 0; return true; }
    // split mode: full_tiles is a multiple of num_clusters (or 0), so every cluster reaches this point at the same i
    if (P.split > 1 && t - cluster_id == P.full_tiles) {
        const int lt = cluster_id / P.split, sl = cluster_id - lt * P.split;
        if (P.full_tiles + lt >= P.total_tiles) return false;
        w->tile = P.full_tiles + lt; w->slice = sl;
        w->kb0 = sl * P.kb_slice; w->kb1 = min(nkb, w->kb0 + P.kb_slice);
        return w->kb0 < w->kb1;
    }
    return false;
}

template <int CG>
__global__ void __launch_bounds__(TCP_THREADS, 1)
conv_tcp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcpParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[8], empty_bar[8], tfull_bar[8], tempty_bar[8];
    __shared__ uint32_t tmem_slot;
    __shared__ float sstat[2][256][2];          // statistics staging, double-buffered by tile parity
    __shared__ uint4 sstage[TCP_EPI_WARPS][32 * 4];   // per epilogue warp: 32 rows x 64 B, for the coalesced store

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int b_rows = P.BN / CG;
    const int stage_bytes = A_STAGE_BYTES + b_rows * 128;
    const int tiles_per_phase = P.n_tiles * P.m_tiles;
    const int nkb_tile = (P.mode == 0 ? P.k * P.k : (P.k / P.s) * (P.k / P.s)) * P.cblocks;   // equal for every phase

    if (threadIdx.x == 32) {          // fetch both TMA descriptors while the barriers and TMEM are being set up
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < P.nbuf; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], (P.epi_alt ? TCP_EPI_WARPS / 2 : TCP_EPI_WARPS) * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (P.stats != nullptr)
        for (int i = threadIdx.x; i < 2 * 256 * 2; i += TCP_THREADS) (&sstat[0][0][0])[i] = 0.f;
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                         "r"((uint32_t)P.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                         "r"((uint32_t)P.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    SG_PDL_SYNC();        // barriers, TMEM and the stats staging are set up: from here on global memory is touched

    if (warp == 0) {
        if (lane == 0) {
            int st = 0;
            uint32_t par = 1;                         // parity to wait for on empty[st]: first round passes
            Work w;
            for (int wi = 0; next_work(P, cluster_id, num_clusters, nkb_tile, wi, &w); ++wi) {
                const int phase = w.tile / tiles_per_phase, r = w.tile - phase * tiles_per_phase;
                const int nt = r / P.m_tiles, mt = r - nt * P.m_tiles;
                const PhaseGeo g = phase_geo(P, phase);
                const int m0 = (mt * CG + (int)rank) * 128;
                const int w0 = m0 & (P.Wq - 1), h0 = (m0 >> P.lgW) & (P.Hq - 1), n0 = m0 >> P.lgHW;   // Hq, Wq are powers of two
                const int nt0 = nt * P.BN + (int)rank * b_rows;
                int tap = w.kb0 / P.cblocks, cb = w.kb0 - tap * P.cblocks;
                int th = tap / g.ntw, tw = tap - th * g.ntw;
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
                    int ca_w, ca_h, bk;
                    if (P.mode == 0) {
                        ca_w = w0 * P.s - P.p + tw; ca_h = h0 * P.s - P.p + th;
                        bk = (th * P.k + tw) * P.Ck;
                    } else {
                        ca_w = w0 + g.base_w - tw; ca_h = h0 + g.base_h - th;
                        bk = ((g.rh + P.s * th) * P.k + (g.rw + P.s * tw)) * P.Ck;
                    }
                    mbar_wait(&empty_bar[st], par);
                    uint8_t* sa = smem + (size_t)st * stage_bytes;
                    if (CG == 2) {
                        mbar_expect_tx_leader(&full_bar[st], (uint32_t)stage_bytes);
                        tma2_load_4d(&tmA, &full_bar[st], sa, cb * 64, ca_w, ca_h, n0);
                        tma2_load_2d(&tmB, &full_bar[st], sa + A_STAGE_BYTES, bk + cb * 64, nt0);
                    } else {
                        mbar_expect_tx(&full_bar[st], (uint32_t)stage_bytes);
                        tma_load_4d(&tmA, &full_bar[st], sa, cb * 64, ca_w, ca_h, n0);
                        tma_load_2d(&tmB, &full_bar[st], sa + A_STAGE_BYTES, bk + cb * 64, nt0);
                    }
                    if (++st == P.stages) { st = 0; par ^= 1; }
                    if (++cb == P.cblocks) { cb = 0; if (++tw == g.ntw) { tw = 0; ++th; } }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128*CG
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.BN >> 3) << 17) |
                                   ((uint32_t)((128 * CG) >> 4) << 24);
            int st = 0;
            uint32_t par = 0, ab = 0, abpar = 1;       // accumulator buffer ring: P.nbuf buffers of acc_stride TMEM columns
            Work w;
            for (int wi = 0; next_work(P, cluster_id, num_clusters, nkb_tile, wi, &w); ++wi) {
                mbar_wait(&tempty_bar[ab], abpar);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ab * (uint32_t)P.acc_stride;
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
                    mbar_wait(&full_bar[st], par);
                    tc_fence_after();
                    const uint32_t sa = base + (uint32_t)st * stage_bytes;
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        uint64_t ad = make_kmajor_sw128_desc(sa + k4 * 32);
                        uint64_t bd = make_kmajor_sw128_desc(sb + k4 * 32);
                        const uint32_t acc = ((kb - w.kb0) | k4) != 0 ? 1u : 0u;
                        if (CG == 2) tc2_mma_bf16(tacc, ad, bd, idesc, acc);
                        else tc_mma_bf16(tacc, ad, bd, idesc, acc);
                    }
                    if (CG == 2) tc2_commit_mc(&empty_bar[st]); else tc_commit(&empty_bar[st]);
                    if (++st == P.stages) { st = 0; par ^= 1; }
                }
                if (CG == 2) tc2_commit_mc(&tfull_bar[ab]); else tc_commit(&tfull_bar[ab]);
                if (++ab == (uint32_t)P.nbuf) { ab = 0; abpar ^= 1; }
            }
        }
    } else {
        // ---- epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).. and every second 16-column chunk
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const int dn = r / (P.bw * P.bh), rem = r - dn * (P.bw * P.bh);
        const int dh = rem / P.bw, dw = rem - dh * P.bw;
        const bool vec_ok = (P.n_total % 8) == 0;
        const bool bias_vec_ok = (reinterpret_cast<uintptr_t>(P.bias) & 15) == 0 && (P.BN & 3) == 0;   // float4 loads of the bias
        const int et = threadIdx.x - 64;          // 0..255 among the epilogue threads
        const int act = P.act;
        const bool has_stats = P.stats != nullptr;
        const int scol = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        // coalesced store (fast path): a lane's 64 B of a row go through shared memory so that FOUR lanes write one
        // row's 64 contiguous bytes and a store instruction touches 8 rows, not 32 (the drain of a 128x256 tile was
        // bound by its 4096 L1 wavefronts: one per lane and instruction).  Lane l stores rows q*32 + i*8 + l/4, i<4.
        uint4* const stg = sstage[warp - 2];
        int cdn[4], cdh[4], cdw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = q * 32 + i * 8 + (lane >> 2);
            cdn[i] = rr / (P.bw * P.bh);
            const int rm = rr - cdn[i] * (P.bw * P.bh);
            cdh[i] = rm / P.bw; cdw[i] = rm - cdh[i] * P.bw;
        }
        uint32_t ti = 0, ab = 0, abpar = 0;
        Work w;
        for (int wi = 0; next_work(P, cluster_id, num_clusters, nkb_tile, wi, &w);
             ++wi, ++ti, ab = (ab + 1 == (uint32_t)P.nbuf ? 0 : ab + 1), abpar ^= (ab == 0 ? 1u : 0u)) {
            // epi_alt (narrow tiles without statistics / K-split): warps 2-5 drain the even work items, warps 6-9 the odd
            // ones, each warp all columns of its 32 rows -- the per-tile fixed cost (decode, row pointers, barrier round
            // trip: ~350 of the ~450 instructions a warp spends on a 128x64 tile) is paid by four warps instead of eight
            // and two tiles drain concurrently
            if (P.epi_alt && (int)(ti & 1u) != half) continue;
            const int tile = w.tile;
            // tile decode: the divisions are ~30 instructions each and this code runs per tile and warp (the thin layers'
            // 128x64 tiles were bound by exactly this: 445 instructions per warp and tile, 16 of them the bf16 packs) --
            // the single-phase / single-column-tile cases (warp-uniform) skip them
            int phase = 0, rr = tile, nt = 0, ph = 0, pw = 0;
            if (tiles_per_phase != P.total_tiles) { phase = tile / tiles_per_phase; rr = tile - phase * tiles_per_phase; }
            int mt = rr;
            if (P.n_tiles != 1) { nt = rr / P.m_tiles; mt = rr - nt * P.m_tiles; }
            if (phase != 0) { ph = phase / P.s; pw = phase - ph * P.s; }
            const int m0 = (mt * CG + (int)rank) * 128;
            const int w0 = m0 & (P.Wq - 1), h0 = (m0 >> P.lgW) & (P.Hq - 1), n0 = m0 >> P.lgHW;   // Hq, Wq are powers of two
            const int nt0 = nt * P.BN;
            const int n_img = n0 + dn, hh = h0 + dh, ww = w0 + dw;
            const bool row_ok = n_img < P.n_img && (P.out != nullptr || P.out32 != nullptr);
            int oh = hh, ow = ww;
            if (P.mode == 1) { oh = hh * P.s + ph; ow = ww * P.s + pw; }
            // pixel indices fit 32 bits (checked on the host); one widening multiply per pointer
            bf16* orow = P.out + (int64_t)((n_img * P.outH + oh) * P.outW + ow) * P.n_total + nt0;
            const int ncols = min(P.BN, P.n_total - nt0);          // valid columns of this tile
            const uint32_t sb2 = ti & 1;           // statistics staging buffer
            // K-split bookkeeping: partial tiles of leftover tile lt live at ws[((lt*(split-1) + slice-1)*CG + rank)][128][BN]
            const bool is_split = tile >= P.full_tiles && P.split > 1;
            const int lt = tile - P.full_tiles;
            float* wsrow = nullptr;
            if (is_split) {
                const int first = w.slice > 0 ? w.slice - 1 : 0;
                // layout [partial tile][16-column chunk][row][16]: the 32 lanes of a warp (32 rows) touch 2 KB contiguous
                wsrow = P.ws + ((int64_t)(lt * (P.split - 1) + first) * CG + rank) * (128 * P.BN) + (int64_t)r * 16;
            }
            if (nkb_tile >= 8) mbar_wait_backoff(&tfull_bar[ab], abpar);     // long mainloop: sleep between polls
            else mbar_wait(&tfull_bar[ab], abpar);                          // short tiles: the sleep would be the latency
            tc_fence_after();
            const uint32_t tacc = tmem_base + ab * (uint32_t)P.acc_stride + ((uint32_t)(q * 32) << 16);
            if (is_split && w.slice > 0) {
                // ---- non-owner slice: fp32 partial accumulators to the workspace, then raise the flag
                for (int c0 = half * 16; c0 < ncols; c0 += 32) {
                    uint32_t v[16];
                    tc_ld16(tacc + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        reinterpret_cast<uint4*>(wsrow + c0 * 128)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_leader(&tempty_bar[ab]); else mbar_arrive_local(&tempty_bar[ab]);
                }
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et == 0) {
                    int* flag = P.flags + (lt * (P.split - 1) + (w.slice - 1)) * CG + (int)rank;
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
                }
                continue;
            }
            if (is_split) {
                // ---- owner: wait for the other slices of this tile
                if (et < P.split - 1) {
                    const int* flag = P.flags + (lt * (P.split - 1) + et) * CG + (int)rank;
                    int v;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (!v) __nanosleep(64);
                    } while (!v);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            auto process = [&](const int c0, const uint32_t (&v)[16]) {
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                if (is_split) {
                    for (int sl = 0; sl < P.split - 1; ++sl) {
                        const float4* pp = reinterpret_cast<const float4*>(wsrow + (int64_t)sl * CG * (128 * P.BN) + c0 * 128);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t4 = __ldcg(pp + j);
                            f[4 * j] += t4.x; f[4 * j + 1] += t4.y; f[4 * j + 2] += t4.z; f[4 * j + 3] += t4.w;
                        }
                    }
                }
                if (P.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < ncols) f[j] += __ldg(P.bias + nt0 + c0 + j);
                }
                if (P.residual != nullptr && row_ok) {
                    const bf16* rrow = P.residual + (((int64_t)n_img * P.outH + oh) * P.outW + ow) * P.n_total + nt0 + c0;
                    if (vec_ok && c0 + 16 <= ncols) {
                        const uint4 r0 = *reinterpret_cast<const uint4*>(rrow), r1 = *reinterpret_cast<const uint4*>(rrow + 8);
                        const uint32_t rr2[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            f[2 * j] += __uint_as_float(rr2[j] << 16);
                            f[2 * j + 1] += __uint_as_float(rr2[j] & 0xffff0000u);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < ncols) f[j] += __bfloat162float(rrow[j]);
                    }
                }
                if (act == SG_ACT_LRELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.1f * f[j]);
                } else if (act == SG_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                } else if (act == SG_ACT_TANH) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = tanh_approx(f[j]);
                }
                if (P.out32 != nullptr) {
                    if ((P.n_total % 4) == 0 && c0 + 16 <= ncols) {
                        // a lane's 64 B go through shared memory so that four lanes write one row's 64 contiguous bytes
                        // (8 rows per store instruction instead of 32 half-used sectors)
                        const int sw = (lane >> 1) & 3;
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            stg[lane * 4 + (j ^ sw)] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                                                  __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = i * 8 + (lane >> 2);
                            const uint4 val = stg[rr * 4 + ((lane & 3) ^ ((rr >> 1) & 3))];
                            const int ni = n0 + cdn[i];
                            int oh2 = h0 + cdh[i], ow2 = w0 + cdw[i];
                            if (P.mode == 1) { oh2 = oh2 * P.s + ph; ow2 = ow2 * P.s + pw; }
                            if (ni < P.n_img)
                                *reinterpret_cast<uint4*>(P.out32 + (((int64_t)ni * P.outH + oh2) * P.outW + ow2) * P.n_total + nt0 +
                                                          c0 + (lane & 3) * 4) = val;
                        }
                    } else if (row_ok) {
                        float* o32 = P.out32 + (((int64_t)n_img * P.outH + oh) * P.outW + ow) * P.n_total + nt0 + c0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < ncols) o32[j] = f[j];
                    }
                    return;
                }
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                if (row_ok) {
                    if (vec_ok && c0 + 16 <= ncols) {
                        *reinterpret_cast<uint4*>(orow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < ncols) orow[c0 + j] = __ushort_as_bfloat16((unsigned short)(pk[j >> 1] >> ((j & 1) * 16)));
                    }
                }
                if (has_stats) {
                    float sq[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                     // statistics of the STORED (rounded) values
                        f[2 * j] = row_ok ? __uint_as_float(pk[j] << 16) : 0.f;
                        f[2 * j + 1] = row_ok ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) sq[j] = f[j] * f[j];
                    warp_colsum16(f, lane);
                    warp_colsum16(sq, lane);
                    if ((lane & 1) == 0) {
                        atomicAdd(&sstat[sb2][c0 + scol][0], f[0]);
                        atomicAdd(&sstat[sb2][c0 + scol][1], sq[0]);
                    }
                }
            };
            // ---- fast path (every BatchNorm'ed conv of the training step): no bias / residual / fp32 output / K-split,
            // whole 32-column groups.  Straight-line code over 32 columns per iteration -- two independent 16-column
            // chains for the scheduler to interleave, since only two epilogue warps share an SM sub-partition and the
            // drain is issue-latency bound -- with the activation and the statistics decided once per tile.
            const bool plain = P.residual == nullptr && P.out32 == nullptr && !is_split && vec_ok &&
                               (ncols & 31) == 0 && P.out != nullptr && (P.bias == nullptr || bias_vec_ok);
            if (plain) {
                bf16* crow[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ni = n0 + cdn[i];
                    int oh2 = h0 + cdh[i], ow2 = w0 + cdw[i];
                    if (P.mode == 1) { oh2 = oh2 * P.s + ph; ow2 = ow2 * P.s + pw; }
                    crow[i] = ni < P.n_img ? P.out + (int64_t)((ni * P.outH + oh2) * P.outW + ow2) * P.n_total + nt0 + (lane & 3) * 8
                                           : nullptr;
                }
                for (int c0 = P.epi_alt ? 0 : half * 32; c0 < ncols; c0 += P.epi_alt ? 32 : 64) {
                    uint32_t v[32];
                    tc_ld16_nowait(tacc + (uint32_t)c0, v);
                    tc_ld16_nowait(tacc + (uint32_t)c0 + 16, v + 16);
                    tc_wait_ld();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (P.bias != nullptr) {                          // (first layers: the same 32 values for every lane)
                        const float4* bp = reinterpret_cast<const float4*>(P.bias + nt0 + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(bp + j);
                            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                        }
                    }
                    if (act == SG_ACT_LRELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.1f * f[j]);
                    } else if (act == SG_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    } else if (act == SG_ACT_TANH) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = tanh_approx(f[j]);
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                    {
                        // row = lane: 16-byte chunk j goes to slot j ^ ((lane >> 1) & 3) (bank-conflict-free both ways)
                        const int sw = (lane >> 1) & 3;
                        __syncwarp();
                        stg[lane * 4 + (0 ^ sw)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        stg[lane * 4 + (1 ^ sw)] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        stg[lane * 4 + (2 ^ sw)] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
                        stg[lane * 4 + (3 ^ sw)] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = i * 8 + (lane >> 2);
                            const uint4 val = stg[rr * 4 + ((lane & 3) ^ ((rr >> 1) & 3))];
                            if (crow[i] != nullptr) *reinterpret_cast<uint4*>(crow[i] + c0) = val;
                        }
                    }
                    if (has_stats) {
                        const bool ok = n_img < P.n_img;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            float a[16], sq[16];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {             // statistics of the STORED (rounded) values
                                a[2 * j] = ok ? __uint_as_float(pk[h2 * 8 + j] << 16) : 0.f;
                                a[2 * j + 1] = ok ? __uint_as_float(pk[h2 * 8 + j] & 0xffff0000u) : 0.f;
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) sq[j] = a[j] * a[j];
                            warp_colsum16(a, lane);
                            warp_colsum16(sq, lane);
                            if ((lane & 1) == 0) {
                                atomicAdd(&sstat[sb2][c0 + h2 * 16 + scol][0], a[0]);
                                atomicAdd(&sstat[sb2][c0 + h2 * 16 + scol][1], sq[0]);
                            }
                        }
                    }
                }
            } else
            {
                for (int c0 = P.epi_alt ? 0 : half * 16; c0 < ncols; c0 += P.epi_alt ? 16 : 32) {
                    uint32_t v[16];
                    tc_ld16(tacc + (uint32_t)c0, v);
                    process(c0, v);
                }
            }
            // accumulator buffer drained: hand it back to the MMA warp (of the leader CTA)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_leader(&tempty_bar[ab]); else mbar_arrive_local(&tempty_bar[ab]);
            }
            if (is_split) {
                // every epilogue thread has read its partials: clear the flags for the next launch on this stream
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et < P.split - 1) P.flags[(lt * (P.split - 1) + et) * CG + (int)rank] = 0;
            }
            if (has_stats) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int grp = n0 / P.imgs_per_group;
                for (int i = et; i < ncols * 2; i += 32 * TCP_EPI_WARPS) {
                    const int col = i >> 1, which = i & 1;
                    atomicAdd(P.stats + ((int64_t)grp * P.n_total + nt0 + col) * 2 + which, (double)sstat[sb2][col][which]);
                    sstat[sb2][col][which] = 0.f;
                }
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        if (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dW[co][ci][tap] += sum_pix dy[pix][co] * x[pix@tap][ci]  as  D[co][(tap,ci)] = A^T B with the pixel
// index as the reduction: both operands are "MN-major" (channels contiguous, pixels strided), which
// UMMA reads directly through MN-major 128B-swizzled descriptors -- no transposes.  One CTA owns a
// 128-channel slab of co, one 64-channel block of ci and up to 8 taps (8 x 64 fp32 columns = all 512
// TMEM columns) and walks a slice of the pixels 64 at a time; slices are combined with vector
// fp32 reductions (red.global.add.v4.f32) into the PyTorch-layout gradient.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // between 64-element blocks along M/N
    d |= (uint64_t)(1024 >> 4) << 32;                    // between 8-row groups along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TcWParams {
    int Mpix, Ho, Wo;      // pixels of dy
    int bw, bh, bn;        // pixel box of one 64-pixel k-block
    int Co, Ci, kk, k, s, p;
    int tpg;               // taps per group (<= 8)
    int tap_groups;
    int kb_per_split;      // 64-pixel blocks per split
    float* dw;
};

constexpr int W_A_BYTES = 2 * 64 * 128;     // dy: two 64-channel blocks x 64 pixels
constexpr int W_B_BYTES = 64 * 128;         // x at one tap: 64 channels x 64 pixels
constexpr int W_STAGES = 2;

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const TcWParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[W_STAGES], empty_bar[W_STAGES], accum_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int stage_bytes = W_A_BYTES + P.tpg * W_B_BYTES;

    const int co0 = blockIdx.x * 128;
    const int cib = blockIdx.y / P.tap_groups, tg = blockIdx.y - cib * P.tap_groups;
    const int ci0 = cib * 64, tap0 = tg * P.tpg;
    const int ntap = min(P.tpg, P.kk - tap0);
    const int total_kb = (P.Mpix + 63) / 64;
    const int kb0 = blockIdx.z * P.kb_per_split;
    const int nkb = min(P.kb_per_split, total_kb - kb0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < W_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (nkb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int i = 0; i < nkb; ++i) {
                    const int st = i % W_STAGES, it = i / W_STAGES;
                    mbar_wait(&empty_bar[st], (it & 1) ^ 1);
                    mbar_expect_tx(&full_bar[st], (uint32_t)(W_A_BYTES + ntap * W_B_BYTES));
                    const int pix0 = (kb0 + i) * 64;
                    const int ow0 = pix0 % P.Wo, oh0 = (pix0 / P.Wo) % P.Ho, n0 = pix0 / (P.Wo * P.Ho);
                    uint8_t* sa = smem + (size_t)st * stage_bytes;
                    tma_load_2d(&tmDy, &full_bar[st], sa, co0, pix0);
                    tma_load_2d(&tmDy, &full_bar[st], sa + 64 * 128, co0 + 64, pix0);
                    for (int t = 0; t < ntap; ++t) {
                        const int tap = tap0 + t, kh = tap / P.k, kw = tap - kh * P.k;
                        tma_load_4d(&tmX, &full_bar[st], sa + W_A_BYTES + t * W_B_BYTES, ci0, ow0 * P.s - P.p + kw,
                                    oh0 * P.s - P.p + kh, n0);
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                // D=f32, A=B=bf16, both MN-major, N=64, M=128
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) |
                                       ((128u >> 4) << 24);
                for (int i = 0; i < nkb; ++i) {
                    const int st = i % W_STAGES, it = i / W_STAGES;
                    mbar_wait(&full_bar[st], it & 1);
                    tc_fence_after();
                    const uint32_t sa = base + (uint32_t)st * stage_bytes;
                    for (int t = 0; t < ntap; ++t) {
                        const uint32_t sb = sa + W_A_BYTES + t * W_B_BYTES;
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
                            uint64_t ad = make_mnmajor_sw128_desc(sa + k16 * 2048, 64 * 128);
                            uint64_t bd = make_mnmajor_sw128_desc(sb + k16 * 2048, 64 * 128);
                            tc_mma_bf16(tmem_base + t * 64, ad, bd, idesc, (i | k16) != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty_bar[st]);
                }
                tc_commit(&accum_bar);
            }
        } else {
            const int q = warp & 3;
            const int co = co0 + q * 32 + lane;
            mbar_wait(&accum_bar, 0);
            tc_fence_after();
            const bool vec = (P.kk == 16) && (P.tpg == 8);
            for (int c16 = 0; c16 < 64; c16 += 16) {
                if (ci0 + c16 >= P.Ci) break;            // warp-uniform
                for (int t = 0; t < ntap; t += 4) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (t + u < ntap) tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t + u) * 64 + c16), v[u]);
                    if (co >= P.Co) continue;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ci = ci0 + c16 + j;
                        if (ci >= P.Ci) break;
                        float* dst = P.dw + ((int64_t)co * P.Ci + ci) * P.kk + tap0 + t;
                        if (vec) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v[0][j])),
                                         "f"(__uint_as_float(v[1][j])), "f"(__uint_as_float(v[2][j])),
                                         "f"(__uint_as_float(v[3][j]))
                                         : "memory");
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (t + u < ntap) atomicAdd(dst + u, __uint_as_float(v[u][j]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ wgrad, version 2
// Same GEMM (D[co][(ci,tap)] += dy^T x over pixels, both operands MN-major), re-tiled:
//   * the N dimension is a list of UNITS u = cib*kk + tap (one 64-channel block of ci at one tap = 64 TMEM
//     columns); a CTA owns up to 8 consecutive units whatever k is, so a 3x3 layer (9 taps) is 8+8+...
//     instead of an 8-tap and a 1-tap CTA, and units are dealt out evenly over the CTAs;
//   * PIX = 32-pixel k-blocks: a stage is 8 KB of dy + 4 KB per unit = 40 KB -> 5 stages in flight
//     (the 64-pixel version had 2), which is what hides the latency of ten TMA loads per stage;
//   * split-K sized so that tiles x splits fills the 148 SMs in whole waves.
__device__ __forceinline__ void tma_load_4d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

struct TcW2Params {
    int Mpix, Ho, Wo;
    int Co, Ci, kk, k, s, p;
    int cblocks, units;        // ci blocks of 64, total units = cblocks*kk
    int ugroups, ubase, urem;  // unit groups: group g has ubase + (g < urem) units
    int kb_per_split, total_kb;
    int stages, nun_max;       // pipeline depth; units per stage buffer
    float* dw;
    int dw_cl;                 // 0: dw[Co][Ci][kk] (PyTorch), 1: dw[Co][kk][Ci] (channels-last accumulation buffer)
};

// MC = true: clusters of two CTAs along the co-tile axis (same units, same pixel range, different 128 output channels).
// The x tiles are the same for both, so each CTA fetches every second unit and TMA-multicasts it into both CTAs' shared
// memory: x is 80 % of a stage, so the L2 -> SM operand traffic per CTA drops from 40 KB to 24 KB per k-block.  A stage
// may be refilled once BOTH CTAs' MMAs have retired it (empty barriers count 2, commits are multicast).  When the number
// of co tiles is odd the last cluster's second CTA owns no channels: it only keeps the barrier protocol going.
template <int PIX, bool MC>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const TcW2Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[8], empty_bar[8], accum_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    constexpr int A_BYTES = 2 * PIX * 128;      // dy: two 64-channel blocks x PIX pixels
    constexpr int B_BYTES = PIX * 128;          // x at one unit: 64 channels x PIX pixels

    const int co0 = blockIdx.x * 128;
    const int ug = blockIdx.y;
    const int u0 = ug * P.ubase + min(ug, P.urem);
    const int nun = P.ubase + (ug < P.urem ? 1 : 0);
    const int stage_bytes = A_BYTES + P.nun_max * B_BYTES;
    const int kb0 = blockIdx.z * P.kb_per_split;
    const int nkb = min(P.kb_per_split, P.total_kb - kb0);
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    const bool has_rows = co0 < P.Co;                // false only for the padding CTA of an odd co-tile count (MC)

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], MC ? 2 : 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (nkb > 0) {
        if (warp == 0) {
            // producer warp: lane 0 owns the barriers, lanes 0..nun-1 each issue the TMA load of one unit and lanes
            // nun, nun+1 the two dy blocks -- ten loads per stage leave in parallel instead of one after another
            int uc = 0, ukh = 0, ukw = 0;
            if (lane < nun) {
                const int u = u0 + lane, cib = u / P.kk, tap = u - cib * P.kk;
                uc = cib * 64; ukh = tap / P.k; ukw = tap - ukh * P.k;
            }
            int st = 0;
            uint32_t par = 1;
            int pix0 = kb0 * PIX;
            int ow0 = pix0 % P.Wo, oh0 = (pix0 / P.Wo) % P.Ho, n0 = pix0 / (P.Wo * P.Ho);
            for (int i = 0; i < nkb; ++i) {
                if (lane == 0) {
                    mbar_wait(&empty_bar[st], par);
                    mbar_expect_tx(&full_bar[st], (uint32_t)((has_rows ? A_BYTES : 0) + nun * B_BYTES));
                }
                __syncwarp();
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                if (lane < nun) {
                    if (!MC)
                        tma_load_4d(&tmX, &full_bar[st], sa + A_BYTES + lane * B_BYTES, uc, ow0 * P.s - P.p + ukw,
                                    oh0 * P.s - P.p + ukh, n0);
                    else if ((lane & 1) == (int)rank)            // every second unit, delivered to both CTAs
                        tma_load_4d_mc(&tmX, &full_bar[st], sa + A_BYTES + lane * B_BYTES, uc, ow0 * P.s - P.p + ukw,
                                       oh0 * P.s - P.p + ukh, n0, (uint16_t)3);
                } else if (lane < nun + 2 && has_rows) {
                    tma_load_2d(&tmDy, &full_bar[st], sa + (lane - nun) * (PIX * 128), co0 + (lane - nun) * 64, pix0);
                }
                if (++st == P.stages) { st = 0; par ^= 1; }
                // advance the pixel block (blocks never straddle images: Ho*Wo % PIX == 0 or PIX % (Ho*Wo) == 0)
                pix0 += PIX;
                ow0 += PIX;
                if (ow0 >= P.Wo) {
                    const int rows = ow0 / P.Wo;
                    ow0 -= rows * P.Wo; oh0 += rows;
                    if (oh0 >= P.Ho) { const int imgs = oh0 / P.Ho; oh0 -= imgs * P.Ho; n0 += imgs; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                // D=f32, A=B=bf16, both MN-major, M=128; up to FOUR units (N = 256) per instruction: the unit tiles lie
                // B_BYTES apart in shared memory, which is exactly the descriptor's stride between 64-element blocks
                // along N, so the 128 x 16 slab of dy is read from shared memory twice per k-step instead of 8 times
                // (an N=64 MMA re-reads 4 KB of A for 2 KB of B and is shared-memory bound)
                const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
                const int n_lo = nun < 4 ? nun : 4, n_hi = nun - n_lo;
                const uint32_t idesc_lo = idesc0 | ((uint32_t)((n_lo * 64) >> 3) << 17);
                const uint32_t idesc_hi = idesc0 | ((uint32_t)((n_hi * 64) >> 3) << 17);
                int st = 0;
                uint32_t par = 0;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&full_bar[st], par);
                    tc_fence_after();
                    if (MC && !has_rows) {
                        // padding CTA: nothing to multiply, but both CTAs must release the stage
                        mbar_arrive_cta(&empty_bar[st], 0);
                        mbar_arrive_cta(&empty_bar[st], 1);
                        if (++st == P.stages) { st = 0; par ^= 1; }
                        continue;
                    }
                    const uint32_t sa = base + (uint32_t)st * stage_bytes;
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < PIX / 16; ++k16) {
                        const uint64_t ad = make_mnmajor_sw128_desc(sa + k16 * 2048, PIX * 128);
                        const uint64_t b0 = make_mnmajor_sw128_desc(sb + k16 * 2048, B_BYTES);
                        tc_mma_bf16(tmem_base, ad, b0, idesc_lo, (i | k16) != 0 ? 1u : 0u);
                        if (n_hi > 0) {
                            const uint64_t b1 = make_mnmajor_sw128_desc(sb + 4 * B_BYTES + k16 * 2048, B_BYTES);
                            tc_mma_bf16(tmem_base + 256, ad, b1, idesc_hi, (i | k16) != 0 ? 1u : 0u);
                        }
                    }
                    if (MC) tc_commit_mc(&empty_bar[st], (uint16_t)3); else tc_commit(&empty_bar[st]);
                    if (++st == P.stages) { st = 0; par ^= 1; }
                }
                if (!MC || has_rows) tc_commit(&accum_bar);
            }
        } else if (!MC || has_rows) {
            const int q = warp & 3;
            const int co = co0 + q * 32 + lane;
            mbar_wait(&accum_bar, 0);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            const bool co_ok = co < P.Co;
            if (P.dw_cl) {
                // channels-last accumulation buffer dw[co][tap][ci]: a chunk's 16 columns are contiguous floats
                for (int t = 0; t < nun; ++t) {
                    const int u = u0 + t, cib = u / P.kk, tap = u - cib * P.kk;
                    const int ci0 = cib * 64;
                    for (int c16 = 0; c16 < 64 && ci0 + c16 < P.Ci; c16 += 16) {
                        uint32_t v[16];
                        tc_ld16(trow + (uint32_t)(t * 64 + c16), v);
                        if (!co_ok) continue;
                        float* dst = P.dw + ((int64_t)co * P.kk + tap) * P.Ci + ci0 + c16;
                        if (ci0 + c16 + 16 <= P.Ci) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                                             "f"(__uint_as_float(v[j + 3])) : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ci0 + c16 + j < P.Ci) atomicAdd(dst + j, __uint_as_float(v[j]));
                        }
                    }
                }
            } else if (P.kk == 1) {
                // 1x1: the 16 columns of a chunk are 16 consecutive ci -> contiguous floats of dw[co][:]
                for (int t = 0; t < nun; ++t) {
                    const int ci0 = (u0 + t) * 64;
                    for (int c16 = 0; c16 < 64 && ci0 + c16 < P.Ci; c16 += 16) {
                        uint32_t v[16];
                        tc_ld16(trow + (uint32_t)(t * 64 + c16), v);
                        if (!co_ok) continue;
                        float* dst = P.dw + (int64_t)co * P.Ci + ci0 + c16;
                        if ((P.Ci & 3) == 0 && ci0 + c16 + 16 <= P.Ci) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                                             "f"(__uint_as_float(v[j + 3])) : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ci0 + c16 + j < P.Ci) atomicAdd(dst + j, __uint_as_float(v[j]));
                        }
                    }
                }
            } else if ((P.kk & 3) == 0 && (u0 & 3) == 0) {
                // taps in groups of 4 are contiguous in dw[co][ci][tap]: one vector reduction per (ci, 4 taps)
                for (int t = 0; t < nun; t += 4) {
                    const int u = u0 + t, cib = u / P.kk, tap = u - cib * P.kk;
                    const int ci0 = cib * 64;
                    for (int c16 = 0; c16 < 64 && ci0 + c16 < P.Ci; c16 += 16) {
                        uint32_t v[4][16];
#pragma unroll
                        for (int w = 0; w < 4; ++w) tc_ld16(trow + (uint32_t)((t + w) * 64 + c16), v[w]);
                        if (!co_ok) continue;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int ci = ci0 + c16 + j;
                            if (ci >= P.Ci) break;
                            float* dst = P.dw + ((int64_t)co * P.Ci + ci) * P.kk + tap;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v[0][j])),
                                         "f"(__uint_as_float(v[1][j])), "f"(__uint_as_float(v[2][j])),
                                         "f"(__uint_as_float(v[3][j])) : "memory");
                        }
                    }
                }
            } else {
                for (int t = 0; t < nun; ++t) {
                    const int u = u0 + t, cib = u / P.kk, tap = u - cib * P.kk;
                    const int ci0 = cib * 64;
                    for (int c16 = 0; c16 < 64 && ci0 + c16 < P.Ci; c16 += 16) {
                        uint32_t v[16];
                        tc_ld16(trow + (uint32_t)(t * 64 + c16), v);
                        if (!co_ok) continue;
                        float* dst = P.dw + ((int64_t)co * P.Ci + ci0 + c16) * P.kk + tap;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ci0 + c16 + j < P.Ci) atomicAdd(dst + (int64_t)j * P.kk, __uint_as_float(v[j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();      // MC: no CTA may exit while its peer can still multicast into it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int ensure_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return SG_ERR_UNSUPPORTED;
    }
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

typedef std::tuple<const void*, int, int, int, int, int, int, int, int, int> MapKey;
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

// NHWC bf16 activation tensor [N][H][W][C], box {64, bw*es, bh*es, bn}, element strides {1, es, es, 1}
static int get_act_map(const void* ptr, int N, int H, int W, int C, int bw, int bh, int bn, int es, CUtensorMap* out) {
    MapKey key(ptr, N, H, W, C, bw, bh, bn, es, 4);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(act N=%d H=%d W=%d C=%d box=%d,%d,%d es=%d) failed: %d", N, H, W, C, bw, bh, bn, es,
                  (int)r);
        return SG_ERR_UNSUPPORTED;
    }
    g_maps[key] = m;
    *out = m;
    return 0;
}

// packed weights [rows][Ktot] bf16 (K contiguous), box {64, BN}
static int get_w_map(const void* ptr, int rows, int Ktot, int BN, CUtensorMap* out) {
    MapKey key(ptr, rows, Ktot, BN, 0, 0, 0, 0, 0, 2);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights rows=%d K=%d BN=%d) failed: %d", rows, Ktot, BN, (int)r);
        return SG_ERR_UNSUPPORTED;
    }
    g_maps[key] = m;
    *out = m;
    return 0;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// choose the pixel box of a 128-row tile over an [N][Hq][Wq] grid
static bool choose_box(int Hq, int Wq, int* bw, int* bh, int* bn) {
    if (!is_pow2(Hq) || !is_pow2(Wq)) return false;
    if (Wq >= 128) { *bw = 128; *bh = 1; *bn = 1; return true; }
    *bw = Wq;
    int rows = 128 / Wq;
    if (Hq >= rows) { *bh = rows; *bn = 1; return true; }
    *bh = Hq;
    *bn = rows / Hq;
    return true;
}

static int pick_bn(int n_total) {
    if (n_total <= 256) return (n_total + 15) / 16 * 16;
    return 128;
}

static bool g_attr_set = false;
static bool g_attr2_set = false;
int g_use_tc2 = 1;     // 2-CTA (cta_group::2) tiles for wide layers; sg_set_option("tc2", 0) disables

// mode 0: fprop (act = x [N][H][W][Ck], out = y [N][Ho][Wo][n_total]);
// mode 1: dgrad (act = dy [N][Ho][Wo][Ck], out = dx [N][H][W][n_total])
static int launch_conv_tc(int mode, const void* act, const void* wpack, const float* bias, void* out, int N, int H, int W,
                          int Ci, int Ho, int Wo, int Co, int k, int s, int p, int actf, cudaStream_t st) {
    int e = ensure_encode();
    if (e) return e;
    TcParams P;
    int aH, aW;                 // spatial dims of the operand tensor the TMA reads
    if (mode == 0) {
        P.Hq = Ho; P.Wq = Wo; P.Ck = Ci; P.n_total = Co; P.outH = Ho; P.outW = Wo; aH = H; aW = W;
    } else {
        P.Hq = H / s; P.Wq = W / s; P.Ck = Co; P.n_total = Ci; P.outH = H; P.outW = W; aH = Ho; aW = Wo;
    }
    if (!choose_box(P.Hq, P.Wq, &P.bw, &P.bh, &P.bn)) { set_error("conv_tc: grid %dx%d not tileable", P.Hq, P.Wq); return SG_ERR_UNSUPPORTED; }
    P.M = N * P.Hq * P.Wq;
    P.BN = pick_bn(P.n_total);
    P.cblocks = (P.Ck + 63) / 64;
    P.mode = mode; P.k = k; P.s = s; P.p = p; P.act = actf; P.bias = bias; P.out = (bf16*)out;
    P.stages = P.BN <= 128 ? 3 : 4;
    P.tmem_cols = 32;
    while (P.tmem_cols < P.BN) P.tmem_cols *= 2;
    CUtensorMap tmA, tmB;
    int es = mode == 0 ? s : 1;
    if ((e = get_act_map(act, N, aH, aW, P.Ck, P.bw, P.bh, P.bn, es, &tmA))) return e;
    int rows = P.n_total, Ktot = k * k * P.Ck;
    if ((e = get_w_map(wpack, rows, Ktot, P.BN, &tmB))) return e;
    size_t smem = (size_t)P.stages * (A_STAGE_BYTES + P.BN * 128) + 1024;
    if (!g_attr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return (int)ce; }
        g_attr_set = true;
    }
    int phases = mode == 0 ? 1 : s * s;
    dim3 grid((P.M + 127) / 128, (P.n_total + P.BN - 1) / P.BN, phases);
    conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmA, tmB, P);
    g_launches.fetch_add(1);
    return check_launch("conv_tc");
}

// CTA-pair launch: 256-row x BN tiles (BN = 128 or 256)
static int launch_conv_tc2(int mode, const void* act, const void* wpack, const float* bias, void* out, int N, int H, int W,
                           int Ci, int Ho, int Wo, int Co, int k, int s, int p, int actf, cudaStream_t st) {
    int e = ensure_encode();
    if (e) return e;
    TcParams P;
    int aH, aW;
    if (mode == 0) {
        P.Hq = Ho; P.Wq = Wo; P.Ck = Ci; P.n_total = Co; P.outH = Ho; P.outW = Wo; aH = H; aW = W;
    } else {
        P.Hq = H / s; P.Wq = W / s; P.Ck = Co; P.n_total = Ci; P.outH = H; P.outW = W; aH = Ho; aW = Wo;
    }
    if (!choose_box(P.Hq, P.Wq, &P.bw, &P.bh, &P.bn)) { set_error("conv_tc2: grid not tileable"); return SG_ERR_UNSUPPORTED; }
    P.M = N * P.Hq * P.Wq;
    {   // widest tile (<= 256, multiple of 32) that wastes the fewest padded columns
        int best = 128, best_waste = 1 << 30;
        for (int bn = 256; bn >= 128; bn -= 32) {
            int waste = (P.n_total + bn - 1) / bn * bn - P.n_total;
            if (waste < best_waste) { best_waste = waste; best = bn; }
        }
        P.BN = best;
    }
    P.cblocks = (P.Ck + 63) / 64;
    P.mode = mode; P.k = k; P.s = s; P.p = p; P.act = actf; P.bias = bias; P.out = (bf16*)out;
    P.stages = 3;
    P.tmem_cols = 32;
    while (P.tmem_cols < P.BN) P.tmem_cols *= 2;
    CUtensorMap tmA, tmB;
    int es = mode == 0 ? s : 1;
    if ((e = get_act_map(act, N, aH, aW, P.Ck, P.bw, P.bh, P.bn, es, &tmA))) return e;
    if ((e = get_w_map(wpack, P.n_total, k * k * P.Ck, P.BN / 2, &tmB))) return e;
    size_t smem = (size_t)P.stages * (A_STAGE_BYTES + (P.BN / 2) * 128) + 1024;
    if (!g_attr2_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(tc2): %s", cudaGetErrorString(ce)); return (int)ce; }
        g_attr2_set = true;
    }
    int phases = mode == 0 ? 1 : s * s;
    int mt = (P.M + 127) / 128;
    mt = (mt + 1) / 2 * 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(mt, (P.n_total + P.BN - 1) / P.BN, phases);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t ce = cudaLaunchKernelEx(&cfg, conv_tc2_kernel, tmA, tmB, P);
    if (ce != cudaSuccess) { set_error("conv_tc2 launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    g_launches.fetch_add(1);
    return check_launch("conv_tc2");
}

// ---- persistent launch: pick (CTA group, tile width) with a small cost model --------------------------
// cycles per 64-deep k-block: tcgen05 issue = 2*BN (M=128 per CTA, either group size); operand fetch from L2 =
// (16 KB of A + BN/CG rows of B) at ~44 B/cycle/SM (the ~12 TB/s L2->SM ceiling shared by 148 SMs).
int g_use_persist = 1;
int g_force_cg = 0, g_force_bn = 0, g_force_stages = 0, g_dbg = 0;
// alternate-tile epilogue for narrow tiles: OFF by default (option "epi_alt" / env SG_EPI_ALT=1) until it has been through the
// full GPU test suite
int g_epi_alt = getenv("SG_EPI_ALT") ? atoi(getenv("SG_EPI_ALT")) : 0;
static bool g_pattr_set = false;

// ---- tail-wave K-split scratch: fp32 partial tiles + flags, one slot per stream that launches split kernels (kernels of
// one stream are ordered; kernels of different streams may overlap and must not share partials).  Allocated once per
// device by sg_check_device() -- never inside a launch, so launches stay capturable.
constexpr int WS_SLOTS = 4;
constexpr size_t WS_BYTES = (size_t)SG_NUM_SMS * 128 * 256 * sizeof(float);     // >= (split-1)*R*CG partial tiles of 128 x 256
struct WsSlot {
    float* ws = nullptr;
    int* flags = nullptr;
    cudaStream_t stream = nullptr;
    bool used = false;
};
static WsSlot g_ws[16][WS_SLOTS];
static std::mutex g_ws_mu;
int g_use_split = 1;

int tcp_workspace_init() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 0;
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (int i = 0; i < WS_SLOTS; ++i) {
        if (g_ws[dev][i].ws) continue;
        if (cudaMalloc(&g_ws[dev][i].ws, WS_BYTES) != cudaSuccess) { g_ws[dev][i].ws = nullptr; cudaGetLastError(); return 0; }
        if (cudaMalloc(&g_ws[dev][i].flags, 4 * SG_NUM_SMS * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return 0; }
        cudaMemset(g_ws[dev][i].flags, 0, 4 * SG_NUM_SMS * sizeof(int));
    }
    return 0;
}
static WsSlot* ws_slot_for(cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (int i = 0; i < WS_SLOTS; ++i)
        if (g_ws[dev][i].used && g_ws[dev][i].stream == st) return g_ws[dev][i].ws ? &g_ws[dev][i] : nullptr;
    for (int i = 0; i < WS_SLOTS; ++i)
        if (!g_ws[dev][i].used && g_ws[dev][i].ws) { g_ws[dev][i].used = true; g_ws[dev][i].stream = st; return &g_ws[dev][i]; }
    return nullptr;
}
// slices per leftover tile: as many as there are idle clusters per leftover tile, >= 4 k-blocks each
// returns S and the estimated duration of the split round in cycles (t_kb: cycles per k-block)
static int split_factor(long tiles, int units, int nkb, int bn = 256, double t_kb = 600.0, double* t_round = nullptr) {
    const int R = (int)(tiles % units);
    double best_t = nkb * t_kb;
    int best = 1;
    if (g_use_split && R != 0) {
        int smax = units / R;
        if (smax > nkb / 4) smax = nkb / 4;
        if (smax > 6) smax = 6;
        // a slice costs its share of the mainloop; a non-owner then drains its accumulator to the workspace (measured
        // ~5.5 us for 128 x 256 fp32), the owner's drain grows by ~3 us per partial it adds (critic ds3 trace)
        const double cs = g_use_split == 2 ? 0.0 : 1.0;      // option split=2: ignore the costs (experiments)
        const double c_store = cs * 10000.0 * bn / 256.0, c_read = cs * 5700.0 * bn / 256.0;
        for (int S = 2; S <= smax; ++S) {
            const int kbs = (nkb + S - 1) / S;
            const double t = kbs * t_kb + c_store + (S - 1) * c_read + 1000.0;
            if (t < best_t * 0.9) { best_t = t; best = S; }
        }
    }
    if (t_round) *t_round = best_t;
    return best;
}

static void pick_tcp_config(int M, int n_total, int phases, int nkb, int* cg_out, int* bn_out) {
    double best = 1e30;
    int best_cg = 1, best_bn = 16;
    for (int cg = 1; cg <= 2; ++cg) {
        if (g_force_cg && cg != g_force_cg) continue;
        const int step = 16;
        for (int bn = 16; bn <= 256; bn += step) {
            if (g_force_bn && bn != g_force_bn) continue;
            if (cg == 2 && bn < 32) continue;
            int n_tiles = (n_total + bn - 1) / bn;
            if (!g_force_bn && n_tiles * bn - n_total >= 16 && bn > 16) {
                // a narrower tile with the same tile count wastes less
                int alt = ((n_total + n_tiles - 1) / n_tiles + 15) / 16 * 16;
                if (alt < bn) continue;
            }
            int m_tiles = (M + 128 * cg - 1) / (128 * cg);
            long tiles = (long)m_tiles * n_tiles * phases;
            int units = SG_NUM_SMS / cg;
            double mma = 2.0 * bn;
            double l2 = (16384.0 + (double)(bn / cg) * 128.0) / 44.0;
            double t_kb = mma > l2 ? mma : l2;
            double t_tile = nkb * t_kb + 150.0;
            double epi = 40.0 * bn / 16.0 + 400.0;                  // drain of the last tile, not overlapped
            double t_epi_tile = 40.0 * bn / 16.0 + 100.0;           // epilogue pace per tile
            if (t_epi_tile > t_tile) t_tile = t_epi_tile;
            // rounds of the persistent tile loop; a partly filled last round may be cut along K over the idle clusters
            double t_last = 0.0;
            if (tiles % units) {
                split_factor(tiles, units, nkb, bn, t_kb, &t_last);
                t_last += 150.0;
            }
            double t = (double)(tiles / units) * t_tile + t_last + epi;
            if (t < best) { best = t; best_cg = cg; best_bn = bn; }
        }
    }
    *cg_out = best_cg; *bn_out = best_bn;
}

static int launch_conv_tcp(int mode, const void* act, const void* wpack, const float* bias, void* out, int N, int H, int W,
                           int Ci, int Ho, int Wo, int Co, int k, int s, int p, int actf, double* stats, int groups,
                           cudaStream_t st, float* out32 = nullptr, const void* residual = nullptr) {
    int e = ensure_encode();
    if (e) return e;
    TcpParams P;
    int aH, aW;
    if (mode == 0) {
        P.Hq = Ho; P.Wq = Wo; P.Ck = Ci; P.n_total = Co; P.outH = Ho; P.outW = Wo; aH = H; aW = W;
    } else {
        P.Hq = H / s; P.Wq = W / s; P.Ck = Co; P.n_total = Ci; P.outH = H; P.outW = W; aH = Ho; aW = Wo;
    }
    if (!choose_box(P.Hq, P.Wq, &P.bw, &P.bh, &P.bn)) { set_error("conv_tcp: grid %dx%d not tileable", P.Hq, P.Wq); return SG_ERR_UNSUPPORTED; }
    const int M = N * P.Hq * P.Wq;
    if ((int64_t)(N + 128) * P.outH * P.outW >= (1ll << 31)) {       // the epilogue indexes output pixels with 32 bits
        set_error("conv_tcp: %d x %d x %d output pixels exceed the 32-bit pixel index", N, P.outH, P.outW);
        return SG_ERR_UNSUPPORTED;
    }
    P.n_img = N;
    P.lgW = 0; while ((1 << P.lgW) < P.Wq) ++P.lgW;
    P.lgHW = P.lgW; while ((1 << P.lgHW) < P.Hq * P.Wq) ++P.lgHW;
    P.cblocks = (P.Ck + 63) / 64;
    P.mode = mode; P.k = k; P.s = s; P.p = p; P.act = actf; P.bias = bias; P.out = (bf16*)out;
    P.stats = stats;
    P.out32 = out32;
    P.residual = (const bf16*)residual;
    if (out32) P.out = nullptr;
    P.imgs_per_group = groups > 0 ? N / groups : N;
    const int phases = mode == 0 ? 1 : s * s;
    const int taps = mode == 0 ? k * k : (k / s) * (k / s);
    int cg, bn;
    pick_tcp_config(M, P.n_total, phases, taps * P.cblocks, &cg, &bn);
    P.BN = bn;
    P.n_tiles = (P.n_total + bn - 1) / bn;
    P.m_tiles = (M + 128 * cg - 1) / (128 * cg);
    P.total_tiles = P.m_tiles * P.n_tiles * phases;
    P.acc_stride = bn;
    // accumulator ring in TMEM: 2 buffers for wide tiles, up to 8 for narrow ones -- with short mainloops (thin layers,
    // 1x1 GEMMs) the MMA warp must be able to run several tiles ahead of the epilogue to hide the barrier round trips
    P.nbuf = 512 / bn;
    if (P.nbuf > 8) P.nbuf = 8;
    if (P.nbuf < 2) P.nbuf = 2;
    P.tmem_cols = 32;
    while (P.tmem_cols < P.nbuf * bn) P.tmem_cols *= 2;
    const int stage_bytes = A_STAGE_BYTES + (bn / cg) * 128;
    P.stages = (200 * 1024 - 2048) / stage_bytes;
    if (P.stages > 8) P.stages = 8;
    if (g_force_stages && g_force_stages < P.stages) P.stages = g_force_stages;
    if (g_dbg & 1) P.out = nullptr;
    CUtensorMap tmA, tmB;
    const int es = mode == 0 ? s : 1;
    if ((e = get_act_map(act, N, aH, aW, P.Ck, P.bw, P.bh, P.bn, es, &tmA))) return e;
    if ((e = get_w_map(wpack, P.n_total, k * k * P.Ck, bn / cg, &tmB))) return e;
    size_t smem = (size_t)P.stages * stage_bytes + 1024;
    if (!g_pattr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_tcp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tcp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(tcp): %s", cudaGetErrorString(ce)); return (int)ce; }
        g_pattr_set = true;
    }
    int units = SG_NUM_SMS / cg;
    int clusters = P.total_tiles < units ? P.total_tiles : units;
    P.full_tiles = P.total_tiles; P.split = 1; P.kb_slice = taps * P.cblocks; P.ws = nullptr; P.flags = nullptr;
    {
        const int nkb = taps * P.cblocks;
        const double mma_c = 2.0 * bn, l2_c = (16384.0 + (double)(bn / cg) * 128.0) / 44.0;
        int S = split_factor(P.total_tiles, units, nkb, bn, mma_c > l2_c ? mma_c : l2_c);
        WsSlot* slot = S > 1 ? ws_slot_for(st) : nullptr;
        if (slot != nullptr) {
            const int F = P.total_tiles / units, R = P.total_tiles % units;
            P.kb_slice = (nkb + S - 1) / S;
            S = (nkb + P.kb_slice - 1) / P.kb_slice;
            if (S > 1 && (size_t)(S - 1) * R * cg * 128 * bn * sizeof(float) <= WS_BYTES) {
                P.split = S; P.full_tiles = F * units; P.ws = slot->ws; P.flags = slot->flags;
                clusters = F > 0 ? units : R * S;
            }
        }
    }
    // two tiles in flight in the epilogue need two accumulator buffers beyond the one being filled: nbuf >= 4 (bn <= 128)
    P.epi_alt = (g_epi_alt && stats == nullptr && P.split == 1 && bn <= 64 && P.nbuf >= 4) ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * cg, 1, 1);
    cfg.blockDim = dim3(TCP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cudaError_t ce = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_tcp_kernel<2>, tmA, tmB, P)
                             : cudaLaunchKernelEx(&cfg, conv_tcp_kernel<1>, tmA, tmB, P);
    if (ce != cudaSuccess) { set_error("conv_tcp launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    g_launches.fetch_add(1);
    return check_launch("conv_tcp");
}

static bool choose_box64(int Ho, int Wo, int* bw, int* bh, int* bn) {
    if (!is_pow2(Ho) || !is_pow2(Wo)) return false;
    if (Wo >= 64) { *bw = 64; *bh = 1; *bn = 1; return true; }
    *bw = Wo;
    int rows = 64 / Wo;
    if (Ho >= rows) { *bh = rows; *bn = 1; return true; }
    *bh = Ho;
    *bn = rows / Ho;
    return true;
}

// 2-D [rows][C] bf16 view, box {64, 64}
static int get_rows_map(const void* ptr, int64_t rows, int C, CUtensorMap* out) {
    MapKey key(ptr, (int)rows, C, 64, 64, 0, 0, 0, 0, 22);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rows=%lld C=%d) failed: %d", (long long)rows, C, (int)r); return SG_ERR_UNSUPPORTED; }
    g_maps[key] = m;
    *out = m;
    return 0;
}

static bool g_wattr_set = false;

static int launch_wgrad_tc(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                           int k, int s, int p, cudaStream_t st) {
    int e = ensure_encode();
    if (e) return e;
    TcWParams P;
    P.Mpix = N * Ho * Wo; P.Ho = Ho; P.Wo = Wo; P.Co = Co; P.Ci = Ci; P.kk = k * k; P.k = k; P.s = s; P.p = p; P.dw = dw;
    if (!choose_box64(Ho, Wo, &P.bw, &P.bh, &P.bn)) { set_error("wgrad_tc: grid not tileable"); return SG_ERR_UNSUPPORTED; }
    P.tpg = P.kk < 8 ? P.kk : 8;
    P.tap_groups = (P.kk + P.tpg - 1) / P.tpg;
    CUtensorMap tmDy, tmX;
    if ((e = get_rows_map(dy, P.Mpix, Co, &tmDy))) return e;
    if ((e = get_act_map(x, N, H, W, Ci, P.bw, P.bh, P.bn, s, &tmX))) return e;
    int co_tiles = (Co + 127) / 128, ci_blocks = (Ci + 63) / 64;
    int tiles = co_tiles * ci_blocks * P.tap_groups;
    int total_kb = (P.Mpix + 63) / 64;
    int want = (SG_NUM_SMS + tiles - 1) / tiles;          // about one wave of CTAs
    int max_splits = (total_kb + 3) / 4;                  // at least 4 k-blocks per CTA
    int splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
    if (splits < 1) splits = 1;
    P.kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + P.kb_per_split - 1) / P.kb_per_split;
    size_t smem = (size_t)W_STAGES * (W_A_BYTES + P.tpg * W_B_BYTES) + 1024;
    if (!g_wattr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(ce)); return (int)ce; }
        g_wattr_set = true;
    }
    dim3 grid(co_tiles, ci_blocks * P.tap_groups, splits);
    conv_wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmDy, tmX, P);
    g_launches.fetch_add(1);
    return check_launch("conv_wgrad_tc");
}

static bool choose_box_n(int npix, int Ho, int Wo, int* bw, int* bh, int* bn) {
    if (!is_pow2(Ho) || !is_pow2(Wo)) return false;
    if (Wo >= npix) { *bw = npix; *bh = 1; *bn = 1; return true; }
    *bw = Wo;
    int rows = npix / Wo;
    if (Ho >= rows) { *bh = rows; *bn = 1; return true; }
    *bh = Ho;
    *bn = rows / Ho;
    return true;
}

// 2-D [rows][C] bf16 view, box {64, npix}
static int get_rows_map_n(const void* ptr, int64_t rows, int C, int npix, CUtensorMap* out) {
    MapKey key(ptr, (int)rows, C, 64, npix, 0, 0, 0, 0, 23);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)npix};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rows=%lld C=%d) failed: %d", (long long)rows, C, (int)r); return SG_ERR_UNSUPPORTED; }
    g_maps[key] = m;
    *out = m;
    return 0;
}

int g_use_wgrad2 = 1;
int g_use_wgrad_mc = 1;

template <int PIX>
static int launch_wgrad2_pix(const void* x, const void* dy, TcW2Params P, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                             int s, cudaStream_t st) {
    int bw, bh, bn, e;
    if (!choose_box_n(PIX, Ho, Wo, &bw, &bh, &bn)) { set_error("wgrad2: grid not tileable"); return SG_ERR_UNSUPPORTED; }
    P.total_kb = (P.Mpix + PIX - 1) / PIX;
    const int stage_bytes = PIX * 128 * (2 + P.nun_max);
    P.stages = (200 * 1024) / stage_bytes;
    if (P.stages > 8) P.stages = 8;
    CUtensorMap tmDy, tmX;
    if ((e = get_rows_map_n(dy, P.Mpix, Co, PIX, &tmDy))) return e;
    if ((e = get_act_map(x, N, H, W, Ci, bw, bh, bn, s, &tmX))) return e;
    const int co_tiles = (Co + 127) / 128;
    // pairs of co tiles share their x tiles through TMA multicast when there are at least two of them
    // (even tile counts only: with a padding CTA in the last cluster the launch measured up to 1.9x slower than unicast)
    const bool mc = g_use_wgrad_mc && co_tiles >= 2 && (co_tiles % 2) == 0;
    const int tiles = (mc ? (co_tiles + 1) / 2 * 2 : co_tiles) * P.ugroups;
    // splits: fill whole waves of 148 CTAs.  Fixed cost per CTA in 32-pixel k-block units (0.6 us each): prologue + the
    // atomics epilogue, ~10 us with vector reductions, ~40 us with scalar ones (k*k not a multiple of 4, PyTorch layout)
    const bool vec_epi = P.dw_cl || P.kk == 1 || (P.kk & 3) == 0;
    const double fixed = (vec_epi ? 18.0 : 66.0) * 32.0 / PIX;
    int max_splits = P.total_kb / 4;
    if (max_splits < 1) max_splits = 1;
    int best = 1;
    double best_cost = 1e30;
    for (int sp = 1; sp <= max_splits && sp <= 148; ++sp) {
        long ctas = (long)tiles * sp;
        long waves = (ctas + SG_NUM_SMS - 1) / SG_NUM_SMS;
        int kbs = (P.total_kb + sp - 1) / sp;
        double cost = (double)waves * (kbs + fixed);
        if (cost < best_cost) { best_cost = cost; best = sp; }
    }
    P.kb_per_split = (P.total_kb + best - 1) / best;
    const int splits = (P.total_kb + P.kb_per_split - 1) / P.kb_per_split;
    size_t smem = (size_t)P.stages * stage_bytes + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_wgrad2_kernel<PIX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
        if (ce == cudaSuccess)
            ce = cudaFuncSetAttribute(conv_wgrad2_kernel<PIX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad2): %s", cudaGetErrorString(ce)); return (int)ce; }
        attr_set = true;
    }
    if (mc) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((co_tiles + 1) / 2 * 2, P.ugroups, splits);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t ce = cudaLaunchKernelEx(&cfg, conv_wgrad2_kernel<PIX, true>, tmDy, tmX, P);
        if (ce != cudaSuccess) { set_error("conv_wgrad2 (multicast) launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    } else {
        dim3 grid(co_tiles, P.ugroups, splits);
        conv_wgrad2_kernel<PIX, false><<<grid, TC_THREADS, smem, st>>>(tmDy, tmX, P);
    }
    g_launches.fetch_add(1);
    return check_launch("conv_wgrad2");
}

static bool wgrad2_box_ok(int pix, int Ho, int Wo, int s) {
    int bw, bh, bn;
    return choose_box_n(pix, Ho, Wo, &bw, &bh, &bn) && bw * s <= 256 && bh * s <= 256;
}

static int launch_wgrad2(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                         int k, int s, int p, cudaStream_t st, int dw_cl = 0) {
    int e = ensure_encode();
    if (e) return e;
    TcW2Params P;
    P.Mpix = N * Ho * Wo; P.Ho = Ho; P.Wo = Wo; P.Co = Co; P.Ci = Ci; P.kk = k * k; P.k = k; P.s = s; P.p = p; P.dw = dw;
    P.dw_cl = dw_cl;
    P.cblocks = (Ci + 63) / 64;
    P.units = P.cblocks * P.kk;
    P.ugroups = (P.units + 7) / 8;
    P.ubase = P.units / P.ugroups;
    P.urem = P.units % P.ugroups;
    P.nun_max = P.ubase + (P.urem > 0 ? 1 : 0);
    // pixels per k-block: few units per CTA (thin layers) -> longer blocks, so that a stage stays ~40 KB and the
    // per-stage barrier round trip is amortised
    if (P.nun_max <= 2 && wgrad2_box_ok(128, Ho, Wo, s) && P.Mpix >= 128 * 8)
        return launch_wgrad2_pix<128>(x, dy, P, N, H, W, Ci, Ho, Wo, Co, s, st);
    if (P.nun_max <= 4 && wgrad2_box_ok(64, Ho, Wo, s) && P.Mpix >= 64 * 8)
        return launch_wgrad2_pix<64>(x, dy, P, N, H, W, Ci, Ho, Wo, Co, s, st);
    return launch_wgrad2_pix<32>(x, dy, P, N, H, W, Ci, Ho, Wo, Co, s, st);
}

}  // namespace sg

using namespace sg;

extern "C" {

// 1 if the tcgen05 path can take this operator direction (bf16 only)
int sg_conv_tc_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    int Ck = mode == 0 ? Ci : Co;
    int Hq = mode == 0 ? Ho : H / s, Wq = mode == 0 ? Wo : W / s;
    int bw, bh, bn;
    if (Ck % 8 != 0) return 0;
    if (s < 1 || s > 2 || k % s != 0 || k > 4) return 0;
    if (mode == 1 && (H % s != 0 || W % s != 0)) return 0;
    if (!choose_box(Hq, Wq, &bw, &bh, &bn)) return 0;
    if (bw * s > 256 || bh * s > 256) return 0;
    return 1;
}

int sg_set_option(const char* name, int value) {
    if (name && !strcmp(name, "tc2")) { g_use_tc2 = value; return 0; }
    if (name && !strcmp(name, "persist")) { g_use_persist = value; return 0; }
    if (name && !strcmp(name, "wgrad2")) { g_use_wgrad2 = value; return 0; }
    if (name && !strcmp(name, "split")) { g_use_split = value; return 0; }
    if (name && !strcmp(name, "wgrad_mc")) { g_use_wgrad_mc = value; return 0; }
    if (name && !strcmp(name, "pdl")) { g_use_pdl = value; return 0; }
    if (name && !strcmp(name, "force_cg")) { g_force_cg = value; return 0; }
    if (name && !strcmp(name, "epi_alt")) { g_epi_alt = value; return 0; }
    if (name && !strcmp(name, "force_bn")) { g_force_bn = value; return 0; }
    if (name && !strcmp(name, "force_stages")) { g_force_stages = value; return 0; }
    if (name && !strcmp(name, "dbg")) { g_dbg = value; return 0; }
    set_error("unknown option");
    return SG_ERR_BAD_ARG;
}

static bool want_tc2(int n_total, int M) { return g_use_tc2 && n_total >= 128 && n_total % 32 == 0 && M >= 256; }

int sg_conv_fprop_tc(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho, int Wo,
                     int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_fprop_tc: bf16 only");
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc: unsupported shape");
    if (g_use_persist)
        return launch_conv_tcp(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, nullptr, 1, SG_STREAM(stream));
    if (want_tc2(Co, N * Ho * Wo))
        return launch_conv_tc2(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, SG_STREAM(stream));
    return launch_conv_tc(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, SG_STREAM(stream));
}

int sg_conv_dgrad_tc(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho,
                     int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_dgrad_tc: bf16 only");
    SG_REQUIRE(sg_conv_tc_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_dgrad_tc: unsupported shape");
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k, "conv_dgrad_tc: inconsistent sizes");
    if (g_use_persist)
        return launch_conv_tcp(1, dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, nullptr, 1, SG_STREAM(stream));
    if (want_tc2(Ci, N * (H / s) * (W / s)))
        return launch_conv_tc2(1, dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, SG_STREAM(stream));
    return launch_conv_tc(1, dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, SG_STREAM(stream));
}

int sg_conv_wgrad_tc_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    int bw, bh, bn;
    if (Ci % 8 != 0 || Co % 8 != 0) return 0;
    if (s < 1 || s > 2 || k > 4) return 0;
    if (!choose_box64(Ho, Wo, &bw, &bh, &bn)) return 0;
    if (bw * s > 256 || bh * s > 256) return 0;
    return 1;
}

// accumulate into a channels-last gradient buffer gw[Co][k][k][Ci] (vector reductions for any k); sg_fold_grad_cl
// adds it into the PyTorch-layout gradient
int sg_conv_wgrad_cl_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int dtype) {
    int bw, bh, bn;
    if (dtype != SG_BF16 || !g_use_wgrad2 || Ci % 4 != 0) return 0;
    if (!sg_conv_wgrad_tc_supported(N, H, W, Ci, Ho, Wo, Co, k, s, p)) return 0;
    return choose_box_n(32, Ho, Wo, &bw, &bh, &bn) && bw * s <= 256 && bh * s <= 256;
}
int sg_conv_wgrad_cl(const void* x, const void* dy, float* gw, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k,
                     int s, int p, int dtype, void* stream) {
    SG_REQUIRE(sg_conv_wgrad_cl_supported(N, H, W, Ci, Ho, Wo, Co, k, s, p, dtype), "conv_wgrad_cl: unsupported shape/dtype");
    return launch_wgrad2(x, dy, gw, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_STREAM(stream), 1);
}

int sg_conv_wgrad_tc(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k,
                     int s, int p, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_wgrad_tc: bf16 only");
    SG_REQUIRE(sg_conv_wgrad_tc_supported(N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_wgrad_tc: unsupported shape");
    if (g_use_wgrad2) {
        int bw, bh, bn;
        // 32-pixel blocks must not straddle images unless whole images fit a block
        if (choose_box_n(32, Ho, Wo, &bw, &bh, &bn) && bw * s <= 256 && bh * s <= 256)
            return launch_wgrad2(x, dy, dw, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_STREAM(stream));
    }
    return launch_wgrad_tc(x, dy, dw, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_STREAM(stream));
}

// y = act(conv(x, W) + bias + residual): the closing layer of a residual block (generator_2.py:23-26) in one kernel
int sg_conv_fprop_tc_res(const void* x, const void* pf, const float* bias, const void* residual, void* y, int N, int H, int W,
                         int Ci, int Ho, int Wo, int Co, int k, int s, int p, int act, void* stream) {
    SG_REQUIRE(g_use_persist, "conv_fprop_tc_res needs the persistent kernel");
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc_res: unsupported shape");
    return launch_conv_tcp(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, nullptr, 1, SG_STREAM(stream), nullptr,
                           residual);
}

// y (FP32) = conv(x, W) with bf16 operands: the un-rounded accumulators, for results that are summed again (col2im)
int sg_conv_fprop_tc_f32out(const void* x, const void* pf, float* y, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                            int k, int s, int p, void* stream) {
    SG_REQUIRE(g_use_persist, "conv_fprop_tc_f32out needs the persistent kernel");
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc_f32out: unsupported shape");
    return launch_conv_tcp(0, x, pf, nullptr, nullptr, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, nullptr, 1,
                           SG_STREAM(stream), y);
}

// conv + per-channel (sum, sum^2) of the stored output accumulated into stats[groups][C][2] (the statistics
// of the BatchNorm that follows), fused into the epilogue.  Returns SG_ERR_UNSUPPORTED (without launching)
// when the shape cannot be fused; the dispatcher then runs the conv and sg_col_stats separately.
int sg_conv_tc_stats_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups) {
    if (!g_use_persist || groups < 1 || N % groups != 0) return 0;
    if (!sg_conv_tc_supported(mode, N, H, W, Ci, Ho, Wo, Co, k, s, p)) return 0;
    int Hq = mode == 0 ? Ho : H / s, Wq = mode == 0 ? Wo : W / s;
    long rows_per_group = (long)(N / groups) * Hq * Wq;
    return rows_per_group % 128 == 0 ? 1 : 0;
}
int sg_conv_fprop_tc_stats(const void* x, const void* pf, void* y, double* stats, int groups, int N, int H, int W, int Ci,
                           int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_fprop_tc_stats: bf16 only");
    SG_REQUIRE(sg_conv_tc_stats_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups), "conv_fprop_tc_stats: unsupported shape");
    return launch_conv_tcp(0, x, pf, nullptr, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, stats, groups, SG_STREAM(stream));
}
int sg_conv_dgrad_tc_stats(const void* dy, const void* pd, void* dx, double* stats, int groups, int N, int H, int W, int Ci,
                           int Ho, int Wo, int Co, int k, int s, int p, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_dgrad_tc_stats: bf16 only");
    SG_REQUIRE(sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups), "conv_dgrad_tc_stats: unsupported shape");
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k, "conv_dgrad_tc_stats: inconsistent sizes");
    return launch_conv_tcp(1, dy, pd, nullptr, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, stats, groups, SG_STREAM(stream));
}

}  // extern "C"
