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
#include <atomic>
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
// one lane of a CONVERGED warp (elect.sync): unlike `lane == 0`, the compiler knows the guarded code runs on exactly one
// lane and issues the TMA / tcgen05 instructions straight from uniform registers -- under `if (lane == 0)` ptxas wraps
// every one of them in an ELECT / BRA.U.ANY loop (4 extra instructions and a convergence branch per MMA on the thread
// whose instruction stream paces the mainloop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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

constexpr int TC_THREADS = 192;
constexpr int A_STAGE_BYTES = 128 * 128;

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
// the same with shared-memory addresses already converted (the producer keeps them as 32-bit integers and advances them
// by adds: no generic-to-shared conversion on its critical path)
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_leader_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar & PEER_BIT_MASK), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_u32(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma2_load_4d_u32(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d_u32(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
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

// ------------------------------------------------------------------------------------------------ persistent kernel
// conv_tcp_kernel<CG>: PERSISTENT implicit-GEMM convolution -- one CTA (CG=1) or CTA pair (CG=2, cta_group::2) per SM /
// SM pair walks a list of work items (output tiles).  The accumulator is multi-buffered in TMEM, so the epilogue warps
// drain tile i (tcgen05.ld -> bias/act -> bf16 NHWC rows, plus the per-channel (sum, sum^2) of the BatchNorm that
// follows) while the MMA warp already accumulates tile i+1.
//
//   warp 0    : TMA producer             full[s] / empty[s] ring
//   warp 1    : MMA issuer (leader CTA)  tcgen05.commit -> empty[s], -> tfull[buf]
//   warps 2-9 : epilogue                 wait tfull[buf] ... arrive tempty[buf] (on the leader)
//
// What bounds this kernel is the L2 -> SM operand stream: at full occupancy the 148 SMs share
// ~6.3 KB/cycle, i.e. ~44 B/cycle/SM, and a 128 x 256 x 64 k-block wants 32 KB in the 512 cycles its MMAs take.  Two
// things cut the bytes per FLOP:
//
//  * SHARED ACTIVATION SLABS.  The taps of a filter column that hit the same input-row parity read the SAME input rows,
//    shifted by one output row: for k4 s2 p1, tap kh=2 at output row oh reads what tap kh=0 reads at oh+1.  A stage is
//    therefore not one (tap, 64-channel block) but one SLAB: the (bh + nv - 1) input rows that nv vertically adjacent
//    taps need, loaded by ONE TMA box, plus the nv weight tiles.  The tile's 128 rows are ordered (h, n, w) -- the tensor
//    map's dimensions are (C, W, N, H) -- so that "one output row further down" is the same byte offset bn*bw*128 for
//    every row of the tile even when the tile spans several images; that offset is a multiple of the 1024-byte swizzle
//    atom (bn*bw >= 8), so tap i's A operand is the SAME shared-memory slab read through a UMMA descriptor whose start
//    address is advanced by i*bn*bw*128 bytes.  A bytes: x9/16 (k4 s2, 8-row tiles), x10/24 (3x3 s1), x(bh+1)/2bh
//    (ConvTranspose phases).
//  * COLUMN SLICES FOR THE LAST ROUND.  With T tiles on U clusters the last round holds R = T mod U tiles; when R <= U/2
//    those tiles are cut into S = U/R column slices of BN2 columns (a second weight tensor map with a BN2/CG-row box and
//    a narrower instruction descriptor), so the round's critical path shrinks from one full tile to one slice without
//    any partial-sum exchange (the K-split this replaces paid ~8 us per partial tile in workspace traffic).
struct SlabEnt {
    int16_t dw, dh;       // first input column / row of the slab, relative to (w0*es, h0*es)
    int16_t nv, pad_;     // taps sharing this slab (<= 4)
    int32_t bk[4];        // K offset of tap i's weights in the packed matrix (elements)
    int32_t roff[4];      // byte offset of tap i's first row inside the slab (multiples of 1024)
};

struct TcpParams {
    int Hq, Wq, n_img;    // output grid handled per phase, images
    int bw, bh, bn;       // pixel box of one 128-row tile; rows ordered (h, n, w)
    int lgBW, lgNW;       // log2(bw), log2(bn*bw)
    int n_total, BN, BN2; // output channels, tile width, width of a last-round column slice
    int Ck, cblocks;
    int mode, s, es;      // es: element stride of the activation tensor map (s for fprop, 1 for dgrad)
    int outH, outW;
    int act;
    int stages, stage_bytes, slab_bytes, slab_pad, b_bytes;   // slab_pad: slab_bytes rounded up to 1024; b_bytes: one full-width weight tile
    int tmem_cols, acc_stride, nbuf;      // TMEM columns allocated, columns per accumulator buffer, buffers in the ring
    int m_tiles, n_tiles, total_tiles, tiles_per_phase;   // m_tiles counts CG*128-row cluster tiles
    int full_tiles, nslices, total_work;  // work items: full_tiles whole tiles, then (total_tiles - full_tiles) * nslices slices
    int kiters;           // stages consumed per work item, per phase equal: nslab * cblocks
    int rotate;           // clusters start the reduction at different slabs (spreads the weight reads over L2)
    int imgs_per_group;
    int lgW, lgHW;        // log2(Wq), log2(Hq*Wq)
    int epi_alt;          // narrow tiles: the two epilogue warp groups drain ALTERNATE tiles (see the epilogue)
    int nslab[4];         // slabs per phase
    const float* bias;
    bf16* out;
    double* stats;        // [groups][n_total][2] or NULL
    float* out32;         // fp32 result instead of bf16 `out` (col2im input of the thin layers) or NULL
    const bf16* residual; // added before the activation (same layout as out) or NULL
    unsigned long long* trace;   // profiling hook (sg_debug_conv_trace): 16 globaltimer stamps per CTA, or NULL
    unsigned int* sched;         // dynamic work distribution: {next item, clusters done} of this launch, or NULL = static lists
    int sched_chunk;             // items per draw (consecutive numbers): > 1 for launches of many short items
    // BatchNorm-BACKWARD statistics fused into the epilogue (sg_conv_*_bstats): this launch's result is da = d loss / d a of
    // the layer below, a = act(bn(y)); with bs_y != NULL `stats` receives (S1, S2) = (sum dz, sum dz * xhat) per (group, channel),
    // dz = da * act'(gamma * xhat + beta), xhat = (y - mean) * rstd -- what sg_bn_bwd_reduce_y computes in a pass of its own
    const bf16* bs_y;            // pre-BN tensor of that layer, same layout as `out`
    const float* bs_mr;          // [groups][n_total][2] mean, rstd
    const float* bs_gamma;
    const float* bs_beta;
    float bs_slope;              // activation slope for negative pre-activations (0 ReLU, 0.1 LeakyReLU, 1 none)
    int bs_mask_out;             // store dz (the masked gradient) instead of da: conv + activation backward in one kernel
    SlabEnt slab[4][16];
};

__device__ __forceinline__ void trace_stamp(const TcpParams& P, int slot) {
    if (P.trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.trace[(size_t)blockIdx.x * 16 + slot] = t;
    }
}

__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// default (.release.cta) semantics on the cluster address, as CUTLASS's ClusterBarrier::arrive does: the explicit
// .release.cluster form costs a cluster-scope MEMBAR (~2600 cycles per tile, measured), and nothing but the TMEM reads --
// already complete after tcgen05.wait::ld and ordered by tcgen05.fence::before_thread_sync -- is handed over here
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

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

constexpr int TCP_EPI_WARPS = 8;
constexpr int TCP_THREADS = 64 + 32 * TCP_EPI_WARPS;

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Work item i of a cluster: whole tiles  c, c+C, c+2C, ...  below P.full_tiles, then the column slices of the leftover
// tiles (slice sl of leftover tile lt is item  full_tiles + lt*nslices + sl: the slices of one tile run on neighbouring
// clusters at the same time, so its activation slabs are fetched from DRAM once and hit L2 after that).
struct Work {
    int tile, nt0, width, sliced;     // tile index; first output column; columns this item computes; 1 = a column slice
};

__device__ __forceinline__ void decode_work(const TcpParams& P, int t, Work* w) {
    int tile = t, sl = 0;
    w->sliced = 0;
    if (t >= P.full_tiles) {
        const int idx = t - P.full_tiles;
        const int lt = idx / P.nslices;
        sl = idx - lt * P.nslices;
        tile = P.full_tiles + lt;
        w->sliced = P.nslices > 1 ? 1 : 0;
    }
    w->tile = tile;
    int r = tile;
    if (P.tiles_per_phase != P.total_tiles) r = tile % P.tiles_per_phase;
    const int nt = P.n_tiles != 1 ? r / P.m_tiles : 0;
    w->nt0 = nt * P.BN + sl * P.BN2;
    w->width = w->sliced ? min(P.BN2, P.BN - sl * P.BN2) : P.BN;
}

// ---- dynamic work distribution (TcpParams::sched != NULL).  The static schedule gives cluster c the items c, c+C, c+2C, ...:
// fine on an empty machine, but the captured step runs these kernels NEXT TO the weight-gradient kernels of the side stream,
// whose CTAs hold an SM for 20-60 us each -- a conv CTA that lands late still owns its whole list and the launch ends when it
// does.  Here the leader CTA's producer warp draws item numbers from a global counter and hands them to every other role
// through a 4-entry shared-memory queue (and to the peer CTA of a pair with st.async + complete_tx): a CTA that lands late
// simply finds less -- or nothing -- left.  The last cluster to finish re-arms the counter for the next launch / graph replay.
constexpr int WQ = 8;
// wait on a queue barrier; bounded (~seconds), so that a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}

// BS: instantiation with the BatchNorm-backward statistics in the epilogue (TcpParams::bs_*).  A template parameter, not a
// run-time flag: carrying that code cost the plain launches 30 registers and ~4 % (critic ds3: 27.7 -> 29.2 us).
// DYN: instantiation with the dynamic work distribution (TcpParams::sched) -- also a template parameter: as a run-time flag the
// queue code cost the static launches ~2 % (sampling 56.0 -> 52.7 k images/s).
template <int CG, bool BS, bool DYN>
__global__ void __launch_bounds__(TCP_THREADS, 1)
conv_tcp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ TcpParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[8], empty_bar[8], tfull_bar[8], tempty_bar[8];
    __shared__ uint32_t tmem_slot;
    __shared__ float sstat[2][256][2];          // statistics staging, double-buffered by tile parity
    __shared__ uint4 sstage[TCP_EPI_WARPS][32 * 4];   // per epilogue warp: 32 rows x 64 B, for the coalesced store
    __shared__ float4 sconst[BS ? 256 : 1];     // backward statistics: (mean, rstd*gamma, beta, rstd) of the tile's columns
    __shared__ int wq_id[DYN ? WQ : 1];         // dynamic schedule: queue of work-item numbers (deep enough for the producer's lead)
    __shared__ uint64_t wq_full[DYN ? WQ : 1], wq_empty[DYN ? WQ : 1];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    if (threadIdx.x == 0) trace_stamp(P, 0);

    if (threadIdx.x == 32) {          // fetch the TMA descriptors while the barriers and TMEM are being set up
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (P.nslices > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < P.nbuf; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], (P.epi_alt ? TCP_EPI_WARPS / 2 : TCP_EPI_WARPS) * CG); }
        // queue: full = the leader producer's publication (peer CTA: its remote expect_tx + the st.async bytes); empty (leader's
        // copy only) = every consumer of the pair: MMA warp + 8 epilogue warps of the leader, producer + 8 epilogue warps of the peer
        if (DYN)
            for (int i = 0; i < WQ; ++i) { mbar_init(&wq_full[i], 1); mbar_init(&wq_empty[i], (1 + TCP_EPI_WARPS) * CG); }
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
    if (threadIdx.x == 0) trace_stamp(P, 1);
    SG_PDL_SYNC();        // barriers, TMEM and the stats staging are set up: from here on global memory is touched
    if (threadIdx.x == 0) trace_stamp(P, 2);
    constexpr bool dyn = DYN;
    // consumer side of the work queue (every role but the leader CTA's producer): item number of this role's next round
    int q_tail = 0;
    uint32_t q_par = 0;
    auto pop_work = [&]() -> int {
        mbar_wait_bounded(&wq_full[q_tail], q_par);
        const int t = *reinterpret_cast<volatile int*>(&wq_id[q_tail]);
        __syncwarp();
        if (lane == 0) {                      // the number is in registers: hand the slot back (to the leader's barrier)
            if (CG == 2) mbar_arrive_leader(&wq_empty[q_tail]); else mbar_arrive_local(&wq_empty[q_tail]);
        }
        if (++q_tail == WQ) { q_tail = 0; q_par ^= 1; }
        return t;
    };

    if (warp == 0) {
        {
            // ONE thread issues every TMA of this CTA, and its instruction stream is on the critical path: measured with
            // the kernel's own timeline (tools/exp_conv_trace.py), a stage cost ~1100 cycles whatever its bytes while this
            // loop re-derived coordinates, re-read the slab table from constant memory and walked a prefetch cursor per
            // stage -- the producer, not L2, paced the mainloop.  So: everything that does not change between stages is
            // hoisted to the work item / slab, the inner loop is wait -> expect_tx -> 1 + nv TMA issues, and shared-memory
            // addresses advance by adds.
            int st = 0;
            uint32_t par = 1;                         // parity to wait for on empty[st]: first round passes
            const uint32_t ring = smem_u32(smem);
            Work w;
            // leader producer under the dynamic schedule: draws the item numbers (one ahead, so that the atomic's round
            // trip hides behind the current item's loads) and publishes them
            const bool fetcher = dyn && rank == 0;
            // the first chunk of a cluster is static (items c*chunk ..: no round trip before the first load); every further
            // chunk is  (C + draw) * chunk ..  with draw = atomicAdd(counter, 1), fetched one chunk ahead
            const int chunk = P.sched_chunk;
            int q_head = 0, t_base = cluster_id * chunk, t_in = 0, t_next = 0;
            uint32_t q_epar = 1;
            if (fetcher && lane == 0) t_next = (num_clusters + (int)atomicAdd(P.sched, 1u)) * chunk;
            for (int wi = 0;; ++wi) {
                int t;
                if (!dyn) t = cluster_id + wi * num_clusters;
                else if (!fetcher) t = pop_work();
                else {
                    if (t_in == chunk) {                 // chunk used up: switch to the prefetched one, draw the one after
                        t_base = __shfl_sync(0xffffffffu, t_next, 0);
                        t_in = 0;
                        if (t_base < P.total_work && lane == 0) t_next = (num_clusters + (int)atomicAdd(P.sched, 1u)) * chunk;
                    }
                    t = t_base + t_in;
                    ++t_in;
                    mbar_wait_bounded(&wq_empty[q_head], q_epar);
                    if (elect_one()) {
                        wq_id[q_head] = t;
                        if (CG == 2) {
                            const uint32_t pf = mapa_u32(smem_u32(&wq_full[q_head]), 1u), pi = mapa_u32(smem_u32(&wq_id[q_head]), 1u);
                            asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(pf), "r"(4u) : "memory");
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(pi), "r"(t), "r"(pf) : "memory");
                        }
                        mbar_arrive_local(&wq_full[q_head]);
                    }
                    __syncwarp();
                    if (++q_head == WQ) { q_head = 0; q_epar ^= 1; }
                }
                if (t >= P.total_work) break;
                decode_work(P, t, &w);
                int phase = 0, r = w.tile;
                if (P.tiles_per_phase != P.total_tiles) { phase = w.tile / P.tiles_per_phase; r = w.tile - phase * P.tiles_per_phase; }
                const int mt = P.n_tiles != 1 ? r % P.m_tiles : r;
                const int m0 = (mt * CG + (int)rank) * 128;
                const int w0 = (m0 & (P.Wq - 1)) * P.es, h0 = ((m0 >> P.lgW) & (P.Hq - 1)) * P.es, n0 = m0 >> P.lgHW;   // Hq, Wq are powers of two
                const int b_rows = w.sliced ? P.BN2 / CG : P.BN / CG;      // weight rows this CTA stages per tap
                const int nt0 = w.nt0 + (int)rank * b_rows;                // CTA 1 of a pair holds the upper half of the MMA's N
                const CUtensorMap* mapB = w.sliced ? &tmB2 : &tmB;
                const int nsl = P.nslab[phase];
                // every cluster starts the reduction at a different slab (and walks it cyclically): otherwise all 74-148
                // CTAs ask L2 for the SAME weight lines at the same moment, a hot spot on a few L2 slices
                int sbr = P.rotate ? (cluster_id + wi) % nsl : 0;
                for (int sb = 0; sb < nsl; ++sb, sbr = (sbr + 1 == nsl ? 0 : sbr + 1)) {
                    const SlabEnt& e = P.slab[phase][sbr];
                    const int ca_w = w0 + e.dw, ca_h = h0 + e.dh, nv = e.nv;
                    const int bk0 = e.bk[0], bk1 = e.bk[1], bk2 = e.bk[2], bk3 = e.bk[3];
                    const uint32_t tx = (uint32_t)(P.slab_bytes + nv * b_rows * 128);
                    for (int cb = 0; cb < P.cblocks; ++cb) {
                        const int c0 = cb * 64;
                        mbar_wait(&empty_bar[st], par);
                        const uint32_t sa = ring + (uint32_t)st * (uint32_t)P.stage_bytes, sb0 = sa + (uint32_t)P.slab_pad;
                        const uint32_t fb = smem_u32(&full_bar[st]);
                        if (!elect_one()) {
                        } else if (CG == 2) {
                            mbar_expect_tx_leader_u32(fb, tx);
                            tma2_load_4d_u32(&tmA, fb, sa, c0, ca_w, n0, ca_h);
                            tma2_load_2d_u32(mapB, fb, sb0, bk0 + c0, nt0);
                            if (nv > 1) tma2_load_2d_u32(mapB, fb, sb0 + (uint32_t)P.b_bytes, bk1 + c0, nt0);
                            if (nv > 2) tma2_load_2d_u32(mapB, fb, sb0 + 2u * (uint32_t)P.b_bytes, bk2 + c0, nt0);
                            if (nv > 3) tma2_load_2d_u32(mapB, fb, sb0 + 3u * (uint32_t)P.b_bytes, bk3 + c0, nt0);
                        } else {
                            mbar_expect_tx_u32(fb, tx);
                            tma_load_4d_u32(&tmA, fb, sa, c0, ca_w, n0, ca_h);
                            tma_load_2d_u32(mapB, fb, sb0, bk0 + c0, nt0);
                            if (nv > 1) tma_load_2d_u32(mapB, fb, sb0 + (uint32_t)P.b_bytes, bk1 + c0, nt0);
                            if (nv > 2) tma_load_2d_u32(mapB, fb, sb0 + 2u * (uint32_t)P.b_bytes, bk2 + c0, nt0);
                            if (nv > 3) tma_load_2d_u32(mapB, fb, sb0 + 3u * (uint32_t)P.b_bytes, bk3 + c0, nt0);
                        }
                        if (++st == P.stages) { st = 0; par ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N = tile or slice width, M=128*CG
            const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((128 * CG) >> 4) << 24);
            // K-major 128B-swizzled operand descriptors (make_kmajor_sw128_desc) differ only in the 14-bit start address
            // (16-byte units) of their low word; everything per-MMA is an integer add on that word.  This loop is what
            // paces narrow tiles: measured ~157 cycles per tcgen05.mma whatever its N when each descriptor was rebuilt with
            // shifts and masks and the slab table was re-read from constant memory per tap (tools/exp_conv_trace.py).
            constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            const uint32_t dlo_ring = ((base >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t stage16 = (uint32_t)P.stage_bytes >> 4, slab16 = (uint32_t)P.slab_pad >> 4, b16 = (uint32_t)P.b_bytes >> 4;
            auto mma = [&](uint32_t tacc, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
                const uint64_t ad = ((uint64_t)DESC_HI << 32) | alo, bd = ((uint64_t)DESC_HI << 32) | blo;
                if (CG == 2) tc2_mma_bf16(tacc, ad, bd, idesc, acc);
                else tc_mma_bf16(tacc, ad, bd, idesc, acc);
            };
            int st = 0;
            uint32_t par = 0, ab = 0, abpar = 1;       // accumulator buffer ring: P.nbuf buffers of acc_stride TMEM columns
            Work w;
            for (int wi = 0;; ++wi) {
                const int t = dyn ? pop_work() : cluster_id + wi * num_clusters;
                if (t >= P.total_work) break;
                decode_work(P, t, &w);
                const int phase = P.tiles_per_phase != P.total_tiles ? w.tile / P.tiles_per_phase : 0;
                // the MMA's N: the slice width rounded up to what the staged weight rows cover (a trailing slice narrower
                // than BN2 still multiplies BN2 staged rows -- rows past the layer's channels are TMA zero fill -- and the
                // epilogue ignores the surplus columns)
                const uint32_t nmma = (uint32_t)(w.sliced ? P.BN2 : P.BN);
                const uint32_t idesc = idesc0 | ((nmma >> 3) << 17);
                mbar_wait(&tempty_bar[ab], abpar);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ab * (uint32_t)P.acc_stride;
                const int nsl = P.nslab[phase];
                uint32_t acc = 0;
                int sbr = P.rotate ? (cluster_id + wi) % nsl : 0;        // the producer's slab order
                for (int sb = 0; sb < nsl; ++sb, sbr = (sbr + 1 == nsl ? 0 : sbr + 1)) {
                    const SlabEnt& e = P.slab[phase][sbr];
                    const int nv = e.nv;
                    const uint32_t ro1 = (uint32_t)e.roff[1] >> 4, ro2 = (uint32_t)e.roff[2] >> 4, ro3 = (uint32_t)e.roff[3] >> 4;
                    for (int cb = 0; cb < P.cblocks; ++cb) {
                        mbar_wait(&full_bar[st], par);
                        const uint32_t a0 = dlo_ring + (uint32_t)st * stage16, b0 = a0 + slab16;
                        if (elect_one()) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) mma(tacc, a0 + 2 * k4, b0 + 2 * k4, idesc, k4 == 0 ? acc : 1u);
                            if (nv > 1) {
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) mma(tacc, a0 + ro1 + 2 * k4, b0 + b16 + 2 * k4, idesc, 1u);
                            }
                            if (nv > 2) {
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) mma(tacc, a0 + ro2 + 2 * k4, b0 + 2 * b16 + 2 * k4, idesc, 1u);
                            }
                            if (nv > 3) {
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) mma(tacc, a0 + ro3 + 2 * k4, b0 + 3 * b16 + 2 * k4, idesc, 1u);
                            }
                            if (CG == 2) tc2_commit_mc(&empty_bar[st]); else tc_commit(&empty_bar[st]);
                        }
                        acc = 1u;
                        __syncwarp();
                        if (++st == P.stages) { st = 0; par ^= 1; }
                    }
                }
                if (elect_one()) {
                    if (CG == 2) tc2_commit_mc(&tfull_bar[ab]); else tc_commit(&tfull_bar[ab]);
                }
                __syncwarp();
                if (++ab == (uint32_t)P.nbuf) { ab = 0; abpar ^= 1; }
                if (lane == 0) trace_stamp(P, wi == 0 ? 4 : 5);
            }
        }
    } else {
        // ---- epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).. and every second 32-column chunk
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        // rows of a tile are ordered (h, n, w)
        const int dw = r & (P.bw - 1), dn = (r >> P.lgBW) & (P.bn - 1), dh = r >> P.lgNW;
        const bool vec_ok = (P.n_total % 8) == 0;
        const bool bias_vec_ok = (reinterpret_cast<uintptr_t>(P.bias) & 15) == 0 && (P.BN & 3) == 0 && (P.BN2 & 3) == 0;   // float4 loads of the bias
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
            cdw[i] = rr & (P.bw - 1); cdn[i] = (rr >> P.lgBW) & (P.bn - 1); cdh[i] = rr >> P.lgNW;
        }
        uint32_t ti = 0, ab = 0, abpar = 0;
        Work w;
        for (int wi = 0;; ++wi, ++ti, ab = (ab + 1 == (uint32_t)P.nbuf ? 0 : ab + 1), abpar ^= (ab == 0 ? 1u : 0u)) {
            const int t = dyn ? pop_work() : cluster_id + wi * num_clusters;
            if (t >= P.total_work) break;
            decode_work(P, t, &w);
            // epi_alt (narrow tiles without statistics): warps 2-5 drain the even work items, warps 6-9 the odd
            // ones, each warp all columns of its 32 rows -- the per-tile fixed cost (decode, row pointers, barrier round
            // trip: ~350 of the ~450 instructions a warp spends on a 128x64 tile) is paid by four warps instead of eight
            // and two tiles drain concurrently
            if (P.epi_alt && (int)(ti & 1u) != half) continue;
            const int tile = w.tile;
            // tile decode: the divisions are ~30 instructions each and this code runs per tile and warp -- the
            // single-phase / single-column-tile cases (warp-uniform) skip them
            int phase = 0, rr = tile, ph = 0, pw = 0;
            if (P.tiles_per_phase != P.total_tiles) { phase = tile / P.tiles_per_phase; rr = tile - phase * P.tiles_per_phase; }
            int mt = rr;
            if (P.n_tiles != 1) mt = rr % P.m_tiles;
            if (phase != 0) { ph = phase / P.s; pw = phase - ph * P.s; }
            const int m0 = (mt * CG + (int)rank) * 128;
            const int w0 = m0 & (P.Wq - 1), h0 = (m0 >> P.lgW) & (P.Hq - 1), n0 = m0 >> P.lgHW;   // Hq, Wq are powers of two
            const int nt0 = w.nt0;
            const int n_img = n0 + dn, hh = h0 + dh, ww = w0 + dw;
            const bool row_ok = n_img < P.n_img && (P.out != nullptr || P.out32 != nullptr);
            int oh = hh, ow = ww;
            if (P.mode == 1) { oh = hh * P.s + ph; ow = ww * P.s + pw; }
            // pixel indices fit 32 bits (checked on the host); one widening multiply per pointer
            bf16* orow = P.out + (int64_t)((n_img * P.outH + oh) * P.outW + ow) * P.n_total + nt0;
            const int ncols = min(w.width, P.n_total - nt0);          // valid columns of this work item
            const uint32_t sb2 = ti & 1;           // statistics staging buffer
            constexpr bool bwd = BS;
            if (bwd) {
                // per-column constants of this tile's image group; the bar.sync that closed the previous tile guarantees that
                // no warp still reads the previous tile's table
                if (et < ncols) {
                    const int c = nt0 + et;
                    const float* mq = P.bs_mr + ((int64_t)(n0 / P.imgs_per_group) * P.n_total + c) * 2;
                    const float mean = __ldg(mq), rstd = __ldg(mq + 1);
                    sconst[et] = make_float4(mean, rstd * __ldg(P.bs_gamma + c), __ldg(P.bs_beta + c), rstd);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const bf16* yrow = bwd ? P.bs_y + (int64_t)((n_img * P.outH + oh) * P.outW + ow) * P.n_total + nt0 : nullptr;
            const float slope = P.bs_slope;
            if (P.kiters >= 8) mbar_wait_backoff(&tfull_bar[ab], abpar);     // long mainloop: sleep between polls
            else mbar_wait(&tfull_bar[ab], abpar);                          // short tiles: the sleep would be the latency
            tc_fence_after();
            if (et == 0) trace_stamp(P, wi == 0 ? 6 : 8);
            const uint32_t tacc = tmem_base + ab * (uint32_t)P.acc_stride + ((uint32_t)(q * 32) << 16);
            auto process = [&](const int c0, const uint32_t (&v)[16]) {
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
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
                            const int r8 = i * 8 + (lane >> 2);
                            const uint4 val = stg[r8 * 4 + ((lane & 3) ^ ((r8 >> 1) & 3))];
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
                if (bwd && P.bs_mask_out) {                            // the stored result is dz = da * act'(gamma * xhat + beta)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float yv = 0.f;
                        if (row_ok && c0 + j < ncols) yv = __bfloat162float(yrow[c0 + j]);
                        const float4 cst = sconst[min(c0 + j, 255)];
                        f[j] = ((yv - cst.x) * cst.y + cst.z > 0.f) ? f[j] : slope * f[j];
                    }
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
                    if (bwd) {
                        // f = da (as stored); turn it into dz and pair it with xhat
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float yv = 0.f;
                            if (row_ok && c0 + j < ncols) yv = __bfloat162float(yrow[c0 + j]);
                            const float4 cst = sconst[min(c0 + j, 255)];
                            const float t = yv - cst.x;
                            const float dz = (P.bs_mask_out || t * cst.y + cst.z > 0.f) ? f[j] : slope * f[j];   // (masked already)
                            const bool live = c0 + j < ncols;          // columns past the layer's channels: table entries unset
                            f[j] = live ? dz : 0.f;
                            sq[j] = live ? dz * (t * cst.w) : 0.f;
                        }
                    } else
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
            // ---- fast path (every BatchNorm'ed conv of the training step): no bias / residual / fp32 output,
            // whole 32-column groups.  Straight-line code over 32 columns per iteration -- two independent 16-column
            // chains for the scheduler to interleave, since only two epilogue warps share an SM sub-partition and the
            // drain is issue-latency bound -- with the activation and the statistics decided once per tile.
            const bool plain = P.residual == nullptr && P.out32 == nullptr && vec_ok &&
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
                    uint32_t yk[16];
                    if (bwd) {
                        // the row's 32 pre-BN values (64 contiguous bytes per lane)
                        uint4 y0 = make_uint4(0u, 0u, 0u, 0u), y1 = y0, y2 = y0, y3 = y0;
                        if (n_img < P.n_img) {
                            const uint4* yp = reinterpret_cast<const uint4*>(yrow + c0);
                            y0 = __ldg(yp); y1 = __ldg(yp + 1); y2 = __ldg(yp + 2); y3 = __ldg(yp + 3);
                        }
                        yk[0] = y0.x; yk[1] = y0.y; yk[2] = y0.z; yk[3] = y0.w; yk[4] = y1.x; yk[5] = y1.y; yk[6] = y1.z; yk[7] = y1.w;
                        yk[8] = y2.x; yk[9] = y2.y; yk[10] = y2.z; yk[11] = y2.w; yk[12] = y3.x; yk[13] = y3.y; yk[14] = y3.z; yk[15] = y3.w;
                        if (P.bs_mask_out) {                          // the stored result is dz = da * act'(gamma * xhat + beta)
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const uint32_t w2 = yk[j >> 1];
                                const float yv = __uint_as_float((j & 1) ? (w2 & 0xffff0000u) : (w2 << 16));
                                const float4 cst = sconst[c0 + j];
                                f[j] = ((yv - cst.x) * cst.y + cst.z > 0.f) ? f[j] : slope * f[j];
                            }
                        }
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
                            const int r8 = i * 8 + (lane >> 2);
                            const uint4 val = stg[r8 * 4 + ((lane & 3) ^ ((r8 >> 1) & 3))];
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
                            if (bwd) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const uint32_t w2 = yk[h2 * 8 + (j >> 1)];
                                    const float yv = __uint_as_float((j & 1) ? (w2 & 0xffff0000u) : (w2 << 16));
                                    const float4 cst = sconst[c0 + h2 * 16 + j];
                                    const float t = yv - cst.x;
                                    a[j] = (P.bs_mask_out || t * cst.y + cst.z > 0.f) ? a[j] : slope * a[j];   // (masked already)
                                    sq[j] = a[j] * (t * cst.w);
                                }
                            } else
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
            if (et == 0) trace_stamp(P, wi == 0 ? 7 : 9);
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
    if (threadIdx.x == 0) trace_stamp(P, 10);
    if (dyn && threadIdx.x == 0 && rank == 0) {
        // every cluster draws (at least) its terminating number before it gets here: the last one to arrive re-arms both words
        __threadfence();
        if (atomicAdd(P.sched + 1, 1u) == (unsigned)(num_clusters - 1)) {
            P.sched[0] = 0u; P.sched[1] = 0u;
            __threadfence();
        }
    }
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
// UMMA reads directly through MN-major 128B-swizzled descriptors -- no transposes.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // between 64-element blocks along M/N
    d |= (uint64_t)(1024 >> 4) << 32;                    // between 8-row groups along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
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
__device__ __forceinline__ void tma_load_4d_mc_u32(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3,
                                                   uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
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
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
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
    int cs;                    // multicast launches: CTAs per cluster (2..5 co tiles sharing their x tiles)
};

// MC = true: clusters of P.cs = 2..5 CTAs along the co-tile axis (same units, same pixel range, different 128 output
// channels).  The x tiles are the same for all of them, so CTA r fetches the units u with u % cs == r and TMA-multicasts
// them into every CTA's shared memory: x is 80 % of a stage, so the L2 -> SM operand traffic per CTA drops from 40 KB per
// k-block to 8 + 32/cs KB (24 KB for pairs, 18.7 KB for the three co tiles of the 320-channel residual convs, 14.4 KB for
// the five of the 640-channel ones: below the ~22 KB that 512 cycles of MMA can absorb at the L2 -> SM ceiling).  A stage
// may be refilled once ALL CTAs' MMAs have retired it (empty barriers count cs, commits are multicast).  A cluster CTA
// whose co tile lies past the layer's channels (only when the host pads) just keeps the barrier protocol going.
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

    // FIVE warps feed a stage: warp 0 the two dy blocks, warps 2-5 (the epilogue warps, idle during the mainloop) two x units
    // each.  One thread issuing all ten TMA loads of a k-block was the critical path of this kernel (its instruction stream
    // took longer than the k-block's MMAs); every feeding warp waits for the stage itself and posts its own expect_tx.
    constexpr int FEEDERS = 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], FEEDERS); mbar_init(&empty_bar[i], MC ? P.cs : 1); }
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
        if (warp != 1) {
            // ---- feeding warps (converged; the TMA issue is done by one elected lane)
            const int fw = warp == 0 ? -1 : warp - 2;          // -1: dy; 0..3: x units 2*fw, 2*fw + 1
            const int ua = fw < 0 ? 0 : 2 * fw, ub = ua + 1;
            const bool has_a = fw >= 0 && ua < nun, has_b = fw >= 0 && ub < nun;
            int ca = 0, ha = 0, wa = 0, cb2 = 0, hb = 0, wb = 0;
            if (has_a) { const int u = u0 + ua, cib = u / P.kk, tap = u - cib * P.kk; ca = cib * 64; ha = tap / P.k; wa = tap - ha * P.k; ha -= P.p; wa -= P.p; }
            if (has_b) { const int u = u0 + ub, cib = u / P.kk, tap = u - cib * P.kk; cb2 = cib * 64; hb = tap / P.k; wb = tap - hb * P.k; hb -= P.p; wb -= P.p; }
            // bytes this warp's loads put into a stage (multicast: the peer CTA delivers the units of the other parity)
            const uint32_t my_tx = fw < 0 ? (has_rows ? (uint32_t)A_BYTES : 0u) : (uint32_t)(((has_a ? 1 : 0) + (has_b ? 1 : 0)) * B_BYTES);
            const bool issue_a = has_a && (!MC || (ua % P.cs) == (int)rank), issue_b = has_b && (!MC || (ub % P.cs) == (int)rank);
            const uint16_t mc_mask = (uint16_t)((1u << P.cs) - 1u);
            const uint32_t ring = smem_u32(smem);
            int st = 0;
            uint32_t par = 1;
            int pix0 = kb0 * PIX;
            int ow0 = pix0 % P.Wo, oh0 = (pix0 / P.Wo) % P.Ho, n0 = pix0 / (P.Wo * P.Ho);
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&empty_bar[st], par);
                if (elect_one()) {
                    const uint32_t fb = smem_u32(&full_bar[st]);
                    const uint32_t sa = ring + (uint32_t)st * (uint32_t)stage_bytes;
                    mbar_expect_tx_u32(fb, my_tx);
                    if (fw < 0) {
                        if (has_rows) {
                            tma_load_2d_u32(&tmDy, fb, sa, co0, pix0);
                            tma_load_2d_u32(&tmDy, fb, sa + PIX * 128, co0 + 64, pix0);
                        }
                    } else {
                        const int w0 = ow0 * P.s, h0 = oh0 * P.s;
                        if (issue_a) {
                            if (!MC) tma_load_4d_u32(&tmX, fb, sa + A_BYTES + ua * B_BYTES, ca, w0 + wa, h0 + ha, n0);
                            else tma_load_4d_mc_u32(&tmX, fb, sa + A_BYTES + ua * B_BYTES, ca, w0 + wa, h0 + ha, n0, mc_mask);
                        }
                        if (issue_b) {
                            if (!MC) tma_load_4d_u32(&tmX, fb, sa + A_BYTES + ub * B_BYTES, cb2, w0 + wb, h0 + hb, n0);
                            else tma_load_4d_mc_u32(&tmX, fb, sa + A_BYTES + ub * B_BYTES, cb2, w0 + wb, h0 + hb, n0, mc_mask);
                        }
                    }
                }
                __syncwarp();
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
        } else {
            // ---- MMA warp (converged; issue by one elected lane)
            // D=f32, A=B=bf16, both MN-major, M=128; up to FOUR units (N = 256) per instruction: the unit tiles lie
            // B_BYTES apart in shared memory, which is exactly the descriptor's stride between 64-element blocks
            // along N, so the 128 x 16 slab of dy is read from shared memory twice per k-step instead of 8 times
            // (an N=64 MMA re-reads 4 KB of A for 2 KB of B and is shared-memory bound)
            const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
            const int n_lo = nun < 4 ? nun : 4, n_hi = nun - n_lo;
            const uint32_t idesc_lo = idesc0 | ((uint32_t)((n_lo * 64) >> 3) << 17);
            const uint32_t idesc_hi = idesc0 | ((uint32_t)((n_hi * 64) >> 3) << 17);
            // MN-major 128B-swizzled descriptors (make_mnmajor_sw128_desc): only the 14-bit start address in the low word
            // changes from MMA to MMA -- everything per instruction is an add on that word
            constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t LBO16 = (uint32_t)((PIX * 128) >> 4) << 16;          // stride between 64-element blocks along M/N
            const uint32_t dlo_ring = ((base >> 4) & 0x3FFFu) | LBO16;
            const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
            auto mma = [&](uint32_t tacc, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
                tc_mma_bf16(tacc, ((uint64_t)DESC_HI << 32) | alo, ((uint64_t)DESC_HI << 32) | blo, idesc, acc);
            };
            int st = 0;
            uint32_t par = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&full_bar[st], par);
                if (MC && !has_rows) {
                    // padding CTA: nothing to multiply, but every CTA of the cluster must see this one release the stage
                    if (elect_one()) {
                        for (int c = 0; c < P.cs; ++c) mbar_arrive_cta(&empty_bar[st], (uint32_t)c);
                    }
                    __syncwarp();
                    if (++st == P.stages) { st = 0; par ^= 1; }
                    continue;
                }
                const uint32_t a0 = dlo_ring + (uint32_t)st * stage16, b0 = a0 + (uint32_t)(A_BYTES >> 4), b1 = b0 + (uint32_t)((4 * B_BYTES) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < PIX / 16; ++k16) {
                        const uint32_t acc = (i | k16) != 0 ? 1u : 0u;
                        mma(tmem_base, a0 + k16 * 128, b0 + k16 * 128, idesc_lo, acc);
                        if (n_hi > 0) mma(tmem_base + 256, a0 + k16 * 128, b1 + k16 * 128, idesc_hi, acc);
                    }
                    if (MC) tc_commit_mc(&empty_bar[st], (uint16_t)((1u << P.cs) - 1u)); else tc_commit(&empty_bar[st]);
                }
                __syncwarp();
                if (++st == P.stages) { st = 0; par ^= 1; }
            }
            if ((!MC || has_rows) && elect_one()) tc_commit(&accum_bar);
            __syncwarp();
        }
        if (warp >= 2 && (!MC || has_rows)) {
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

// The same tensor seen as (C, W, N, H) -- H outermost -- so that a box {64, bw*es, bn, rows*es} lands in shared memory as
// [row][image][column][64 ch]: the slab layout of conv_tcp_kernel (rows of a tile ordered (h, n, w)).
static int get_slab_map(const void* ptr, int N, int H, int W, int C, int bw, int rows, int bn, int es, CUtensorMap* out) {
    MapKey key(ptr, N, H, W, C, bw, rows, bn, es, 5);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)N, (cuuint64_t)H};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)bn, (cuuint32_t)(rows * es)};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, 1, (cuuint32_t)es};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(slab N=%d H=%d W=%d C=%d box=%d,%d,%d es=%d) failed: %d", N, H, W, C, bw, rows, bn, es,
                  (int)r);
        return SG_ERR_UNSUPPORTED;
    }
    g_maps[key] = m;
    *out = m;
    return 0;
}

// direct_tc.cu: NHWC bf16 tensor [N][H][W][C] (C = 16 / 32 / 64 channels), box {C, bw*es, bh, 1} with element stride es along W
// only: bw pixels es apart x bh consecutive rows -> shared memory [row][pixel][C], swizzled with a span of swz bytes (>= C*2)
int get_direct_map(const void* ptr, int N, int H, int W, int C, int bw, int bh, int es, int swz, CUtensorMap* out) {
    if (int e = ensure_encode()) return e;
    MapKey key(ptr, N, H, W, C, bw, bh, swz, es, 7);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(bw * es), (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, 1, 1};
    const CUtensorMapSwizzle sw = swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(direct N=%d H=%d W=%d C=%d box=%d,%d es=%d) failed: %d", N, H, W, C, bw, bh, es, (int)r);
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
static int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

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

// ---- persistent launch: pick (CTA group, tile width) with a small cost model --------------------------
// cycles per 64-deep k-block: tcgen05 issue = 2*BN (M=128 per CTA, either group size); operand fetch from L2 =
// (the tap's share of the activation slab + BN/CG rows of B) at ~44 B/cycle/SM (the ~12 TB/s L2->SM ceiling shared by
// 148 SMs).
int g_force_cg = 0, g_force_bn = 0, g_force_stages = 0, g_dbg = 0;
extern int g_use_narrow, g_narrow_cfg; // narrow_conv.cu
extern int g_dtc_diag, g_dtc_wide;     // direct_tc.cu
int g_use_slab = 1;      // option "slab": 0 = one activation box per tap (no sharing), for A/B measurements
// option "dyn_sched" / env SG_DYN_SCHED: dynamic work distribution in the persistent conv kernel (see decode_work); needs the
// counter pool of sg_init_workspace().  OFF by default: measured on B200 (round 2, bench.py) Stage-I 5.42 -> 5.60 ms and
// Stage-II 35.2 -> 36.1 ms per step with it, critic ds3 fprop alone 27.7 -> 29.7 us, many-tile layers up to +28 % (D1 ds2
// dgrad 42.1 -> 53.9 us): the queue hand-off sits in front of every item of every role, and with the main chain on a
// high-priority stream the late-landing CTAs it was meant to cure are rare.  Kept for machines shared with other work.
int g_dyn_sched = getenv("SG_DYN_SCHED") ? atoi(getenv("SG_DYN_SCHED")) : 0;
constexpr unsigned SCHED_SLOTS = 4096;
unsigned int* g_sched_pool = nullptr;        // SCHED_SLOTS x {next item, clusters done}, zero between launches
std::atomic<unsigned> g_sched_seq{0};
// option "bstats_min_k" / env SG_BSTATS_MIN_K: the BatchNorm-backward statistics ride in a conv's epilogue only when its
// reduction is at least this deep (see conv_dispatch.cu); 0 = always, 1 << 30 = never.  Measured on B200 (bench.py, Stage-II
// B=64, ms per outer step): always 36.55, >= 1500: 35.63, >= 2500: 35.41, >= 4000: 35.32, never: 35.40 (Stage-I: 5.76-5.81
// whatever the setting).  The fused epilogue costs what the separate HBM-bound reduce pass (6 TB/s) saves: with two epilogue
// warp groups per SM the y tile loads, the per-column constants and two more butterflies per 16 columns put the drain on the
// critical path of every tile whose mainloop is shorter than ~4000 reduction elements.  Default: the deep layers only.
int g_bstats_min_k = getenv("SG_BSTATS_MIN_K") ? atoi(getenv("SG_BSTATS_MIN_K")) : 4000;
int g_use_nsplit = 1;    // option "nsplit": 0 = no column slices in the last round
unsigned long long* g_trace = nullptr;   // sg_debug_conv_trace
int g_rotate = 0;        // option "rotate": measured neutral on B200 (gpurun_out/bench_conv_r2i.txt), off
// alternate-tile epilogue for narrow tiles (option "epi_alt" / env SG_EPI_ALT)
int g_epi_alt = getenv("SG_EPI_ALT") ? atoi(getenv("SG_EPI_ALT")) : 1;
static bool g_pattr_set = false;
constexpr int TCP_SMEM_BYTES = 205 * 1024;      // + ~20 KB static (statistics staging, store transposes) <= 227 KB
constexpr int TCP_SMEM_BYTES_BS = 201 * 1024;   // the instantiation with the 4 KB BatchNorm constant table

// column slices for the leftover tiles of the last round: returns S (1 = none) and the slice width
static int nsplit_plan(long tiles, int units, int bn, int cg, int* bn2) {
    *bn2 = bn;
    if (!g_use_nsplit) return 1;
    const int R = (int)(tiles % units);
    if (R == 0) return 1;
    int S = units / R;
    if (S > bn / 32) S = bn / 32;
    if (S < 2) return 1;
    int w = ((bn + S - 1) / S + 31) / 32 * 32;      // multiple of 32: whole 32-column epilogue groups, 16*CG for the MMA
    S = (bn + w - 1) / w;
    if (S < 2) return 1;
    *bn2 = w;
    return S;
}

static void pick_tcp_config(int M, int n_total, int phases, int nkb, double a_share, int* cg_out, int* bn_out) {
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
            auto t_kblock = [&](int width) {
                double mma = 2.0 * width;
                double l2 = (16384.0 * a_share + (double)(width / cg) * 128.0) / 44.0;
                return mma > l2 ? mma : l2;
            };
            double t_tile = nkb * t_kblock(bn) + 150.0;
            double epi = 40.0 * bn / 16.0 + 400.0;                  // drain of the last tile, not overlapped
            double t_epi_tile = 40.0 * bn / 16.0 + 100.0;           // epilogue pace per tile
            if (t_epi_tile > t_tile) t_tile = t_epi_tile;
            // rounds of the persistent tile loop; a partly filled last round may be cut into column slices
            double t_last = 0.0;
            if (tiles % units) {
                int bn2;
                const int S = nsplit_plan(tiles, units, bn, cg, &bn2);
                t_last = nkb * t_kblock(S > 1 ? bn2 : bn) + 150.0;
                if (S > 1) epi = 40.0 * bn2 / 16.0 + 400.0;
            }
            double t = (double)(tiles / units) * t_tile + t_last + epi;
            if (t < best) { best = t; best_cg = cg; best_bn = bn; }
        }
    }
    *cg_out = best_cg; *bn_out = best_bn;
}

// mode 0: fprop (act = x [N][H][W][Ck], out = y [N][Ho][Wo][n_total]);
// mode 1: dgrad (act = dy [N][Ho][Wo][Ck], out = dx [N][H][W][n_total])
struct BsArgs {            // BatchNorm-backward statistics in the epilogue (TcpParams::bs_*)
    const void* y; const float* mr; const float* gamma; const float* beta; float slope; int mask_out;
};

static int launch_conv_tcp(int mode, const void* act, const void* wpack, const float* bias, void* out, int N, int H, int W,
                           int Ci, int Ho, int Wo, int Co, int k, int s, int p, int actf, double* stats, int groups,
                           cudaStream_t st, float* out32 = nullptr, const void* residual = nullptr, const BsArgs* bs = nullptr) {
    int e = ensure_encode();
    if (e) return e;
    static TcpParams Pz;        // zero-initialised template (the slab table has padding the compiler would not clear)
    TcpParams P = Pz;
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
    P.lgW = ilog2(P.Wq); P.lgHW = ilog2(P.Hq * P.Wq);
    P.lgBW = ilog2(P.bw); P.lgNW = ilog2(P.bn * P.bw);
    P.cblocks = (P.Ck + 63) / 64;
    P.mode = mode; P.s = s; P.es = mode == 0 ? s : 1; P.act = actf; P.bias = bias; P.out = (bf16*)out;
    P.stats = stats;
    P.out32 = out32;
    P.residual = (const bf16*)residual;
    if (bs != nullptr) { P.bs_y = (const bf16*)bs->y; P.bs_mr = bs->mr; P.bs_gamma = bs->gamma; P.bs_beta = bs->beta; P.bs_slope = bs->slope; P.bs_mask_out = bs->mask_out; }
    P.trace = g_trace;
    P.sched = (g_dyn_sched && g_sched_pool != nullptr && bs == nullptr) ? g_sched_pool + 2 * (g_sched_seq.fetch_add(1) % SCHED_SLOTS) : nullptr;
    P.sched_chunk = 1;
    if (out32) P.out = nullptr;
    P.imgs_per_group = groups > 0 ? N / groups : N;
    const int phases = mode == 0 ? 1 : s * s;
    // ---- slab table: which taps share one activation slab (see the kernel's header comment)
    const int row_bytes = P.bn * P.bw * 128;                  // one output row of the tile, all its images
    const bool share = g_use_slab && (P.bn * P.bw) % 8 == 0;  // tap offsets must be whole 1024-byte swizzle atoms
    int nv_max = 1, taps_per_phase = 0;
    for (int ph = 0; ph < phases; ++ph) {
        int n = 0, taps = 0;
        if (mode == 0) {
            const int nclass = share ? (s < k ? s : k) : k;   // sharing: one class per input-row parity; else one per tap row
            for (int c = 0; c < nclass; ++c) {
                const int nv = share ? (k - c + s - 1) / s : 1;
                for (int kw = 0; kw < k; ++kw) {
                    SlabEnt& E = P.slab[ph][n++];
                    E.dw = (int16_t)(-p + kw); E.dh = (int16_t)(-p + c); E.nv = (int16_t)nv;
                    for (int i = 0; i < nv; ++i) {
                        const int kh = share ? c + i * s : c;
                        E.bk[i] = (kh * k + kw) * P.Ck;
                        E.roff[i] = i * row_bytes;
                    }
                    taps += nv;
                }
                if (nv > nv_max) nv_max = nv;
            }
        } else {
            const int phh = ph / s, pww = ph - phh * s;
            const int rh = (phh + p) % s, rw = (pww + p) % s;
            const int nth = (k - rh + s - 1) / s, ntw = (k - rw + s - 1) / s;
            const int base_h = (phh + p - rh) / s, base_w = (pww + p - rw) / s;
            const int nclass = share ? 1 : nth;
            for (int c = 0; c < nclass; ++c) {
                const int nv = share ? nth : 1;
                for (int tw = 0; tw < ntw; ++tw) {
                    SlabEnt& E = P.slab[ph][n++];
                    E.dw = (int16_t)(base_w - tw);
                    E.dh = (int16_t)(share ? base_h - (nth - 1) : base_h - c);
                    E.nv = (int16_t)nv;
                    for (int i = 0; i < nv; ++i) {
                        // sharing: tap th reads rows h0 + base_h - th + j = slab row (nth - 1 - th) + j; store the taps in
                        // slab-row order (i = nth - 1 - th)
                        const int th = share ? nth - 1 - i : c;
                        E.bk[i] = ((rh + s * th) * k + (rw + s * tw)) * P.Ck;
                        E.roff[i] = i * row_bytes;
                    }
                    taps += nv;
                }
                if (nv > nv_max) nv_max = nv;
            }
        }
        if (n > 16) { set_error("conv_tcp: %d slabs per phase exceed the table", n); return SG_ERR_UNSUPPORTED; }
        P.nslab[ph] = n;
        if (ph == 0) taps_per_phase = taps;
        else if (taps != taps_per_phase || n != P.nslab[0]) { set_error("conv_tcp: phases with unequal tap counts"); return SG_ERR_UNSUPPORTED; }
    }
    const int slab_rows = P.bh + nv_max - 1;
    P.slab_bytes = slab_rows * row_bytes;
    P.slab_pad = (P.slab_bytes + 1023) / 1024 * 1024;
    P.kiters = P.nslab[0] * P.cblocks;
    P.rotate = g_rotate;
    const int nkb = taps_per_phase * P.cblocks;               // 64-deep k-blocks per tile
    int cg, bn;
    pick_tcp_config(M, P.n_total, phases, nkb, (double)slab_rows / (double)(nv_max * P.bh), &cg, &bn);
    P.BN = bn;
    P.n_tiles = (P.n_total + bn - 1) / bn;
    P.m_tiles = (M + 128 * cg - 1) / (128 * cg);
    P.tiles_per_phase = P.m_tiles * P.n_tiles;
    P.total_tiles = P.tiles_per_phase * phases;
    P.acc_stride = bn;
    // accumulator ring in TMEM: 2 buffers for wide tiles, up to 8 for narrow ones -- with short mainloops (thin layers,
    // 1x1 GEMMs) the MMA warp must be able to run several tiles ahead of the epilogue to hide the barrier round trips
    P.nbuf = 512 / bn;
    if (P.nbuf > 8) P.nbuf = 8;
    if (P.nbuf < 2) P.nbuf = 2;
    P.tmem_cols = 32;
    while (P.tmem_cols < P.nbuf * bn) P.tmem_cols *= 2;
    P.b_bytes = (bn / cg) * 128;
    P.stage_bytes = P.slab_pad + nv_max * P.b_bytes;
    P.stages = ((bs != nullptr ? TCP_SMEM_BYTES_BS : TCP_SMEM_BYTES) - 1024) / P.stage_bytes;
    if (P.stages > 8) P.stages = 8;
    if (g_force_stages && g_force_stages < P.stages) P.stages = g_force_stages;
    if (P.stages < 2) { set_error("conv_tcp: a %d-byte stage does not fit twice", P.stage_bytes); return SG_ERR_UNSUPPORTED; }
    if (g_dbg & 1) P.out = nullptr;
    // ---- work list: whole tiles, then the column slices of the last round's leftover tiles
    const int units = SG_NUM_SMS / cg;
    P.full_tiles = P.total_tiles; P.nslices = 1; P.BN2 = bn;
    {
        int bn2;
        const int S = nsplit_plan(P.total_tiles, units, bn, cg, &bn2);
        if (S > 1) {
            P.full_tiles = P.total_tiles - P.total_tiles % units;
            P.nslices = S; P.BN2 = bn2;
        }
    }
    P.total_work = P.full_tiles + (P.total_tiles - P.full_tiles) * P.nslices;
    const int clusters = P.total_work < units ? P.total_work : units;
    if (P.sched != nullptr) {                  // ~8 draws per cluster at least; never more than 8 items per draw
        int c = P.total_work / (clusters * 8);
        P.sched_chunk = c < 1 ? 1 : (c > 8 ? 8 : c);
    }
    CUtensorMap tmA, tmB, tmB2;
    if ((e = get_slab_map(act, N, aH, aW, P.Ck, P.bw, slab_rows, P.bn, P.es, &tmA))) return e;
    if ((e = get_w_map(wpack, P.n_total, k * k * P.Ck, bn / cg, &tmB))) return e;
    if ((e = get_w_map(wpack, P.n_total, k * k * P.Ck, P.BN2 / cg, &tmB2))) return e;
    size_t smem = (size_t)P.stages * P.stage_bytes + 1024;
    if (!g_pattr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_tcp_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tcp_kernel<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tcp_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tcp_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tcp_kernel<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES_BS);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv_tcp_kernel<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES_BS);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(tcp): %s", cudaGetErrorString(ce)); return (int)ce; }
        g_pattr_set = true;
    }
    // two tiles in flight in the epilogue need two accumulator buffers beyond the one being filled: nbuf >= 4 (bn <= 128)
    P.epi_alt = (g_epi_alt && stats == nullptr && bn <= 64 && P.nbuf >= 4) ? 1 : 0;
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
    cudaError_t ce;
    if (bs != nullptr)
        ce = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_tcp_kernel<2, true, false>, tmA, tmB, tmB2, P)
                     : cudaLaunchKernelEx(&cfg, conv_tcp_kernel<1, true, false>, tmA, tmB, tmB2, P);
    else if (P.sched != nullptr)
        ce = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_tcp_kernel<2, false, true>, tmA, tmB, tmB2, P)
                     : cudaLaunchKernelEx(&cfg, conv_tcp_kernel<1, false, true>, tmA, tmB, tmB2, P);
    else
        ce = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_tcp_kernel<2, false, false>, tmA, tmB, tmB2, P)
                     : cudaLaunchKernelEx(&cfg, conv_tcp_kernel<1, false, false>, tmA, tmB, tmB2, P);
    if (ce != cudaSuccess) { set_error("conv_tcp launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    g_launches.fetch_add(1);
    return check_launch("conv_tcp");
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

int g_use_wgrad_mc = 1;
// option "wgrad_smem_kb": shared memory the wgrad pipeline may take (stages = that / stage bytes).  200 = the whole SM; ~150
// leaves room for the BatchNorm kernels' CTAs next to a wgrad CTA (co-residency experiment, bn_fast.cu SG_BN_MAXREG)
int g_wgrad_smem_kb = 200;
// option "wgrad_mc_max": largest co-tile cluster the launcher may pick.  The kernel takes 2..8; the default stays at PAIRS:
// measured on B200 (profiles/bench_wgrad_r2c.txt, bench_wgrad_r2d.txt) clusters of 3-5 co tiles never beat unicast even when
// the whole grid is co-resident (res1: 75.9 us as triples, 64.4 us unicast; res3 / up0 as quintuples 65.5 / 107.6 vs
// 62.6 / 102.7) -- every stage then waits for the slowest of cs producers and is released by the slowest of cs consumers
int g_wgrad_mc_max = 2;
int g_wgrad_mc_odd = 1;      // option "wgrad_mc_odd": 0 = pairs only (the round-1 behaviour), for A/B measurements

template <int PIX>
static int launch_wgrad2_pix(const void* x, const void* dy, TcW2Params P, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                             int s, cudaStream_t st) {
    int bw, bh, bn, e;
    if (!choose_box_n(PIX, Ho, Wo, &bw, &bh, &bn)) { set_error("wgrad2: grid not tileable"); return SG_ERR_UNSUPPORTED; }
    P.total_kb = (P.Mpix + PIX - 1) / PIX;
    const int stage_bytes = PIX * 128 * (2 + P.nun_max);
    P.stages = (g_wgrad_smem_kb * 1024) / stage_bytes;
    if (P.stages > 8) P.stages = 8;
    if (P.stages < 1) { P.stages = (200 * 1024) / stage_bytes; }      // a stage that does not fit the reduced budget keeps the full one
    CUtensorMap tmDy, tmX;
    if ((e = get_rows_map_n(dy, P.Mpix, Co, PIX, &tmDy))) return e;
    if ((e = get_act_map(x, N, H, W, Ci, bw, bh, bn, s, &tmX))) return e;
    const int co_tiles = (Co + 127) / 128;
    size_t smem = (size_t)P.stages * stage_bytes + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ce = cudaFuncSetAttribute(conv_wgrad2_kernel<PIX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
        if (ce == cudaSuccess)
            ce = cudaFuncSetAttribute(conv_wgrad2_kernel<PIX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
        if (ce == cudaSuccess)
            ce = cudaFuncSetAttribute(conv_wgrad2_kernel<PIX, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad2): %s", cudaGetErrorString(ce)); return (int)ce; }
        attr_set = true;
    }
    // ---- cluster size x split-K.  The co tiles of a layer share their x tiles: a cluster of cs co tiles multicasts them
    // (kernel header), which cuts the operand bytes per k-block from 8 + 4*nun KB to 8 + 4*nun/cs KB.  But a cluster needs cs
    // free SMs in ONE GPC (16-20 SMs each, one CTA per SM at this shared-memory footprint): only cap(cs) clusters are
    // co-resident (measured through the occupancy API: ~46 triples, 33 quads), and a grid one cluster over that runs a
    // second wave -- round 2's first cut (always cluster all co tiles) made the 320-channel layers 1.7x SLOWER that way.
    // So: cost(cs, splits) = waves(cs) * (k-blocks per split * max(MMA, L2) cycles + fixed), minimised over both.
    static int cap_cache[9] = {0};
    auto cluster_cap = [&](int c) -> int {
        if (c == 1) return SG_NUM_SMS;
        if (cap_cache[c] > 0) return cap_cache[c];
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(c * 64, 1, 1); q.blockDim = dim3(TC_THREADS); q.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        q.attrs = a; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_wgrad2_kernel<PIX, true>, &q) != cudaSuccess || n < 1) {
            cudaGetLastError();
            n = SG_NUM_SMS / c * 3 / 4;          // conservative guess if the query is unavailable
        }
        cap_cache[c] = n;
        return n;
    };
    const bool vec_epi = P.dw_cl || P.kk == 1 || (P.kk & 3) == 0;
    // fixed cost per CTA (prologue + the atomics epilogue): ~10 us with vector reductions, ~40 us with scalar ones
    const double fixed_cycles = (vec_epi ? 18.0 : 66.0) * 1150.0;
    const double mma_cycles = 2.0 * PIX * P.nun_max;                       // (PIX/16) MMAs of N = 64*nun columns at N/2 cycles
    int max_splits = P.total_kb / 4;
    if (max_splits < 1) max_splits = 1;
    int best = 1, cs = 1;
    double best_cost = 1e30;
    for (int c = 1; c <= 8 && c <= co_tiles; ++c) {
        if (co_tiles % c != 0) continue;
        if (c > 1 && (!g_use_wgrad_mc || c > g_wgrad_mc_max || (!g_wgrad_mc_odd && c != 2))) continue;
        const int cap = cluster_cap(c);
        const double l2_cycles = (double)PIX * 128.0 * (2.0 + (double)P.nun_max / c) / 44.0;
        const double t_kb = mma_cycles > l2_cycles ? mma_cycles : l2_cycles;
        const long clusters1 = (long)(co_tiles / c) * P.ugroups;
        for (int sp = 1; sp <= max_splits && sp <= 148; ++sp) {
            const long waves = (clusters1 * sp + cap - 1) / cap;
            const int kbs = (P.total_kb + sp - 1) / sp;
            const double cost = (double)waves * (kbs * t_kb + fixed_cycles);
            if (cost < best_cost * 0.999) { best_cost = cost; best = sp; cs = c; }
        }
    }
    const bool mc = cs > 1;
    P.cs = cs;
    P.kb_per_split = (P.total_kb + best - 1) / best;
    const int splits = (P.total_kb + P.kb_per_split - 1) / P.kb_per_split;
    if (mc) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(co_tiles, P.ugroups, splits);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
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
    P.dw_cl = dw_cl; P.cs = 1;
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
    if (bw * s > 256 || (bh + 3) * s > 256) return 0;          // TMA box extents (a slab is up to bh + 3 rows)
    return 1;
}

// Profiling hook: while ``buf`` (device memory, >= 296 * 16 uint64) is set, every persistent conv launch writes 16
// %globaltimer stamps per CTA into it (kernel entry, set-up done, first operands landed, first / last work item issued,
// drained, exit -- see trace_stamp() call sites); NULL switches it off.  tools/exp_conv_trace.py prints the timeline.
int sg_debug_conv_trace(void* buf) {
    g_trace = reinterpret_cast<unsigned long long*>(buf);
    return 0;
}

// One-time device allocations of the library (the launch entry points themselves never allocate): the counter pool of the
// dynamic conv schedule.  Call once per process and device before the first launch, outside stream capture.
int sg_init_workspace(void) {
    if (g_sched_pool != nullptr) return 0;
    unsigned int* p = nullptr;
    cudaError_t e = cudaMalloc(&p, SCHED_SLOTS * 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p, 0, SCHED_SLOTS * 2 * sizeof(unsigned int));
    if (e != cudaSuccess) { set_error("sg_init_workspace: %s", cudaGetErrorString(e)); return (int)e; }
    g_sched_pool = p;
    return 0;
}

int sg_set_option(const char* name, int value) {
    if (name && !strcmp(name, "slab")) { g_use_slab = value; return 0; }
    if (name && !strcmp(name, "nsplit")) { g_use_nsplit = value; return 0; }
    if (name && !strcmp(name, "narrow")) { g_use_narrow = value; return 0; }
    if (name && !strcmp(name, "narrow_cfg")) { g_narrow_cfg = value; return 0; }
    if (name && !strcmp(name, "dtc_diag")) { g_dtc_diag = value; return 0; }
    if (name && !strcmp(name, "dtc_wide")) { g_dtc_wide = value; return 0; }
    if (name && !strcmp(name, "wgrad_smem_kb")) { g_wgrad_smem_kb = value < 48 ? 48 : (value > 200 ? 200 : value); return 0; }
    if (name && !strcmp(name, "dyn_sched")) { g_dyn_sched = value; return 0; }
    if (name && !strcmp(name, "bstats_min_k")) { g_bstats_min_k = value; return 0; }
    if (name && !strcmp(name, "rotate")) { g_rotate = value; return 0; }
    if (name && !strcmp(name, "wgrad_mc")) { g_use_wgrad_mc = value; return 0; }
    if (name && !strcmp(name, "wgrad_mc_max")) { g_wgrad_mc_max = value; return 0; }
    if (name && !strcmp(name, "wgrad_mc_odd")) { g_wgrad_mc_odd = value; return 0; }
    if (name && !strcmp(name, "pdl")) { g_use_pdl = value; return 0; }
    if (name && !strcmp(name, "force_cg")) { g_force_cg = value; return 0; }
    if (name && !strcmp(name, "epi_alt")) { g_epi_alt = value; return 0; }
    if (name && !strcmp(name, "force_bn")) { g_force_bn = value; return 0; }
    if (name && !strcmp(name, "force_stages")) { g_force_stages = value; return 0; }
    if (name && !strcmp(name, "dbg")) { g_dbg = value; return 0; }
    if (name && !strcmp(name, "bn_fused")) { g_bn_fused = value; return 0; }
    if (name && !strcmp(name, "bn_fused_keep_pct")) { g_bn_fused_keep_pct = value; return 0; }
    if (name && !strcmp(name, "dbg_skip_memset")) { g_dbg_skip_memset = value; return 0; }
    if (name && !strcmp(name, "bn_act_bulk")) { g_bn_act_bulk = value; return 0; }
    if (name && !strcmp(name, "gp_bn_fused")) { g_gp_bn_fused = value; return 0; }
    if (name && !strcmp(name, "bn_fused_dbg")) { g_bn_fused_dbg = value; return 0; }
    if (name && !strcmp(name, "bn_fused_steal_ns")) { g_bn_fused_steal_ns = value; return 0; }
    set_error("unknown option");
    return SG_ERR_BAD_ARG;
}

int sg_conv_fprop_tc(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho, int Wo,
                     int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_fprop_tc: bf16 only");
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc: unsupported shape");
    return launch_conv_tcp(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, nullptr, 1, SG_STREAM(stream));
}

int sg_conv_dgrad_tc(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho,
                     int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_dgrad_tc: bf16 only");
    SG_REQUIRE(sg_conv_tc_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_dgrad_tc: unsupported shape");
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k, "conv_dgrad_tc: inconsistent sizes");
    return launch_conv_tcp(1, dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, nullptr, 1, SG_STREAM(stream));
}

int sg_conv_wgrad_tc_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    int bw, bh, bn;
    if (Ci % 8 != 0 || Co % 8 != 0) return 0;
    if (s < 1 || s > 2 || k > 4) return 0;
    if (!choose_box_n(32, Ho, Wo, &bw, &bh, &bn)) return 0;       // 32-pixel k-blocks must not straddle images
    if (bw * s > 256 || bh * s > 256) return 0;
    return 1;
}

// accumulate into a channels-last gradient buffer gw[Co][k][k][Ci] (vector reductions for any k); sg_fold_grad_cl
// adds it into the PyTorch-layout gradient
int sg_conv_wgrad_cl_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int dtype) {
    int bw, bh, bn;
    if (dtype != SG_BF16 || Ci % 4 != 0) return 0;
    (void)bw; (void)bh; (void)bn;
    return sg_conv_wgrad_tc_supported(N, H, W, Ci, Ho, Wo, Co, k, s, p);
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
    return launch_wgrad2(x, dy, dw, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_STREAM(stream));
}

// y = act(conv(x, W) + bias + residual): the closing layer of a residual block (generator_2.py:23-26) in one kernel
int sg_conv_fprop_tc_res(const void* x, const void* pf, const float* bias, const void* residual, void* y, int N, int H, int W,
                         int Ci, int Ho, int Wo, int Co, int k, int s, int p, int act, void* stream) {
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc_res: unsupported shape");
    return launch_conv_tcp(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, nullptr, 1, SG_STREAM(stream), nullptr,
                           residual);
}

// y (FP32) = conv(x, W) with bf16 operands: the un-rounded accumulators, for results that are summed again (col2im)
int sg_conv_fprop_tc_f32out(const void* x, const void* pf, float* y, int N, int H, int W, int Ci, int Ho, int Wo, int Co,
                            int k, int s, int p, void* stream) {
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc_f32out: unsupported shape");
    return launch_conv_tcp(0, x, pf, nullptr, nullptr, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, nullptr, 1,
                           SG_STREAM(stream), y);
}

// conv + per-channel (sum, sum^2) of the stored output accumulated into stats[groups][C][2] (the statistics
// of the BatchNorm that follows), fused into the epilogue.  Returns SG_ERR_UNSUPPORTED (without launching)
// when the shape cannot be fused; the dispatcher then runs the conv and sg_col_stats separately.
int sg_conv_tc_stats_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups) {
    if (groups < 1 || N % groups != 0) return 0;
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

// conv + the BatchNorm-BACKWARD statistics of the layer below in the epilogue: the result da (T) is d loss / d a for
// a = act(bn(ybn)); sums[groups][C][2] (zeroed here) += (sum dz, sum dz * xhat), dz = da * act'(gamma * xhat + beta).  Replaces the
// sg_bn_bwd_reduce_y pass over (da, ybn) that followed every data-gradient conv of a BatchNorm'ed layer.
static float bs_slope_of(int act) { return act == SG_ACT_RELU ? 0.f : act == SG_ACT_LRELU ? 0.1f : 1.f; }
int sg_conv_fprop_tc_bstats(const void* x, const void* pf, void* y, const void* ybn, const float* mr, const float* gamma,
                            const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo,
                            int Co, int k, int s, int p, void* stream) {
    SG_REQUIRE(sg_conv_tc_stats_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups) && Co <= 2048, "conv_fprop_tc_bstats: unsupported shape");
    SG_REQUIRE(act == SG_ACT_NONE || act == SG_ACT_RELU || act == SG_ACT_LRELU, "conv_fprop_tc_bstats: act in {none, relu, lrelu}");
    cudaMemsetAsync(sums, 0, (size_t)groups * Co * 2 * sizeof(double), SG_STREAM(stream));
    BsArgs bs{ybn, mr, gamma, beta, bs_slope_of(act), 0};
    return launch_conv_tcp(0, x, pf, nullptr, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, sums, groups, SG_STREAM(stream), nullptr,
                           nullptr, &bs);
}
int sg_conv_dgrad_tc_bstats(const void* dy, const void* pd, void* dx, const void* ybn, const float* mr, const float* gamma,
                            const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho, int Wo,
                            int Co, int k, int s, int p, void* stream) {
    SG_REQUIRE(sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups) && Ci <= 2048, "conv_dgrad_tc_bstats: unsupported shape");
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k, "conv_dgrad_tc_bstats: inconsistent sizes");
    SG_REQUIRE(act == SG_ACT_NONE || act == SG_ACT_RELU || act == SG_ACT_LRELU, "conv_dgrad_tc_bstats: act in {none, relu, lrelu}");
    cudaMemsetAsync(sums, 0, (size_t)groups * Ci * 2 * sizeof(double), SG_STREAM(stream));
    BsArgs bs{ybn, mr, gamma, beta, bs_slope_of(act), 0};
    return launch_conv_tcp(1, dy, pd, nullptr, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, sums, groups, SG_STREAM(stream), nullptr,
                           nullptr, &bs);
}

// The same with the MASKED gradient stored: dx = dz = conv result * act'(gamma * xhat + beta) -- data-gradient conv + activation
// (+ BatchNorm-statistics) backward in one kernel, sums[.][.][0] = the column sums of dz.  With the identity table (mean 0, rstd 1,
// gamma 1, beta 0) and ybn = the stored activation of a conv + bias + activation layer (discrminator_1.py:17-18) this replaces the
// separate activation-backward pass over the critics' largest activation, and sums[.][.][0] is that layer's bias gradient.
int sg_conv_dgrad_tc_bstats_masked(const void* dy, const void* pd, void* dx, const void* ybn, const float* mr, const float* gamma,
                                   const float* beta, double* sums, int groups, int act, int N, int H, int W, int Ci, int Ho,
                                   int Wo, int Co, int k, int s, int p, int sums_zeroed, void* stream) {
    SG_REQUIRE(sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups) && Ci <= 256 && Ci % 32 == 0,
               "conv_dgrad_tc_bstats_masked: unsupported shape");
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k, "conv_dgrad_tc_bstats_masked: inconsistent sizes");
    SG_REQUIRE(act == SG_ACT_NONE || act == SG_ACT_RELU || act == SG_ACT_LRELU, "conv_dgrad_tc_bstats_masked: act in {none, relu, lrelu}");
    if (!sums_zeroed) cudaMemsetAsync(sums, 0, (size_t)groups * Ci * 2 * sizeof(double), SG_STREAM(stream));
    BsArgs bs{ybn, mr, gamma, beta, bs_slope_of(act), 1};
    return launch_conv_tcp(1, dy, pd, nullptr, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, SG_ACT_NONE, sums, groups, SG_STREAM(stream), nullptr,
                           nullptr, &bs);
}
int sg_conv_dgrad_tc_bstats_masked_supported(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p, int groups) {
    return sg_conv_tc_stats_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p, groups) && Ci <= 256 && Ci % 32 == 0;
}

}  // extern "C"
