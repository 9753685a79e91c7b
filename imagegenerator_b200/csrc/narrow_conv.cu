// narrow_conv.cu -- the k4 s2 p1 convolutions with 16..32 input and <= 64 output channels as direct kernels.
//
// The Stage-II critic's second and third layers (discriminator_2.py:13-18: Conv2d(16 -> 32) and (32 -> 64) on 128x128 / 64x64
// maps, three image groups of 64) move 150 MB and 75 MB per pass for 13 and 6 GFLOP: HBM-bound shapes.  On the tcgen05 kernel
// they are 128 x 32 / 128 x 16 tiles whose per-tile epilogue costs more than the whole mainloop (4 k-blocks): 123 us forward,
// 186 us data gradient for ds2 -- 0.8-1.2 TB/s.  Here, like thin_conv.cu, a CTA stages a spatial tile in shared memory and
// contracts it with warp-level mma.sync; the result leaves through a per-warp staging buffer as 16-byte stores, and the
// forward kernel can reduce the BatchNorm batch statistics of its tiles on the way (sg_conv_fprop_stats).
//
// Second version: PERSISTENT CTAs (one contiguous range of tiles each, the weights staged once per CTA instead of once per
// tile -- they were 23-45 % of a tile's bytes) and the input tiles brought in with cp.async into a two-deep ring, so the next
// tile is in flight while this one is contracted and stored.  The first version loaded each tile with ld.global -> st.shared
// in a 19-trip loop (one DRAM latency per trip, nothing else running): 15 us per tile and CTA, 1 TB/s.
//
//   narrow_fprop_kernel  y[n,oh,ow,co]  = sum_{kh,kw,ci} x[n,2oh-1+kh,2ow-1+kw,ci] * pf[co][kh][kw][ci]          (+ bias, act)
//   narrow_dgrad_kernel  dx[n,oh,ow,ci] = sum_{kh,kw,co} dy[n,(oh+1-kh)/2,(ow+1-kw)/2,co] * pd[ci][kh][kw][co]    (+ bias, act)
#include "common.cuh"

namespace sg {

__device__ __forceinline__ void nmma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int NC_TH = 8, NC_TW = 32;        // output pixels per tile (fprop) / input (q) pixels per tile (dgrad): warp = row

__device__ __forceinline__ uint32_t nsaddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// asynchronous global -> shared copies; bytes = 0 writes zeros (the conv's padding) without touching src
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------ forward
// CI in {16, 32}, NT = Co / 8 in {1..8}, ST = depth of the input-tile ring, MB = CTAs per SM the registers are capped for.  Input
// tile: 18 x 66 pixels, pixel stride PS = CI + 8 elements; weights [Co][16 taps * CI + 16].  CTA b owns tiles [b * T / G,
// (b + 1) * T / G) in (n, tile row, tile column) order.
// Fragments: the MMA's K slots are a PERMUTATION of the 16 channels of a chunk -- lane q feeds channels 4q .. 4q+3 as its slots
// (2q, 2q+1, 2q+8, 2q+9) of A and of B alike, so each fragment half is ONE 8-byte shared-memory load instead of two 4-byte ones;
// the strides above make the 16 lanes of an 8-byte load phase (4 pixels two apart / 4 weight rows x 32 bytes) hit distinct banks.
template <int CI, int NT, int ST, int MB>
__global__ void __launch_bounds__(256, MB)
narrow_fprop_kernel(const bf16* __restrict__ x, const bf16* __restrict__ pf, const float* __restrict__ bias, bf16* __restrict__ y,
                    double* __restrict__ stats, int H, int W, int act, int imgs_per_group, int tiles_w, int tiles_h, int total) {
    constexpr int CO = NT * 8, PS = CI + 8, IR = 2 * NC_TH + 2, IC = 2 * NC_TW + 2, WS = 16 * CI + 16, SS = CO + 8;
    constexpr int TILE = IR * IC * PS;
    extern __shared__ __align__(16) uint8_t nsmem[];
    bf16* tile0 = reinterpret_cast<bf16*>(nsmem);                // [ST][IR][IC][PS]
    bf16* wsm = tile0 + ST * TILE;                                // [CO][WS]
    bf16* stage = wsm + CO * WS;                                  // [8 warps][32][SS]
    __shared__ float sstat[CO][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int t_begin = (int)(((int64_t)blockIdx.x * total) / gridDim.x), t_end = (int)(((int64_t)(blockIdx.x + 1) * total) / gridDim.x);
    const int Ho = H >> 1, Wo = W >> 1;
    if (stats != nullptr && tid < CO * 2) (&sstat[0][0])[tid] = 0.f;

    // input tile t -> ring slot: 16-byte chunks (8 channels); out-of-image pixels are zero (the conv's padding)
    auto issue_tile = [&](int t, int slot) {
        constexpr int CH = CI / 8;
        const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
        const int ih0 = 2 * th * NC_TH - 1, iw0 = 2 * tw * NC_TW - 1;
        const bf16* xin = x + (size_t)n * H * W * CI;
        const uint32_t dst0 = nsaddr(tile0 + slot * TILE);
#pragma unroll 4
        for (int i = tid; i < IR * IC * CH; i += 256) {
            const int p = i / CH, c = i - p * CH;
            const int r = p / IC, cc = p - r * IC;
            const int ih = ih0 + r, iw = iw0 + cc;
            const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
            const bf16* src = ok ? xin + ((size_t)ih * W + iw) * CI + c * 8 : x;
            cp_async16(dst0 + (uint32_t)(p * PS + c * 8) * 2u, src, ok ? 16 : 0);
        }
    };

    SG_PDL_SYNC();
    if (t_begin < t_end) issue_tile(t_begin, 0);
    cp_async_commit();
    {   // weights, once per CTA: rows of 16 * CI elements, 16-byte chunks
        constexpr int CH = 16 * CI / 8;
        for (int i = tid; i < CO * CH; i += 256) {
            const int co = i / CH, c = i - co * CH;
            *reinterpret_cast<uint4*>(wsm + co * WS + c * 8) = __ldg(reinterpret_cast<const uint4*>(pf + (size_t)co * 16 * CI + c * 8));
        }
    }
    float bz[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        bz[nt][0] = bias != nullptr ? __ldg(bias + nt * 8 + 2 * q) : 0.f;
        bz[nt][1] = bias != nullptr ? __ldg(bias + nt * 8 + 2 * q + 1) : 0.f;
    }
    int cur_grp = -1;
    auto flush_stats = [&](int grp) {       // all threads; sstat holds the sums of group grp
        __syncthreads();
        if (tid < CO * 2) {
            atomicAdd(stats + ((size_t)grp * CO) * 2 + tid, (double)(&sstat[0][0])[tid]);
            (&sstat[0][0])[tid] = 0.f;
        }
        __syncthreads();
    };

    for (int t = t_begin; t < t_end; ++t) {
        const int slot = ST == 2 ? ((t - t_begin) & 1) : 0;
        if (ST == 2) {
            if (t + 1 < t_end) issue_tile(t + 1, slot ^ 1);       // that slot was released by the barrier closing tile t - 1
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
        const int oh0 = th * NC_TH, ow0 = tw * NC_TW;
        if (stats != nullptr) {
            const int grp = n / imgs_per_group;
            if (grp != cur_grp) {
                if (cur_grp >= 0) flush_stats(cur_grp);
                cur_grp = grp;
            }
        }
        const bf16* tile = tile0 + slot * TILE;
        float acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;
        // output pixel p of this warp's row reads input tile pixel (2*warp + kh, 2*p + kw)
        const bf16* arow = tile + (2 * warp) * IC * PS + 4 * q;
        const bf16* wrow = wsm + g * WS + 4 * q;
#pragma unroll 1
        for (int kh = 0; kh < 4; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 4; ++kw) {
#pragma unroll
                for (int cs = 0; cs < CI / 16; ++cs) {
                    const int koff = (kh * 4 + kw) * CI + cs * 16;
                    uint32_t bf[NT][2];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint2 b = *reinterpret_cast<const uint2*>(wrow + nt * 8 * WS + koff);
                        bf[nt][0] = b.x; bf[nt][1] = b.y;
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const bf16* a0p = arow + (kh * IC + 2 * (mt * 16 + g) + kw) * PS + cs * 16;
                        const bf16* a1p = a0p + 16 * PS;                              // output pixel g + 8: input pixel + 16
                        const uint2 lo = *reinterpret_cast<const uint2*>(a0p), hi = *reinterpret_cast<const uint2*>(a1p);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) nmma16816(acc[mt][nt], lo.x, hi.x, lo.y, hi.y, bf[nt][0], bf[nt][1]);
                    }
                }
            }
        }
        if (ST == 1) {      // single slot: everybody is done reading it -> refill it under the epilogue
            __syncthreads();
            if (t + 1 < t_end) issue_tile(t + 1, 0);
            cp_async_commit();
        }
        // epilogue: bias / activation, bf16, per-warp staging, 16-byte stores; optional BatchNorm statistics of the STORED values
        bf16* stw = stage + warp * 32 * SS;
        float s1[NT][2], s2[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t lo = pack_bf16x2(act_fwd(acc[mt][nt][0] + bz[nt][0], act), act_fwd(acc[mt][nt][1] + bz[nt][1], act));
                const uint32_t hi = pack_bf16x2(act_fwd(acc[mt][nt][2] + bz[nt][0], act), act_fwd(acc[mt][nt][3] + bz[nt][1], act));
                *reinterpret_cast<uint32_t*>(stw + (mt * 16 + g) * SS + nt * 8 + 2 * q) = lo;
                *reinterpret_cast<uint32_t*>(stw + (mt * 16 + g + 8) * SS + nt * 8 + 2 * q) = hi;
                const float v0 = __uint_as_float(lo << 16), v1 = __uint_as_float(lo & 0xffff0000u);
                const float v2 = __uint_as_float(hi << 16), v3 = __uint_as_float(hi & 0xffff0000u);
                s1[nt][0] += v0 + v2; s1[nt][1] += v1 + v3;
                s2[nt][0] += v0 * v0 + v2 * v2; s2[nt][1] += v1 * v1 + v3 * v3;
            }
        }
        if (stats != nullptr) {
            // column sums over the warp's 32 pixels: lanes with the same q hold the same columns -> reduce over g (lane bits 2..4)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float a = s1[nt][j], c = s2[nt][j];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
                    if (g == 0) { atomicAdd(&sstat[nt * 8 + 2 * q + j][0], a); atomicAdd(&sstat[nt * 8 + 2 * q + j][1], c); }
                }
        }
        __syncwarp();
        {
            constexpr int C8 = CO / 8, TOTAL = 32 * C8;
            uint4* dst = reinterpret_cast<uint4*>(y + ((size_t)(n * Ho + oh0 + warp) * Wo + ow0) * CO);
#pragma unroll
            for (int j = lane; j < TOTAL; j += 32) {
                const int row = j / C8, col = j - row * C8;
                dst[j] = *reinterpret_cast<const uint4*>(stw + row * SS + col * 8);
            }
        }
        if (ST == 2) __syncthreads();      // slot `slot` and the staging rows are free for tile t + 1 / t + 2
        else __syncwarp();
    }
    if (stats != nullptr && cur_grp >= 0) flush_stats(cur_grp);
}

// ------------------------------------------------------------------------------------------------ data gradient
// ConvTranspose2d(CO -> Ci, k4 s2 p1): tile = 8 x 32 input (q) pixels + one-pixel halo -> 16 x 64 output pixels; warp = q row.
// Output row parity ph: kh = ph + 1 from input row q, kh = ph + 3 (ph = 0) / ph - 1 (ph = 1) from row q - 1 / q + 1; the same along
// columns.  Per row parity the warp accumulates both column parities (2 x 4 taps, K = 4 * CO each) in registers.  Persistent
// like the forward kernel: weights once per CTA, dy tiles through a two-deep cp.async ring.
template <int CO, int NT, int MB>
__global__ void __launch_bounds__(256, MB)
narrow_dgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pd, const float* __restrict__ bias, bf16* __restrict__ dx,
                    int Hi, int Wi, int act, int tiles_w, int tiles_h, int total) {
    constexpr int CI = NT * 8, PS = CO + 16, HW = NC_TW + 2, PIX = (NC_TH + 2) * HW, WS = 16 * CO + 16, SS = CI + 8;
    constexpr int TILE = PIX * PS;
    extern __shared__ __align__(16) uint8_t nsmem[];
    bf16* xs0 = reinterpret_cast<bf16*>(nsmem);                  // [2][PIX][PS]
    bf16* wsm = xs0 + 2 * TILE;                                   // [CI][WS]
    bf16* stage = wsm + CI * WS;                                  // [8 warps][64 output pixels of one row][SS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int t_begin = (int)(((int64_t)blockIdx.x * total) / gridDim.x), t_end = (int)(((int64_t)(blockIdx.x + 1) * total) / gridDim.x);

    auto issue_tile = [&](int t, int slot) {
        constexpr int CH = CO / 8;
        const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
        const int qh0 = th * NC_TH, qw0 = tw * NC_TW;
        const bf16* xin = dy + (size_t)n * Hi * Wi * CO;
        const uint32_t dst0 = nsaddr(xs0 + slot * TILE);
#pragma unroll 4
        for (int i = tid; i < PIX * CH; i += 256) {
            const int p = i / CH, c = i - p * CH;
            const int r = p / HW, cc = p - r * HW;
            const int ih = qh0 - 1 + r, iw = qw0 - 1 + cc;
            const bool ok = ih >= 0 && ih < Hi && iw >= 0 && iw < Wi;
            const bf16* src = ok ? xin + ((size_t)ih * Wi + iw) * CO + c * 8 : dy;
            cp_async16(dst0 + (uint32_t)(p * PS + c * 8) * 2u, src, ok ? 16 : 0);
        }
    };

    SG_PDL_SYNC();
    if (t_begin < t_end) issue_tile(t_begin, 0);
    cp_async_commit();
    {
        constexpr int CH = 16 * CO / 8;
        for (int i = tid; i < CI * CH; i += 256) {
            const int ci = i / CH, c = i - ci * CH;
            *reinterpret_cast<uint4*>(wsm + ci * WS + c * 8) = __ldg(reinterpret_cast<const uint4*>(pd + (size_t)ci * 16 * CO + c * 8));
        }
    }
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    bf16* stw = stage + warp * 64 * SS;
    const bf16* wrow = wsm + g * WS + 4 * q;          // K slots = permuted channels, 8-byte fragment loads (see the forward kernel)
    float bz[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        bz[nt][0] = bias != nullptr ? __ldg(bias + nt * 8 + 2 * q) : 0.f;
        bz[nt][1] = bias != nullptr ? __ldg(bias + nt * 8 + 2 * q + 1) : 0.f;
    }
    for (int t = t_begin; t < t_end; ++t) {
        const int slot = (t - t_begin) & 1;
        if (t + 1 < t_end) issue_tile(t + 1, slot ^ 1);           // released by the barrier closing tile t - 1
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
        const int qh0 = th * NC_TH, qw0 = tw * NC_TW;
        const bf16* xs = xs0 + slot * TILE;
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
            // (kh, input row offset) of this output row parity
            const int khA = ph + 1, khB = ph ? 0 : 3, dhB = ph ? 1 : -1;
            float acc[2][2][NT][4];                                   // [m-tile][column parity][n-tile]
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int pw = 0; pw < 2; ++pw)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[mt][pw][nt][j] = 0.f;
#pragma unroll
            for (int rsel = 0; rsel < 2; ++rsel) {                    // the two input rows of this parity
                const int kh = rsel ? khB : khA;
                const bf16* xrow = xs + ((warp + 1 + (rsel ? dhB : 0)) * HW + 1) * PS + 4 * q;     // input pixel (row, column qw = 0)
#pragma unroll
                for (int cs = 0; cs < CO / 16; ++cs) {
                    // A fragments of the three input columns qw - 1, qw, qw + 1 for both m-tiles
                    uint32_t a[2][3][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const bf16* p0 = xrow + (mt * 16 + g + d - 1) * PS + cs * 16;
                            const uint2 lo = *reinterpret_cast<const uint2*>(p0), hi = *reinterpret_cast<const uint2*>(p0 + 8 * PS);
                            a[mt][d][0] = lo.x; a[mt][d][1] = hi.x; a[mt][d][2] = lo.y; a[mt][d][3] = hi.y;
                        }
                    // column parity 0: (kw 1, column qw), (kw 3, qw - 1);  parity 1: (kw 2, qw), (kw 0, qw + 1)
#pragma unroll
                    for (int pw = 0; pw < 2; ++pw)
#pragma unroll
                        for (int csel = 0; csel < 2; ++csel) {
                            const int kw = csel ? (pw ? 0 : 3) : pw + 1;
                            const int d = csel ? (pw ? 2 : 0) : 1;
                            const int koff = (kh * 4 + kw) * CO + cs * 16;
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) {
                                const uint2 b = *reinterpret_cast<const uint2*>(wrow + nt * 8 * WS + koff);
#pragma unroll
                                for (int mt = 0; mt < 2; ++mt)
                                    nmma16816(acc[mt][pw][nt], a[mt][d][0], a[mt][d][1], a[mt][d][2], a[mt][d][3], b.x, b.y);
                            }
                        }
                }
            }
            // stage the output row 2*(qh0 + warp) + ph: pixel 2*qw + pw, then 16-byte stores (64 pixels x CI channels, contiguous)
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int pw = 0; pw < 2; ++pw)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float* c = acc[mt][pw][nt];
                        const int px0 = 2 * (mt * 16 + g) + pw, px1 = px0 + 16;
                        *reinterpret_cast<uint32_t*>(stw + px0 * SS + nt * 8 + 2 * q) =
                            pack_bf16x2(act_fwd(c[0] + bz[nt][0], act), act_fwd(c[1] + bz[nt][1], act));
                        *reinterpret_cast<uint32_t*>(stw + px1 * SS + nt * 8 + 2 * q) =
                            pack_bf16x2(act_fwd(c[2] + bz[nt][0], act), act_fwd(c[3] + bz[nt][1], act));
                    }
            __syncwarp();
            constexpr int C8 = CI / 8, TOTAL = 64 * C8;
            uint4* dst = reinterpret_cast<uint4*>(dx + ((size_t)(n * Ho + 2 * (qh0 + warp) + ph) * Wo + 2 * qw0) * CI);
#pragma unroll
            for (int j = lane; j < TOTAL; j += 32) {
                const int row = j / C8, col = j - row * C8;
                dst[j] = *reinterpret_cast<const uint4*>(stw + row * SS + col * 8);
            }
        }
        __syncthreads();       // slot `slot` is free for tile t + 2
    }
}

static int g_nf_grid[4] = {0, 0, 0, 0}, g_nd_grid[4] = {0, 0, 0, 0};     // CTAs per launch (SMs x resident CTAs), per instantiation
// option "narrow": which supported shapes sg_conv_fprop(_stats) / sg_conv_dgrad send here instead of to the tcgen05 kernel, as a
// bit mask -- 1: forward 16 -> 32, 2: data gradient 16 <- 32, 4: forward 32 -> 64, 8: data gradient 32 <- 64 (15 = all, 0 = none).
// Default = the ones measured faster on B200 (tools/bench_conv.py, profiles/bench_conv_r2k_narrow.txt, bench_conv_r2l_direct_tc.txt):
// both directions of 16 <-> 32 and the forward 32 -> 64, all three on the tcgen05 kernels of direct_tc.cu.
int g_use_narrow = 7;
// option "narrow_cfg" (A/B), bits: 4 = the mma.sync kernels of this file instead of the tcgen05 ones of direct_tc.cu wherever both take
// the shape; for the mma.sync kernels: 1 = forward 16->32 as one CTA per SM with a two-deep ring, 2 = data gradient 16<-32 at one CTA per SM
int g_narrow_cfg = 0;

template <typename K>
static int persistent_grid(K kernel, size_t smem, int* cache) {
    if (*cache == 0) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        *cache = SG_NUM_SMS * per_sm;
    }
    return *cache;
}

template <int CI, int NT, int ST, int MB>
static cudaError_t launch_nf(const void* x, const void* pf, const float* bias, void* y, double* stats, int N, int H, int W, int act,
                             int groups, cudaStream_t st) {
    constexpr int CO = NT * 8;
    const int tiles_w = (W / 2) / NC_TW, tiles_h = (H / 2) / NC_TH, total = N * tiles_w * tiles_h;
    const size_t smem = (size_t)ST * (2 * NC_TH + 2) * (2 * NC_TW + 2) * (CI + 8) * 2 + (size_t)CO * (16 * CI + 16) * 2 + (size_t)8 * 32 * (CO + 8) * 2;
    const int cap = persistent_grid(narrow_fprop_kernel<CI, NT, ST, MB>, smem, &g_nf_grid[(CI == 32) + 2 * (ST - 1)]);
    return launch_pdl(narrow_fprop_kernel<CI, NT, ST, MB>, dim3((unsigned)(total < cap ? total : cap)), dim3(256), smem, st, (const bf16*)x,
                      (const bf16*)pf, bias, (bf16*)y, stats, H, W, act, groups > 0 ? N / groups : N, tiles_w, tiles_h, total);
}
template <int CO, int NT, int MB>
static cudaError_t launch_nd(const void* dy, const void* pd, const float* bias, void* dx, int N, int Ho, int Wo, int act,
                             cudaStream_t st) {
    constexpr int CI = NT * 8;
    const int tiles_w = Wo / NC_TW, tiles_h = Ho / NC_TH, total = N * tiles_w * tiles_h;
    const size_t smem = (size_t)2 * (NC_TH + 2) * (NC_TW + 2) * (CO + 16) * 2 + (size_t)CI * (16 * CO + 16) * 2 + (size_t)8 * 64 * (CI + 8) * 2;
    const int cap = persistent_grid(narrow_dgrad_kernel<CO, NT, MB>, smem, &g_nd_grid[(CO == 64) + 2 * (MB - 1)]);
    return launch_pdl(narrow_dgrad_kernel<CO, NT, MB>, dim3((unsigned)(total < cap ? total : cap)), dim3(256), smem, st, (const bf16*)dy,
                      (const bf16*)pd, bias, (bf16*)dx, Ho, Wo, act, tiles_w, tiles_h, total);
}

// direct_tc.cu: the same operators on tcgen05 (staged spatial tile as the UMMA operand)
bool direct_tc_supported(int mode, int Ci, int Co, int Hs, int Ws);
cudaError_t direct_tc_fprop(const void* x, const void* pf, const float* bias, void* y, double* stats, int groups, int N, int H, int W, int Ci,
                            int Co, int act, cudaStream_t st);
cudaError_t direct_tc_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Hi, int Wi, int Ci, int Co, int act,
                            cudaStream_t st);

static bool mma_sync_supported(int Ho, int Wo) { return Ho % NC_TH == 0 && Wo % NC_TW == 0; }
// which implementation a supported shape runs on: tcgen05 (direct_tc.cu) unless option "narrow_cfg" bit 4 asks for mma.sync
static bool use_direct_tc(int mode, int Ci, int Co, int Hs, int Ws, int Ho, int Wo) {
    if (!direct_tc_supported(mode, Ci, Co, Hs, Ws)) return false;
    return !(g_narrow_cfg & 4) || !mma_sync_supported(Ho, Wo);
}

}  // namespace sg

using namespace sg;

extern "C" {

// 1 if the direct narrow-channel kernels take this operator (Conv2d orientation: x [N,H,W,Ci] <-> y [N,Ho,Wo,Co], k4 s2 p1):
// (Ci, Co) in {(16, 32), (32, 64)}, output grid rows % 8 == 0 and columns % 32 == 0.  mode 0 = forward, 1 = data gradient.
int sg_conv_narrow_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    (void)mode;
    if (k != 4 || s != 2 || p != 1 || Ho * 2 != H || Wo * 2 != W || N < 1) return 0;
    if (!((Ci == 16 && Co == 32) || (Ci == 32 && Co == 64))) return 0;
    return (mma_sync_supported(Ho, Wo) || direct_tc_supported(mode, Ci, Co, mode ? Ho : H, mode ? Wo : W)) ? 1 : 0;
}

// Routing policy of sg_conv_fprop / sg_conv_fprop_stats / sg_conv_dgrad: supported AND selected by option "narrow" (above).
int sg_conv_narrow_routed(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    if (!sg_conv_narrow_supported(mode, N, H, W, Ci, Ho, Wo, Co, k, s, p)) return 0;
    const int bit = (mode ? 2 : 1) << (Ci == 32 ? 2 : 0);
    return (g_use_narrow & bit) ? 1 : 0;
}

// y = act(conv(x, W) + bias); stats != NULL: stats[groups][Co][2] += (sum, sum^2) of the stored y per image group
int sg_conv_narrow_fprop(const void* x, const void* pf, const float* bias, void* y, double* stats, int groups, int N, int H, int W,
                         int Ci, int Co, int act, void* stream) {
    SG_REQUIRE(sg_conv_narrow_supported(0, N, H, W, Ci, H / 2, W / 2, Co, 4, 2, 1), "conv_narrow_fprop: unsupported shape N=%d %dx%d %d->%d", N, H, W, Ci, Co);
    SG_REQUIRE(stats == nullptr || (groups >= 1 && N % groups == 0), "conv_narrow_fprop: N %% groups != 0");
    cudaError_t ce = use_direct_tc(0, Ci, Co, H, W, H / 2, W / 2) ? direct_tc_fprop(x, pf, bias, y, stats, groups, N, H, W, Ci, Co, act, SG_STREAM(stream))
                     : Ci == 32 ? launch_nf<32, 8, 1, 1>(x, pf, bias, y, stats, N, H, W, act, groups, SG_STREAM(stream))
                     : (g_narrow_cfg & 1) ? launch_nf<16, 4, 2, 1>(x, pf, bias, y, stats, N, H, W, act, groups, SG_STREAM(stream))
                                          : launch_nf<16, 4, 1, 2>(x, pf, bias, y, stats, N, H, W, act, groups, SG_STREAM(stream));
    if (ce != cudaSuccess) { set_error("conv_narrow_fprop launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    SG_LAUNCHED("conv_narrow_fprop");
    return 0;
}

// dx = act(convT(dy, W) + bias): dy [N,Ho,Wo,Co], pd [Ci][4][4][Co], dx [N,2Ho,2Wo,Ci]
int sg_conv_narrow_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Ho, int Wo, int Ci, int Co, int act,
                         void* stream) {
    SG_REQUIRE(sg_conv_narrow_supported(1, N, 2 * Ho, 2 * Wo, Ci, Ho, Wo, Co, 4, 2, 1), "conv_narrow_dgrad: unsupported shape N=%d %dx%d %d<-%d", N, Ho, Wo, Ci, Co);
    cudaError_t ce = use_direct_tc(1, Ci, Co, Ho, Wo, Ho, Wo) ? direct_tc_dgrad(dy, pd, bias, dx, N, Ho, Wo, Ci, Co, act, SG_STREAM(stream))
                     : Co == 64 ? launch_nd<64, 4, 1>(dy, pd, bias, dx, N, Ho, Wo, act, SG_STREAM(stream))
                     : (g_narrow_cfg & 2) ? launch_nd<32, 2, 1>(dy, pd, bias, dx, N, Ho, Wo, act, SG_STREAM(stream))
                                          : launch_nd<32, 2, 2>(dy, pd, bias, dx, N, Ho, Wo, act, SG_STREAM(stream));
    if (ce != cudaSuccess) { set_error("conv_narrow_dgrad launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    SG_LAUNCHED("conv_narrow_dgrad");
    return 0;
}

}  // extern "C"
