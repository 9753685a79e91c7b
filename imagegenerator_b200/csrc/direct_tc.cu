// direct_tc.cu -- the narrow k4 s2 p1 convolutions on tcgen05, fed from a spatial tile staged ONCE in shared memory.
//
// Why a third conv kernel.  The Stage-II critic's 16 -> 32 (and 32 -> 64) channel layers on 128x128 / 64x64 maps
// (discriminator_2.py:13-18) are HBM-bound: 151 MB for 12.9 GFLOP.  The implicit-GEMM kernel (conv_tc.cu) fetches every tap's
// operand rows with TMA -- 16 channels = 32-byte rows, thousands of them per 128-row tile -- and lands at 1.2 TB/s; the warp-level
// mma.sync kernels (narrow_conv.cu) stop at ~175 TFLOP/s whatever their shape, the legacy tensor path's own rate on this part
// (profiles/bench_conv_r2k_narrow.txt).  Here the input tile is brought in once (cp.async, 16-byte chunks, zero-filled padding)
// in a layout that IS the canonical K-major no-swizzle UMMA operand layout for every tap at once, so one tcgen05.mma per
// (tap, 16 channels) reads its 128 x 16 operand straight out of the staged tile through a shared-memory descriptor:
//
//   forward   M = 16 output rows x 8 output columns.  Row m = (r, c) of tap (kh, kw) is input pixel (2r + kh, 2c + kw) of the
//             tile.  The tile is stored [input row][column parity][8-channel chunk][column / 2][8 channels]: the 8 pixels of a
//             core matrix (fixed r, c = 0..7) are 8 consecutive 16-byte entries (column parity kw & 1, starting at kw >> 1),
//             the next 8-channel chunk is one plane further (LBO), the next output row two input rows further (SBO).
//   data      M = 16 x 8 INPUT (dy) pixels q; the four output parities are the N dimension: N = (ph, pw, ci), K runs over the
//   gradient  nine neighbours (dr, dc) of q times the dy channels, the weight matrix holds tap (ph + 1 - 2dr, pw + 1 - 2dc) or
//             zero.  One accumulator row is then two output rows x two output pixels x ci -- 2 x 64 contiguous bytes.
//
// Persistent CTAs (one per SM, a contiguous range of tiles each): warp 0 issues the MMAs, warps 1-4 drain the accumulators
// (double-buffered in TMEM: tcgen05.ld -> bias / activation -> bf16 -> 16-byte stores, BatchNorm statistics by a transposing
// butterfly into per-lane running sums), warps 5-8 stage the tiles (two-deep ring).  Weights are staged once per CTA.
#include "common.cuh"
#include <type_traits>

namespace sg {
namespace {

__device__ __forceinline__ uint32_t dsaddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsaddr(bar)), "r"(count));
}
__device__ __forceinline__ void dbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dsaddr(bar)) : "memory");
}
// bounded wait: a protocol bug traps (a CUDA error on the host) instead of hanging the device
__device__ __forceinline__ void dbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = dsaddr(bar);
    uint32_t done;
    for (uint32_t spins = 0;; ++spins) {
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
        if (spins > (1u << 24)) __trap();
        if (spins > 64) __nanosleep(64);
    }
}
__device__ __forceinline__ bool dtc_elect() {
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
__device__ __forceinline__ void dcp16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dcp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes (cp.async, st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void dfence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void dtc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void dtc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void dtc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dsaddr(bar)) : "memory");
}
__device__ __forceinline__ void dtc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void dtc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no swizzle ("interleave"): core matrices of 8 rows x 16 bytes, contiguous; lbo = bytes between the two core
// matrices of an MMA's K = 16, sbo = bytes between 8-row groups (cute/atom/mma_traits_sm100.hpp: ((8,n),2):((1,SBO),LBO))
__device__ __forceinline__ uint64_t dtc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    return d;                          // layout type 0 = no swizzle
}
// K-major, 128-byte swizzle: rows of 128 bytes, atoms of 8 rows (1024 bytes, 1024-aligned), sbo = bytes between atoms along M/N
__device__ __forceinline__ uint64_t dtc_desc_sw128(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
// 32 columns x 32 lanes -> lane l keeps the warp total of column l (reduce-scatter: 31 shuffles instead of 160)
__device__ __forceinline__ float dtc_colsum(float (&a)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float keep = up ? a[i + o] : a[i];
            const float send = up ? a[i] : a[i + o];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return a[0];
}

constexpr int DT_THREADS = 288;      // warp 0: MMA, warps 1-4: epilogue, warps 5-8: tile producers
constexpr int DT_ROWS = 16;          // M = 16 rows x 8 columns per MMA

// MODE 0 = forward (x [n,H,W,CI] -> y [n,H/2,W/2,CO]), 1 = data gradient (dy [n,Hi,Wi,CK] -> dx [n,2Hi,2Wi,CN]).
// CK = channels contracted (the staged tensor's), CN = channels produced, NMT = M tiles side by side (tile = 16 x 8*NMT).
template <int MODE, int CK, int CN, int NMT, bool STATS>
__global__ void __launch_bounds__(DT_THREADS, 1)
direct_tc_kernel(const bf16* __restrict__ src, const bf16* __restrict__ wgt, const float* __restrict__ bias, bf16* __restrict__ dst,
                 double* __restrict__ stats, int Hs, int Ws, int act, int imgs_per_group, int tiles_w, int tiles_h, int total) {
    // SW (forward, 16 channels): 128-byte-swizzled operand rows instead of the 16-byte core-matrix rows of the no-swizzle layout.
    // The tensor core fetches operand ROWS, about one row slice per cycle whatever its width -- measured with the no-swizzle
    // layout: 64 MMAs of 128 x 32 x 16 took 17 k cycles per tile = 128 x 2 + 32 x 2 fetches each (profiles/bench_conv_r2l_direct_tc.txt).
    // A 128-byte row here = the four kw taps of one output pixel = four consecutive input pixels x 16 channels, contiguous in NHWC
    // memory.  Rows of neighbouring output columns overlap by two pixels, so the tile is staged twice: copy 0 holds the quads of
    // the EVEN output columns (input columns 4e .. 4e+3), copy 1 those of the odd ones (4e+2 .. 4e+5); an M tile = 16 output rows x
    // the 8 even (odd) columns of a 16-column tile, one swizzle atom per output row.  K = 64 per filter row kh: four MMAs.
    constexpr bool SW = MODE == 0 && CK == 16;
    static_assert(!SW || NMT == 2, "swizzled forward layout: M tiles = the even and the odd columns");
    constexpr int NCH = CK / 8, TCOLS = 8 * NMT;
    constexpr int IR = MODE == 0 ? 2 * DT_ROWS + 2 : DT_ROWS + 2;            // staged rows
    constexpr int IC = MODE == 0 ? 2 * TCOLS + 2 : TCOLS + 2;                // staged columns
    constexpr int CP = (MODE == 0 ? TCOLS + 1 : TCOLS + 2) * 16;             // bytes of one (row, [parity,] chunk) plane
    constexpr int RB = SW ? 2048 : (MODE == 0 ? 2 : 1) * NCH * CP;           // bytes of one staged row
    constexpr int TILE = IR * RB;
    constexpr int NN = MODE == 0 ? CN : 4 * CN;                              // MMA N
    constexpr int KPOS = MODE == 0 ? 16 : 9;                                 // taps / neighbours
    constexpr int KT = KPOS * CK;                                            // weight matrix K
    constexpr int ACC = NMT * NN;                                            // TMEM columns per accumulator buffer
    static_assert(2 * ACC <= 512 && NN % 32 == 0 && NN <= 256, "accumulators do not fit");
    extern __shared__ __align__(128) uint8_t dsm_raw[];
    uint8_t* dsm = dsm_raw + ((1024u - (dsaddr(dsm_raw) & 1023u)) & 1023u);   // swizzle atoms are 1024-byte aligned
    uint8_t* wsm = dsm + 2 * TILE;                                            // [NN / 8][KT / 8][8 rows][16 bytes]; SW: [kh][NN / 8][8 rows][128 B]
    __shared__ uint64_t full_bar[2], empty_bar[2], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float sbias[CN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = (int)(((int64_t)blockIdx.x * total) / gridDim.x), t_end = (int)(((int64_t)(blockIdx.x + 1) * total) / gridDim.x);
    const int ntiles = t_end - t_begin;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            dbar_init(&full_bar[i], 128); dbar_init(&empty_bar[i], 1);
            dbar_init(&tfull_bar[i], 1); dbar_init(&tempty_bar[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dsaddr(&tmem_slot)), "r"((uint32_t)(2 * ACC))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    SG_PDL_SYNC();
    // ---- weights, once per CTA, into the K-major core-matrix layout
    if (SW) {
        // wgt = pf [CN][kh][kw][16]: per kh a K-major block of CN rows x 128 bytes (kw, ci), 128-byte swizzled
        for (int i = tid; i < NN * 4 * 8; i += DT_THREADS) {
            const int ch = i & 7, kh = (i >> 3) & 3, n = i >> 5;
            *reinterpret_cast<uint4*>(wsm + kh * (NN * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((ch ^ (n & 7)) << 4)) =
                __ldg(reinterpret_cast<const uint4*>(wgt + (size_t)n * KT + kh * 64 + ch * 8));
        }
    } else if (MODE == 0) {
        // wgt = pf [CN][16 taps][CK]: row n, K index tap * CK + ck
        for (int i = tid; i < NN * (KT / 8); i += DT_THREADS) {
            const int n = i / (KT / 8), k8 = i - n * (KT / 8);
            *reinterpret_cast<uint4*>(wsm + ((n >> 3) * (KT / 8) + k8) * 128 + (n & 7) * 16) =
                __ldg(reinterpret_cast<const uint4*>(wgt + (size_t)n * KT + k8 * 8));
        }
    } else {
        // wgt = pd [CN][16 taps][CK]; row n = (ph, pw, ci), K index pos * CK + ck, pos = (dr + 1) * 3 + (dc + 1):
        // tap (ph + 1 - 2 dr, pw + 1 - 2 dc) or zero
        for (int i = tid; i < NN * (KT / 8); i += DT_THREADS) {
            const int n = i / (KT / 8), k8 = i - n * (KT / 8);
            const int pos = k8 / NCH, j = k8 - pos * NCH;
            const int ph = n / (2 * CN), pw = (n / CN) & 1, ci = n % CN;
            const int kh = ph + 1 - 2 * (pos / 3 - 1), kw = pw + 1 - 2 * (pos % 3 - 1);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (kh >= 0 && kh < 4 && kw >= 0 && kw < 4) v = __ldg(reinterpret_cast<const uint4*>(wgt + ((size_t)ci * 16 + kh * 4 + kw) * CK + j * 8));
            *reinterpret_cast<uint4*>(wsm + ((n >> 3) * (KT / 8) + k8) * 128 + (n & 7) * 16) = v;
        }
    }
    if (tid < CN) sbias[tid] = bias != nullptr ? __ldg(bias + tid) : 0.f;
    dfence_async();
    dtc_before();
    __syncthreads();
    dtc_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        // The whole warp waits; ONE elected lane issues a tile's MMAs back to back.  Descriptors differ only in the 14-bit start
        // address of their low word: everything per MMA is an integer add (conv_tc.cu measured ~157 cycles per tcgen05.mma when
        // each descriptor was rebuilt with shifts and masks under a divergent `lane == 0`).
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t A_SBO = SW ? 2 * RB : (MODE == 0 ? 2 * RB : RB), B_SBO = SW ? 1024 : (KT / 8) * 128;
        constexpr uint32_t A_HI = (A_SBO >> 4) | (1u << 14) | (SW ? (2u << 29) : 0u);
        constexpr uint32_t B_HI = (B_SBO >> 4) | (1u << 14) | (SW ? (2u << 29) : 0u);
        constexpr uint32_t A_LBO = SW ? 1u : (uint32_t)(CP >> 4), B_LBO = SW ? 1u : (uint32_t)(128 >> 4);
        const uint32_t blo = ((dsaddr(wsm) >> 4) & 0x3FFFu) | (B_LBO << 16);
        auto mma = [&](uint32_t tacc, uint32_t alo, uint32_t bl, uint32_t acc) {
            dtc_mma(tacc, ((uint64_t)A_HI << 32) | alo, ((uint64_t)B_HI << 32) | bl, idesc, acc);
        };
        for (int it = 0; it < ntiles; ++it) {
            const int s = it & 1, k = it >> 1;
            if (k >= 1) dbar_wait(&tempty_bar[s], (k - 1) & 1);
            dbar_wait(&full_bar[s], k & 1);
            dtc_after();
            const uint32_t alo = ((dsaddr(dsm + s * TILE) >> 4) & 0x3FFFu) | (A_LBO << 16);
            if (dtc_elect()) {
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt) {
                    const uint32_t tacc = tmem_base + (uint32_t)(s * ACC + mt * NN);
                    if (SW) {
#pragma unroll
                        for (int kh = 0; kh < 4; ++kh)
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                mma(tacc, alo + (uint32_t)((kh * RB + mt * 1024 + k4 * 32) >> 4), blo + (uint32_t)((kh * (NN * 128) + k4 * 32) >> 4),
                                    (kh | k4) != 0 ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int pos = 0; pos < KPOS; ++pos) {
                            uint32_t aoff;
                            if (MODE == 0) {
                                const int kh = pos >> 2, kw = pos & 3;
                                aoff = kh * RB + (kw & 1) * NCH * CP + (8 * mt + (kw >> 1)) * 16;
                            } else {
                                const int dr = pos / 3, dc = pos % 3;             // 0..2 = offset + 1
                                aoff = dr * RB + (8 * mt + dc) * 16;
                            }
#pragma unroll
                            for (int kc = 0; kc < CK / 16; ++kc)
                                mma(tacc, alo + ((aoff + 2 * kc * CP) >> 4), blo + (uint32_t)(((pos * NCH + 2 * kc) * 128) >> 4), (pos | kc) != 0 ? 1u : 0u);
                        }
                    }
                }
                dtc_commit(&empty_bar[s]);
                dtc_commit(&tfull_bar[s]);
            }
            __syncwarp();
        }
    } else if (warp <= 4) {
        // ------------------------------------------------------------------------------------------ epilogue
        const int qd = warp & 3;                                   // the TMEM lane quarter this warp may read
        const int r = 4 * qd + (lane >> 3), c = lane & 7;          // accumulator row = (r, c) of the M tile
        const int Hd = MODE == 0 ? Hs >> 1 : Hs << 1, Wd = MODE == 0 ? Ws >> 1 : Ws << 1;
        float rs1[NN / 32], rs2[NN / 32];
#pragma unroll
        for (int i = 0; i < NN / 32; ++i) rs1[i] = rs2[i] = 0.f;
        int cur_grp = -1;
        auto flush = [&](int grp) {
#pragma unroll
            for (int i = 0; i < NN / 32; ++i) {
                atomicAdd(stats + ((size_t)grp * CN + i * 32 + lane) * 2, (double)rs1[i]);
                atomicAdd(stats + ((size_t)grp * CN + i * 32 + lane) * 2 + 1, (double)rs2[i]);
                rs1[i] = rs2[i] = 0.f;
            }
        };
        for (int it = 0; it < ntiles; ++it) {
            const int t = t_begin + it, s = it & 1, k = it >> 1;
            const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
            if (STATS) {
                const int grp = n / imgs_per_group;
                if (grp != cur_grp) {
                    if (cur_grp >= 0) flush(cur_grp);
                    cur_grp = grp;
                }
            }
            dbar_wait(&tfull_bar[s], k & 1);
            dtc_after();
#pragma unroll 1
            for (int mt = 0; mt < NMT; ++mt) {
#pragma unroll
                for (int cc = 0; cc < NN / 32; ++cc) {
                    uint32_t v[32];
                    dtc_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(s * ACC + mt * NN + cc * 32), v);
                    uint32_t pk[16];
                    auto finish = [&](auto actc) {          // act is uniform: one branch per chunk, not a switch per element
                        constexpr int A = decltype(actc)::value;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c0 = (cc * 32 + 2 * j) % CN;
                            pk[j] = pack_bf16x2(act_fwd(__uint_as_float(v[2 * j]) + sbias[c0], A),
                                                act_fwd(__uint_as_float(v[2 * j + 1]) + sbias[c0 + 1], A));
                        }
                    };
                    if (act == SG_ACT_NONE) finish(std::integral_constant<int, SG_ACT_NONE>{});
                    else if (act == SG_ACT_LRELU) finish(std::integral_constant<int, SG_ACT_LRELU>{});
                    else if (act == SG_ACT_RELU) finish(std::integral_constant<int, SG_ACT_RELU>{});
                    else finish(std::integral_constant<int, SG_ACT_TANH>{});
                    bf16* o;
                    if (SW) {               // M tile mt = the even (odd) columns of the tile
                        o = dst + ((size_t)(n * Hd + th * DT_ROWS + r) * Wd + tw * TCOLS + 2 * c + mt) * CN + cc * 32;
                    } else if (MODE == 0) {
                        o = dst + ((size_t)(n * Hd + th * DT_ROWS + r) * Wd + tw * TCOLS + 8 * mt + c) * CN + cc * 32;
                    } else {
                        const int ph = (cc * 32) / (2 * CN), off = (cc * 32) % (2 * CN);
                        o = dst + ((size_t)(n * Hd + 2 * (th * DT_ROWS + r) + ph) * Wd + 2 * (tw * TCOLS + 8 * mt + c)) * CN + off;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        reinterpret_cast<uint4*>(o)[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    if (STATS) {
                        float a[32], b[32];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            a[2 * j] = __uint_as_float(pk[j] << 16);
                            a[2 * j + 1] = __uint_as_float(pk[j] & 0xffff0000u);
                            b[2 * j] = a[2 * j] * a[2 * j];
                            b[2 * j + 1] = a[2 * j + 1] * a[2 * j + 1];
                        }
                        rs1[cc] += dtc_colsum(a, lane);
                        rs2[cc] += dtc_colsum(b, lane);
                    }
                }
            }
            dtc_before();
            dbar_arrive(&tempty_bar[s]);
        }
        if (STATS && cur_grp >= 0) flush(cur_grp);
    } else {
        // ------------------------------------------------------------------------------------------ tile producers
        const int ptid = tid - 160;
        for (int it = 0; it < ntiles; ++it) {
            const int t = t_begin + it, s = it & 1, k = it >> 1;
            const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
            const int ih0 = MODE == 0 ? 2 * th * DT_ROWS - 1 : th * DT_ROWS - 1;
            const int iw0 = MODE == 0 ? 2 * tw * TCOLS - 1 : tw * TCOLS - 1;
            const bf16* sin = src + (size_t)n * Hs * Ws * CK;
            if (k >= 1) dbar_wait(&empty_bar[s], (k - 1) & 1);
            const uint32_t d0 = dsaddr(dsm + s * TILE);
            // a thread owns the same 16-byte column chunk(s) in every staged row: column, bounds and destination once per tile,
            // then one cp.async and a handful of adds per row (the first version recomputed everything per chunk: 45 instructions
            // per cp.async, 6 k of the kernel's 8 k warp instructions per tile -- ncu: issue-bound at 1.2 IPC)
            constexpr int PER_ROW = SW ? 128 : IC * NCH, NCC = (PER_ROW + 127) / 128;
            int goff[NCC];
            uint32_t doff[NCC];
            bool okc[NCC], have[NCC];
#pragma unroll
            for (int q = 0; q < NCC; ++q) {
                const int cc = ptid + 128 * q;
                have[q] = cc < PER_ROW;
                int iw;
                if (SW) {       // chunk ch of quad e of copy cp: input column 2 cp + 4 e + ch / 2 (1 KB contiguous per (row, copy))
                    const int ch = cc & 7, e = (cc >> 3) & 7, cp = cc >> 6;
                    iw = iw0 + 2 * cp + 4 * e + (ch >> 1);
                    goff[q] = iw * CK + (ch & 1) * 8;
                    doff[q] = cp * 1024 + e * 128 + ((ch ^ e) << 4);
                } else {
                    const int cx = cc / NCH, j = cc - cx * NCH;
                    iw = iw0 + cx;
                    goff[q] = iw * CK + j * 8;
                    doff[q] = MODE == 0 ? ((cx & 1) * NCH + j) * CP + (cx >> 1) * 16 : j * CP + cx * 16;
                }
                okc[q] = have[q] && iw >= 0 && iw < Ws;
            }
#pragma unroll 2
            for (int R = 0; R < IR; ++R) {
                const int ih = ih0 + R;
                const bool okr = ih >= 0 && ih < Hs;
                const bf16* rowp = sin + (size_t)(okr ? ih : 0) * Ws * CK;
#pragma unroll
                for (int q = 0; q < NCC; ++q) {
                    const bool ok = okr && okc[q];
                    if (have[q]) dcp16(d0 + R * RB + doff[q], ok ? rowp + goff[q] : src, ok ? 16 : 0);
                }
            }
            dcp_commit();
            if (it > 0) {                  // the previous tile has landed: publish it while this one is in flight
                dcp_wait<1>();
                dfence_async();
                dbar_arrive(&full_bar[s ^ 1]);
            }
        }
        if (ntiles > 0) {
            dcp_wait<0>();
            dfence_async();
            dbar_arrive(&full_bar[(ntiles - 1) & 1]);
        }
    }
    dtc_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * ACC)) : "memory");
    }
}

template <int MODE, int CK, int CN, int NMT>
constexpr size_t direct_tc_smem() {
    constexpr int NCH = CK / 8, TCOLS = 8 * NMT;
    constexpr int IR = MODE == 0 ? 2 * DT_ROWS + 2 : DT_ROWS + 2;
    constexpr int CP = (MODE == 0 ? TCOLS + 1 : TCOLS + 2) * 16;
    constexpr int RB = (MODE == 0 && CK == 16) ? 2048 : (MODE == 0 ? 2 : 1) * NCH * CP;
    constexpr int NN = MODE == 0 ? CN : 4 * CN, KT = (MODE == 0 ? 16 : 9) * CK;
    return (size_t)2 * IR * RB + (size_t)NN * KT * 2 + 1024;
}

template <int MODE, int CK, int CN, int NMT, bool STATS>
cudaError_t launch_direct_tc(const void* src, const void* wgt, const float* bias, void* dst, double* stats, int groups, int N, int Hs, int Ws,
                             int act, cudaStream_t st) {
    constexpr size_t smem = direct_tc_smem<MODE, CK, CN, NMT>();
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(direct_tc_kernel<MODE, CK, CN, NMT, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const int Hm = MODE == 0 ? Hs / 2 : Hs, Wm = MODE == 0 ? Ws / 2 : Ws;        // the grid the M tiles cover
    const int tiles_w = Wm / (8 * NMT), tiles_h = Hm / DT_ROWS, total = N * tiles_w * tiles_h;
    return launch_pdl(direct_tc_kernel<MODE, CK, CN, NMT, STATS>, dim3((unsigned)(total < SG_NUM_SMS ? total : SG_NUM_SMS)), dim3(DT_THREADS),
                      smem, st, (const bf16*)src, (const bf16*)wgt, bias, (bf16*)dst, stats, Hs, Ws, act, groups > 0 ? N / groups : N, tiles_w,
                      tiles_h, total);
}

}  // namespace

// Entry points for narrow_conv.cu (which owns the C ABI of these shapes).  Return cudaErrorInvalidValue for a shape this file
// has no instantiation for.  mode 0: Hs x Ws = input size, (Ci, Co); mode 1: Hs x Ws = dy size, dy has Co channels, dx Ci.
bool direct_tc_supported(int mode, int Ci, int Co, int Hs, int Ws) {
    const int Hm = mode == 0 ? Hs / 2 : Hs, Wm = mode == 0 ? Ws / 2 : Ws;
    if (Hm % DT_ROWS != 0) return false;
    if (mode == 0 && Ci == 16 && Co == 32) return Wm % 16 == 0;
    if (mode == 0 && Ci == 32 && Co == 64) return Wm % 16 == 0;
    if (mode == 1 && Ci == 16 && Co == 32) return Wm % 32 == 0;
    return false;
}
cudaError_t direct_tc_fprop(const void* x, const void* pf, const float* bias, void* y, double* stats, int groups, int N, int H, int W, int Ci,
                            int Co, int act, cudaStream_t st) {
    if (Ci == 16 && Co == 32)
        return stats ? launch_direct_tc<0, 16, 32, 2, true>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 16, 32, 2, false>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 32 && Co == 64)
        return stats ? launch_direct_tc<0, 32, 64, 2, true>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 32, 64, 2, false>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    return cudaErrorInvalidValue;
}
cudaError_t direct_tc_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Hi, int Wi, int Ci, int Co, int act,
                            cudaStream_t st) {
    if (Ci == 16 && Co == 32) return launch_direct_tc<1, 32, 16, 4, false>(dy, pd, bias, dx, nullptr, 1, N, Hi, Wi, act, st);
    return cudaErrorInvalidValue;
}

}  // namespace sg
