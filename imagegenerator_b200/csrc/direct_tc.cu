// direct_tc.cu -- the narrow k4 s2 p1 convolutions on tcgen05, fed from spatial tiles staged by TMA.
//
// Why a third conv kernel.  The Stage-II critic's 16 -> 32 (and 32 -> 64) channel layers on 128x128 / 64x64 maps
// (discriminator_2.py:13-18) are HBM-bound: 151 MB for 12.9 GFLOP.  The implicit-GEMM kernel (conv_tc.cu) brings every tap's
// operand in 64-channel k-blocks -- with 16 channels a k-block is four taps = four TMA boxes of 32-byte rows per 128-row tile --
// and lands at 1.2 TB/s; the warp-level mma.sync kernels (narrow_conv.cu) plateau at 2 TB/s (profiles/bench_conv_r2k_narrow.txt).
// Here a CTA stages a SPATIAL tile of the input and every tap's operand is a window of it, read by tcgen05.mma through a
// shared-memory descriptor -- no im2col anywhere, each input byte crosses L2 -> SM once per column phase:
//
//   forward   M tile = 16 output rows x 8 output columns.  Row m = (r, c) of tap (kh, kw) is input pixel (2r + kh, 2c + kw - 1).
//             One TMA box per kw with element stride 2 along W brings [34 input rows][16 pixels two apart][C channels]; a
//             pixel's C channels are one swizzle span (32 / 64 bytes), eight neighbouring output columns one swizzle atom,
//             the next output row two staged rows further (SBO).  Tap (kh, kw) = copy kw, start row kh.  K = C per tap.
//   data      M tile = 16 x 8 INPUT (dy) pixels q; the four output parities are the N dimension: N = (ph, pw, ci), K runs over
//   gradient  the nine neighbours (dr, dc) of q times the dy channels; the weight matrix holds tap (ph + 1 - 2dr, pw + 1 - 2dc)
//             or zero.  One box per dc (shifted by a pixel, so every window starts on an atom).  An accumulator row is two
//             output rows x two output pixels x ci: 2 x 64 contiguous bytes.
//
// The conv's zero padding is TMA's out-of-bounds fill.  Persistent CTAs (one per SM, a contiguous range of tiles each): warp 0
// issues the MMAs (one elected lane, descriptors advanced by integer adds), warps 1-4 drain the accumulators (double-buffered in
// TMEM: tcgen05.ld -> bias / activation -> bf16 -> 16-byte stores, BatchNorm statistics by a transposing butterfly into per-lane
// running sums), one lane of warp 5 issues the TMA boxes of a two-deep tile ring.  Weights are staged once per CTA.
//
// History (profiles/bench_conv_r2l_direct_tc.txt): the first versions staged the tile with cp.async (16 bytes per thread and
// instruction) -- correct, and stuck at 2.0-2.4 TB/s like the mma.sync kernels whatever the MMA layout: the per-thread copies, not
// the contraction, were the common ceiling.
#include "common.cuh"
#include <cuda.h>
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
__device__ __forceinline__ void dbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsaddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dtma_4d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(map), "r"(dsaddr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
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

template <int MODE, int CK, int CN, int NMT, int ONE>
constexpr size_t direct_tc_tile_bytes() {
    constexpr int IR = MODE == 0 ? 2 * 16 + 2 : 16 + 2, NCOPY = ONE == 2 ? 1 : ONE ? (MODE == 0 ? 2 : 1) : (MODE == 0 ? 4 : 3);
    constexpr int BOXW = ONE == 2 ? 16 * NMT + 4 : ONE ? (MODE == 0 ? 8 * NMT + 1 : 8 * NMT + 2) : 8 * NMT;
    constexpr int ROWB = BOXW * CK * 2, COPYB = (IR * ROWB + 1023) / 1024 * 1024;
    return (size_t)NCOPY * COPYB;
}
template <int MODE, int CK, int CN, int NMT, int ONE>
constexpr size_t direct_tc_weight_bytes() { return (size_t)(MODE == 0 ? CN : 4 * CN) * (MODE == 0 ? 16 : 9) * CK * 2; }
template <int MODE, int CK, int CN, int NMT, int ONE>
constexpr int direct_tc_slots() {
    constexpr size_t room = 226 * 1024 - 1024 - direct_tc_weight_bytes<MODE, CK, CN, NMT, ONE>();
    constexpr size_t n = room / direct_tc_tile_bytes<MODE, CK, CN, NMT, ONE>();
    return n > 4 ? 4 : (int)n;
}
template <int MODE, int CK, int CN, int NMT, int ONE>
constexpr size_t direct_tc_smem() {
    return direct_tc_slots<MODE, CK, CN, NMT, ONE>() * direct_tc_tile_bytes<MODE, CK, CN, NMT, ONE>() + direct_tc_weight_bytes<MODE, CK, CN, NMT, ONE>() + 1024;
}

constexpr int DT_THREADS = 192;      // warp 0: MMA, warps 1-4: epilogue, warp 5: TMA producer
constexpr int DT_ROWS = 16;          // M = 16 rows x 8 columns per MMA

// MODE 0 = forward (x [n,H,W,CK] -> y [n,H/2,W/2,CN]), 1 = data gradient (dy [n,Hi,Wi,CK] -> dx [n,2Hi,2Wi,CN]).
// CK = channels contracted (the staged tensor's: 16 or 32), CN = channels produced, NMT = M tiles side by side (tile = 16 x 8*NMT).
// ONE = 1: windows that start INSIDE a swizzle atom -- the forward kernel stages two copies (pixel parities, one column more) instead
// of four, the gradient kernel one (two columns more) instead of three; tap kw / neighbour dc then shifts the window by a pixel.
// Legal because the swizzle XOR is taken from absolute address bits by TMA and tensor core alike (bit-identical results, measured).
// Forward 16 -> 32: 61.8 -> 47.2 us (half the staged bytes); gradient: 51.3 -> 57.4 us, so it keeps its three copies.
template <int MODE, int CK, int CN, int NMT, bool STATS, int ONE>
__global__ void __launch_bounds__(DT_THREADS, 1)
direct_tc_kernel(const __grid_constant__ CUtensorMap tm, const bf16* __restrict__ wgt, const float* __restrict__ bias, bf16* __restrict__ dst,
                 double* __restrict__ stats, int Hs, int Ws, int act, int imgs_per_group, int tiles_w, int tiles_h, int total, int diag) {
    constexpr int NCH = CK / 8, TCOLS = 8 * NMT, PIXB = CK * 2;             // bytes of a staged pixel = the swizzle span
    constexpr int IR = MODE == 0 ? 2 * DT_ROWS + 2 : DT_ROWS + 2;            // staged rows
    // ONE = 2 (forward): ONE contiguous box per tile over the tensor seen as PIXEL PAIRS [n][H][W/2][2C]: [34 rows][TCOLS + 2 pairs],
    // starting one pair left of the tile (pixel 2 ow0 - 2; the pair grid is the memory's, so out-of-bounds fill stays exact).  An
    // operand row is a pair (= the swizzle span), tap kw starts kw + 1 pixels into the staged row and its K slice is that pixel's
    // half of the pair -- no strided gather in the TMA unit.  (A 16-channel inner box under a 64-byte swizzle does NOT land densely:
    // garbage and an illegal address, measured.)
    static_assert(ONE != 2 || MODE == 0, "pixel-pair rows are the forward kernel's");
    constexpr int NCOPY = ONE == 2 ? 1 : ONE ? (MODE == 0 ? 2 : 1) : (MODE == 0 ? 4 : 3);   // boxes per tile
    constexpr int BOXW = ONE == 2 ? 2 * TCOLS + 4 : ONE ? (MODE == 0 ? TCOLS + 1 : TCOLS + 2) : TCOLS;  // staged pixels per row
    constexpr int SPAN = ONE == 2 ? 2 * PIXB : PIXB;                          // swizzle span = pitch of the operand rows
    constexpr int ROWB = BOXW * PIXB, COPYB = (IR * ROWB + 1023) / 1024 * 1024, TILE = NCOPY * COPYB;
    constexpr int NN = MODE == 0 ? CN : 4 * CN;                              // MMA N
    constexpr int KPOS = MODE == 0 ? 16 : 9;                                 // taps / neighbours
    constexpr int KT = KPOS * CK;                                            // weight matrix K
    constexpr int ACC = NMT * NN;                                            // TMEM columns per accumulator buffer
    constexpr int TM_COLS = 2 * ACC < 32 ? 32 : 2 * ACC;
    // depth of the tile ring: as many slots as fit next to the weights (up to 4).  With two, the kernels were bound by the LATENCY of
    // the strided TMA boxes -- 37 us with neither MMAs nor stores (option dtc_diag = 6), 2 tiles in flight per SM
    constexpr int NST = direct_tc_slots<MODE, CK, CN, NMT, ONE>();
    static_assert(SPAN == 32 || SPAN == 64 || SPAN == 128, "operand rows must be one swizzle span");
    static_assert(2 * ACC <= 512 && (TM_COLS & (TM_COLS - 1)) == 0 && NN % 32 == 0 && NN <= 256, "accumulators do not fit");
    extern __shared__ __align__(128) uint8_t dsm_raw[];
    uint8_t* dsm = dsm_raw + ((1024u - (dsaddr(dsm_raw) & 1023u)) & 1023u);   // swizzle atoms: 1024-byte aligned bases
    uint8_t* wsm = dsm + NST * TILE;                                          // [NN / 8][KT / 8][8 rows][16 bytes] (no swizzle)
    __shared__ uint64_t full_bar[NST], empty_bar[NST], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float sbias[CN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = (int)(((int64_t)blockIdx.x * total) / gridDim.x), t_end = (int)(((int64_t)(blockIdx.x + 1) * total) / gridDim.x);
    const int ntiles = t_end - t_begin;

    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { dbar_init(&full_bar[i], 1); dbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { dbar_init(&tfull_bar[i], 1); dbar_init(&tempty_bar[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dsaddr(&tmem_slot)), "r"((uint32_t)TM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    SG_PDL_SYNC();
    // ---- weights, once per CTA, into the K-major core-matrix layout (8 rows x 16 bytes, K chunks 128 bytes apart)
    if (MODE == 0) {
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
        constexpr uint32_t A_SBO = (MODE == 0 ? 2 : 1) * ROWB;                 // the next output (dy) row: two (one) staged rows further
        constexpr uint32_t A_HI = (A_SBO >> 4) | (1u << 14) | ((SPAN == 32 ? 6u : SPAN == 64 ? 4u : 2u) << 29);      // SWIZZLE_32B / 64B / 128B
        constexpr uint32_t B_HI = ((uint32_t)((KT / 8) * 128) >> 4) | (1u << 14);                  // no swizzle
        const uint32_t blo = ((dsaddr(wsm) >> 4) & 0x3FFFu) | ((uint32_t)(128 >> 4) << 16);
        auto mma = [&](uint32_t tacc, uint32_t alo, uint32_t bl, uint32_t acc) {
            // descriptor base offset (bits 49-51): the phase of the swizzle pattern at a start address that is not atom-aligned
            const uint32_t bo = (ONE && (diag & 16)) ? ((alo >> 3) & 7u) << 17 : 0u;
            dtc_mma(tacc, ((uint64_t)(A_HI | bo) << 32) | alo, ((uint64_t)B_HI << 32) | bl, idesc, acc);
        };
        for (int it = 0; it < ntiles; ++it) {
            const int s = it % NST, k = it / NST, b = it & 1, kb = it >> 1;        // ring slot, its use count; TMEM buffer, its use count
            if (kb >= 1) dbar_wait(&tempty_bar[b], (kb - 1) & 1);
            dbar_wait(&full_bar[s], k & 1);
            dtc_after();
            const uint32_t alo = ((dsaddr(dsm + s * TILE) >> 4) & 0x3FFFu) | (1u << 16);
            if (dtc_elect()) {
                if (!(diag & 4))
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt) {
                    const uint32_t tacc = tmem_base + (uint32_t)(b * ACC + mt * NN);
#pragma unroll
                    for (int pos = 0; pos < KPOS; ++pos) {
                        // forward: tap (kh, kw) = copy kw from staged row kh;  gradient: neighbour (dr, dc) = copy dc from staged row dr
                        const int col = MODE == 0 ? (pos & 3) : pos % 3, row = MODE == 0 ? (pos >> 2) : pos / 3;
                        const int cp = ONE == 2 ? 0 : ONE ? (MODE == 0 ? (col & 1) : 0) : col, shift = ONE == 2 ? col + 1 : ONE ? (MODE == 0 ? (col >> 1) : col) : 0;
                        const uint32_t aoff = cp * COPYB + row * ROWB + mt * 8 * SPAN + shift * PIXB;
#pragma unroll
                        for (int kc = 0; kc < CK / 16; ++kc)
                            mma(tacc, alo + ((aoff + kc * 32) >> 4), blo + (uint32_t)(((pos * NCH + 2 * kc) * 128) >> 4), (pos | kc) != 0 ? 1u : 0u);
                    }
                }
                dtc_commit(&empty_bar[s]);
                dtc_commit(&tfull_bar[b]);
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
                    if (MODE == 0) {
                        o = dst + ((size_t)(n * Hd + th * DT_ROWS + r) * Wd + tw * TCOLS + 8 * mt + c) * CN + cc * 32;
                    } else {
                        const int ph = (cc * 32) / (2 * CN), off = (cc * 32) % (2 * CN);
                        o = dst + ((size_t)(n * Hd + 2 * (th * DT_ROWS + r) + ph) * Wd + 2 * (tw * TCOLS + 8 * mt + c)) * CN + off;
                    }
                    if (!(diag & 2))
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
        // ------------------------------------------------------------------------------------------ TMA producer
        for (int it = 0; it < ntiles; ++it) {
            const int t = t_begin + it, s = it % NST, k = it / NST;
            const int tw = t % tiles_w, r2 = t / tiles_w, th = r2 % tiles_h, n = r2 / tiles_h;
            if (k >= 1) dbar_wait(&empty_bar[s], (k - 1) & 1);
            if (dtc_elect()) {
                const uint32_t d0 = dsaddr(dsm + s * TILE);
                dbar_expect_tx(&full_bar[s], (uint32_t)(NCOPY * IR * ROWB));
#pragma unroll
                for (int cp = 0; cp < NCOPY; ++cp) {
                    // forward: pixels 2 (ow0 + c) - 1 + kw, rows 2 oh0 - 1 ..;  gradient: pixels qw0 + c + dc - 1, rows qh0 - 1 ..
                    const int c1 = ONE == 2 ? tw * TCOLS - 1 : MODE == 0 ? 2 * tw * TCOLS - 1 + cp : tw * TCOLS - 1 + cp;      // ONE == 2: pair index
                    const int c2 = MODE == 0 ? 2 * th * DT_ROWS - 1 : th * DT_ROWS - 1;
                    dtma_4d(&tm, &full_bar[s], d0 + cp * COPYB, 0, c1, c2, n);
                }
            }
            __syncwarp();
        }
    }
    dtc_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TM_COLS) : "memory");
    }
}

}  // namespace

int get_direct_map(const void* ptr, int N, int H, int W, int C, int bw, int bh, int es, int swz, CUtensorMap* out);      // conv_tc.cu
// option "dtc_diag": timing experiments.  Bits 1, 2, 4, 16 give WRONG results: 1 = contiguous TMA boxes instead of every second pixel,
// 2 = the epilogue does not store, 4 = no MMAs are issued, 16 = descriptor base offset on windows that start inside a swizzle atom
// (measured wrong: the tensor core's swizzle is a function of the absolute shared-memory address, like TMA's, so such windows
// need NO base offset -- tools/exp_dtc_onecopy.py).  Bit 8 (correct results) swaps the copy scheme: forward one copy per kw instead
// of one per pixel parity, gradient a single copy instead of one per dc.
int g_dtc_diag = 0;
// option "dtc_wide": 1 (default) = tiles of two M tiles (16 x 16 pixels, one CTA per SM) for the 16 <-> 32 layers wherever the column
// count allows, 0 = always one M tile (16 x 8 pixels, two CTAs per SM: the same staged bytes per pixel and twice the independent
// pipelines per SM -- measured SLOWER, ds2 forward 63.9 vs 58.0 us, data gradient 59.4 vs 50.9: TMA boxes of 256-byte rows)
int g_dtc_wide = 1;

namespace {

template <int MODE, int CK, int CN, int NMT, bool STATS, int ONE = 0>
cudaError_t launch_direct_tc(const void* src, const void* wgt, const float* bias, void* dst, double* stats, int groups, int N, int Hs, int Ws,
                             int act, cudaStream_t st) {
    constexpr size_t smem = direct_tc_smem<MODE, CK, CN, NMT, ONE>();
    static_assert(smem <= 227 * 1024 && direct_tc_slots<MODE, CK, CN, NMT, ONE>() >= 2, "tile ring + weights exceed the SM's shared memory");
    static int per_sm = 0;               // resident CTAs per SM (shared memory decides: 2 with one M tile per tile, else 1)
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(direct_tc_kernel<MODE, CK, CN, NMT, STATS, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int n = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, direct_tc_kernel<MODE, CK, CN, NMT, STATS, ONE>, DT_THREADS, smem) != cudaSuccess || n < 1) n = 1;
        per_sm = n > 2 ? 2 : n;
    }
    CUtensorMap tm;
    if (ONE == 2 ? get_direct_map(src, N, Hs, Ws / 2, 2 * CK, 8 * NMT + 2, 2 * DT_ROWS + 2, 1, CK * 4, &tm)         // pixel pairs
                 : get_direct_map(src, N, Hs, Ws, CK, ONE ? (MODE == 0 ? 8 * NMT + 1 : 8 * NMT + 2) : 8 * NMT, MODE == 0 ? 2 * DT_ROWS + 2 : DT_ROWS + 2,
                                  (MODE == 0 && !(g_dtc_diag & 1)) ? 2 : 1, CK * 2, &tm))
        return cudaErrorInvalidValue;
    const int Hm = MODE == 0 ? Hs / 2 : Hs, Wm = MODE == 0 ? Ws / 2 : Ws;        // the grid the M tiles cover
    const int tiles_w = Wm / (8 * NMT), tiles_h = Hm / DT_ROWS, total = N * tiles_w * tiles_h;
    return launch_pdl(direct_tc_kernel<MODE, CK, CN, NMT, STATS, ONE>, dim3((unsigned)(total < SG_NUM_SMS * per_sm ? total : SG_NUM_SMS * per_sm)), dim3(DT_THREADS),
                      smem, st, tm, (const bf16*)wgt, bias, (bf16*)dst, stats, Hs, Ws, act, groups > 0 ? N / groups : N, tiles_w, tiles_h, total, g_dtc_diag);
}

}  // namespace

// Entry points for narrow_conv.cu (which owns the C ABI of these shapes).  Return cudaErrorInvalidValue for a shape this file
// has no instantiation for.  mode 0: Hs x Ws = input size, (Ci, Co); mode 1: Hs x Ws = dy size, dy has Co channels, dx Ci.
bool direct_tc_supported(int mode, int Ci, int Co, int Hs, int Ws) {
    const int Hm = mode == 0 ? Hs / 2 : Hs, Wm = mode == 0 ? Ws / 2 : Ws;
    if (Hm % DT_ROWS != 0) return false;
    if (mode == 0 && Ci == 16 && Co == 32) return Wm % 8 == 0;
    if (mode == 0 && Ci == 32 && Co == 64) return Wm % 8 == 0;
    if (mode == 1 && Ci == 16 && Co == 32) return Wm % 8 == 0;
    return false;
}
cudaError_t direct_tc_fprop(const void* x, const void* pf, const float* bias, void* y, double* stats, int groups, int N, int H, int W, int Ci,
                            int Co, int act, cudaStream_t st) {
    const bool wide = g_dtc_wide && (W / 2) % 16 == 0;
    if (Ci == 16 && Co == 32 && wide && (g_dtc_diag & 8))        // A/B: one copy per kw (four) instead of one per pixel parity (two)
        return stats ? launch_direct_tc<0, 16, 32, 2, true, 0>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 16, 32, 2, false, 0>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 16 && Co == 32 && wide && (g_dtc_diag & 32))       // A/B: two copies (pixel parities, strided boxes)
        return stats ? launch_direct_tc<0, 16, 32, 2, true, 1>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 16, 32, 2, false, 1>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 16 && Co == 32 && wide)
        return stats ? launch_direct_tc<0, 16, 32, 2, true, 2>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 16, 32, 2, false, 2>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 16 && Co == 32)
        return stats ? launch_direct_tc<0, 16, 32, 1, true, 0>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 16, 32, 1, false, 0>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 32 && Co == 64 && wide && (g_dtc_diag & 32))
        return stats ? launch_direct_tc<0, 32, 64, 2, true, 1>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 32, 64, 2, false, 1>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 32 && Co == 64 && wide && !(g_dtc_diag & 8))
        return stats ? launch_direct_tc<0, 32, 64, 2, true, 2>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 32, 64, 2, false, 2>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    if (Ci == 32 && Co == 64)
        return stats ? launch_direct_tc<0, 32, 64, 1, true, 0>(x, pf, bias, y, stats, groups, N, H, W, act, st)
                     : launch_direct_tc<0, 32, 64, 1, false, 0>(x, pf, bias, y, stats, groups, N, H, W, act, st);
    return cudaErrorInvalidValue;
}
cudaError_t direct_tc_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Hi, int Wi, int Ci, int Co, int act,
                            cudaStream_t st) {
    const bool wide = g_dtc_wide && Wi % 16 == 0;
    // the single-copy variant is SLOWER here (57.4 vs 51.3 us): A/B only
    if (Ci == 16 && Co == 32 && wide && (g_dtc_diag & 8)) return launch_direct_tc<1, 32, 16, 2, false, 1>(dy, pd, bias, dx, nullptr, 1, N, Hi, Wi, act, st);
    if (Ci == 16 && Co == 32 && wide) return launch_direct_tc<1, 32, 16, 2, false, 0>(dy, pd, bias, dx, nullptr, 1, N, Hi, Wi, act, st);
    if (Ci == 16 && Co == 32) return launch_direct_tc<1, 32, 16, 1, false, 0>(dy, pd, bias, dx, nullptr, 1, N, Hi, Wi, act, st);
    return cudaErrorInvalidValue;
}

}  // namespace sg
