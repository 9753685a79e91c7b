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

int sg_conv_fprop_tc(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Ci, int Ho, int Wo,
                     int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_fprop_tc: bf16 only");
    SG_REQUIRE(sg_conv_tc_supported(0, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_fprop_tc: unsupported shape");
    return launch_conv_tc(0, x, pf, bias, y, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, SG_STREAM(stream));
}

int sg_conv_dgrad_tc(const void* dy, const void* pd, const float* bias, void* dx, int N, int H, int W, int Ci, int Ho,
                     int Wo, int Co, int k, int s, int p, int act, int dtype, void* stream) {
    SG_REQUIRE(dtype == SG_BF16, "conv_dgrad_tc: bf16 only");
    SG_REQUIRE(sg_conv_tc_supported(1, N, H, W, Ci, Ho, Wo, Co, k, s, p), "conv_dgrad_tc: unsupported shape");
    SG_REQUIRE(H == (Ho - 1) * s - 2 * p + k && W == (Wo - 1) * s - 2 * p + k, "conv_dgrad_tc: inconsistent sizes");
    return launch_conv_tc(1, dy, pd, bias, dx, N, H, W, Ci, Ho, Wo, Co, k, s, p, act, SG_STREAM(stream));
}

}  // extern "C"
