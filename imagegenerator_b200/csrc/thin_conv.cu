// thin_conv.cu -- the 3-channel (image-side) convolutions of both critics and generators as direct kernels.
//
// StackGAN touches RGB images in four places: the critics' first Conv2d(3 -> C, k4 s2 p1) (discrminator_1.py:17-18,
// discriminator_2.py:11-12), its input gradient (the gradient penalty differentiates the critic w.r.t. the image,
// utils.py:15-21), the generators' last ConvTranspose2d(C -> 3, k4 s2 p1) + Tanh (generator_1.py:30-33,
// generator_2.py:55-57) and that layer's input gradient.  With N = 3 these are HBM-bound: 16-80 activation channels per
// pixel on one side, 3 on the other.  The tcgen05 path ran them as 1x1 GEMMs over a [pix][48] patch matrix (4x the
// image, written and re-read) resp. onto a [pix][48] fp32 col matrix (201 MB for G2's output layer, written and
// re-read by a col2im pass) -- 3-4x the algorithmic traffic, and 128-row tiles whose epilogue cost more than their
// mainloop.  Here each CTA stages one spatial tile in shared memory, multiplies it with warp-level mma.sync (the
// contraction is tiny: K = 48 or K = C <= 128, so the legacy tensor path is more than enough) and writes the result
// once; every activation byte is read from HBM once (tile halos hit L2).
//
//   conv3_k4s2_kernel   y[n,oh,ow,co]  = act(bias[co] + sum_{kh,kw,ci} x[n,2oh-1+kh,2ow-1+kw,ci] * wp[co][kh][kw][ci])
//   convt3_k4s2_kernel  out[n,oh,ow,c] = act(bias[c] + sum_{kh,kw,ct} x[n,(oh+1-kh)/2,(ow+1-kw)/2,ct] * pd[c][kh][kw][ct])
#include "common.cuh"

namespace sg {

// D (16x8, fp32) += A (16x16, bf16, row-major) * B (16x8, bf16, "col": k contiguous per column)
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// activation of the thin kernels: tanh through the MUFU approximation like the tcgen05 epilogue (2^-11 relative error, below
// the 2^-9 of the bf16 result; tanhf() was ~25 of the 100 instructions a thread spent per output pixel pair)
template <int ACT>
__device__ __forceinline__ float thin_act(float x) {
    if (ACT == SG_ACT_TANH) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    return act_fwd(x, ACT);
}

// ------------------------------------------------------------------------------------------------ Conv2d(3 -> Co)
// CTA = 16 x 32 output pixels of one image; warp w owns output rows 2w, 2w+1 (two 16-pixel m-tiles each).  The 34 x 66-pixel input
// tile sits in shared memory as 34 rows of C3_RS bf16, copied with 8-byte loads: element s of a row is element
// 6*ow0 - 4 + s of the image row (a multiple of 4, so every 4-element chunk is aligned and lies wholly inside or outside
// the row).  The 12 values (kw, ci) a tap row of output pixel p needs are s = 6p + 1 .. 6p + 12; the reduction is laid
// out as K = 4 x 16: k = kh*16 + jj covers s = 6p + jj, with ZERO weights at jj = 0, 13, 14, 15 -- one k16 step per tap
// row, every (k, k+1) pair of the mma A fragment one aligned 32-bit shared-memory load, no per-element address math
// (the first version gathered K = 48 exactly and spent its time on 2-byte copies and index arithmetic: 840 warp
// instructions per warp, 1.7 IPC, 175 us on the Stage-II critic's first layer; ncu in profiles/).  Results are staged
// per warp in shared memory and leave as 16-byte stores: a tile row's 32 pixels x Co channels are contiguous in NHWC.
constexpr int C3_RW = 2, C3_TH = 8 * C3_RW, C3_TW = 32, C3_IR = 2 * C3_TH + 2, C3_RS = 208, C3_CH = 51, C3_WS = 72;

template <int ACT>
__global__ void __launch_bounds__(256)
conv3_k4s2_kernel(const bf16* __restrict__ x, const bf16* __restrict__ wp, const float* __restrict__ bias,
                  bf16* __restrict__ y, int H, int W, int Co, int tiles_w, int tiles_h) {
    extern __shared__ __align__(16) uint8_t thin_smem[];
    bf16* tile = reinterpret_cast<bf16*>(thin_smem);          // [C3_IR][C3_RS]
    bf16* wsm = tile + C3_IR * C3_RS;                          // [Co][C3_WS]
    const int SS = Co + 8;                                     // staging row stride (elements)
    bf16* stage = wsm + Co * C3_WS;                            // [8 warps][32][SS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int n = b / tiles_h;
    const int Ho = H >> 1, Wo = W >> 1;
    const int oh0 = th * C3_TH, ow0 = tw * C3_TW;
    const int ih0 = 2 * oh0 - 1, ge0 = 6 * ow0 - 4, rowlen = W * 3;
    SG_PDL_SYNC();
    {
        unsigned short* w16 = reinterpret_cast<unsigned short*>(wsm);
        const unsigned short* src = reinterpret_cast<const unsigned short*>(wp);
        for (int i = tid; i < Co * 48; i += 256) {
            const int co = i / 48, r = i - co * 48, kh = r / 12, j = r - kh * 12;
            w16[co * C3_WS + kh * 16 + 1 + j] = __ldg(src + i);
        }
        for (int i = tid; i < Co * 16; i += 256) {
            const int co = i >> 4, r = i & 15, kh = r >> 2, d = r & 3;
            w16[co * C3_WS + kh * 16 + (d == 0 ? 0 : 12 + d)] = 0;
        }
    }
    {
        const bf16* xin = x + (size_t)n * H * rowlen;
        for (int i = tid; i < C3_IR * C3_CH; i += 256) {
            const int r = i / C3_CH, c = i - r * C3_CH;
            const int ih = ih0 + r, e = ge0 + 4 * c;
            uint2 v = make_uint2(0u, 0u);
            if (ih >= 0 && ih < H && e >= 0 && e < rowlen) v = __ldg(reinterpret_cast<const uint2*>(xin + (size_t)ih * rowlen + e));
            *reinterpret_cast<uint2*>(tile + r * C3_RS + 4 * c) = v;
        }
    }
    __syncthreads();
    bf16* stw = stage + warp * 32 * SS;
    const int c8 = Co >> 3, total = 32 * c8;
    const int row0 = lane / c8, col0 = lane - row0 * c8, drow = 32 / c8, dcol = 32 - drow * c8;
#pragma unroll 1
    for (int rr = 0; rr < C3_RW; ++rr) {
        const int ohl = warp * C3_RW + rr;
        // A fragments of this row's two m-tiles: k-step kh reads tile row 2*ohl + kh, elements 6p + 2q (+1), 6p + 2q + 8 (+1)
        uint32_t a[2][4][4];
#pragma unroll
        for (int kh = 0; kh < 4; ++kh) {
            const bf16* base = tile + (2 * ohl + kh) * C3_RS + 2 * q;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int p0 = mt * 16 + g;
                a[mt][kh][0] = *reinterpret_cast<const uint32_t*>(base + 6 * p0);
                a[mt][kh][1] = *reinterpret_cast<const uint32_t*>(base + 6 * (p0 + 8));
                a[mt][kh][2] = *reinterpret_cast<const uint32_t*>(base + 6 * p0 + 8);
                a[mt][kh][3] = *reinterpret_cast<const uint32_t*>(base + 6 * (p0 + 8) + 8);
            }
        }
        for (int nt = 0; nt < c8; ++nt) {
            const bf16* wr = wsm + (nt * 8 + g) * C3_WS + 2 * q;
            uint32_t bf[4][2];
#pragma unroll
            for (int kh = 0; kh < 4; ++kh) {
                bf[kh][0] = *reinterpret_cast<const uint32_t*>(wr + kh * 16);
                bf[kh][1] = *reinterpret_cast<const uint32_t*>(wr + kh * 16 + 8);
            }
            float b0 = 0.f, b1 = 0.f;
            if (bias != nullptr) { b0 = __ldg(bias + nt * 8 + 2 * q); b1 = __ldg(bias + nt * 8 + 2 * q + 1); }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float c[4] = {b0, b1, b0, b1};
#pragma unroll
                for (int kh = 0; kh < 4; ++kh) mma16816(c, a[mt][kh][0], a[mt][kh][1], a[mt][kh][2], a[mt][kh][3], bf[kh][0], bf[kh][1]);
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = thin_act<ACT>(c[j]);
                *reinterpret_cast<uint32_t*>(stw + (mt * 16 + g) * SS + nt * 8 + 2 * q) = pack_bf16x2(c[0], c[1]);
                *reinterpret_cast<uint32_t*>(stw + (mt * 16 + g + 8) * SS + nt * 8 + 2 * q) = pack_bf16x2(c[2], c[3]);
            }
        }
        __syncwarp();
        // copy-out: the row's 32 pixels x Co channels are one contiguous run of the NHWC output
        uint4* dst = reinterpret_cast<uint4*>(y + ((size_t)(n * Ho + oh0 + ohl) * Wo + ow0) * Co);
        int row = row0, col = col0;
        for (int j = lane; j < total; j += 32) {
            dst[j] = *reinterpret_cast<const uint4*>(stw + row * SS + col * 8);
            row += drow; col += dcol;
            if (col >= c8) { col -= c8; ++row; }
        }
        __syncwarp();                  // the staging rows are rewritten by the next output row
    }
}

// ------------------------------------------------------------------------------------------------ ConvT(C -> 3)
// CTA = 8 x 32 input pixels (+ a one-pixel halo: 10 x 34 = 340) -> 16 x 64 output pixels of one image.
//   1. the halo'd activation tile [340][C] goes to shared memory (16-byte loads, pixel stride C + 8 elements: conflict-free
//      fragment loads);
//   2. col[pix][tap*3 + c] = x[pix][:] . pd[c][tap][:]   (M = 340 -> 22 m-tiles over 8 warps, N = 48, K = C) with mma.sync,
//      accumulators in registers;
//   3. the fp32 col tile replaces the activation tile in shared memory and every thread sums the 2 x 2 taps that land on
//      its output pixels (k4 s2 p1: output row 2q+ph takes kh = ph+1 from input row q and kh = ph+3 / ph-1 from row
//      q-1 / q+1), adds the bias, applies the activation and stores two horizontally adjacent pixels (12 bytes).
// The col matrix never leaves the SM (it was a 201 MB fp32 round trip through HBM for G2's output layer).
constexpr int CT_QH = 8, CT_QW = 32, CT_HW = CT_QW + 2, CT_PIX = (CT_QH + 2) * CT_HW, CT_MT = (CT_PIX + 15) / 16, CT_CS = 50;

template <int ACT>
__global__ void __launch_bounds__(256, 2)
convt3_k4s2_kernel(const bf16* __restrict__ x, const bf16* __restrict__ pd, const float* __restrict__ bias,
                   bf16* __restrict__ out, int Hi, int Wi, int C, int CP, int lg_tpp, int tiles_w, int tiles_h, int xs_bytes) {
    extern __shared__ __align__(16) uint8_t thin_smem[];
    const int PS = CP + 8;
    bf16* xs = reinterpret_cast<bf16*>(thin_smem);                    // [CT_PIX][PS]   (phase 1-2)
    float* col = reinterpret_cast<float*>(thin_smem);                 // [CT_PIX][CT_CS] (phase 3, same memory)
    bf16* wsm = reinterpret_cast<bf16*>(thin_smem + xs_bytes);        // [48][PS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int n = b / tiles_h;
    const int qh0 = th * CT_QH, qw0 = tw * CT_QW;
    // 2^lg_tpp >= CP/8 threads share a pixel (each one 16-byte chunk of its channels): no division by a run-time channel count
    const int ch = (tid & ((1 << lg_tpp) - 1)) * 8, psub = tid >> lg_tpp, pstep = 256 >> lg_tpp;
    SG_PDL_SYNC();
    if (ch < CP) {
        for (int nn = psub; nn < 48; nn += pstep) {
            const int tap = nn / 3, c = nn - tap * 3;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (ch < C) v = __ldg(reinterpret_cast<const uint4*>(pd + (size_t)(c * 16 + tap) * C + ch));
            *reinterpret_cast<uint4*>(wsm + nn * PS + ch) = v;
        }
        const bf16* xin = x + (size_t)n * Hi * Wi * C + ch;
        for (int p = psub; p < CT_PIX; p += pstep) {
            const int r = p / CT_HW, cc = p - r * CT_HW;
            const int ih = qh0 - 1 + r, iw = qw0 - 1 + cc;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (ch < C && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi)
                v = __ldg(reinterpret_cast<const uint4*>(xin + ((size_t)ih * Wi + iw) * C));
            *reinterpret_cast<uint4*>(xs + p * PS + ch) = v;
        }
    }
    __syncthreads();
    float acc[3][6][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int nt = 0; nt < 6; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][nt][j] = 0.f;
    {
        const bf16* wr0 = wsm + g * PS + 2 * q;
        const bf16* ar[3][2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int mt = warp + 8 * i;
            ar[i][0] = xs + min(mt * 16 + g, CT_PIX - 1) * PS + 2 * q;
            ar[i][1] = xs + min(mt * 16 + g + 8, CT_PIX - 1) * PS + 2 * q;
        }
        for (int cs = 0; cs < (CP >> 4); ++cs) {
            uint32_t bf[6][2];
#pragma unroll
            for (int nt = 0; nt < 6; ++nt) {
                bf[nt][0] = *reinterpret_cast<const uint32_t*>(wr0 + nt * 8 * PS + cs * 16);
                bf[nt][1] = *reinterpret_cast<const uint32_t*>(wr0 + nt * 8 * PS + cs * 16 + 8);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (warp + 8 * i < CT_MT) {
                    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(ar[i][0] + cs * 16);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(ar[i][1] + cs * 16);
                    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(ar[i][0] + cs * 16 + 8);
                    const uint32_t a3 = *reinterpret_cast<const uint32_t*>(ar[i][1] + cs * 16 + 8);
#pragma unroll
                    for (int nt = 0; nt < 6; ++nt) mma16816(acc[i][nt], a0, a1, a2, a3, bf[nt][0], bf[nt][1]);
                }
            }
        }
    }
    __syncthreads();                   // every warp is done reading the activation tile: col may overwrite it
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int mt = warp + 8 * i;
        if (mt < CT_MT) {
            const int p0 = mt * 16 + g, p1 = p0 + 8;
            float* c0 = col + p0 * CT_CS + 2 * q;
            float* c1 = col + p1 * CT_CS + 2 * q;
            if (p0 < CT_PIX) {
#pragma unroll
                for (int nt = 0; nt < 6; ++nt) *reinterpret_cast<float2*>(c0 + nt * 8) = make_float2(acc[i][nt][0], acc[i][nt][1]);
            }
            if (p1 < CT_PIX) {
#pragma unroll
                for (int nt = 0; nt < 6; ++nt) *reinterpret_cast<float2*>(c1 + nt * 8) = make_float2(acc[i][nt][2], acc[i][nt][3]);
            }
        }
    }
    __syncthreads();
    const int Wo = Wi * 2;
    float bz[3] = {0.f, 0.f, 0.f};
    if (bias != nullptr) { bz[0] = __ldg(bias); bz[1] = __ldg(bias + 1); bz[2] = __ldg(bias + 2); }
    // thread = one input column qw of the tile, output rows ohl = tid/32 and tid/32 + 8: both output pixels 2qw, 2qw + 1.
    //   column parity 0: (kw 1, input column qw), (kw 3, qw - 1);   parity 1: (kw 2, qw), (kw 0, qw + 1)
    const int qw = tid & 31;
    bf16* obase = out + ((size_t)n * (2 * Hi) * Wo + (size_t)(2 * qh0) * Wo + 2 * qw0 + 2 * qw) * 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ohl = (tid >> 5) + 8 * i, ph = ohl & 1, qh = ohl >> 1;
        const int kh0 = ph + 1, kh1 = ph ? 0 : 3, dh1 = ph ? 1 : -1;
        const float* rA = col + ((qh + 1) * CT_HW + qw + 1) * CT_CS;       // input row qh, column qw
        const float* rB = rA + dh1 * CT_HW * CT_CS;                        // the other input row of this output row parity
        float s0[3], s1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s0[c] = bz[c] + ((rA[(kh0 * 4 + 1) * 3 + c] + rA[(kh0 * 4 + 3) * 3 + c - CT_CS]) +
                             (rB[(kh1 * 4 + 1) * 3 + c] + rB[(kh1 * 4 + 3) * 3 + c - CT_CS]));
            s1[c] = bz[c] + ((rA[(kh0 * 4 + 2) * 3 + c] + rA[(kh0 * 4 + 0) * 3 + c + CT_CS]) +
                             (rB[(kh1 * 4 + 2) * 3 + c] + rB[(kh1 * 4 + 0) * 3 + c + CT_CS]));
        }
        uint32_t* o = reinterpret_cast<uint32_t*>(obase + (size_t)ohl * Wo * 3);
        o[0] = pack_bf16x2(thin_act<ACT>(s0[0]), thin_act<ACT>(s0[1]));
        o[1] = pack_bf16x2(thin_act<ACT>(s0[2]), thin_act<ACT>(s1[0]));
        o[2] = pack_bf16x2(thin_act<ACT>(s1[1]), thin_act<ACT>(s1[2]));
    }
}

static bool g_thin_attr = false, g_thin_attr3 = false;

}  // namespace sg

using namespace sg;

extern "C" {

// 1 if the direct kernels take this operator direction: mode 0 = Conv2d(3 -> Co) forward, mode 1 = its data gradient
// (ConvTranspose2d(Co -> 3) forward).  Conv2d orientation like every sg_conv_* entry point: x [N,H,W,3], y [N,Ho,Wo,Co].
int sg_conv_thin_supported(int mode, int N, int H, int W, int Ci, int Ho, int Wo, int Co, int k, int s, int p) {
    if (Ci != 3 || k != 4 || s != 2 || p != 1 || Ho * 2 != H || Wo * 2 != W || N < 1) return 0;
    if (Co % 8 != 0 || Co < 8 || Co > 128) return 0;
    if (mode == 0) return (Ho % C3_TH == 0 && Wo % C3_TW == 0) ? 1 : 0;
    return (Ho % CT_QH == 0 && Wo % CT_QW == 0) ? 1 : 0;
}

// y = act(conv(x, W) + bias), x [N,H,W,3] bf16, pf [Co][4][4][3] bf16 (the packed fprop operand), y [N,H/2,W/2,Co] bf16
int sg_conv_thin_fprop(const void* x, const void* pf, const float* bias, void* y, int N, int H, int W, int Co, int act,
                       void* stream) {
    SG_REQUIRE(sg_conv_thin_supported(0, N, H, W, 3, H / 2, W / 2, Co, 4, 2, 1), "conv_thin_fprop: unsupported shape N=%d %dx%d Co=%d", N, H, W, Co);
    const int tiles_w = (W / 2) / C3_TW, tiles_h = (H / 2) / C3_TH;
    const size_t smem = (size_t)C3_IR * C3_RS * 2 + (size_t)Co * C3_WS * 2 + (size_t)8 * 32 * (Co + 8) * 2;
    if (!g_thin_attr3) {
        cudaError_t ce = cudaFuncSetAttribute(conv3_k4s2_kernel<SG_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv3_k4s2_kernel<SG_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv3_k4s2_kernel<SG_ACT_LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(conv3_k4s2_kernel<SG_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(conv3): %s", cudaGetErrorString(ce)); return (int)ce; }
        g_thin_attr3 = true;
    }
    const dim3 grid((unsigned)(N * tiles_w * tiles_h));
    cudaError_t ce;
#define SG_C3(A) launch_pdl(conv3_k4s2_kernel<A>, grid, dim3(256), smem, SG_STREAM(stream), (const bf16*)x, (const bf16*)pf, bias, \
                            (bf16*)y, H, W, Co, tiles_w, tiles_h)
    switch (act) {
        case SG_ACT_NONE: ce = SG_C3(SG_ACT_NONE); break;
        case SG_ACT_RELU: ce = SG_C3(SG_ACT_RELU); break;
        case SG_ACT_LRELU: ce = SG_C3(SG_ACT_LRELU); break;
        case SG_ACT_TANH: ce = SG_C3(SG_ACT_TANH); break;
        default: set_error("conv_thin_fprop: bad activation %d", act); return SG_ERR_BAD_ARG;
    }
#undef SG_C3
    if (ce != cudaSuccess) { set_error("conv_thin_fprop launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    SG_LAUNCHED("conv_thin_fprop");
    return 0;
}

// dx = act(convT(dy, W) + bias): dy [N,Ho,Wo,Co] bf16, pd [3][4][4][Co] bf16 (the packed dgrad operand), dx [N,2Ho,2Wo,3]
int sg_conv_thin_dgrad(const void* dy, const void* pd, const float* bias, void* dx, int N, int Ho, int Wo, int Co, int act,
                       void* stream) {
    SG_REQUIRE(sg_conv_thin_supported(1, N, 2 * Ho, 2 * Wo, 3, Ho, Wo, Co, 4, 2, 1), "conv_thin_dgrad: unsupported shape N=%d %dx%d Co=%d", N, Ho, Wo, Co);
    const int CP = (Co + 15) / 16 * 16, PS = CP + 8;
    size_t xs_bytes = (size_t)CT_PIX * PS * 2;
    const size_t col_bytes = (size_t)CT_PIX * CT_CS * 4;
    if (xs_bytes < col_bytes) xs_bytes = col_bytes;
    xs_bytes = (xs_bytes + 15) / 16 * 16;
    const size_t smem = xs_bytes + (size_t)48 * PS * 2;
    if (!g_thin_attr) {
        cudaError_t ce = cudaFuncSetAttribute(convt3_k4s2_kernel<SG_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(convt3_k4s2_kernel<SG_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (ce != cudaSuccess) { set_error("cudaFuncSetAttribute(convt3): %s", cudaGetErrorString(ce)); return (int)ce; }
        g_thin_attr = true;
    }
    SG_REQUIRE(smem <= 112 * 1024, "conv_thin_dgrad: tile does not fit shared memory");
    SG_REQUIRE(act == SG_ACT_NONE || act == SG_ACT_TANH, "conv_thin_dgrad: activation must be none or tanh");
    const int tiles_w = Wo / CT_QW, tiles_h = Ho / CT_QH;
    int lg_tpp = 0;
    while ((1 << lg_tpp) < CP / 8) ++lg_tpp;
    const dim3 grid((unsigned)(N * tiles_w * tiles_h));
    cudaError_t ce = act == SG_ACT_TANH
        ? launch_pdl(convt3_k4s2_kernel<SG_ACT_TANH>, grid, dim3(256), smem, SG_STREAM(stream), (const bf16*)dy, (const bf16*)pd, bias,
                     (bf16*)dx, Ho, Wo, Co, CP, lg_tpp, tiles_w, tiles_h, (int)xs_bytes)
        : launch_pdl(convt3_k4s2_kernel<SG_ACT_NONE>, grid, dim3(256), smem, SG_STREAM(stream), (const bf16*)dy, (const bf16*)pd, bias,
                     (bf16*)dx, Ho, Wo, Co, CP, lg_tpp, tiles_w, tiles_h, (int)xs_bytes);
    if (ce != cudaSuccess) { set_error("conv_thin_dgrad launch: %s", cudaGetErrorString(ce)); return (int)ce; }
    SG_LAUNCHED("conv_thin_dgrad");
    return 0;
}

}  // extern "C"
