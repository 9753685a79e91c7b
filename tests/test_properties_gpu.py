"""Size-independent properties of the conv kernels at the FULL sizes of BASELINE.json's configs (where an fp64 reference
conv would take minutes): the three directions of one operator are mutually adjoint,

    < dy, fprop(x, W) >  ==  < dgrad(dy, W), x >  ==  < wgrad(x, dy), W >,

the fused BatchNorm statistics equal the statistics of the tensor that was stored, and a conv is linear in its input."""
import pytest
import torch

pytestmark = pytest.mark.gpu

FULL = [
    # name, N, H, Ci, Co, k, s, p, groups                    (Conv2d orientation, as tools/bench_conv.py)
    ("stage1 critic ds3, batch 128 x 3 groups", 384, 16, 128, 256, 4, 2, 1, 3),
    ("stage1 critic ds4", 384, 8, 256, 512, 4, 2, 1, 3),
    ("stage2 G2 residual 640->320 3x3, batch 64", 64, 16, 640, 320, 3, 1, 1, 1),
    ("stage2 G2 up1 operator (ConvT 320->160)", 64, 64, 160, 320, 4, 2, 1, 1),
    ("stage2 critic ds3 (32->64 channels)", 192, 64, 32, 64, 4, 2, 1, 3),
]


def _dot(a, b):
    return (a.double().flatten() * b.double().flatten()).sum().item()


@pytest.mark.parametrize("case", FULL, ids=[c[0] for c in FULL])
def test_adjoint_identity_and_stats_at_full_size(case):
    from imagegenerator_b200.ops import CudaOps
    _, N, H, Ci, Co, k, s, p, G = case
    ops = CudaOps("bf16")
    Ho = (H + 2 * p - k) // s + 1
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, H, Ci, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(N, Ho, Ho, Co, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Co, Ci, k, k, device="cuda", generator=g) * (Ci * k * k) ** -0.5).to(torch.bfloat16).float()
    pf, pd = ops.empty((Co, k, k, Ci)), ops.empty((Ci, k, k, Co))
    ops.pack_weight(w, pf, pd)
    y = ops.empty((N, Ho, Ho, Co))
    stats = torch.zeros(G, Co, 2, dtype=torch.float64, device="cuda")
    ops.conv_fprop_stats(x, pf, y, stats, G, k, s, p)
    dx = ops.empty((N, H, H, Ci))
    ops.conv_dgrad(dy, pd, None, dx, k, s, p)
    dw = torch.zeros(Co, Ci, k, k, device="cuda")
    ops.conv_wgrad(x, dy, dw, k, s, p)
    torch.cuda.synchronize()
    a, b, c = _dot(dy, y), _dot(dx, x), _dot(dw, w)
    # y and dx are rounded to bf16 element-wise (relative 2^-9 each, random signs): the inner products of ~1e7..1e8 terms
    # agree far better than the element tolerance; the wgrad accumulates in fp32 and is exact up to summation order
    scale = (dy.double().norm() * y.double().norm()).item()
    assert abs(a - c) <= 2e-4 * scale, (a, c, scale)
    assert abs(b - c) <= 2e-4 * scale, (b, c, scale)
    # fused statistics == statistics of the stored tensor
    v = y.double().reshape(G, -1, Co)
    want = torch.stack([v.sum(1), (v * v).sum(1)], dim=-1)
    assert torch.allclose(stats, want, rtol=1e-4, atol=1e-4 * want.abs().max().item())
    # linearity in the input: conv(2x) == 2 conv(x) exactly (powers of two commute with every rounding)
    y2 = ops.empty((N, Ho, Ho, Co))
    ops.conv_fprop((x.float() * 2).to(torch.bfloat16), pf, None, y2, k, s, p)
    torch.cuda.synchronize()
    assert torch.equal(y2.float(), y.float() * 2)


@pytest.mark.parametrize("rows,C,G", [(64 * 128 * 128, 80, 1), (128 * 16 * 16, 128, 3)], ids=["G2 up2 activation (168 MB)", "critic ds2, 3 groups"])
def test_batchnorm_invariants_at_full_size(rows, C, G):
    """Train-mode BatchNorm: the normalised tensor has zero mean / unit variance per (group, channel), and the backward
    output is orthogonal to 1 and to xhat per (group, channel) -- whatever the incoming gradient."""
    from imagegenerator_b200.ops import CudaOps, ACT_NONE
    ops = CudaOps("bf16")
    g = torch.Generator(device="cuda").manual_seed(1)
    y = (torch.randn(G * rows, C, device="cuda", generator=g) * 1.7 + 0.3).to(torch.bfloat16)
    stats = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
    ops.col_stats(y, stats, G)
    mr = torch.empty(G, C, 2, device="cuda")
    rm, rv, nbt = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    ops.bn_finalize(stats, rows, mr, rm, rv, nbt, 1, True)
    ones, zeros = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    a = torch.empty_like(y)
    ops.bn_act(y, mr, ones, zeros, a, G, ACT_NONE)
    v = a.double().reshape(G, rows, C)
    assert v.mean(1).abs().max().item() < 5e-3 and (v.var(1, unbiased=False) - 1).abs().max().item() < 1e-2
    assert int(nbt) == G
    da = torch.randn(G * rows, C, device="cuda", generator=g).to(torch.bfloat16)
    sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
    dyo = torch.empty_like(y)
    ops.bn_bwd_reduce(da, a, y, mr, sums, G, ACT_NONE)
    ops.bn_bwd_apply(da, a, y, mr, ones, sums, dyo, G, ACT_NONE)
    torch.cuda.synchronize()
    d = dyo.double().reshape(G, rows, C)
    xh = (y.double().reshape(G, rows, C) - mr[:, None, :, 0].double()) * mr[:, None, :, 1].double()
    scale = d.abs().sum(1)                                          # per (group, channel)
    assert (d.sum(1).abs() / scale).max().item() < 2e-3             # bf16 rounding of ~1e6 terms with random signs
    assert ((d * xh).sum(1).abs() / scale).max().item() < 2e-3


THIN_FULL = [
    # name, N, H (image side), Co
    ("stage2 critic ds0, batch 64 x 3 groups (256x256x3 <-> 128x128x16)", 192, 256, 16),
    ("stage2 G2 up3, batch 64 (128x128x80 <-> 256x256x3)", 64, 256, 80),
    ("stage1 critic ds0, batch 128 x 3 groups (64x64x3 <-> 32x32x64)", 384, 64, 64),
]


@pytest.mark.parametrize("case", THIN_FULL, ids=[c[0] for c in THIN_FULL])
def test_thin_kernels_adjoint_and_linear_at_full_size(case):
    """The direct 3-channel kernels (thin_conv.cu) at BASELINE.json's full sizes: Conv2d(3->C) and ConvTranspose2d(C->3) of the
    same weights are mutually adjoint, both agree with the weight gradient taken through the patch matrix, and both are linear."""
    from imagegenerator_b200.ops import CudaOps
    _, N, H, Co = case
    ops = CudaOps("bf16")
    assert ops.lib.sg_conv_thin_supported(0, N, H, H, 3, H // 2, H // 2, Co, 4, 2, 1)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(N, H, H, 3, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(N, H // 2, H // 2, Co, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Co, 3, 4, 4, device="cuda", generator=g) * 48 ** -0.5).to(torch.bfloat16).float()
    pf, pd = ops.empty((Co, 4, 4, 3)), ops.empty((3, 4, 4, Co))
    ops.pack_weight(w, pf, pd)
    y, dx = ops.empty((N, H // 2, H // 2, Co)), ops.empty((N, H, H, 3))
    ops.conv_fprop(x, pf, None, y, 4, 2, 1)
    ops.conv_dgrad(dy, pd, None, dx, 4, 2, 1)
    P = ops.empty((N, H // 2, H // 2, 48))
    ops.patchify(x, P, 4, 2, 1)
    dw = torch.zeros(Co, 48, 1, 1, device="cuda")
    ops.conv_wgrad(P, dy, dw, 1, 1, 0)                                  # PyTorch weight order (ci, kh, kw)
    torch.cuda.synchronize()
    a, b, c = _dot(dy, y), _dot(dx, x), _dot(dw, w.reshape(Co, 48, 1, 1))
    scale = (dy.double().norm() * y.double().norm()).item()
    assert abs(a - c) <= 2e-4 * scale and abs(b - c) <= 2e-4 * scale, (a, b, c, scale)
    y2, dx2 = ops.empty(y.shape), ops.empty(dx.shape)
    ops.conv_fprop((x.float() * 2).to(torch.bfloat16), pf, None, y2, 4, 2, 1)
    ops.conv_dgrad((dy.float() * 2).to(torch.bfloat16), pd, None, dx2, 4, 2, 1)
    torch.cuda.synchronize()
    assert torch.equal(y2.float(), y.float() * 2) and torch.equal(dx2.float(), dx.float() * 2)


def test_fused_bn_backward_statistics_at_full_size():
    """sg_conv_dgrad_bstats on the deepest Stage-II residual conv (640 -> 320, batch 64): the statistics that leave the conv
    epilogue equal what sg_bn_bwd_reduce computes from the stored gradient, and the gradient equals the plain dgrad's."""
    from imagegenerator_b200.ops import CudaOps, ACT_RELU
    ops = CudaOps("bf16")
    N, H, Ci, Co = 64, 16, 320, 640                                     # dgrad: dy [N,16,16,640] -> dx [N,16,16,320]
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = torch.randn(N, H, H, Co, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) * (Co * 9) ** -0.5
    pd = ops.empty((Ci, 3, 3, Co))
    ops.pack_weight(w, None, pd)
    ybn = (torch.randn(N, H, H, Ci, device="cuda", generator=g) * 1.3).to(torch.bfloat16)
    mr = torch.stack([torch.randn(1, Ci, device="cuda", generator=g) * 0.2, torch.rand(1, Ci, device="cuda", generator=g) + 0.5], -1)
    gamma, beta = torch.rand(Ci, device="cuda", generator=g) + 0.5, torch.randn(Ci, device="cuda", generator=g) * 0.3
    dx, dx_ref = ops.empty((N, H, H, Ci)), ops.empty((N, H, H, Ci))
    sums = torch.full((1, Ci, 2), 3.0, dtype=torch.float64, device="cuda")
    want = torch.zeros_like(sums)
    ops.set_option("bstats_min_k", 0)
    ops.conv_dgrad_bstats(dy, pd, dx, ybn, mr, gamma, beta, sums, 1, ACT_RELU, 3, 1, 1)
    ops.set_option("bstats_min_k", 4000)
    ops.conv_dgrad(dy, pd, None, dx_ref, 3, 1, 1)
    ops.bn_bwd_reduce(dx, None, ybn, mr, want, 1, ACT_RELU, gamma=gamma, beta=beta)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_ref)
    scale = dx.double().abs().sum(dim=(0, 1, 2))[None, :, None] * 8.0
    assert ((sums - want).abs() <= 2e-4 * scale + 1e-6).all(), ((sums - want).abs() / scale).max().item()
