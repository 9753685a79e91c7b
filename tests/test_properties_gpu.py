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
