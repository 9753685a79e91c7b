"""Every C-ABI kernel against the torch statement of the same name in tests/emu_ops.py (fp64),
in both storage modes.  Runs on the B200 box: ``pytest -m gpu``."""
import pytest
import torch

from emu_ops import EmuOps, ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH

pytestmark = pytest.mark.gpu

MODES = ["fp32", "bf16"]
TOL = {"fp32": dict(rtol=1e-4, atol=1e-5), "bf16": dict(rtol=2e-2, atol=1e-3)}


class T:      # activation-typed tensor (storage type of the mode)
    def __init__(self, t): self.t = t
class F:      # fp32 tensor
    def __init__(self, t): self.t = t
class D:      # fp64 tensor
    def __init__(self, t): self.t = t
class I64:
    def __init__(self, t): self.t = t


def _ops(mode):
    from imagegenerator_b200.ops import CudaOps
    return CudaOps(mode)


def run_pair(mode, name, args, outs, kwargs=None, scale_atol=True, tol=None):
    """Call ops.<name>(*args) on the emulator (fp64, CPU) and on CUDA; compare tensors at positions ``outs``."""
    kwargs = kwargs or {}
    ops = _ops(mode)
    emu = EmuOps(torch.float64)
    sd = ops.act_dtype

    def conv(a, side):
        if isinstance(a, T):
            q = a.t.to(sd)                       # both sides see the storage-rounded values
            return q.double().clone() if side == "emu" else q.cuda().contiguous()
        if isinstance(a, F):
            return a.t.float().double().clone() if side == "emu" else a.t.float().cuda().contiguous()
        if isinstance(a, D):
            return a.t.double().clone() if side == "emu" else a.t.double().cuda().contiguous()
        if isinstance(a, I64):
            return a.t.clone() if side == "emu" else a.t.cuda()
        return a

    ea = [conv(a, "emu") for a in args]
    ca = [conv(a, "cuda") for a in args]
    ek = {k: conv(v, "emu") for k, v in kwargs.items()}
    ck = {k: conv(v, "cuda") for k, v in kwargs.items()}
    getattr(emu, name)(*ea, **ek)
    getattr(ops, name)(*ca, **ck)
    torch.cuda.synchronize()
    t = dict(TOL[mode])
    if tol:
        t.update(tol)
    for i in outs:
        ref, got = ea[i].double(), ca[i].double().cpu()
        atol = t["atol"] * (max(ref.abs().max().item(), 1e-30) if scale_atol else 1.0)
        if not torch.allclose(got, ref, rtol=t["rtol"], atol=atol):
            err = (got - ref).abs().max().item()
            raise AssertionError(f"{name}[{mode}] arg {i}: max abs err {err:.3e}, ref max {ref.abs().max().item():.3e}")
    return ea, ca


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g) * scale


CONV_CASES = [
    # N, H, Ci, Co, k, s, p
    (2, 64, 3, 64, 4, 2, 1),        # critic ds0
    (2, 32, 64, 128, 4, 2, 1),      # critic ds2
    (3, 8, 256, 512, 4, 2, 1),      # critic ds4
    (2, 16, 48, 96, 4, 2, 1),       # G1 up2 operator (Co=convT in, Ci=convT out)
    (2, 64, 3, 24, 4, 2, 1),        # G1 up4 operator
    (5, 4, 192, 228, 4, 1, 0),      # G1 up0 operator (1x1 <-> 4x4)
    (2, 16, 40, 24, 3, 1, 1),       # residual-block style 3x3
    (1, 256, 3, 16, 4, 2, 1),       # stage-II critic ds0
]


def _impl_for(mode, direction, case):
    """bf16 mode has ONE backend: shapes the tcgen05 kernels cannot take are an error at the dispatcher (checked by
    test_bf16_dispatch_refuses_unsupported_shapes), and the CUDA-core kernels are only reachable by their own entry points --
    which is how their bf16-storage instantiation stays covered here."""
    if mode != "bf16":
        return ""
    from imagegenerator_b200.ops import CudaOps
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    lib = CudaOps("bf16").lib
    ok = (lib.sg_conv_wgrad_tc_supported(N, H, H, Ci, Ho, Ho, Co, k, s, p) if direction == 2
          else (lib.sg_conv_tc_supported(direction, N, H, H, Ci, Ho, Ho, Co, k, s, p) or
                lib.sg_conv_thin_supported(direction, N, H, H, Ci, Ho, Ho, Co, k, s, p)))
    return "" if ok else "_ffma"


# the 3-channel image side: direct kernels of thin_conv.cu behind sg_conv_fprop / sg_conv_dgrad (bf16 mode)
THIN_CASES = [
    # N, H (image side), Co
    (3, 64, 64),        # Stage-I critic ds0 / its input gradient
    (2, 128, 16),       # Stage-II critic ds0 (4 x 2 tiles per image)
    (1, 256, 16),       # ... at full size
    (2, 64, 24),        # G1 output layer (channels padded 24 -> 32 inside the transposed kernel)
    (1, 128, 80),       # G2 output layer
    (2, 64, 128),       # G2 ds0
    (1, 64, 8),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", THIN_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_LRELU, True), (ACT_TANH, True)])
def test_thin_conv_fprop_direct(case, act, use_bias):
    N, H, Co = case
    from imagegenerator_b200.ops import CudaOps
    assert CudaOps("bf16").lib.sg_conv_thin_supported(0, N, H, H, 3, H // 2, H // 2, Co, 4, 2, 1)
    x, w = rnd(N, H, H, 3), rnd(Co, 3, 4, 4, scale=48 ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    bias = F(rnd(Co)) if use_bias else None
    run_pair("bf16", "conv_fprop", [T(x), T(pf), bias, T(torch.zeros(N, H // 2, H // 2, Co)), 4, 2, 1], [3], dict(act=act))


@pytest.mark.gpu
@pytest.mark.parametrize("case", THIN_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_TANH, True)])
def test_thin_conv_dgrad_direct(case, act, use_bias):
    N, H, Co = case
    from imagegenerator_b200.ops import CudaOps
    assert CudaOps("bf16").lib.sg_conv_thin_supported(1, N, H, H, 3, H // 2, H // 2, Co, 4, 2, 1)
    dy, w = rnd(N, H // 2, H // 2, Co), rnd(Co, 3, 4, 4, scale=(Co * 4) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    bias = F(rnd(3)) if use_bias else None
    run_pair("bf16", "conv_dgrad", [T(dy), T(pd), bias, T(torch.zeros(N, H, H, 3)), 4, 2, 1], [3], dict(act=act))


# 16 / 32 input channels on large maps: the direct kernels of narrow_conv.cu, by name (persistent CTAs: a CTA walks several tiles
# as soon as there are more tiles than SMs) and through the dispatcher's routing mask
NARROW_CASES = [
    # N, H (input side), Ci, Co, groups
    (3, 64, 16, 32, 3),       # one 8 x 32 tile column, four tile rows: 12 tiles, one per CTA
    (2, 128, 16, 32, 1),      # Stage-II critic ds1 at half size (2 x 8 tiles per image)
    (1, 256, 16, 32, 1),      # ... at full size
    (24, 128, 16, 32, 3),     # 384 tiles on 148 CTAs: 2-3 tiles per CTA, ranges crossing image and group boundaries
    (6, 64, 32, 64, 3),       # Stage-II critic ds2 (64 x 64 -> 32 x 32)
    (2, 128, 32, 64, 2),
    (45, 64, 32, 64, 3),      # 180 tiles: one or two per CTA
]


class _narrow_impl:
    """option "narrow_cfg" bit 4: the mma.sync kernels of narrow_conv.cu instead of the tcgen05 ones of direct_tc.cu"""

    def __init__(self, impl):
        self.v = 4 if impl == "mma_sync" else 0

    def __enter__(self):
        _ops("bf16").set_option("narrow_cfg", self.v)

    def __exit__(self, *a):
        _ops("bf16").set_option("narrow_cfg", 0)


NARROW_IMPLS = ["tcgen05", "mma_sync"]


@pytest.mark.gpu
@pytest.mark.parametrize("impl", NARROW_IMPLS)
@pytest.mark.parametrize("case", NARROW_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_LRELU, True)])
def test_narrow_conv_fprop_direct(case, act, use_bias, impl):
    N, H, Ci, Co, G = case
    x, w = rnd(N, H, H, Ci), rnd(Co, Ci, 4, 4, scale=(16 * Ci) ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    bias = F(rnd(Co)) if use_bias else None
    with _narrow_impl(impl):
        run_pair("bf16", "conv_narrow_fprop", [T(x), T(pf), bias, T(torch.zeros(N, H // 2, H // 2, Co))], [3], dict(act=act))


@pytest.mark.gpu
@pytest.mark.parametrize("impl", NARROW_IMPLS)
@pytest.mark.parametrize("case", NARROW_CASES)
def test_narrow_conv_fprop_stats(case, impl):
    N, H, Ci, Co, G = case
    x, w = rnd(N, H, H, Ci), rnd(Co, Ci, 4, 4, scale=(16 * Ci) ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    st0 = torch.ones(G, Co, 2, dtype=torch.float64) * 0.25
    with _narrow_impl(impl):
        ea, ca = run_pair("bf16", "conv_narrow_fprop", [T(x), T(pf), None, T(torch.zeros(N, H // 2, H // 2, Co)), ACT_NONE, D(st0), G], [3])
    _check_stats(ca[3], ca[5], st0, G)


@pytest.mark.gpu
@pytest.mark.parametrize("impl", NARROW_IMPLS)
@pytest.mark.parametrize("case", NARROW_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_LRELU, True)])
def test_narrow_conv_dgrad_direct(case, act, use_bias, impl):
    N, H, Ci, Co, G = case
    dy, w = rnd(N, H // 2, H // 2, Co), rnd(Co, Ci, 4, 4, scale=(Co * 4) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    bias = F(rnd(Ci)) if use_bias else None
    with _narrow_impl(impl):
        run_pair("bf16", "conv_narrow_dgrad", [T(dy), T(pd), bias, T(torch.zeros(N, H, H, Ci))], [3], dict(act=act))


@pytest.mark.gpu
def test_narrow_one_m_tile_variant():
    """Option "dtc_wide" = 0: the tcgen05 kernels with one M tile per tile (16 x 8 pixels, two CTAs per SM) give the same results."""
    ops = _ops("bf16")
    N, H, Ci, Co = 5, 128, 16, 32
    x = rnd(N, H, H, Ci).to(torch.bfloat16).cuda()
    pf = rnd(Co, 4, 4, Ci, scale=(16 * Ci) ** -0.5).to(torch.bfloat16).cuda()
    dy = rnd(N, H // 2, H // 2, Co).to(torch.bfloat16).cuda()
    pd = rnd(Ci, 4, 4, Co, scale=(4 * Co) ** -0.5).to(torch.bfloat16).cuda()
    st = [torch.zeros(1, Co, 2, dtype=torch.float64, device="cuda") for _ in range(2)]
    y = [ops.empty((N, H // 2, H // 2, Co)) for _ in range(2)]
    dx = [ops.empty((N, H, H, Ci)) for _ in range(2)]
    for i, wide in enumerate((1, 0)):
        ops.set_option("dtc_wide", wide)
        try:
            ops.conv_narrow_fprop(x, pf, None, y[i], ACT_LRELU, st[i], 1)
            ops.conv_narrow_dgrad(dy, pd, None, dx[i], ACT_NONE)
        finally:
            ops.set_option("dtc_wide", 1)
    torch.cuda.synchronize()
    assert torch.equal(y[0], y[1]) and torch.equal(dx[0], dx[1])
    assert torch.allclose(st[0], st[1], rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("mask", [0, 15])
def test_narrow_routing_mask(mask):
    """Option "narrow" (bit mask) decides which supported shapes the dispatcher sends to narrow_conv.cu; both backends agree."""
    ops = _ops("bf16")
    N, H, Ci, Co = 2, 128, 16, 32
    dims = (N, H, H, Ci, H // 2, H // 2, Co, 4, 2, 1)
    x = (rnd(N, H, H, Ci)).to(torch.bfloat16).cuda()
    pf = rnd(Co, 4, 4, Ci, scale=(16 * Ci) ** -0.5).to(torch.bfloat16).cuda()
    ya, yb = ops.empty((N, H // 2, H // 2, Co)), ops.empty((N, H // 2, H // 2, Co))
    ops.conv_narrow_fprop(x, pf, None, ya)
    default = [m for m in range(16) if all(bool(ops.lib.sg_conv_narrow_routed(md, N, H, H, ci, H // 2, H // 2, 2 * ci, 4, 2, 1)) == bool(m & ((2 if md else 1) << (2 if ci == 32 else 0)))
                                           for md in (0, 1) for ci in (16, 32))][0]
    ops.set_option("narrow", mask)
    try:
        assert ops.lib.sg_conv_narrow_supported(0, *dims) == 1
        assert ops.lib.sg_conv_narrow_routed(0, *dims) == (1 if mask else 0)
        assert ops.lib.sg_conv_narrow_routed(1, *dims) == (1 if mask else 0)
        n0 = ops.launch_count()
        ops.conv_fprop(x, pf, None, yb, 4, 2, 1)
        assert ops.launch_count() == n0 + 1
    finally:
        ops.set_option("narrow", default)
    torch.cuda.synchronize()
    assert torch.allclose(ya.float(), yb.float(), rtol=2e-2, atol=2e-2)


def test_bf16_dispatch_refuses_unsupported_shapes():
    from imagegenerator_b200.ops import CudaOps
    ops = CudaOps("bf16")
    x = torch.zeros(2, 8, 8, 3, device="cuda", dtype=torch.bfloat16)          # 3 input channels: not a TMA shape
    pf = torch.zeros(16, 4, 4, 3, device="cuda", dtype=torch.bfloat16)
    y = torch.zeros(2, 4, 4, 16, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CUDA-core fallback in bf16 mode"):
        ops.conv_fprop(x, pf, None, y, 4, 2, 1)
    with pytest.raises(RuntimeError, match="no CUDA-core fallback in bf16 mode"):
        ops.conv_wgrad(x, y, torch.zeros(16, 3, 4, 4, device="cuda"), 4, 2, 1)
    pd = torch.zeros(3, 4, 4, 16, device="cuda", dtype=torch.bfloat16)
    dy228 = torch.zeros(2, 1, 1, 228, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CUDA-core fallback in bf16 mode"):
        ops.conv_dgrad(dy228, torch.zeros(192, 4, 4, 228, device="cuda", dtype=torch.bfloat16), None,
                       torch.zeros(2, 4, 4, 192, device="cuda", dtype=torch.bfloat16), 4, 1, 0)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_LRELU, True), (ACT_TANH, True)])
def test_conv_fprop(mode, case, act, use_bias):
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    x, w = rnd(N, H, H, Ci), rnd(Co, Ci, k, k, scale=(Ci * k * k) ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    bias = F(rnd(Co)) if use_bias else None
    run_pair(mode, "conv_fprop", [T(x), T(pf), bias, T(torch.zeros(N, Ho, Ho, Co)), k, s, p], [3],
             dict(act=act, impl=_impl_for(mode, 0, case)))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_TANH, True)])
def test_conv_dgrad(mode, case, act, use_bias):
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    dy, w = rnd(N, Ho, Ho, Co), rnd(Co, Ci, k, k, scale=(Co * k * k / (s * s)) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    bias = F(rnd(Ci)) if use_bias else None
    run_pair(mode, "conv_dgrad", [T(dy), T(pd), bias, T(torch.zeros(N, H, H, Ci)), k, s, p], [3],
             dict(act=act, impl=_impl_for(mode, 1, case)))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_wgrad(mode, case):
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    x, dy = rnd(N, H, H, Ci), rnd(N, Ho, Ho, Co, scale=(N * Ho * Ho) ** -0.5)
    dw0 = rnd(Co, Ci, k, k, scale=0.1)          # accumulate semantics
    run_pair(mode, "conv_wgrad", [T(x), T(dy), F(dw0), k, s, p], [2], dict(impl=_impl_for(mode, 2, case)),
             tol=dict(rtol=2e-3) if mode == "fp32" else None)


@pytest.mark.parametrize("mode", MODES)
def test_layout_and_pack(mode):
    x = rnd(3, 5, 8, 12)
    run_pair(mode, "nchw_to_nhwc", [F(x), T(torch.zeros(3, 8, 12, 5))], [1])
    run_pair(mode, "nhwc_to_nchw", [T(rnd(3, 8, 12, 5)), F(torch.zeros(3, 5, 8, 12))], [1])
    w = rnd(24, 10, 4, 4)
    run_pair(mode, "pack_weight", [F(w), T(torch.zeros(24, 4, 4, 10)), T(torch.zeros(10, 4, 4, 24))], [1, 2])
    for Co, Ci, k in ((512, 256, 4), (320, 640, 3), (64, 48, 1), (37, 70, 3)):      # both store directions, every tap count
        w = rnd(Co, Ci, k, k).bfloat16().float()                                     # representable: the pack is exact
        run_pair(mode, "pack_weight", [F(w), T(torch.zeros(Co, k, k, Ci)), T(torch.zeros(Ci, k, k, Co))], [1, 2],
                 tol=dict(rtol=0, atol=0))
        run_pair(mode, "pack_weight", [F(w), T(torch.zeros(Co, k, k, Ci)), None], [1], tol=dict(rtol=0, atol=0))
        run_pair(mode, "pack_weight", [F(w), None, T(torch.zeros(Ci, k, k, Co))], [2], tol=dict(rtol=0, atol=0))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,rows,G", [(3, 4096, 1), (24, 1024, 2), (128, 512, 3), (512, 64, 3), (640, 256, 1), (1280, 96, 3)])
def test_bn_chain(mode, C, rows, G):
    y = rnd(G * rows, C) * 2 + 0.5
    stats = torch.zeros(G, C, 2, dtype=torch.float64)
    ea, ca = run_pair(mode, "col_stats", [T(y.reshape(G, rows, 1, C)), D(stats), G], [1], tol=dict(rtol=1e-5, atol=1e-6))
    stats = ea[1]
    rm, rv, nbt = rnd(C) * 0.1, torch.rand(C) + 0.5, torch.tensor(3)
    mr = torch.zeros(G, C, 2)
    ea, ca = run_pair(mode, "bn_finalize", [D(stats), rows, F(mr), F(rm), F(rv), I64(nbt), 2, True], [2, 3, 4],
                      tol=dict(rtol=1e-4, atol=1e-5))
    assert int(ca[5]) == 3 + G + 1
    mr = ea[2].float()
    gamma, beta = rnd(C) * 0.5 + 1, rnd(C) * 0.2
    y4 = y.reshape(G * rows, 1, 1, C)
    res = rnd(G * rows, 1, 1, C)
    for act, r in ((ACT_LRELU, None), (ACT_RELU, res), (ACT_NONE, None)):
        out = torch.zeros_like(y4)
        ea, _ = run_pair(mode, "bn_act", [T(y4), F(mr), F(gamma), F(beta), T(out), G, act], [4],
                         dict(residual=T(r)) if r is not None else None)
        # the same in one launch, from the raw sums (C = 3 takes the two-launch fallback inside the library)
        ea2, ca2 = run_pair(mode, "bn_finalize_act", [D(stats), rows, F(torch.zeros(G, C, 2)), F(rm), F(rv), I64(nbt), 2, T(y4),
                                                      F(gamma), F(beta), T(torch.zeros_like(y4)), act], [2, 3, 4, 10],
                            dict(residual=T(r)) if r is not None else None, tol=dict(rtol=1e-4, atol=1e-5)
                            if mode == "fp32" else None)
        assert int(ca2[5]) == 3 + G + 1
    # backward pieces use a consistent a_out
    emu = EmuOps(torch.float64)
    a_out = torch.zeros_like(y4).double()
    sd = torch.bfloat16 if mode == "bf16" else torch.float32
    emu.bn_act(y4.to(sd).double(), mr.double(), gamma.double(), beta.double(), a_out, G, ACT_LRELU)
    a_out = a_out.float()
    da = rnd(G * rows, 1, 1, C, seed=5)
    sums = torch.zeros(G, C, 2, dtype=torch.float64)
    ea, _ = run_pair(mode, "bn_bwd_reduce", [T(da), T(a_out), T(y4), F(mr), D(sums), G, ACT_LRELU], [4],
                     tol=dict(rtol=1e-3, atol=1e-4))
    sums = ea[4]
    inj = rnd(rows, 1, 1, C, seed=9)
    run_pair(mode, "bn_bwd_apply", [T(da), T(a_out), T(y4), F(mr), F(gamma), D(sums), T(torch.zeros_like(y4)), G, ACT_LRELU],
             [6], dict(inject=T(inj), inject_group=G - 1))
    run_pair(mode, "bn_param_grad", [D(sums), F(rnd(C)), F(rnd(C))], [1, 2], tol=dict(rtol=1e-4, atol=1e-5))
    run_pair(mode, "act_bwd", [T(da), T(a_out), T(torch.zeros_like(y4)), ACT_LRELU], [2])
    run_pair(mode, "act_bwd", [T(da), T(torch.tanh(y4)), T(torch.zeros_like(y4)), ACT_TANH], [2])
    run_pair(mode, "colsum", [T(y4), F(rnd(C))], [1], tol=dict(rtol=1e-3, atol=1e-4))
    if G == 1:
        v = rnd(rows, 1, 1, C, seed=11)
        ts = torch.zeros(C, 3, dtype=torch.float64)
        ea, _ = run_pair(mode, "gp_bn_reduce", [T(v), T(da), T(a_out), T(y4), F(mr), D(ts), ACT_LRELU], [5],
                         tol=dict(rtol=1e-3, atol=1e-4))
        ts = ea[5]
        run_pair(mode, "gp_bn_apply", [T(v), T(da), T(a_out), T(y4), F(mr), F(gamma), D(sums), D(ts),
                                       T(torch.zeros_like(y4)), T(torch.zeros_like(y4)), F(rnd(C)), ACT_LRELU], [8, 9, 10],
                 tol=dict(rtol=3e-2, atol=2e-3) if mode == "bf16" else dict(rtol=1e-3, atol=1e-4))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,rows,G,act", [(24, 1024, 1, ACT_RELU), (128, 512, 3, ACT_LRELU), (640, 256, 1, ACT_RELU), (80, 4096, 1, ACT_NONE)])
def test_bn_backward_without_activation_tensor(mode, C, rows, G, act):
    """sg_bn_bwd_reduce_y / sg_bn_bwd_apply_y: the activation's sign recomputed from y == the stored activation's."""
    y = rnd(G * rows, C)
    mr = torch.stack([rnd(G, C, scale=0.1), torch.rand(G, C) + 0.5], dim=-1)
    gamma, beta = torch.rand(C) + 0.5, rnd(C, scale=0.3)
    da = rnd(G * rows, C, seed=3)
    inj = rnd(rows, C, seed=5)
    sums0 = torch.zeros(G, C, 2, dtype=torch.float64)
    ea, ca = run_pair(mode, "bn_bwd_reduce", [T(da), None, T(y), F(mr), D(sums0), G, act], [4],
                      dict(gamma=F(gamma), beta=F(beta)), tol=dict(rtol=1e-3, atol=1e-4))
    run_pair(mode, "bn_bwd_apply", [T(da), None, T(y), F(mr), F(gamma), D(ea[4]), T(torch.zeros(G * rows, C)), G, act], [6],
             dict(inject=T(inj), inject_group=G - 1, beta=F(beta)))


BN_BWD_CASES = [
    # C, rows per group, groups, act, stored activation tensor?, launches expected (1 = parked in shared memory, 2 = two kernels)
    (24, 1024, 1, ACT_RELU, False, 1), (128, 512, 3, ACT_LRELU, False, 1), (640, 256, 1, ACT_RELU, False, 1),
    (80, 4096, 1, ACT_NONE, False, 1), (64, 2048, 3, ACT_LRELU, True, 1), (512, 37, 3, ACT_LRELU, False, 1),
    (128, 100000, 1, ACT_LRELU, False, 1),        # 25.6 MB in bf16: only part of every range is parked
    (128, 33000, 3, ACT_LRELU, False, 1),         # the Stage-I critic's largest BatchNorm layer, ragged group size
    (80, 400000, 1, ACT_RELU, False, 2),          # 64 MB: too large to profit, reduce + apply
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,rows,G,act,with_a,launches", BN_BWD_CASES)
def test_bn_backward_one_call(mode, C, rows, G, act, with_a, launches):
    """sg_bn_bwd (sums + rendezvous + apply in one launch when the tensor fits the SMs' shared memory) == reduce, then apply;
    the work words re-arm themselves: the same call site repeated gives the same answer."""
    y = rnd(G * rows, C)
    mr = torch.stack([rnd(G, C, scale=0.1), torch.rand(G, C) + 0.5], dim=-1)
    gamma, beta = torch.rand(C) + 0.5, rnd(C, scale=0.3)
    da = rnd(G * rows, C, seed=3)
    inj = rnd(rows, C, seed=5)
    emu, ops = EmuOps(torch.float64), _ops(mode)
    sd = ops.act_dtype
    q = lambda t: t.to(sd)
    a_out = None
    if with_a:
        a_out = torch.zeros(G * rows, C, dtype=torch.float64)
        emu.bn_act(q(y).double(), mr.double(), gamma.double(), beta.double(), a_out, G, act)
        a_out = q(a_out.float())
    kw = dict(inject_group=G - 1) if with_a else dict(inject_group=G - 1)
    want_sums, want = torch.zeros(G, C, 2, dtype=torch.float64), torch.zeros(G * rows, C, dtype=torch.float64)
    emu.bn_bwd(q(da).double(), None if a_out is None else a_out.double(), q(y).double(), mr.double(), gamma.double(), want_sums,
               want, G, act, inject=q(inj).double(), beta=None if with_a else beta.double(), **kw)
    c = lambda t: t.cuda().contiguous()
    cda, cy, cmr, cg, cb, cinj = c(q(da)), c(q(y)), c(mr.float()), c(gamma), c(beta), c(q(inj))
    ca = None if a_out is None else c(a_out)
    sums = torch.full((G, C, 2), 7.0, dtype=torch.float64, device="cuda")       # the call zeroes them itself
    dy = torch.zeros(G * rows, C, dtype=sd, device="cuda")
    first = None
    try:
        ops.set_option("bn_fused", 1)                  # the one-launch kernel (an option: off by default)
        for rep in range(4):
            if rep == 3:
                ops.set_option("bn_fused", 0)          # the default: memset + reduce + apply behind the same call
            n0 = ops.launch_count()
            ops.bn_bwd(cda, ca, cy, cmr, cg, sums, dy, G, act, inject=cinj, beta=None if with_a else cb, **kw)
            torch.cuda.synchronize()
            if rep == 3:
                assert ops.launch_count() - n0 == 2
            seen = ops.launch_count() - n0
            if rep < 3:
                assert seen == launches if mode == "bf16" else seen in (1, 2)  # fp32 storage: twice the bytes, the big cases do not fit
            t = TOL[mode]
            got_s, got = sums.cpu(), dy.double().cpu()
            # (per-thread fp32 partial sums: the 32 M-element case sits at ~1e-4 of the largest sum, run to run)
            assert torch.allclose(got_s, want_sums, rtol=1e-3, atol=1e-3 * want_sums.abs().max().item())
            assert torch.allclose(got, want, rtol=t["rtol"], atol=t["atol"] * want.abs().max().item()), (got - want).abs().max().item()
            if first is None:
                first = dy.clone()
            else:
                # (the order of the fp64 atomics differs run to run: a result on a bf16 rounding boundary may move by one ulp)
                assert torch.allclose(dy.float(), first.float(), rtol=t["rtol"], atol=t["atol"] * want.abs().max().item())
    finally:
        ops.set_option("bn_fused", 0)


def test_bn_backward_one_call_under_contention():
    """The rendezvous counts finished ranges, not CTAs: with a side stream hogging the SMs (some CTAs of the launch start
    late and find their range taken over) the result is unchanged and nothing hangs."""
    ops = _ops("bf16")
    C, rows, G = 256, 128 * 64, 3
    y, da = rnd(G * rows, C).cuda().bfloat16(), rnd(G * rows, C, seed=3).cuda().bfloat16()
    mr = torch.stack([rnd(G, C, scale=0.1), torch.rand(G, C) + 0.5], dim=-1).cuda()
    gamma, beta = (torch.rand(C) + 0.5).cuda(), rnd(C, scale=0.3).cuda()
    sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
    dy = torch.zeros_like(y)
    try:
        ops.set_option("bn_fused", 1)
        n0 = ops.launch_count()
        ops.bn_bwd(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)
        torch.cuda.synchronize()
        assert ops.launch_count() - n0 == 1
        want, want_s = dy.clone(), sums.clone()
        side = torch.cuda.Stream()
        big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        for rep in range(30):
            with torch.cuda.stream(side):
                for _ in range(3):
                    big @ big
            dy.zero_()
            ops.bn_bwd(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)
        torch.cuda.synchronize()
    finally:
        ops.set_option("bn_fused", 0)
    assert (dy.float() - want.float()).abs().max().item() <= 1e-2 * want.float().abs().max().item()
    assert torch.allclose(sums, want_s, rtol=1e-6, atol=1e-6 * want_s.abs().max().item())


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,rows,act", [(64, 3 * 128 * 1024, ACT_LRELU), (24, 1000, ACT_RELU), (640, 77, ACT_NONE), (3, 4096, ACT_LRELU)])
def test_act_bwd_with_column_sums(mode, C, rows, act):
    """sg_act_bwd_colsum: the activation backward and the column sums of its result (the bias gradient of a conv + bias +
    activation layer) in one pass == act_bwd, then colsum; the sums are ADDED to what the buffer held."""
    ops, emu = _ops(mode), EmuOps(torch.float64)
    sd = ops.act_dtype
    da, a = rnd(rows, C, seed=1).to(sd), rnd(rows, C, seed=2).to(sd)
    cs0 = rnd(C, seed=3)
    want, want_cs = torch.zeros(rows, C, dtype=torch.float64), cs0.double().clone()
    emu.act_bwd(da.double(), a.double(), want, act)
    emu.colsum(want.to(sd).double(), want_cs)                     # of the stored values
    out, cs = torch.zeros(rows, C, dtype=sd, device="cuda"), cs0.cuda()
    n0 = ops.launch_count()
    ops.act_bwd(da.cuda(), a.cuda(), out, act, colsum=cs)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == (1 if C % 8 == 0 else 2)
    t = TOL[mode]
    assert torch.allclose(out.double().cpu(), want, rtol=t["rtol"], atol=t["atol"] * want.abs().max().item())
    assert torch.allclose(cs.double().cpu(), want_cs, rtol=1e-3, atol=1e-3 * want_cs.abs().max().item())


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,rows,launches", [(128, 128 * 256, 1), (256, 128 * 64, 1), (512, 2048, 1), (64, 50000, 1), (24, 3000, 1),
                                             (128, 400000, 2)])
def test_gp_bn_one_call(mode, C, rows, launches):
    """sg_gp_bn: the penalty's second-order pass through a train-mode BN as one call (ONE launch where the four tensors fit the
    SMs' shared memory: sums, rendezvous, apply) == gp_bn_reduce, then gp_bn_apply; repeated on the same call site."""
    ops, emu = _ops(mode), EmuOps(torch.float64)
    sd = ops.act_dtype
    q = lambda t: t.to(sd)
    y, da, v = rnd(rows, C, seed=1), rnd(rows, C, seed=2), rnd(rows, C, seed=3)
    mr = torch.stack([rnd(1, C, scale=0.1), torch.rand(1, C) + 0.5], dim=-1)
    gamma, beta = torch.rand(C) + 0.5, rnd(C, scale=0.3)
    a = torch.zeros(rows, C, dtype=torch.float64)
    emu.bn_act(q(y).double(), mr.double(), gamma.double(), beta.double(), a, 1, ACT_LRELU)
    a = q(a.float())
    sums = torch.zeros(1, C, 2, dtype=torch.float64)
    emu.bn_bwd_reduce(q(da).double(), a.double(), q(y).double(), mr.double(), sums, 1, ACT_LRELU)
    dg0 = rnd(C, seed=4)
    want_t, want_w, want_gy, want_dg = (torch.zeros(C, 3, dtype=torch.float64), torch.zeros(rows, C, dtype=torch.float64),
                                        torch.zeros(rows, C, dtype=torch.float64), dg0.double().clone())
    emu.gp_bn(q(v).double(), q(da).double(), a.double(), q(y).double(), mr.double(), gamma.double(), sums, want_t, want_w, want_gy,
              want_dg, ACT_LRELU)
    c = lambda t: t.cuda().contiguous()
    cv, cda, ca, cy, cmr, cg, cs = c(q(v)), c(q(da)), c(a), c(q(y)), c(mr.float()), c(gamma), c(sums)
    ts = torch.full((C, 3), 7.0, dtype=torch.float64, device="cuda")
    w, gy = torch.zeros(rows, C, dtype=sd, device="cuda"), torch.zeros(rows, C, dtype=sd, device="cuda")
    tol = dict(rtol=3e-2, atol=2e-3) if mode == "bf16" else dict(rtol=1e-3, atol=1e-4)
    for rep in range(3):
        dg = dg0.cuda()
        if rep == 2:
            ts.zero_()
        n0 = ops.launch_count()
        ops.gp_bn(cv, cda, ca, cy, cmr, cg, cs, ts, w, gy, dg, ACT_LRELU, zeroed=rep == 2)
        torch.cuda.synchronize()
        seen = ops.launch_count() - n0
        assert seen == launches if mode == "bf16" else seen in (1, 2, 3), seen     # (two kernels + dgamma's fallback = 2 or 3)
        assert torch.allclose(ts.cpu(), want_t, rtol=1e-3, atol=1e-3 * want_t.abs().max().item())
        for got, want in ((w, want_w), (gy, want_gy)):
            assert torch.allclose(got.double().cpu(), want, rtol=tol["rtol"], atol=tol["atol"] * want.abs().max().item()), \
                (got.double().cpu() - want).abs().max().item()
        assert torch.allclose(dg.double().cpu(), want_dg, rtol=1e-3, atol=1e-3 * want_dg.abs().max().item())


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,rows,G,act", [(128, 32768, 3, ACT_LRELU), (256, 8192, 3, ACT_LRELU), (512, 2048, 3, ACT_LRELU),
                                          (64, 100000, 1, ACT_RELU), (24, 131072, 1, ACT_NONE)])
def test_bn_finalize_act_bulk(mode, C, rows, G, act):
    """Option bn_act_bulk: finalize + apply with every CTA's range brought in by bulk async copies (one CTA per SM) == the
    register-staged kernel's contract (mr, running statistics, batch counter, output)."""
    g = torch.Generator().manual_seed(C + rows)
    y = torch.randn(G * rows, C, generator=g) * 2 + 0.5
    emu = EmuOps(torch.float64)
    sd = torch.bfloat16 if mode == "bf16" else torch.float32
    stats = torch.zeros(G, C, 2, dtype=torch.float64)
    emu.col_stats(y.to(sd).double().reshape(G, rows, 1, C), stats, G)
    rm, rv, nbt = rnd(C) * 0.1, torch.rand(C) + 0.5, torch.tensor(3)
    gamma, beta = rnd(C) * 0.5 + 1, rnd(C) * 0.2
    y4 = y.reshape(G * rows, 1, 1, C)
    ops = _ops(mode)
    try:
        ops.set_option("bn_act_bulk", 1)
        n0 = ops.launch_count()
        ea, ca = run_pair(mode, "bn_finalize_act", [D(stats), rows, F(torch.zeros(G, C, 2)), F(rm), F(rv), I64(nbt), 2, T(y4),
                                                    F(gamma), F(beta), T(torch.zeros_like(y4)), act], [2, 3, 4, 10],
                          tol=dict(rtol=1e-4, atol=1e-5) if mode == "fp32" else None)
        assert int(ca[5]) == 3 + G + 1
        assert ops.launch_count() - n0 == 1
    finally:
        ops.set_option("bn_act_bulk", 0)


def test_zero_multi_and_accumulating_reductions():
    """sg_zero_multi zeroes up to 32 buffers in one launch (more: one launch per 32); the *_acc reductions / sg_bn_bwd with
    sums_zeroed ADD to what the caller zeroed -- twice the call, twice the sums."""
    ops = _ops("bf16")
    bufs = [torch.full((n,), 3.0, dtype=dt, device="cuda") for n, dt in
            [(1, torch.float32), (7, torch.float32), (512 * 3 * 2, torch.float64), (1000, torch.float64), (128, torch.float32)] * 8]
    n0 = ops.launch_count()
    ops.zero_multi(bufs + [None])
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == 2 and all(float(b.abs().max()) == 0.0 for b in bufs)
    C, rows = 64, 4096
    y, da, v = (rnd(rows, C, seed=i).cuda().bfloat16() for i in (1, 2, 3))
    a = torch.nn.functional.leaky_relu(y.float(), 0.1).bfloat16()
    mr = torch.stack([rnd(1, C, scale=0.1), torch.rand(1, C) + 0.5], dim=-1).cuda()
    gamma, beta = (torch.rand(C) + 0.5).cuda(), rnd(C, scale=0.3).cuda()
    ts1, ts2 = torch.full((C, 3), 5.0, dtype=torch.float64, device="cuda"), torch.zeros(C, 3, dtype=torch.float64, device="cuda")
    ops.gp_bn_reduce(v, da, a, y, mr, ts1, ACT_LRELU)
    ops.gp_bn_reduce(v, da, a, y, mr, ts2, ACT_LRELU, zeroed=True)
    ops.gp_bn_reduce(v, da, a, y, mr, ts2, ACT_LRELU, zeroed=True)
    assert torch.allclose(ts2, 2 * ts1, rtol=1e-6, atol=1e-6 * float(ts1.abs().max()))
    g = rnd(8, 1024, seed=4).cuda().bfloat16()
    q1, q2 = torch.full((8,), 5.0, device="cuda"), torch.zeros(8, device="cuda")
    ops.sample_sqnorm(g, q1)
    ops.sample_sqnorm(g, q2, zeroed=True)
    ops.sample_sqnorm(g, q2, zeroed=True)
    assert torch.allclose(q2, 2 * q1, rtol=1e-5)
    s1, s2 = torch.full((1, C, 2), 5.0, dtype=torch.float64, device="cuda"), torch.zeros(1, C, 2, dtype=torch.float64, device="cuda")
    dy = torch.zeros_like(y)
    ops.bn_bwd(da, None, y, mr, gamma, s1, dy, 1, ACT_LRELU, beta=beta)
    ops.bn_bwd(da, None, y, mr, gamma, s2, dy, 1, ACT_LRELU, beta=beta, zeroed=True)
    torch.cuda.synchronize()
    assert torch.allclose(s2, s1, rtol=1e-6, atol=1e-6 * float(s1.abs().max()))


@pytest.mark.parametrize("mode", MODES)
def test_bn_eval_mr(mode):
    C = 96
    run_pair(mode, "bn_eval_mr", [F(rnd(C)), F(torch.rand(C) + 0.1), F(torch.zeros(1, C, 2))], [2],
             tol=dict(rtol=1e-5, atol=1e-6))


@pytest.mark.parametrize("mode", ["fp32"])
@pytest.mark.parametrize("N,K,M", [(9, 512, 256), (256, 512, 128), (128, 256, 128), (70, 300, 37)])
def test_linear(mode, N, K, M):
    x, w, b = rnd(N, K), rnd(M, K, scale=K ** -0.5), rnd(M)
    ea, _ = run_pair(mode, "linear_fwd", [F(x), F(w), F(b), F(torch.zeros(N, M))], [3], dict(relu=True))
    h = ea[3].float()
    dout = rnd(N, M, seed=3)
    run_pair(mode, "linear_bwd", [F(x), F(w), F(dout), F(rnd(M, K)), F(rnd(M)), F(rnd(N, K))], [3, 4, 5],
             dict(dx_acc=True, relu_out=F(h)))
    run_pair(mode, "linear_bwd", [F(x), F(w), F(dout), None, None, F(rnd(N, K))], [5], dict(dx_acc=False))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("K,Cx,Nd", [(128, 512, 128), (160, 512, 128)])
def test_head(mode, K, Cx, Nd):
    N = 6
    wcr, bcr = rnd(K, Cx + Nd, 1, 1, scale=0.04), rnd(K, scale=0.04)
    wcs, bcs = rnd(1, K * 16, scale=0.02), rnd(1, scale=0.02)
    ea, _ = run_pair(mode, "head_prepare", [F(wcr), F(bcr), F(wcs), F(bcs), F(torch.zeros(16, Cx)), F(torch.zeros(Nd)),
                                            F(torch.zeros(1))], [4, 5, 6], tol=dict(rtol=1e-4, atol=1e-5))
    A, Bv, c0 = ea[4].float(), ea[5].float(), ea[6].float()
    a4, ce = rnd(N, 4, 4, Cx), rnd(N, Nd)
    run_pair(mode, "head_fwd", [T(a4), F(ce), F(A), F(Bv), F(c0), F(torch.zeros(N))], [5],
             tol=dict(rtol=1e-3, atol=1e-4))
    # the four score rows of a critic forward in one launch (real, mismatched, fake, interpolated)
    a_all, ce2 = rnd(3 * N, 4, 4, Cx, seed=7), rnd(2 * N, Nd, seed=8)
    jobs = [(0, 0, 0), (N, 0, 2 * N), (2 * N, 0, 3 * N), (0, N, N)]
    run_pair(mode, "head_fwd_multi", [T(a_all), F(ce2), F(A), F(Bv), F(c0), F(torch.zeros(4, N)), jobs, N], [5],
             tol=dict(rtol=1e-3, atol=1e-4))
    run_pair(mode, "head_fwd_multi", [T(a_all), F(ce2), F(A), F(Bv), F(c0), F(torch.zeros(4, N)), jobs[2:3], N], [5],
             tol=dict(rtol=1e-3, atol=1e-4))
    coef = rnd(N)
    coef[2] = 0.0
    run_pair(mode, "head_bwd_data", [F(coef), F(A), T(torch.zeros(N, 4, 4, Cx))], [2])
    run_pair(mode, "head_bwd_data", [F(coef), F(Bv), F(torch.zeros(N, Nd))], [2])
    run_pair(mode, "head_bwd_reduce", [F(coef), T(a4), F(rnd(16, Cx))], [2], tol=dict(rtol=1e-3, atol=1e-4))
    run_pair(mode, "head_param_grads", [F(rnd(16, Cx)), F(rnd(Nd)), F(rnd(1)), F(wcr), F(bcr), F(wcs),
                                        F(torch.zeros(K, Cx + Nd, 1, 1)), F(torch.zeros(K)), F(torch.zeros(1, K * 16)),
                                        F(torch.zeros(1))], [6, 7, 8, 9], tol=dict(rtol=1e-3, atol=1e-4))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("N", [7, 128])
def test_ca_fused(mode, N):
    """con_augment.py:13-22 in one launch / its backward in two, against the composition of the per-op statements."""
    Tm, Hd, C, nz, ld = 512, 256, 128, 100, 256
    tem, eps, z = rnd(N, Tm), rnd(N, C, seed=2), rnd(N, nz, seed=3)
    Wh, bh = rnd(Hd, Tm, scale=Tm ** -0.5), rnd(Hd, seed=5, scale=0.1)
    Wmu, bmu = rnd(C, Hd, seed=6, scale=Hd ** -0.5), rnd(C, seed=7, scale=0.1)
    Wsg, bsg = rnd(C, Hd, seed=8, scale=Hd ** -0.5), rnd(C, seed=9, scale=0.1) + 1.0
    z0 = lambda *sh: torch.zeros(*sh)
    t = dict(rtol=1e-4, atol=1e-5)
    ea, _ = run_pair(mode, "ca_forward", [F(tem), F(Wh), F(bh), F(Wmu), F(bmu), F(Wsg), F(bsg), F(eps), F(z), F(z0(N, Hd)),
                                          F(z0(N, C)), F(z0(N, C)), F(z0(N, C)), T(torch.ones(N, 1, 1, ld))], [9, 10, 11, 12], tol=t)
    run_pair(mode, "ca_forward", [F(tem), F(Wh), F(bh), F(Wmu), F(bmu), F(Wsg), F(bsg), F(eps), F(z), F(z0(N, Hd)),
                                  F(z0(N, C)), F(z0(N, C)), F(z0(N, C)), T(torch.ones(N, 1, 1, ld))], [13])
    # encode only (eps None), no generator row
    run_pair(mode, "ca_forward", [F(tem), F(Wh), F(bh), F(Wmu), F(bmu), F(Wsg), F(bsg), None, None, F(z0(N, Hd)),
                                  F(z0(N, C)), F(z0(N, C)), F(z0(N, C)), None], [9, 10, 11], tol=t)
    h, mu, sigma = ea[9].float(), ea[10].float(), ea[11].float()
    dcg = rnd(N, 1, 1, ld, seed=11)
    g0 = lambda *sh: rnd(*sh, seed=13, scale=0.1)                    # gradients accumulate into non-zero buffers
    for dtem_acc in (False, True):
        run_pair(mode, "ca_backward", [T(dcg), F(eps), F(mu), F(sigma), 0.7, F(h), F(tem), F(Wmu), F(Wsg), F(Wh), F(z0(N, C)),
                                       F(z0(N, C)), F(z0(N, Hd)), F(g0(C, Hd)), F(g0(C)), F(g0(C, Hd)), F(g0(C)), F(g0(Hd, Tm)),
                                       F(g0(Hd)), F(g0(N, Tm)), dtem_acc], [10, 11, 12, 13, 14, 15, 16, 17, 18, 19],
                 tol=dict(rtol=1e-4, atol=2e-5))
    # Stage-II form: d loss / d c_hat given as fp32 [N, C], text side frozen (no dtem)
    run_pair(mode, "ca_backward", [F(rnd(N, C, seed=12)), F(eps), F(mu), F(sigma), 0.0, F(h), F(tem), F(Wmu), F(Wsg), F(Wh),
                                   F(z0(N, C)), F(z0(N, C)), F(z0(N, Hd)), F(g0(C, Hd)), F(g0(C)), F(g0(C, Hd)), F(g0(C)),
                                   F(g0(Hd, Tm)), F(g0(Hd)), None, False], [10, 11, 12, 13, 14, 15, 16, 17, 18],
             tol=dict(rtol=1e-4, atol=2e-5))


@pytest.mark.parametrize("mode", MODES)
def test_ca_and_losses(mode):
    N, C, nz = 7, 128, 100
    mu, sigma, eps, z = rnd(N, C), rnd(N, C, seed=1), rnd(N, C, seed=2), rnd(N, nz)
    run_pair(mode, "ca_reparam", [F(mu), F(sigma), F(eps), F(z), F(torch.zeros(N, C)), T(torch.zeros(N, 1, 1, C + nz))], [4, 5])
    dcg = rnd(N, 1, 1, C + nz, seed=4)
    run_pair(mode, "ca_bwd_seed", [T(dcg), F(eps), F(mu), F(sigma), 1.0, F(torch.zeros(N, C)), F(torch.zeros(N, C))], [5, 6],
             tol=dict(rtol=1e-4, atol=1e-6))
    real, fake = rnd(N, 8, 8, 3), rnd(N, 8, 8, 3, seed=3)
    e = torch.rand(N)
    run_pair(mode, "interp", [T(real), T(fake), F(e), T(torch.zeros(N, 8, 8, 3))], [3])
    g = rnd(N, 64, 64, 3, scale=0.02)
    ea, _ = run_pair(mode, "sample_sqnorm", [T(g), F(torch.zeros(N))], [1], tol=dict(rtol=1e-4, atol=1e-6))
    sq = ea[1].float()
    run_pair(mode, "gp_seed", [T(g), F(sq), 2.5, T(torch.zeros_like(g))], [3])
    sr, sm, sf = rnd(N), rnd(N, seed=1), rnd(N, seed=2)
    run_pair(mode, "critic_loss", [F(sr), F(sm), F(sf), F(sq), 10.0, F(torch.zeros(2))], [5], tol=dict(rtol=1e-5, atol=1e-6))
    run_pair(mode, "gen_loss", [F(sf), F(mu), F(sigma), F(torch.zeros(2))], [3], tol=dict(rtol=1e-5, atol=1e-6))
    run_pair(mode, "scale_rows_add", [T(real), F(e), T(fake), True], [2])


def test_adam_matches_torch():
    from imagegenerator_b200.ops import CudaOps
    ops = CudaOps("fp32")
    n = 4096 + 8
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 0.0, 0.0, 0.0], device="cuda")
    for step in range(5):
        g = torch.randn(n, device="cuda") * (0.1 ** step)
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, hyper)
    torch.cuda.synchronize()
    assert float(hyper[4]) == 5
    assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6)


TC_CASES = CONV_CASES + [
    (128, 16, 128, 256, 4, 2, 1),    # critic ds3 at the bench batch
    (16, 32, 64, 128, 4, 2, 1),
    (24, 8, 256, 512, 4, 2, 1),      # critic ds4, batch not a multiple of the 8-image tile
    (4, 16, 640, 320, 3, 1, 1),      # Stage-II residual block layer1 operator
    (2, 128, 16, 32, 4, 2, 1),       # Stage-II critic ds2 operator (Wo = 64)
    (2, 256, 8, 16, 4, 2, 1),        # wide rows: Wo = 128 -> one-row tiles
]


def _tc_ok(mode, case):
    from imagegenerator_b200.ops import CudaOps
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    return bool(CudaOps("bf16").lib.sg_conv_tc_supported(mode, N, H, H, Ci, Ho, Ho, Co, k, s, p))


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_LRELU, True)])
def test_conv_fprop_tcgen05(case, act, use_bias):
    if not _tc_ok(0, case):
        pytest.skip("shape routed to the CUDA-core kernel")
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    x, w = rnd(N, H, H, Ci), rnd(Co, Ci, k, k, scale=(Ci * k * k) ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    bias = F(rnd(Co)) if use_bias else None
    run_pair("bf16", "conv_fprop", [T(x), T(pf), bias, T(torch.zeros(N, Ho, Ho, Co)), k, s, p], [3], dict(act=act, impl="_tc"))


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("act,use_bias", [(ACT_NONE, False), (ACT_TANH, True)])
def test_conv_dgrad_tcgen05(case, act, use_bias):
    if not _tc_ok(1, case):
        pytest.skip("shape routed to the CUDA-core kernel")
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    dy, w = rnd(N, Ho, Ho, Co), rnd(Co, Ci, k, k, scale=(Co * k * k / (s * s)) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    bias = F(rnd(Ci)) if use_bias else None
    run_pair("bf16", "conv_dgrad", [T(dy), T(pd), bias, T(torch.zeros(N, H, H, Ci)), k, s, p], [3], dict(act=act, impl="_tc"))


STATS_CASES = [
    # N, H, Ci, Co, k, s, p, groups
    (24, 8, 256, 512, 4, 2, 1, 3),      # critic ds4: 8 images x 16 pixels = one 128-row tile per group -> fused
    (48, 16, 128, 256, 4, 2, 1, 3),     # critic ds3 (CTA pairs), 3 groups
    (6, 32, 64, 128, 4, 2, 1, 3),       # critic ds2
    (4, 16, 640, 320, 3, 1, 1, 1),      # Stage-II residual block
    (4, 16, 48, 96, 4, 2, 1, 1),        # generator up-layer operator
    (3, 8, 256, 512, 4, 2, 1, 3),       # 1 image / group: 16 rows -> not fusable, conv + col_stats
    (2, 16, 40, 24, 3, 1, 1, 2),
]


def _check_stats(y_dev, stats_dev, st0, G):
    """The statistics must be those of the tensor the kernel STORED (what bn_act will normalise)."""
    C = y_dev.shape[-1]
    v = y_dev.double().cpu().reshape(G, -1, C)
    want = st0.clone()
    want[:, :, 0] += v.sum(1)
    want[:, :, 1] += (v * v).sum(1)
    got = stats_dev.cpu()
    scale = want.abs().amax(dim=(0, 1), keepdim=True)
    assert ((got - want).abs() <= 1e-4 * scale + 1e-6).all(), (got - want).abs().max().item()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", STATS_CASES)
def test_conv_fprop_stats(mode, case):
    N, H, Ci, Co, k, s, p, G = case
    Ho = (H + 2 * p - k) // s + 1
    x, w = rnd(N, H, H, Ci), rnd(Co, Ci, k, k, scale=(Ci * k * k) ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    st0 = torch.ones(G, Co, 2, dtype=torch.float64) * 0.25            # accumulate semantics
    ea, ca = run_pair(mode, "conv_fprop_stats", [T(x), T(pf), T(torch.zeros(N, Ho, Ho, Co)), D(st0), G, k, s, p], [2])
    _check_stats(ca[2], ca[3], st0, G)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", STATS_CASES)
def test_conv_dgrad_stats(mode, case):
    N, H, Ci, Co, k, s, p, G = case
    Ho = (H + 2 * p - k) // s + 1
    dy, w = rnd(N, Ho, Ho, Co), rnd(Co, Ci, k, k, scale=(Co * k * k / (s * s)) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    st0 = torch.zeros(G, Ci, 2, dtype=torch.float64)
    ea, ca = run_pair(mode, "conv_dgrad_stats", [T(dy), T(pd), T(torch.zeros(N, H, H, Ci)), D(st0), G, k, s, p], [2])
    _check_stats(ca[2], ca[3], st0, G)


def _bstats_inputs(G, C, shape, act, seed):
    """mean / rstd / gamma / beta and a pre-BN tensor whose activation argument gamma*xhat+beta stays 0.25 away from 0, so that
    fp32 vs fp64 evaluation of the mask cannot disagree (a flipped mask moves a sum by a whole element)."""
    g = torch.Generator().manual_seed(seed)
    mr = torch.stack([torch.randn(G, C, generator=g) * 0.3, torch.rand(G, C, generator=g) + 0.5], dim=-1)
    gamma = (torch.rand(C, generator=g) + 0.5) * torch.where(torch.rand(C, generator=g) < 0.3, -1.0, 1.0)
    beta = torch.randn(C, generator=g) * 0.2
    u = torch.randn(*shape, generator=g)
    u = torch.sign(u) * (0.25 + u.abs())
    N = shape[0]
    mean = mr[:, :, 0].repeat_interleave(N // G, dim=0)[:, None, None, :]
    rstd = mr[:, :, 1].repeat_interleave(N // G, dim=0)[:, None, None, :]
    ybn = mean + (u - beta) / (rstd * gamma)
    return mr, gamma, beta, ybn


def _check_bstats(da_dev, ybn_dev, mr, gamma, beta, sums_dev, G, act):
    """The statistics must be those of the STORED gradient (what bn_bwd_apply will read), recomputed in fp64."""
    emu = EmuOps(torch.float64)
    want = torch.zeros(sums_dev.shape, dtype=torch.float64)
    da, ybn = da_dev.double().cpu(), ybn_dev.double().cpu()
    emu.bn_bwd_reduce(da, None, ybn, mr.double(), want, G, act, gamma=gamma.double(), beta=beta.double())
    C = da.shape[-1]
    scale = da.abs().reshape(G, -1, C).sum(1)[:, :, None]
    xmax = 12.0                                                       # |xhat| of the constructed data stays well below this
    got = sums_dev.cpu()
    assert ((got - want).abs() <= 2e-4 * scale * xmax + 1e-6).all(), ((got - want).abs() / (scale + 1e-9)).max().item()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", STATS_CASES)
@pytest.mark.parametrize("act", [ACT_RELU, ACT_LRELU])
def test_conv_fprop_bstats(mode, case, act):
    """conv + the BatchNorm-backward statistics of the layer below in its epilogue == conv, then bn_bwd_reduce."""
    N, H, Ci, Co, k, s, p, G = case
    if Co % 8:
        pytest.skip("the *_y BatchNorm kernels need C % 8 == 0")
    Ho = (H + 2 * p - k) // s + 1
    x, w = rnd(N, H, H, Ci), rnd(Co, Ci, k, k, scale=(Ci * k * k) ** -0.5)
    pf = w.permute(0, 2, 3, 1).contiguous()
    mr, gamma, beta, ybn = _bstats_inputs(G, Co, (N, Ho, Ho, Co), act, 5)
    sums = torch.full((G, Co, 2), 7.0, dtype=torch.float64)           # must be overwritten, not accumulated
    _ops(mode).set_option("bstats_min_k", 0)                          # the fused epilogue whatever the reduction depth
    ea, ca = run_pair(mode, "conv_fprop_bstats", [T(x), T(pf), T(torch.zeros(N, Ho, Ho, Co)), T(ybn), F(mr), F(gamma), F(beta),
                                                  D(sums), G, act, k, s, p], [2])
    _ops(mode).set_option("bstats_min_k", 4000)
    _check_bstats(ca[2], ca[3], mr, gamma, beta, ca[7], G, act)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", STATS_CASES)
@pytest.mark.parametrize("act", [ACT_RELU, ACT_LRELU])
def test_conv_dgrad_bstats(mode, case, act):
    N, H, Ci, Co, k, s, p, G = case
    if Ci % 8:
        pytest.skip("the *_y BatchNorm kernels need C % 8 == 0")
    Ho = (H + 2 * p - k) // s + 1
    dy, w = rnd(N, Ho, Ho, Co), rnd(Co, Ci, k, k, scale=(Co * k * k / (s * s)) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    mr, gamma, beta, ybn = _bstats_inputs(G, Ci, (N, H, H, Ci), act, 6)
    sums = torch.full((G, Ci, 2), 7.0, dtype=torch.float64)
    _ops(mode).set_option("bstats_min_k", 0)
    ea, ca = run_pair(mode, "conv_dgrad_bstats", [T(dy), T(pd), T(torch.zeros(N, H, H, Ci)), T(ybn), F(mr), F(gamma), F(beta),
                                                  D(sums), G, act, k, s, p], [2])
    _ops(mode).set_option("bstats_min_k", 4000)
    _check_bstats(ca[2], ca[3], mr, gamma, beta, ca[7], G, act)


MASKED_CASES = [(24, 8, 256, 512, 4, 2, 1, 3), (48, 16, 128, 256, 4, 2, 1, 3), (6, 32, 64, 128, 4, 2, 1, 3), (6, 32, 64, 128, 4, 2, 1, 1),
                (4, 16, 96, 192, 4, 2, 1, 1)]


@pytest.mark.parametrize("case", MASKED_CASES)
@pytest.mark.parametrize("act", [ACT_RELU, ACT_LRELU])
@pytest.mark.parametrize("identity", [False, True])
def test_conv_dgrad_masked(case, act, identity):
    """sg_conv_dgrad_tc_bstats_masked: the data-gradient conv stores dz = conv * act'(gamma * xhat + beta) itself (conv + activation
    backward in one kernel) == conv_dgrad, then the mask; sums = (sum dz, sum dz * xhat) of the stored dz.  ``identity``: the table
    the engines use for the first critic layer (mean 0, rstd 1, gamma 1, beta 0, ybn = the stored activation)."""
    N, H, Ci, Co, k, s, p, G = case
    Ho = (H + 2 * p - k) // s + 1
    dy, w = rnd(N, Ho, Ho, Co), rnd(Co, Ci, k, k, scale=(Co * k * k / (s * s)) ** -0.5)
    pd = w.permute(1, 2, 3, 0).contiguous()
    mr, gamma, beta, ybn = _bstats_inputs(G, Ci, (N, H, H, Ci), act, 6)
    if identity:
        mr = torch.stack([torch.zeros(G, Ci), torch.ones(G, Ci)], dim=-1)
        gamma, beta = torch.ones(Ci), torch.zeros(Ci)
        u = rnd(N, H, H, Ci, seed=8)
        ybn = torch.sign(u) * (0.25 + u.abs())
    ops = _ops("bf16")
    assert ops.conv_dgrad_masked_supported(torch.empty(N, Ho, Ho, Co, dtype=torch.bfloat16), torch.empty(N, H, H, Ci), k, s, p, G)
    for zeroed in (False, True):
        sums = torch.zeros(G, Ci, 2, dtype=torch.float64) if zeroed else torch.full((G, Ci, 2), 7.0, dtype=torch.float64)
        ea, ca = run_pair("bf16", "conv_dgrad_masked", [T(dy), T(pd), T(torch.zeros(N, H, H, Ci)), T(ybn), F(mr), F(gamma), F(beta),
                                                        D(sums), G, act, k, s, p], [2], dict(zeroed=zeroed))
        # the statistics are those of the STORED masked gradient: re-mask it with slope 1 (none) to re-use the checker
        _check_bstats(ca[2], ca[3], mr, gamma, beta, ca[7], G, ACT_NONE)


def _wtc_ok(case):
    from imagegenerator_b200.ops import CudaOps
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    return bool(CudaOps("bf16").lib.sg_conv_wgrad_tc_supported(N, H, H, Ci, Ho, Ho, Co, k, s, p))


@pytest.mark.parametrize("case", TC_CASES + [(5, 8, 192, 96, 4, 2, 1), (7, 32, 24, 48, 4, 2, 1)])
def test_conv_wgrad_tcgen05(case):
    if not _wtc_ok(case):
        pytest.skip("shape routed to the CUDA-core kernel")
    N, H, Ci, Co, k, s, p = case
    Ho = (H + 2 * p - k) // s + 1
    x, dy = rnd(N, H, H, Ci), rnd(N, Ho, Ho, Co, scale=(N * Ho * Ho) ** -0.5)
    dw0 = rnd(Co, Ci, k, k, scale=0.1)
    run_pair("bf16", "conv_wgrad", [T(x), T(dy), F(dw0), k, s, p], [2], dict(impl="_tc"), tol=dict(rtol=2e-3, atol=2e-4))


@pytest.mark.parametrize("case", [(4, 16, 640, 320, 3, 1, 1), (4, 16, 320, 320, 3, 1, 1), (3, 16, 128, 256, 4, 2, 1),
                                  (2, 32, 40, 24, 3, 1, 1)])
def test_conv_wgrad_channels_last_and_fold(case):
    """wgrad into the channels-last buffer + fold == wgrad into the PyTorch layout (accumulate semantics on both)."""
    N, H, Ci, Co, k, s, p = case
    ops = _ops("bf16")
    Ho = (H + 2 * p - k) // s + 1
    if not ops.lib.sg_conv_wgrad_cl_supported(N, H, H, Ci, Ho, Ho, Co, k, s, p, 1):
        pytest.skip("not eligible")
    x, dy = rnd(N, H, H, Ci), rnd(N, Ho, Ho, Co, scale=(N * Ho * Ho) ** -0.5)
    gw0, dw0 = rnd(Co, k, k, Ci, scale=0.1), rnd(Co, Ci, k, k, scale=0.1)
    run_pair("bf16", "conv_wgrad_cl", [T(x), T(dy), F(gw0), k, s, p], [2], tol=dict(rtol=2e-3, atol=2e-4))
    ea, ca = run_pair("bf16", "fold_grad_cl", [F(gw0), F(dw0)], [0, 1], tol=dict(rtol=1e-6, atol=1e-7))
    assert float(ca[0].abs().max()) == 0.0


@pytest.mark.parametrize("mode", MODES)
def test_patchify(mode):
    x = rnd(3, 16, 16, 3)
    run_pair(mode, "patchify", [T(x), T(torch.zeros(3, 8, 8, 48)), 4, 2, 1], [1])


@pytest.mark.parametrize("mode", MODES)
def test_patchify_generic(mode):
    x = rnd(2, 9, 9, 2)                                   # odd size, 2 channels, k3 s1 p1 -> generic kernel
    run_pair(mode, "patchify", [T(x), T(torch.zeros(2, 9, 9, 18)), 3, 1, 1], [1])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", [(3, 8, 3, 4, 2, 1, ACT_TANH, True), (2, 16, 3, 4, 2, 1, ACT_NONE, False),
                                  (2, 37, 3, 4, 2, 1, ACT_TANH, True), (1, 64, 3, 4, 2, 1, ACT_NONE, False),
                                  (2, 5, 2, 3, 1, 1, ACT_NONE, True), (1, 4, 4, 4, 1, 0, ACT_LRELU, True)])
def test_unpatchify(mode, case):
    N, Hi, C, k, s, p, act, use_bias = case
    Ho = (Hi - 1) * s - 2 * p + k
    col = rnd(N, Hi, Hi, C * k * k)
    bias = F(rnd(C)) if use_bias else None
    run_pair(mode, "unpatchify", [F(col), bias, T(torch.zeros(N, Ho, Ho, C)), k, s, p], [2], dict(act=act))


@pytest.mark.parametrize("mode", MODES)
def test_thin_conv_transpose_via_col2im(mode):
    """ConvTranspose2d(Cin -> 3, k4 s2 p1) as 1x1 GEMM + col2im == the direct data-gradient kernel's definition."""
    N, Hi, Cin = 2, 16, 24
    x, w = rnd(N, Hi, Hi, Cin), rnd(Cin, 3, 4, 4, scale=(Cin * 4) ** -0.5)      # conv op: Co=Cin, Ci=3
    pd = w.permute(1, 2, 3, 0).contiguous()                                    # [3][4][4][Cin]
    bias = rnd(3)
    ops = _ops(mode)
    dev = lambda t, d=None: t.to(d or ops.act_dtype).cuda().contiguous()
    xq, pdq = dev(x), dev(pd)
    col = ops.empty((N, Hi, Hi, 48), torch.float32)
    out = ops.empty((N, 2 * Hi, 2 * Hi, 3))
    ops.conv_fprop_f32out(xq, pdq.view(48, 1, 1, Cin), col, 1, 1, 0)
    ops.unpatchify(col, dev(bias, torch.float32), out, 4, 2, 1, act=ACT_TANH)
    ref = torch.tanh(torch.nn.functional.conv_transpose2d(xq.double().cpu().permute(0, 3, 1, 2),
                                                          pdq.double().cpu().permute(3, 0, 1, 2), bias.double(), 2, 1))
    got = out.double().cpu().permute(0, 3, 1, 2)
    t = TOL[mode]
    assert torch.allclose(got, ref, rtol=t["rtol"], atol=t["atol"] * ref.abs().max().item()), (got - ref).abs().max().item()


@pytest.mark.parametrize("mode", MODES)
def test_concat_split_affine(mode):
    N, H, Cx, Cc = 3, 16, 512, 128
    x, c = rnd(N, H, H, Cx), rnd(N, Cc)
    run_pair(mode, "concat_rep", [T(x), F(c), T(torch.zeros(N, H, H, Cx + Cc))], [2])
    dout = rnd(N, H, H, Cx + Cc, seed=2)
    run_pair(mode, "split_rep_bwd", [T(dout), T(torch.zeros(N, H, H, Cx)), F(torch.zeros(N, Cc))], [1, 2],
             tol=dict(rtol=1e-3, atol=1e-4))
    run_pair("fp32", "affine_f32", [F(torch.rand(37)), -1.0, 1.0, F(torch.zeros(37))], [3])


def test_bn_param_grad_multi():
    """All BatchNorm layers' gamma / beta gradients in one launch == the per-layer kernel."""
    items_e, items_c = [], []
    emu = EmuOps(torch.float64)
    ops = _ops("bf16")
    for i, (G, C) in enumerate([(3, 128), (1, 24), (3, 512), (2, 640), (1, 80)]):
        sums = rnd(G, C, 2, seed=i).double()
        dg, db = rnd(C, seed=10 + i), rnd(C, seed=20 + i)
        items_e.append((sums.clone(), dg.double().clone(), db.double().clone()))
        items_c.append((sums.cuda(), dg.float().cuda(), db.float().cuda()))
    emu.bn_param_grad_multi(items_e)
    ops.bn_param_grad_multi(items_c)
    torch.cuda.synchronize()
    for (_, eg, eb), (_, cg, cb) in zip(items_e, items_c):
        assert torch.allclose(cg.double().cpu(), eg, rtol=1e-5, atol=1e-6) and torch.allclose(cb.double().cpu(), eb, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", [(48, 16, 128, 256, 4, 2, 1, 3), (6, 32, 64, 128, 4, 2, 1, 3), (4, 16, 640, 320, 3, 1, 1, 1),
                                  (24, 32, 48, 64, 1, 1, 0, 1)])
def test_conv_dynamic_schedule(case):
    """The persistent conv kernel with work items drawn from a global counter (option dyn_sched; off by default) computes what
    the static schedule does: many-tile and few-tile launches, CTA pairs, fused statistics, and the counter re-arms itself."""
    N, H, Ci, Co, k, s, p, G = case
    Ho = (H + 2 * p - k) // s + 1
    ops = _ops("bf16")
    ops.set_option("dyn_sched", 1)
    try:
        for rep in range(3):                                              # the same counter slot pool, launch after launch
            x, w = rnd(N, H, H, Ci, seed=rep), rnd(Co, Ci, k, k, scale=(Ci * k * k) ** -0.5, seed=rep)
            pf, pd = w.permute(0, 2, 3, 1).contiguous(), w.permute(1, 2, 3, 0).contiguous()
            run_pair("bf16", "conv_fprop", [T(x), T(pf), F(rnd(Co)), T(torch.zeros(N, Ho, Ho, Co)), k, s, p], [3], dict(act=ACT_LRELU))
            dy = rnd(N, Ho, Ho, Co, seed=rep + 7)
            run_pair("bf16", "conv_dgrad", [T(dy), T(pd), None, T(torch.zeros(N, H, H, Ci)), k, s, p], [3])
            st0 = torch.zeros(G, Co, 2, dtype=torch.float64)
            ea, ca = run_pair("bf16", "conv_fprop_stats", [T(x), T(pf), T(torch.zeros(N, Ho, Ho, Co)), D(st0), G, k, s, p], [2])
            _check_stats(ca[2], ca[3], st0, G)
    finally:
        ops.set_option("dyn_sched", 0)
