"""The captured step (CUDA graph, parameter-gradient kernels on a side stream, next fake batch generated ahead on a
third stream) must compute what the serial eager step computes: a race between the forked streams would show up as
a difference in the gradients the optimizers saw.  Only the order of floating-point atomics differs between the two."""
import os

import pytest
import torch

from test_engine2_emulated import build_all

pytestmark = pytest.mark.gpu


def _inputs(B, hw, stage2):
    g = torch.Generator().manual_seed(11)
    real = torch.randn(B, 3, hw, hw, generator=g).clamp_(-1, 1).cuda()
    tem = torch.randn(B, 512, generator=g).cuda()
    tem_mis = tem[torch.randperm(B, generator=g).cuda()].contiguous()
    z = torch.randn(5, B, 100, generator=g).cuda()
    e1, e2 = torch.randn(5, B, 128, generator=g).cuda(), torch.randn(5, B, 128, generator=g).cuda()
    egp = torch.rand(5, B, generator=g).cuda()
    return (real, tem, tem_mis, z, e1, e2, egp) if stage2 else (real, tem, tem_mis, z, e1, egp)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("stage", [1, 2])
def test_concurrent_graph_step_equals_serial_eager(stage, monkeypatch):
    from imagegenerator_b200.ops import CudaOps
    from imagegenerator_b200.engine import Stage1Engine
    from imagegenerator_b200.engine2 import Stage2Engine
    B = 16 if stage == 1 else 4
    outs = []
    for serial in (True, False):
        if serial:
            monkeypatch.setenv("SG_NO_SIDE_STREAM", "1")
        else:
            monkeypatch.delenv("SG_NO_SIDE_STREAM", raising=False)
        ms = build_all()
        ops = CudaOps("fp32")                         # fp32 storage: no rounding kinks, differences are atomics order only
        if stage == 1:
            eng = Stage1Engine(ms["ca1"], ms["d1"], ms["g1"], B, ops=ops)
            inp = _inputs(B, 64, False)
            mods = (ms["ca1"], ms["d1"], ms["g1"])
        else:
            eng = Stage2Engine(ms["ca1"], ms["g1"], ms["ca2"], ms["d2"], ms["g2"], B, ops=ops)
            inp = _inputs(B, 256, True)
            mods = (ms["ca2"], ms["d2"], ms["g2"])
        assert eng.side.enabled == (not serial)
        fps = (eng.ca.fp, eng.d.fp, eng.g.fp) if stage == 1 else (eng.ca2.fp, eng.d.fp, eng.g2.fp)
        for fp in fps:
            fp.set_lr(0.0)        # frozen weights: every gradient is a pure function of the inputs, no Adam chaos in between
        # the critic's LAST gradients and the generator's accumulated ones are still in the flat buffers after the step
        eng.step(*inp, use_graph=not serial)
        if not serial:
            eng.step(*inp, use_graph=True)            # second replay: same static buffers, advanced weights
        torch.cuda.synchronize()
        if serial:
            eng.step(*inp, use_graph=False)
            torch.cuda.synchronize()
        if stage == 2:
            eng.sync_grads()
        bufs = [m.state_dict()[k].clone() for m in mods for k in m.state_dict() if "running" in k]
        outs.append((eng.losses.clone(), [fp.grad.clone() for fp in fps[:2]], bufs))
    (l0, g0, b0), (l1, g1, b1) = outs
    assert torch.allclose(l0, l1, rtol=1e-4, atol=1e-5), (l0.tolist(), l1.tolist())
    # CA / critic gradient buffers of the last backward pass (the generator's are cleared after its step in Stage-II)
    worst = max(_rel(a, b) for a, b in zip(g1, g0))
    assert worst < 1e-4, worst
    assert all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(b1, b0))
