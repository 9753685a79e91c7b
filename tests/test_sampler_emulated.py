"""Forward-only sampling path (imagegenerator_b200/sampler.py; reference stage_2_train_fn.py:181-195) on CPU through
the kernel emulator in fp64 against the oracle: BN folded into the packed weights == eval-mode BatchNorm."""
import pytest
import torch

from oracle import stackgan_oracle as O
from emu_ops import EmuOps
from test_engine2_emulated import build_all, _close
from imagegenerator_b200.sampler import StackGANSampler


def _randomise_running_stats(p, seed, dt):
    g = torch.Generator().manual_seed(seed)
    for k, v in p.items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g, dtype=dt) * 0.1)
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g, dtype=dt) + 0.5)
        if k.endswith("1.weight") and v.dim() == 1:             # BN gamma / beta away from (1, 0)
            v.copy_(1 + 0.2 * torch.randn(v.shape, generator=g, dtype=dt))
        if k.endswith("1.bias") and v.dim() == 1:
            v.copy_(0.1 * torch.randn(v.shape, generator=g, dtype=dt))


@pytest.mark.parametrize("batch_stats", [False, True])
def test_sampler_fp64_matches_oracle(batch_stats):
    dt, B = torch.float64, 2
    ms = build_all()
    ps = O.init_all(42)
    p = {k: O.to_dtype(ps[k], dt) for k in ps}
    _randomise_running_stats(p["gen_1"], 3, dt)
    _randomise_running_stats(p["gen_2"], 4, dt)
    for key, m in (("gen_1", "g1"), ("gen_2", "g2"), ("con_augment_1", "ca1"), ("con_augment_2", "ca2")):
        ms[m].double()
        ms[m].load_state_dict(p[key])
    g = torch.Generator().manual_seed(0)
    tem = torch.randn(B, 512, generator=g, dtype=dt)
    z, e1, e2 = (torch.randn(B, n, generator=g, dtype=dt) for n in (100, 128, 128))
    ref64, ref256 = O.sample(p["con_augment_1"], p["gen_1"], p["con_augment_2"], p["gen_2"], tem, z, e1, e2,
                             g2_training=batch_stats)
    smp = StackGANSampler(ms["ca1"], ms["g1"], ms["ca2"], ms["g2"], B, ops=EmuOps(dt), bn_batch_stats=batch_stats)
    f64, f256 = smp.sample(tem, z, e1, e2)
    _close(f64, ref64, 1e-9, 1e-10, "fake_64")
    _close(f256, ref256, 1e-8, 1e-9, "fake_256")
    assert f256.shape == (B, 3, 256, 256) and f64.shape == (B, 3, 64, 64)


def test_stage2_engine_preview_fp64_matches_oracle():
    """Stage2Engine.preview (the in-loop sample of stage_2_train_fn.py:181-195, gen_2 in train mode) on the emulator."""
    from imagegenerator_b200.engine2 import Stage2Engine
    dt, B = torch.float64, 2
    ms = build_all()
    ps = O.init_all(42)
    p = {k: O.to_dtype(ps[k], dt) for k in ps}
    _randomise_running_stats(p["gen_1"], 3, dt)
    eng = Stage2Engine(ms["ca1"], ms["g1"], ms["ca2"], ms["d2"], ms["g2"], B, ops=EmuOps(dt))
    for key, m in (("gen_1", "g1"), ("gen_2", "g2"), ("con_augment_1", "ca1"), ("con_augment_2", "ca2")):
        ms[m].load_state_dict(p[key])                      # exact fp64 copies into the engine's flat buffers
    eng.g1.refresh_weights()
    eng.g2.refresh_weights()
    g = torch.Generator().manual_seed(1)
    tem = torch.randn(B, 512, generator=g, dtype=dt)
    z, e1, e2 = (torch.randn(B, n, generator=g, dtype=dt) for n in (100, 128, 128))
    nbt0 = int(ms["g2"].state_dict()["up_sampler.0.1.num_batches_tracked"])
    ref64, ref256 = O.sample(p["con_augment_1"], p["gen_1"], p["con_augment_2"], p["gen_2"], tem, z, e1, e2, g2_training=True)
    f64, f256 = eng.preview(tem, z, e1, e2)
    _close(f64, ref64, 1e-9, 1e-10, "preview fake_64")
    _close(f256, ref256, 1e-8, 1e-9, "preview fake_256")
    # train-mode forward: gen_2's running statistics moved once (as in the reference), gen_1's did not
    assert int(ms["g2"].state_dict()["up_sampler.0.1.num_batches_tracked"]) == nbt0 + 1
    assert int(ms["g1"].state_dict()["upsampling.0.1.num_batches_tracked"]) == 0
