"""Host logic of the Stage-II step (imagegenerator_b200/engine2.py) on CPU through the kernel emulator,
fp64, against the autograd oracle -- including the reference's quirk that G2/CA2 gradients accumulate
over the five critic backward passes (stage_2_train_fn.py:131,154,163-168)."""
import pytest
import torch

from oracle import stackgan_oracle as O
from emu_ops import EmuOps
from imagegenerator_b200.con_augment import ConditioningAugmentation
from imagegenerator_b200.discrminator_1 import StageIDiscriminator
from imagegenerator_b200.discriminator_2 import StageIIDiscriminator
from imagegenerator_b200.generator_1 import StageIGenerator
from imagegenerator_b200.generator_2 import StageIIGenerator
from imagegenerator_b200.engine2 import Stage2Engine


def build_all(seed=42):
    torch.manual_seed(seed)
    return dict(ca1=ConditioningAugmentation(512, 256, 128), d1=StageIDiscriminator(512, 128), g1=StageIGenerator(128, 100),
                ca2=ConditioningAugmentation(512, 256, 128), d2=StageIIDiscriminator(512, 128), g2=StageIIGenerator())


def _close(a, b, rtol, atol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if not torch.allclose(a, b, rtol=rtol, atol=atol):
        raise AssertionError(f"{what}: max abs err {(a - b).abs().max().item():.3e}, ref max {b.abs().max().item():.3e}")


@pytest.mark.slow
def test_stage2_outer_step_fp64_matches_oracle():
    dt, B = torch.float64, 2
    ms = build_all()
    ps = O.init_all(42)
    p = {k: O.to_dtype(ps[k], dt) for k in ps}
    # give the frozen Stage-I generator non-trivial running statistics (eval-mode BN uses them)
    g = torch.Generator().manual_seed(3)
    for k, v in p["gen_1"].items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g, dtype=dt) * 0.1)
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g, dtype=dt) + 0.5)
    ms["g1"].load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in p["gen_1"].items()})
    b = O.synthetic_batch(B, 2, 0, dtype=dt)
    tr = dict(ca2=O.Trainer(p["con_augment_2"]), d2=O.Trainer(p["critic_2"]), g2=O.Trainer(p["gen_2"]))
    ref = O.stage2_step(p["con_augment_1"], p["gen_1"], p["con_augment_2"], p["critic_2"], p["gen_2"], b["real"], b["tem"],
                        b["perm"], b["z"], b["eps_ca"], b["eps_ca2"], b["eps_gp"], tr)

    ops = EmuOps(dt)
    # emulator keeps parameters in fp64: reload exact fp64 copies of the frozen nets' buffers
    eng = Stage2Engine(ms["ca1"], ms["g1"], ms["ca2"], ms["d2"], ms["g2"], B, ops=ops)
    ms["g1"].load_state_dict(p["gen_1"])
    eng.g1.refresh_weights()
    eng.load_batch(b["real"], b["tem"], b["tem"][b["perm"]])
    for it in range(5):
        eng.critic_iteration(b["z"][it], b["eps_ca"][it], b["eps_ca2"][it], b["eps_gp"][it])
        lc = ref["loss_critic"][it].item()
        assert abs(eng.losses[0].item() - lc) < 1e-7 * max(1, abs(lc)), (it, eng.losses[0].item(), lc)
        for k, v in ms["d2"].named_parameters():
            _close(v.grad, ref["critic_grads"][it][k], 1e-5, 1e-9, f"critic2 grad it{it} {k}")
        if it == 0:
            _close(eng.g1.out.permute(0, 3, 1, 2), ref["first"]["fake_64"], 1e-8, 1e-9, "fake_64")
            _close(eng.g2.out.permute(0, 3, 1, 2), ref["first"]["fake"], 1e-8, 1e-9, "fake_256")
    # gradients G2/CA2 are stepped with (accumulated over the 5 critic backwards + lossG)
    d, ops_ = eng.d, eng.ops
    d.forward(1, 1, dup_first=1, training=True)
    ops_.gen_loss(d.score[2], eng.ca2.st.mu, eng.ca2.st.sigma, eng.losses[2:4])
    d.backward(1, 1, d.coef_gen, inject=False, param_grads=False, need_input_grad=True)
    eng._generator_backward(d.group_view(d.dx, 1, 1), 1.0)
    eng.sync_grads()
    assert abs(eng.losses[2].item() - ref["lossG"].item()) < 1e-8 * abs(ref["lossG"].item())
    for k, v in ms["g2"].named_parameters():
        _close(v.grad, ref["g2_grads"][k], 1e-5, 1e-8, f"g2 grad {k}")
    for k, v in ms["ca2"].named_parameters():
        _close(v.grad, ref["ca2_grads"][k], 1e-5, 1e-7, f"ca2 grad {k}")
    eng.optimizer_step(eng.g2.fp)
    eng.optimizer_step(eng.ca2.fp)
    for m, key in ((ms["ca2"], "ca2"), (ms["d2"], "d2"), (ms["g2"], "g2")):
        sd = m.state_dict()
        for k, v in ref["after"][key].items():
            if v.is_floating_point():
                _close(sd[k], v, 1e-6, 1e-8, f"after {key}.{k}")
            else:
                assert int(sd[k]) == int(v), (key, k)
