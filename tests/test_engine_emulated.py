"""Host logic of the fused train step (imagegenerator_b200/engine.py) on CPU.

The engine is driven through tests/emu_ops.py -- a torch-CPU statement of each kernel -- in fp64
and one whole Stage-I outer step (5 critic updates with the hand-derived WGAN-GP second-order
backward + generator/CA update + Adam) is compared with the autograd oracle.  This pins the
orchestration and the hand-derived math without a GPU; the CUDA kernels are then checked against
the same emulator methods one by one in tests/test_kernels_gpu.py."""
import torch
import pytest

from oracle import stackgan_oracle as O
from emu_ops import EmuOps
from imagegenerator_b200.con_augment import ConditioningAugmentation
from imagegenerator_b200.discrminator_1 import StageIDiscriminator
from imagegenerator_b200.generator_1 import StageIGenerator
from imagegenerator_b200.engine import Stage1Engine


def build_modules(seed=42):
    torch.manual_seed(seed)
    ca = ConditioningAugmentation(512, 256, 128)
    d1 = StageIDiscriminator(512, 128)
    g1 = StageIGenerator(128, 100)
    return ca, d1, g1


def test_state_dict_layout_and_default_init_equal_reference():
    ca, d1, g1 = build_modules()
    ps = O.init_all(42, with_stage2=False)
    for m, p in ((ca, ps["con_augment_1"]), (d1, ps["critic_1"]), (g1, ps["gen_1"])):
        sd = m.state_dict()
        assert list(sd.keys()) == list(p.keys())
        for k in sd:
            assert sd[k].shape == p[k].shape, k
            assert torch.equal(sd[k], p[k]), k


def _close(a, b, rtol, atol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if not torch.allclose(a, b, rtol=rtol, atol=atol):
        err = (a - b).abs().max().item()
        raise AssertionError(f"{what}: max abs err {err:.3e}, ref max {b.abs().max().item():.3e}")


@pytest.mark.parametrize("B", [4])
def test_stage1_outer_step_fp64_matches_oracle(B):
    dt = torch.float64
    ca, d1, g1 = build_modules()
    ps = O.init_all(42, with_stage2=False)
    pca, pd1, pg1 = (O.to_dtype(ps[k], dt) for k in ("con_augment_1", "critic_1", "gen_1"))
    b = O.synthetic_batch(B, 1, 0, dtype=dt)
    tr = dict(ca=O.Trainer(pca), d1=O.Trainer(pd1), g1=O.Trainer(pg1))
    tem = b["tem"].clone().requires_grad_(True)
    ref = O.stage1_step(pca, pd1, pg1, b["real"], tem, b["perm"], b["z"], b["eps_ca"], b["eps_gp"], tr)

    ops = EmuOps(dt)
    eng = Stage1Engine(ca, d1, g1, B, ops=ops)
    eng.load_batch(b["real"], b["tem"], b["tem"][b["perm"]])
    grads = []
    for it in range(5):
        eng.critic_iteration(b["z"][it], b["eps_ca"][it], b["eps_gp"][it])
        # the grads Adam just consumed are still in the flat buffer
        grads.append({k: v.grad.clone() for k, v in d1.named_parameters()})
        assert abs(eng.losses[0].item() - ref["loss_critic"][it].item()) < 1e-8 * max(1, abs(ref["loss_critic"][it].item()))
        if it == 0:
            _close(eng.d.score[0], ref["first"]["s_real"], 1e-9, 1e-10, "s_real")
            _close(eng.d.score[1], ref["first"]["s_mis"], 1e-9, 1e-10, "s_mis")
            _close(eng.d.score[2], ref["first"]["s_fake"], 1e-9, 1e-10, "s_fake")
            _close(eng.losses[1], ref["first"]["gp"], 1e-9, 1e-12, "gp")
            fake = eng.d.group_view(eng.d.a[0], 1, 1).permute(0, 3, 1, 2)
            _close(fake, ref["first"]["fake"], 1e-9, 1e-10, "fake")
    for it in range(5):
        for k, g in grads[it].items():
            _close(g, ref["critic_grads"][it][k], 1e-6, 1e-10, f"critic grad it{it} {k}")
    eng.generator_step()
    assert abs(eng.losses[2].item() - ref["lossG"].item()) < 1e-8 * abs(ref["lossG"].item())
    for k, v in g1.named_parameters():
        _close(v.grad, ref["g1_grads"][k], 1e-6, 1e-10, f"g1 grad {k}")
    for k, v in ca.named_parameters():
        _close(v.grad, ref["ca_grads"][k], 1e-6, 1e-9, f"ca grad {k}")
    _close(eng.d.dtem, ref["dtem"], 1e-6, 1e-10, "dtem")
    for m, key in ((ca, "ca"), (d1, "d1"), (g1, "g1")):
        sd = m.state_dict()
        for k, v in ref["after"][key].items():
            if v.is_floating_point():
                _close(sd[k], v, 1e-6, 1e-9, f"after {key}.{k}")
            else:
                assert int(sd[k]) == int(v), (key, k, int(sd[k]), int(v))


def test_outer_step_off_chain_bookkeeping_equals_the_plain_iterations():
    """outer_step runs the critic updates with the main-chain shortcuts -- gradient buffers, the per-channel sums of all three
    passes and the head seeds prepared behind the PREVIOUS optimizer step (grads_zeroed / zero_after / prezeroed / seeded), the
    generator's gradients cleared at the start of the step -- and the reductions ADD to what was zeroed.  Two outer steps must
    leave exactly the parameters that plain critic_iteration / generator_step calls (each zeroing for itself) leave."""
    B, dt = 4, torch.float64
    params = []
    for shortcut in (False, True):
        ca, d1, g1 = build_modules()
        eng = Stage1Engine(ca, d1, g1, B, ops=EmuOps(dt))
        for step in range(2):
            b = O.synthetic_batch(B, 1, step, dtype=dt)
            eng.load_batch(b["real"], b["tem"], b["tem"][b["perm"]])
            if shortcut:
                eng.outer_step(b["z"], b["eps_ca"], b["eps_gp"])
            else:
                eng._ce_ready = False
                for it in range(5):
                    eng.critic_iteration(b["z"][it], b["eps_ca"][it], b["eps_gp"][it])
                eng.generator_step()
        params.append({f"{n}.{k}": v.detach().clone() for n, m in (("ca", ca), ("d1", d1), ("g1", g1))
                       for k, v in m.state_dict().items()})
    for k, v in params[0].items():
        if v.is_floating_point():
            assert torch.allclose(params[1][k].double(), v.double(), rtol=1e-7, atol=1e-9), k   # fp32 masters: a few ulps from the order of += terms
        else:
            assert int(params[1][k]) == int(v), k


def test_compressed_text_is_recomputed_for_every_new_batch():
    """The critic's compressed text is computed on the weight re-pack stream after each critic update and reused by the
    following forwards -- but never across batches: the first critic forward of every outer step (and of every
    load_batch) must compress the new text itself."""
    B = 2
    ca, d1, g1 = build_modules()
    eng = Stage1Engine(ca, d1, g1, B, ops=EmuOps(torch.float64))
    seen = []
    orig = eng.d.forward

    def spy(*a, **k):
        seen.append(bool(k.get("ce_ready", False)))
        return orig(*a, **k)
    eng.d.forward = spy
    for step in range(2):
        b = O.synthetic_batch(B, 1, step, dtype=torch.float64)
        eng.load_batch(b["real"], b["tem"], b["tem"][b["perm"]])
        if step == 1:                                    # the graph path writes the text buffer directly, then outer_step
            eng.d.tem_all[:B].copy_(b["tem"])
        eng.outer_step(b["z"], b["eps_ca"], b["eps_gp"])
        # what the last forward used == compress(current text) with the current weights
        want = b["tem"] @ d1.compress.weight.data.double().t() + d1.compress.bias.data.double()
        assert torch.allclose(eng.d.ce[:B], want, rtol=1e-9, atol=1e-12)
    assert seen == [False, True, True, True, True, True] * 2, seen


def test_optimizer_state_round_trip_through_torch_adam():
    """Checkpoints carry the fused Adam's moments as a torch.optim.Adam state_dict (stage_1_train_fn.py:218-222) and a
    resumed run takes them over again (:69-73): export -> state_dict -> load_state_dict -> import gives a replica that
    continues exactly like the original."""
    from imagegenerator_b200.engine import export_optimizer_state, import_optimizer_state
    dt, B = torch.float64, 2
    b = O.synthetic_batch(B, 1, 0, dtype=dt)

    def fresh():
        ca, d1, g1 = build_modules()
        return ca, d1, g1

    ca, d1, g1 = fresh()
    eng = Stage1Engine(ca, d1, g1, B, ops=EmuOps(dt))
    eng.load_batch(b["real"], b["tem"], b["tem"][b["perm"]])
    eng.critic_iteration(b["z"][0], b["eps_ca"][0], b["eps_gp"][0])
    eng.critic_iteration(b["z"][1], b["eps_ca"][1], b["eps_gp"][1])
    opt = torch.optim.Adam(d1.parameters(), lr=1e-3, betas=(0.9, 0.999))
    export_optimizer_state(opt, eng.d.fp)
    ck = {"critic_1": {k: v.clone() for k, v in d1.state_dict().items()}, "opt_critic_1": opt.state_dict()}
    st = ck["opt_critic_1"]["state"]
    assert len(st) == len(list(d1.parameters())) and float(st[0]["step"]) == 2.0
    assert set(st[0]) == {"step", "exp_avg", "exp_avg_sq"} and st[0]["exp_avg"].shape == d1.down_sampler[0].weight.shape

    ca2, d2, g2 = fresh()
    d2.double()                    # (load_state_dict casts the moments to the parameters' type: fp32 in the product, where
    d2.load_state_dict(ck["critic_1"])         # the fused Adam's moments are fp32 too; the emulator runs in fp64)
    opt2 = torch.optim.Adam(d2.parameters(), lr=1e-3, betas=(0.9, 0.999))
    opt2.load_state_dict(ck["opt_critic_1"])
    eng2 = Stage1Engine(ca2, d2, g2, B, ops=EmuOps(dt))
    assert import_optimizer_state(opt2, eng2.d.fp) == len(list(d2.parameters()))
    assert torch.equal(eng2.d.fp.m, eng.d.fp.m) and torch.equal(eng2.d.fp.v, eng.d.fp.v)
    assert float(eng2.d.fp.hyper[4]) == float(eng.d.fp.hyper[4]) == 2.0
    # an optimizer without state (a fresh run, or a checkpoint written before the state was exported) changes nothing
    assert import_optimizer_state(torch.optim.Adam(g2.parameters()), eng2.g.fp) == 0 and float(eng2.g.fp.hyper[4]) == 0.0
    eng2.load_batch(b["real"], b["tem"], b["tem"][b["perm"]])
    for e in (eng, eng2):
        e._ce_ready = False
        e.critic_iteration(b["z"][2], b["eps_ca"][2], b["eps_gp"][2])
    for (k, v), (_, w) in zip(d1.state_dict().items(), d2.state_dict().items()):
        if v.is_floating_point():
            assert torch.allclose(v, w, rtol=1e-10, atol=1e-12), k


def test_one_flat_buffer_per_module_shared_by_every_runtime():
    """a second runtime on a module a live engine owns must not re-point p.data / p.grad away from the buffers the
    engine's Adam, all-reduce and captured graph use -- every runtime shares the module's one FlatParams."""
    from emu_ops import EmuOps
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.discrminator_1 import StageIDiscriminator
    from imagegenerator_b200.generator_1 import StageIGenerator
    from imagegenerator_b200.engine import CART, CriticRT, GenRT, Stage1Engine
    torch.manual_seed(0)
    ca, d, g = ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)
    ops = EmuOps(torch.float64)
    eng = Stage1Engine(ca, d, g, 2, ops=ops)
    w = d.down_sampler[2][0].weight
    ptr, gptr = w.data.data_ptr(), w.grad.data_ptr()
    rt2 = d.runtime(4, ops)                      # what critic.forward / utils.gradient_penalty build
    g2 = GenRT(ops, g, 4)
    c2 = CART(ops, ca)
    c2.ensure(4)
    assert rt2.fp is eng.d.fp and g2.fp is eng.g.fp and c2.fp is eng.ca.fp
    assert w.data.data_ptr() == ptr and w.grad.data_ptr() == gptr
    assert w.data.data_ptr() >= eng.d.fp.flat.data_ptr() and w.data.data_ptr() < eng.d.fp.flat.data_ptr() + eng.d.fp.flat.numel() * 8
    # an ops object of another precision on the same module is refused instead of silently detaching the engine
    with pytest.raises(RuntimeError, match="already lives in a flat parameter buffer"):
        CriticRT(EmuOps(torch.float32), d, 2)


def test_module_forward_refuses_autograd_with_a_clear_message():
    from imagegenerator_b200.layers import no_autograd
    from imagegenerator_b200.generator_1 import StageIGenerator
    m = StageIGenerator(128, 100)
    out = no_autograd(torch.zeros(2, 3), m)
    assert out.requires_grad
    with pytest.raises(RuntimeError, match="does not record an autograd graph"):
        out.sum().backward()
    with torch.no_grad():
        assert not no_autograd(torch.zeros(2, 3), m).requires_grad
