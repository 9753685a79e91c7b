"""Parity of the CUDA Stage-II train step (generator_2 / discriminator_2 / stage_2_train_fn path) with the fp64 oracle at a
mid-size batch (``pytest -m gpu``); the benchmarked batch 64 and the comparison policy are in
tests/test_parity_config_gpu.py.  The oracle runs on the GPU (tests/gpu_oracle.py); every iteration is compared -- forward
images, scores, GP, losses, the critic's gradients of all five iterations, what the first critic backward leaves in G2 / CA2
and the accumulated gradients G2 / CA2 are stepped with (stage_2_train_fn.py:131,154,163-168) -- then a free-running step
is checked on its losses."""
import pytest
import torch

import gpu_oracle as GO
from test_parity_config_gpu import REPORT_ONLY, _dump, _ref_table2, _run_tf2, compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_stage2_teacher_forced_and_losses(mode):
    from imagegenerator_b200.ops import CudaOps
    from emu_ops import EmuOps
    B = 16
    b, ref = GO.stage2(B, torch.float64)
    want = {k: (v.detach().double().cpu() if torch.is_tensor(v) else v) for k, v in _ref_table2(ref).items()}
    force = {"critic_before": ref["critic_before"]}
    loss_last, loss_g = ref["loss_critic"][-1].item(), ref["lossG"].item()
    del ref
    torch.cuda.empty_cache()
    if mode == "fp32":
        _, r32 = GO.stage2(B, torch.float32, force=force["critic_before"])
        yard = {k: (v.detach().double().cpu() if torch.is_tensor(v) else v) for k, v in _ref_table2(r32).items()}
        del r32
    else:
        yard = _run_tf2(EmuOps(torch.bfloat16, device="cuda"), b, force)
    torch.cuda.empty_cache()
    got = _run_tf2(CudaOps(mode), b, force)
    torch.cuda.synchronize()
    lines, fails = compare(mode, want, got, yard, f"stage2 {mode} B{B}", yard_small=(mode == "fp32"))
    _dump(f"parity_stage2_{mode}_B{B}.txt", lines)
    assert REPORT_ONLY or not fails, "\n".join(fails[:10])
    # the free-running step (no re-synchronisation): the losses of the last critic iteration and of the generator
    from imagegenerator_b200.engine2 import Stage2Engine
    from test_parity_config_gpu import _modules2
    ms = _modules2()
    ops = CudaOps(mode)
    eng = Stage2Engine(ms["ca1"], ms["g1"], ms["ca2"], ms["d2"], ms["g2"], B, ops=ops)
    dv = lambda t: t.to(ops.device).to(ops.f32).contiguous()
    eng.load_batch(dv(b["real"]), dv(b["tem"]), dv(b["tem"][b["perm"]]))
    eng.outer_step(dv(b["z"]), dv(b["eps_ca"]), dv(b["eps_ca2"]), dv(b["eps_gp"]))
    torch.cuda.synchronize()
    rt = 5e-2 if mode == "bf16" else 1e-2
    for key, got_v, r in (("loss_critic_last", eng.losses[0].item(), loss_last), ("lossG", eng.losses[2].item(), loss_g)):
        assert abs(got_v - r) <= rt * abs(r) + 1e-3, (key, got_v, r)
