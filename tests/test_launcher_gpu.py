"""The train.py-equivalent launcher and the checkpoint hand-off (f2): Stage-I for a few synthetic
batches -> checkpoint with the reference's keys -> Stage-II picks the frozen Stage-I nets up from it -> its own
checkpoint -> resuming restores the epoch counter and the weights."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_stage1_then_stage2_with_checkpoints(tmp_path):
    from imagegenerator_b200 import train as T
    dev = torch.device("cuda:0")
    logs = []
    save = str(tmp_path)
    eng1, m1 = T.run(1, dev, epochs=1, batch=8, save_dir=save, synthetic=3, log=logs.append)
    ck = torch.load(os.path.join(save, "Stage1", "latest_checkpoint_stage1.pth"), map_location="cpu", weights_only=False)
    want = {"textEncoder", "projection_head", "con_augment_1", "critic_1", "gen_1", "opt_encoder", "opt_projection_head",
            "opt_con_augment_1", "opt_critic_1", "opt_gen_1", "lr_scheduler_encoder", "lr_scheduler_projection_head",
            "lr_scheduler_con_augment_1", "lr_scheduler_critic_1", "lr_scheduler_gen_1", "epoch"}     # stage_1_train_fn.py:212-229
    assert set(ck) == want and ck["epoch"] == 0
    assert ck["opt_critic_1"]["state"][0]["exp_avg"].abs().sum() > 0          # fused Adam state exported like optim.Adam's
    assert int(ck["critic_1"]["down_sampler.2.1.num_batches_tracked"]) == 3 * 21
    assert len(logs) == 3 and all("Loss D" in l for l in logs)
    g1_trained = {k: v.clone() for k, v in m1["gen_1"].state_dict().items()}

    eng2, m2 = T.run(2, dev, epochs=1, batch=2, save_dir=save, synthetic=2, log=logs.append, preview_every=1)
    # the fixed-noise preview + scalars of stage_2_train_fn.py:181-212 (written for batch 1; batch 0 is skipped like :175)
    pv = torch.load(os.path.join(save, "Stage2", "previews", "preview_000000.pt"), weights_only=False)
    assert pv["fake_256"].shape == (3, 256, 256) and float(pv["fake_256"].min()) >= 0.0 and float(pv["fake_256"].max()) <= 1.0
    assert torch.isfinite(pv["fake_256"]).all() and pv["batch"] == 1
    assert open(os.path.join(save, "Stage2", "previews", "scalars.csv")).read().count("\n") == 1
    for k, v in m2["gen_1"].state_dict().items():                              # frozen Stage-I generator = the trained one
        assert torch.equal(v.cpu(), g1_trained[k].cpu()), k
    ck2 = torch.load(os.path.join(save, "Stage2", "latest_checkpoint_stage2.pth"), map_location="cpu", weights_only=False)
    assert {"con_augment_2", "critic_2", "gen_2", "epoch"} <= set(ck2)          # stage_2_train_fn.py:214-228
    assert all(torch.isfinite(v).all() for v in ck2["gen_2"].values() if v.is_floating_point())
    # resume: nothing left to do for epochs=1, weights come back from the checkpoint
    eng3, m3 = T.run(2, dev, epochs=1, batch=2, save_dir=save, synthetic=1, log=logs.append)
    assert any("Loaded checkpoint" in l for l in logs)
    for k, v in m3["gen_2"].state_dict().items():
        assert torch.equal(v.cpu(), ck2["gen_2"][k]), k
