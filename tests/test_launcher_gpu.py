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
    # resume (:55-82): weights, schedulers AND optimizer state come back -- the fused Adam continues at step 15 / 3
    eng1b, _ = T.run(1, dev, epochs=2, batch=8, save_dir=save, synthetic=3, log=logs.append)
    assert any("Loaded checkpoint at epoch 0" in l for l in logs)
    assert float(eng1b.d.fp.hyper[4]) == 30.0 and float(eng1b.g.fp.hyper[4]) == 6.0 and float(eng1b.ca.fp.hyper[4]) == 6.0

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
    st = ck2["opt_gen_2"]["state"]                                             # the fused Adam's moments, exported (:219-221)
    assert len(st) == len(list(m2["gen_2"].parameters())) and float(st[0]["step"]) == 2.0 and st[0]["exp_avg"].abs().sum() > 0
    assert float(ck2["opt_critic_2"]["state"][0]["step"]) == 10.0
    # resume: nothing left to do for epochs=1, weights come back from the checkpoint
    eng3, m3 = T.run(2, dev, epochs=1, batch=2, save_dir=save, synthetic=1, log=logs.append)
    assert any("Loaded checkpoint" in l for l in logs)
    for k, v in m3["gen_2"].state_dict().items():
        assert torch.equal(v.cpu(), ck2["gen_2"][k]), k
    assert float(eng3.g2.fp.hyper[4]) == 2.0 and float(eng3.d.fp.hyper[4]) == 10.0   # optimizer state restored (:84-86)
    assert torch.equal(eng3.g2.fp.m[:st[0]["exp_avg"].numel()].cpu(), st[0]["exp_avg"].reshape(-1))


def test_stage1_on_captions_with_a_bert_encoder(tmp_path):
    """f3: the reference's real pipeline shape -- COCO-style files -> tokenizer -> BERT CLS state -> projection head ->
    train_1 -- with the encoder and the head trained through d lossG / d tem (stage_1_train_fn.py:117-119, :161-171)."""
    from _util import make_coco_dir, tiny_bert
    from imagegenerator_b200 import train as T
    from imagegenerator_b200.data_loader import get_loader
    dev = torch.device("cuda:0")
    root, ann, tok, _ = make_coco_dir(tmp_path, n_images=6, captions_per_image=2)
    loader = get_loader("local", root, ann, T.image_transform(64), batch_size=4, shuffle=True, tokenizer=tok,
                        num_workers=2, prefetch_factor=2)
    enc = tiny_bert(tok)
    before = {k: v.clone() for k, v in enc.state_dict().items()}
    logs = []
    save = str(tmp_path / "ck")
    eng, m = T.run(1, dev, epochs=1, batch=4, loader=loader, text_encoder=enc, save_dir=save, log=logs.append)
    assert len(logs) == 3 and all("Loss D" in l for l in logs)               # 12 captions / 4, drop_last
    assert torch.isfinite(eng.losses).all()
    after = m["textEncoder"].state_dict()
    moved = [k for k, v in before.items() if v.is_floating_point() and not torch.equal(v, after[k].cpu())]
    assert "embeddings.word_embeddings.weight" in moved and len(moved) > 10  # AdamW stepped on the encoder
    ck = torch.load(os.path.join(save, "Stage1", "latest_checkpoint_stage1.pth"), map_location="cpu", weights_only=False)
    assert set(ck["textEncoder"]) == set(before)
    # Stage-II reads the same loader items at 256x256 and keeps the text side frozen (stage_2_train_fn.py:52-63)
    loader2 = get_loader("local", root, ann, T.image_transform(256), batch_size=2, shuffle=True, tokenizer=tok,
                         num_workers=0)
    enc2 = tiny_bert(tok)
    eng2, m2 = T.run(2, dev, epochs=1, batch=2, loader=loader2, text_encoder=enc2, save_dir=save, log=logs.append,
                     preview_every=10 ** 9)
    assert torch.isfinite(eng2.losses).all()
    for k, v in m2["textEncoder"].state_dict().items():                       # restored from the Stage-I checkpoint
        assert torch.equal(v.cpu(), ck["textEncoder"][k]), k
