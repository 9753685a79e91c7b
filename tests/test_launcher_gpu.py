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


def test_checkpoints_cross_the_boundary_to_and_from_the_unmodified_reference(tmp_path):
    """f2, both directions, against the REAL reference (oracle/ref_harness.py runs it on CPU under stubs):

    1. the unmodified ``train_1`` (stage_1_train_fn.py:19-238) trains one epoch and writes its checkpoint into the (fake)
       bucket; those bytes, dropped where this package's ``train_1`` looks, must resume it: epoch counter, every weight and
       BatchNorm buffer bit-identical, the fused Adam's moments and step counts taken over from the reference's
       ``torch.optim.Adam`` state (:55-82), and training continues from there with finite losses;
    2. the checkpoint THIS package then writes must load into the reference's own modules and optimizers with strict
       ``load_state_dict`` (:63-73) -- a user can go back."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip("neither /root/reference nor oracle/_ref/reference_src.zip is here")
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.discrminator_1 import StageIDiscriminator
    from imagegenerator_b200.engine import Stage1Engine
    from imagegenerator_b200.generator_1 import StageIGenerator
    from imagegenerator_b200.ops import CudaOps
    from imagegenerator_b200.stage_1_train_fn import train_1

    B, dev = 4, torch.device("cuda:0")
    mk = lambda m, lr=1e-3: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.999))
    sched = lambda opts: [torch.optim.lr_scheduler.StepLR(o, step_size=100, gamma=0.5) for o in opts]
    table = torch.randn(B, 512, generator=torch.Generator().manual_seed(3))
    real = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(4)).clamp_(-1, 1)
    loader = [({"idx": torch.arange(B)}, real), ({"idx": torch.arange(B)}, real.flip(0))]

    # ---- 1. the reference writes
    H.reset_store()
    torch.manual_seed(11)
    r_ca = H.load("con_augment").ConditioningAugmentation(512, 256, 128)
    r_d1 = H.load("discrminator_1").StageIDiscriminator(512, 128)
    r_g1 = H.load("generator_1").StageIGenerator(128, 100)
    r_enc, r_head = H.TableEncoder(table), H.IdentityHead()
    r_opts = [mk(r_enc, 0.0), mk(r_head, 0.0), mk(r_ca), mk(r_d1), mk(r_g1)]
    with H.quiet():
        H.load("stage_1_train_fn").train_1([r_enc, r_head, r_ca, r_d1, r_g1], r_opts, sched(r_opts), loader, 1, "cpu", B)
    blob = H.store()["./checkpoints/Stage1/latest_checkpoint_stage1.pth"]
    save = str(tmp_path / "Stage1")
    os.makedirs(save)
    with open(os.path.join(save, "latest_checkpoint_stage1.pth"), "wb") as f:
        f.write(blob)
    ck = torch.load(os.path.join(save, "latest_checkpoint_stage1.pth"), map_location="cpu", weights_only=False)
    assert ck["epoch"] == 0 and float(ck["opt_critic_1"]["state"][0]["step"]) == 10.0      # 2 batches x 5 critic steps

    # ---- ... and this package resumes from it (differently initialised modules: everything must come from the file)
    torch.manual_seed(99)
    ca, d1, g1 = ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)
    enc, head = H.TableEncoder(torch.zeros(B, 512)).to(dev), H.IdentityHead().to(dev)
    opts = [mk(enc, 0.0), mk(head, 0.0), mk(ca), mk(d1), mk(g1)]
    eng = Stage1Engine(ca, d1, g1, B, ops=CudaOps("fp32"))
    logs = []
    train_1([enc, head, ca, d1, g1], opts, sched(opts), loader, 1, dev, B, save_dir=save, log=logs.append, engine=eng,
            use_graph=False)                                    # epochs=1: resume only, nothing left to train
    assert any("Loaded checkpoint at epoch 0" in l for l in logs)
    for ours, key in ((ca, "con_augment_1"), (d1, "critic_1"), (g1, "gen_1"), (enc, "textEncoder")):
        for k, v in ours.state_dict().items():
            assert torch.equal(v.cpu(), ck[key][k]), (key, k)
    for fp, key, steps in ((eng.d.fp, "opt_critic_1", 10.0), (eng.g.fp, "opt_gen_1", 2.0), (eng.ca.fp, "opt_con_augment_1", 2.0)):
        st = ck[key]["state"]
        assert float(fp.hyper[4]) == steps, (key, float(fp.hyper[4]))
        m = torch.cat([st[i]["exp_avg"].reshape(-1) for i in range(len(st))])
        v = torch.cat([st[i]["exp_avg_sq"].reshape(-1) for i in range(len(st))])
        assert torch.equal(fp.m.cpu()[:m.numel()], m) and torch.equal(fp.v.cpu()[:v.numel()], v), key
    # the packed bf16/fp32 operands were refreshed from the loaded weights: one more epoch trains on from here
    w_before = d1.down_sampler[2][0].weight.detach().clone()
    train_1([enc, head, ca, d1, g1], opts, sched(opts), loader, 11, dev, B, save_dir=save, log=logs.append, engine=eng,
            use_graph=False)                                    # epochs 1..10; epoch 10 writes a checkpoint (:211)
    assert sum("Loss D" in l for l in logs) == 20 and all("nan" not in l.lower() for l in logs)
    assert not torch.equal(w_before, d1.down_sampler[2][0].weight.detach())

    # ---- 2. this package writes, the reference reads
    ck2 = torch.load(os.path.join(save, "latest_checkpoint_stage1.pth"), map_location="cpu", weights_only=False)
    assert ck2["epoch"] == 10 and set(ck2) == set(ck)
    torch.manual_seed(5)
    r_ca2 = H.load("con_augment").ConditioningAugmentation(512, 256, 128)
    r_d12 = H.load("discrminator_1").StageIDiscriminator(512, 128)
    r_g12 = H.load("generator_1").StageIGenerator(128, 100)
    r_ca2.load_state_dict(ck2["con_augment_1"])                 # strict
    r_d12.load_state_dict(ck2["critic_1"])
    r_g12.load_state_dict(ck2["gen_1"])
    for m, key in ((r_ca2, "opt_con_augment_1"), (r_d12, "opt_critic_1"), (r_g12, "opt_gen_1")):
        o = mk(m)
        o.load_state_dict(ck2[key])
        st = o.state_dict()["state"]
        assert len(st) == len(list(m.parameters())) and float(st[0]["step"]) == float(ck2[key]["state"][0]["step"])
    assert float(ck2["opt_critic_1"]["state"][0]["step"]) == 10.0 + 10 * 2 * 5
    assert int(ck2["critic_1"]["down_sampler.2.1.num_batches_tracked"]) == 11 * 2 * 21       # :125-154: 21 BN passes per batch
