"""The oracle restatement (oracle/stackgan_oracle.py) against fixtures produced by the REAL
reference (oracle/make_golden.py).  CPU only; runs everywhere."""
import torch
import pytest

from oracle import stackgan_oracle as O
from _util import load_golden, assert_digest, assert_digest_dict, digest

RT, AT = 1e-4, 1e-6      # same torch build on both sides: only thread-count reduction-order noise


def test_init_matches_reference_default_init():
    g = load_golden("modules")
    ps = O.init_all(42)
    for name, dgs in g["init_digest"].items():
        assert_digest_dict(ps[name], dgs, 0, 0, name)


def test_module_forwards():
    g = load_golden("modules")
    ps = O.init_all(42)
    gen = torch.Generator().manual_seed(1)
    B = 3
    tem = torch.randn(B, 512, generator=gen)
    eps = torch.randn(B, 128, generator=gen)
    z = torch.randn(B, 100, generator=gen)
    img64 = torch.randn(B, 3, 64, 64, generator=gen).clamp_(-1, 1)
    img256 = torch.randn(B, 3, 256, 256, generator=gen).clamp_(-1, 1)
    assert torch.equal(tem, g["inputs"]["tem"])
    assert_digest(img256, g["inputs"]["img256_digest"], 0, 0, "img256")
    with torch.no_grad():
        c_hat, mu, sigma = O.ca_forward(ps["con_augment_1"], tem, eps)
        for k, v in dict(c_hat=c_hat, mu=mu, sigma=sigma).items():
            assert torch.allclose(v, g["ca"][k], rtol=RT, atol=AT), k
        fake64 = O.g1_forward(ps["gen_1"], torch.cat((c_hat, z), 1))
        assert_digest(fake64, g["g1"]["digest"], RT, AT, "g1")
        assert torch.allclose(fake64[:, :, :4, :4], g["g1"]["corner"], rtol=RT, atol=AT)
        assert torch.allclose(O.d1_forward(ps["critic_1"], img64, tem), g["d1"]["score"], rtol=RT, atol=1e-5)
        assert torch.allclose(O.d2_forward(ps["critic_2"], img256, tem), g["d2"]["score"], rtol=RT, atol=1e-5)
        fake256 = O.g2_forward(ps["gen_2"], img64, c_hat)
        assert_digest(fake256, g["g2"]["digest"], RT, AT, "g2")
        assert_digest(O.g1_forward(ps["gen_1"], torch.cat((c_hat, z), 1), training=False),
                      g["g1_eval"]["digest"], RT, AT, "g1_eval")
        assert torch.allclose(O.d1_forward(ps["critic_1"], img64, tem, training=False),
                              g["d1_eval"]["score"], rtol=RT, atol=1e-5)
    d1 = ps["critic_1"]
    for v in O.trainable(d1).values():
        v.requires_grad_(True)
    gp = O.gradient_penalty(lambda i, t: O.d1_forward(d1, i, t), img64,
                            fake64.detach().requires_grad_(True), tem, g["gp1"]["eps"])
    assert torch.allclose(gp, g["gp1"]["value"], rtol=RT, atol=AT)
    gp.backward()
    assert_digest_dict({k: v.grad for k, v in O.trainable(d1).items()}, g["gp1"]["grads"], 1e-3, 1e-6, "gp grads")


def _run_stage1(g):
    ps = O.init_all(42, with_stage2=False)
    ca, d1, g1 = ps["con_augment_1"], ps["critic_1"], ps["gen_1"]
    b = O.synthetic_batch(g["B"], 1, g["seed"])
    assert_digest(b["tem"], g["inputs"]["tem_digest"], 0, 0, "tem")
    assert_digest(b["real"], g["inputs"]["real_digest"], 0, 0, "real")
    tr = dict(ca=O.Trainer(ca), d1=O.Trainer(d1), g1=O.Trainer(g1))
    i = g["inputs"]
    tem = b["tem"].clone().requires_grad_(True)
    return O.stage1_step(ca, d1, g1, b["real"], tem, i["perm"], i["z"], i["eps_ca"], i["eps_gp"], tr)


def test_stage1_step_matches_unmodified_train_1():
    g = load_golden("stage1_B4")
    out = _run_stage1(g)
    for it in range(5):
        assert_digest_dict(out["critic_grads"][it], g["critic_grads"][it], 2e-3, 1e-6, f"critic grads it{it}")
    assert_digest_dict(out["g1_grads"], g["g1_grads"], 2e-3, 1e-6, "g1 grads")
    assert_digest_dict(out["ca_grads"], g["ca_grads"], 2e-3, 1e-5, "ca grads")
    assert torch.allclose(out["dtem"], g["dtem"], rtol=2e-3, atol=1e-5)
    # the reference prints "... Loss D: {:.4f}, loss G: {:.4f}" (stage_1_train_fn.py:178-181)
    line = g["printed"]
    ld = float(line.split("Loss D:")[1].split(",")[0])
    lg = float(line.split("loss G:")[1])
    assert abs(out["loss_critic"][-1].item() - ld) <= 2e-3 * max(1, abs(ld))
    assert abs(out["lossG"].item() - lg) <= 1e-4 * abs(lg) + 1e-3
    for k in ("ca", "d1", "g1"):
        assert_digest_dict(out["after"][k], g["after"][k], 2e-3, 2e-5, f"weights after [{k}]")
    assert int(out["after"]["d1"]["down_sampler.2.1.num_batches_tracked"]) == g["nbt"]["d1"] == 21
    assert int(out["after"]["g1"]["upsampling.0.1.num_batches_tracked"]) == g["nbt"]["g1"] == 5


@pytest.mark.slow
def test_stage2_step_matches_fixed_train_2():
    g = load_golden("stage2_B2")
    ps = O.init_all(42)
    b = O.synthetic_batch(g["B"], 2, g["seed"])
    assert_digest(b["real"], g["inputs"]["real_digest"], 0, 0, "real")
    ca1, g1, ca2, d2, g2 = (ps[k] for k in ("con_augment_1", "gen_1", "con_augment_2", "critic_2", "gen_2"))
    tr = dict(ca2=O.Trainer(ca2), d2=O.Trainer(d2), g2=O.Trainer(g2))
    i = g["inputs"]
    out = O.stage2_step(ca1, g1, ca2, d2, g2, b["real"], b["tem"], i["perm"], i["z"],
                        i["eps_ca1"], i["eps_ca2"], i["eps_gp"], tr)
    for it in range(5):
        assert_digest_dict(out["critic_grads"][it], g["critic_grads"][it], 5e-3, 1e-6, f"critic2 grads it{it}")
    assert_digest_dict(out["g2_grads"], g["g2_grads"], 5e-3, 1e-6, "g2 grads (accumulated)")
    assert_digest_dict(out["ca2_grads"], g["ca2_grads"], 5e-3, 1e-5, "ca2 grads")
    for k in ("ca2", "d2", "g2"):
        assert_digest_dict(out["after"][k], g["after"][k], 5e-3, 5e-5, f"weights after [{k}]")


def test_caption_loader_matches_the_reference_loader(tmp_path):
    """tests/golden/loader.pt holds what the UNMODIFIED reference ``data_loader.get_loader`` yields (oracle/
    make_golden_loader.py: fake GCS bucket, 8 workers, DistributedSampler) on the seeded directory
    ``_util.make_coco_dir`` writes; ``imagegenerator_b200.data_loader.get_loader`` must yield the same batches."""
    from _util import make_coco_dir
    from imagegenerator_b200.data_loader import get_loader
    from imagegenerator_b200.train import image_transform
    gold = load_golden("loader")
    root, ann, tok, rows = make_coco_dir(tmp_path, n_images=gold["n_images"], captions_per_image=gold["captions_per_image"])
    assert [tuple(r) for r in rows] == [tuple(r) for r in gold["rows"]]
    for key, shuffle in (("ordered", False), ("shuffled", True)):
        loader = get_loader("data-and-checkpoints-bucket", root, ann, image_transform(64), batch_size=gold["batch"],
                            shuffle=shuffle, tokenizer=tok, num_workers=0)
        assert len(loader) == gold[key + "_len"] and len(loader.dataset) == gold[key + "_dataset_len"]
        for (tokenized, imgs), want in zip(loader, gold[key]):
            assert set(tokenized.keys()) == set(want["tokenized"].keys())
            for k, v in want["tokenized"].items():
                assert torch.equal(tokenized[k], v), (key, k)                     # same captions in the same order
            if want["imgs"] is not None:
                assert torch.allclose(imgs, want["imgs"], rtol=0, atol=1e-6)
            assert abs(imgs.double().sum().item() - want["img_sum"]) <= 1e-4 * want["img_abs"]
            assert abs(imgs.double().abs().sum().item() - want["img_abs"]) <= 1e-5 * want["img_abs"]
