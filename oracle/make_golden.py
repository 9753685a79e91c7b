"""Generate tests/golden/*.pt from the REAL reference -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

What is recorded (all produced by the unmodified reference code under the stubs of
oracle/ref_harness.py; Stage-II with its two one-token fixes):

  * modules.pt   -- forward outputs of the five reference modules on seeded inputs
  * stage1_B4.pt -- one outer step of the unmodified ``train_1`` (B=4): inputs, the
                    recorded noise tape, both losses, and digests of the gradients
                    every optimizer saw at its step and of the weights after it
  * stage2_B2.pt -- same for ``train_2`` (B=2)

Weights are NOT stored (16-100 MB): they are regenerated from the seed by
``stackgan_oracle.init_all`` -- the fixture keeps their digest so a drift in
torch's RNG/init would be caught.  Tensors are stored as digests
(sum, L2 norm, 64 sampled elements) to keep the fixtures small.
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import torch

from . import ref_harness as H
from . import stackgan_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def digest(t, n=64):
    t = t.detach().to(torch.float64).reshape(-1)
    g = torch.Generator().manual_seed(t.numel() % 9973 + 17)
    idx = torch.randint(0, t.numel(), (min(n, t.numel()),), generator=g)
    return dict(numel=t.numel(), sum=t.sum().item(), norm=t.norm().item(), idx=idx, vals=t[idx].clone())


def digest_dict(d):
    return OrderedDict((k, digest(v)) for k, v in d.items() if v is not None and v.is_floating_point())


def _load_into(module, p):
    sd = module.state_dict()
    assert list(sd.keys()) == list(p.keys()), (list(sd.keys())[:5], list(p.keys())[:5])
    module.load_state_dict(p)


def build_reference_models(seed=42):
    """Real reference modules, constructed in train.py:70-75 order under the seed."""
    torch.manual_seed(seed)
    CA = H.load("con_augment").ConditioningAugmentation
    ms = OrderedDict()
    ms["con_augment_1"] = CA(512, 256, 128)
    ms["critic_1"] = H.load("discrminator_1").StageIDiscriminator(512, 128)
    ms["gen_1"] = H.load("generator_1").StageIGenerator(128, 100)
    ms["con_augment_2"] = CA(512, 256, 128)
    ms["critic_2"] = H.load("discriminator_2").StageIIDiscriminator(512, 128)
    ms["gen_2"] = H.load("generator_2").StageIIGenerator()
    return ms


def golden_modules():
    ms = build_reference_models()
    g = torch.Generator().manual_seed(1)
    B = 3
    tem = torch.randn(B, 512, generator=g)
    eps = torch.randn(B, 128, generator=g)
    z = torch.randn(B, 100, generator=g)
    img64 = torch.randn(B, 3, 64, 64, generator=g).clamp_(-1, 1)
    img256 = torch.randn(B, 3, 256, 256, generator=g).clamp_(-1, 1)
    out = OrderedDict(inputs=dict(tem=tem, eps=eps, z=z, img64_digest=digest(img64), img256_digest=digest(img256)))
    out["init_digest"] = OrderedDict((k, digest_dict(m.state_dict())) for k, m in ms.items())
    for m in ms.values():
        m.train()
    # con_augment: eps comes from the global RNG (con_augment.py:20) -> replay it
    with torch.no_grad():
        real_randn_like = torch.randn_like
        torch.randn_like = lambda t, *a, **k: eps.clone()
        try:
            c_hat, mu, sigma = ms["con_augment_1"](tem)
        finally:
            torch.randn_like = real_randn_like
        out["ca"] = dict(c_hat=c_hat, mu=mu, sigma=sigma)
        fake64 = ms["gen_1"](torch.cat((c_hat, z), 1))
        out["g1"] = dict(digest=digest(fake64), corner=fake64[:, :, :4, :4].clone())
        out["d1"] = dict(score=ms["critic_1"](img64, tem))
        out["d2"] = dict(score=ms["critic_2"](img256, tem))
        fake256 = ms["gen_2"](img64, c_hat)
        out["g2"] = dict(digest=digest(fake256), corner=fake256[:, :, :4, :4].clone())
        # eval-mode forwards (running stats after exactly one training forward each)
        for m in ms.values():
            m.eval()
        out["g1_eval"] = dict(digest=digest(ms["gen_1"](torch.cat((c_hat, z), 1))))
        out["d1_eval"] = dict(score=ms["critic_1"](img64, tem))
    # gradient penalty (utils.py:8-26) with the torch.rand draw replayed
    for m in ms.values():
        m.train()
    e = torch.rand(B, 1, 1, 1, generator=g)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: e.clone()
    try:
        gp = H.load("utils").gradient_penalty(ms["critic_1"], img64, fake64.detach().requires_grad_(True), tem, "cpu")
    finally:
        torch.rand = real_rand
    gp.backward()
    out["gp1"] = dict(eps=e.reshape(-1), value=gp.detach(),
                      grads=digest_dict(OrderedDict((k, p.grad) for k, p in ms["critic_1"].named_parameters())))
    return out


def _named(m):
    return list(m.named_parameters())


def golden_stage1(B=4, seed=0):
    H.reset_store()
    H.seed_everything(1234)
    ms = build_reference_models()
    batch = O.synthetic_batch(B, 1, seed)
    enc, head = H.TableEncoder(batch["tem"]), H.IdentityHead()
    ca, d1, g1 = ms["con_augment_1"], ms["critic_1"], ms["gen_1"]
    init_digest = OrderedDict((k, digest_dict(ms[k].state_dict())) for k in ("con_augment_1", "critic_1", "gen_1"))
    mk = lambda m, lr=1e-3: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.999))
    opts = [mk(enc, 0.0), mk(head, 0.0), mk(ca), mk(d1), mk(g1)]
    scheds = [torch.optim.lr_scheduler.StepLR(o, step_size=100, gamma=0.5) for o in opts]
    for o, tag, m in zip(opts, ["enc", "head", "ca", "d1", "g1"], [enc, head, ca, d1, g1]):
        H.RECORDER.register(o, tag, _named(m))
    loader = [({"idx": torch.arange(B)}, batch["real"])]
    s1 = H.load("stage_1_train_fn")
    tape = H.NoiseTape()
    torch.manual_seed(777)            # global RNG state at the start of the step (randint/eps draws)
    with tape.recording(), H.quiet():
        s1.train_1([enc, head, ca, d1, g1], opts, scheds, loader, 1, "cpu", B)
    kinds = [n for n, _ in tape.draws]
    assert kinds == ["randint", "randperm"] + ["randn_like", "randn", "rand"] * 5, kinds
    out = OrderedDict(B=B, seed=seed, init_digest=init_digest)
    # tem/real are regenerated by stackgan_oracle.synthetic_batch(B, 1, seed); only digests are kept
    out["inputs"] = dict(tem_digest=digest(batch["tem"]), real_digest=digest(batch["real"]),
                         perm=tape.by_kind("randperm")[0],
                         z=torch.stack(tape.by_kind("randn")), eps_ca=torch.stack(tape.by_kind("randn_like")),
                         eps_gp=torch.stack([t.reshape(-1) for t in tape.by_kind("rand")]))
    ev = H.RECORDER.events
    tags = [t for t, _ in ev]
    assert tags == ["d1"] * 5 + ["g1", "enc", "head", "ca"], tags
    out["critic_grads"] = [digest_dict(g) for t, g in ev if t == "d1"]
    out["g1_grads"] = digest_dict(ev[5][1])
    out["ca_grads"] = digest_dict(ev[8][1])
    out["dtem"] = ev[6][1]["table"]                      # d lossG / d tem   [B,512]
    line = H.RECORDER.prints[-1]
    out["printed"] = line
    out["after"] = OrderedDict(ca=digest_dict(ca.state_dict()), d1=digest_dict(d1.state_dict()),
                               g1=digest_dict(g1.state_dict()))
    out["nbt"] = dict(d1=int(d1.state_dict()["down_sampler.2.1.num_batches_tracked"]),
                      g1=int(g1.state_dict()["upsampling.0.1.num_batches_tracked"]))
    return out


def golden_stage2(B=2, seed=0):
    H.reset_store()
    H.seed_everything(4321)
    ms = build_reference_models()
    batch = O.synthetic_batch(B, 2, seed)
    enc, head = H.TableEncoder(batch["tem"]), H.IdentityHead()
    ca1, g1, ca2, d2, g2 = (ms[k] for k in ("con_augment_1", "gen_1", "con_augment_2", "critic_2", "gen_2"))
    init_digest = OrderedDict((k, digest_dict(ms[k].state_dict()))
                              for k in ("con_augment_1", "gen_1", "con_augment_2", "critic_2", "gen_2"))
    # Stage-1 checkpoint the reference insists on loading (stage_2_train_fn.py:65-72)
    import tempfile
    with tempfile.NamedTemporaryFile() as tmp:
        torch.save(dict(textEncoder=enc.state_dict(), projection_head=head.state_dict(),
                        con_augment_1=ca1.state_dict(), gen_1=g1.state_dict()), tmp.name)
        with open(tmp.name, "rb") as f:
            H.store()["./checkpoint/Stage1/latest_checkpoint_stage1.pth"] = f.read()
    mk = lambda m: torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.999))
    opts = [mk(ca2), mk(d2), mk(g2)]
    scheds = [torch.optim.lr_scheduler.StepLR(o, step_size=100, gamma=0.5) for o in opts]
    for o, tag, m in zip(opts, ["ca2", "d2", "g2"], [ca2, d2, g2]):
        H.RECORDER.register(o, tag, _named(m))
    loader = [({"idx": torch.arange(B)}, batch["real"])]
    s2 = H.load("stage_2_train_fn")
    tape = H.NoiseTape()
    torch.manual_seed(778)
    import random
    random.seed(99)                    # stage_2_train_fn.py:125 reseeds torch from python's RNG
    with tape.recording(), H.quiet():
        s2.train_2([enc, head, ca1, ca2, g1, d2, g2], opts, scheds, loader, 1, "cpu", B)
    kinds = [n for n, _ in tape.draws]
    assert kinds == ["randint", "randperm"] + ["randn_like", "randn", "randn_like", "rand"] * 5, kinds
    rl = tape.by_kind("randn_like")
    out = OrderedDict(B=B, seed=seed, init_digest=init_digest)
    out["inputs"] = dict(tem_digest=digest(batch["tem"]), real_digest=digest(batch["real"]),
                         perm=tape.by_kind("randperm")[0],
                         z=torch.stack(tape.by_kind("randn")), eps_ca1=torch.stack(rl[0::2]),
                         eps_ca2=torch.stack(rl[1::2]),
                         eps_gp=torch.stack([t.reshape(-1) for t in tape.by_kind("rand")]))
    ev = H.RECORDER.events
    tags = [t for t, _ in ev]
    assert tags == ["d2"] * 5 + ["g2", "ca2"], tags
    out["critic_grads"] = [digest_dict(g) for t, g in ev if t == "d2"]
    out["g2_grads"] = digest_dict(ev[5][1])
    out["ca2_grads"] = digest_dict(ev[6][1])
    out["after"] = OrderedDict(ca2=digest_dict(ca2.state_dict()), d2=digest_dict(d2.state_dict()),
                               g2=digest_dict(g2.state_dict()))
    return out


def main():
    if not H.reference_available():
        sys.exit("reference tree not available; golden fixtures can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    for name, fn in (("modules", golden_modules), ("stage1_B4", golden_stage1), ("stage2_B2", golden_stage2)):
        data = fn()
        data["torch_version"] = torch.__version__
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(data, path)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
