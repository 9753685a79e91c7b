"""CPU restatement of the reference's StackGAN path -- TEST INFRASTRUCTURE ONLY.

The reference's arithmetic lives in un-vendored, un-pinned PyTorch (SURVEY.md
section 8c); this file restates every function on the hot path as a pure function
of (parameter dict keyed by the reference's state_dict names, inputs, explicit
noise) on torch CPU ops, so it can travel to the GPU box where /root/reference
does not exist.  It is pinned against the real reference by the golden fixtures under tests/golden/
(recorded from the unmodified reference by oracle/make_golden*.py through
oracle/ref_harness.py; checked by tests/test_oracle_golden.py everywhere).

All citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

N_CRITIC = 5      # stage_1_train_fn.py:14 / stage_2_train_fn.py:15
LAMBDA_GP = 10    # stage_1_train_fn.py:15
Z_DIM = 100       # stage_1_train_fn.py:16
BN_EPS = 1e-5     # nn.BatchNorm2d default
BN_MOMENTUM = 0.1
LRELU = 0.1       # discrminator_1.py:11,38; discriminator_2.py:10,53; generator_2.py:47,96


# --------------------------------------------------------------------------- parameter construction
def _conv_like(p, key, shape, bias, fan_in):
    w = torch.empty(shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))       # torch _ConvNd/Linear.reset_parameters
    p[key + ".weight"] = w
    if bias:
        b = torch.empty(shape[0] if bias is True else bias)
        bound = 1.0 / math.sqrt(fan_in)
        torch.nn.init.uniform_(b, -bound, bound)
        p[key + ".bias"] = b


def _bn(p, key, c):
    p[key + ".weight"] = torch.ones(c)
    p[key + ".bias"] = torch.zeros(c)
    p[key + ".running_mean"] = torch.zeros(c)
    p[key + ".running_var"] = torch.ones(c)
    p[key + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def init_con_augment(tem_size=512, h_dim=256, c_dim=128):
    """con_augment.py:7-11 (three nn.Linear, default init, in this order)."""
    p = OrderedDict()
    _conv_like(p, "h", (h_dim, tem_size), True, tem_size)
    _conv_like(p, "mu", (c_dim, h_dim), True, h_dim)
    _conv_like(p, "sigma", (c_dim, h_dim), True, h_dim)
    return p


G1_CH = [192, 96, 48, 24]


def init_generator_1(c_dim=128, z_dim=100):
    """generator_1.py:9-22.  ConvTranspose2d weight is (Cin, Cout, kh, kw); torch's fan_in
    for it is size(1)*kh*kw = Cout*16."""
    p = OrderedDict()
    cin = c_dim + z_dim
    for i, co in enumerate(G1_CH):
        _conv_like(p, f"upsampling.{i}.0", (cin, co, 4, 4), False, co * 16)
        _bn(p, f"upsampling.{i}.1", co)
        cin = co
    _conv_like(p, "upsampling.4", (cin, 3, 4, 4), 3, 3 * 16)
    return p


def _init_critic(chs, head_ch, tem_size, Nd):
    p = OrderedDict()
    _conv_like(p, "down_sampler.0", (chs[0], 3, 4, 4), True, 3 * 16)
    cin = chs[0]
    for j, co in enumerate(chs[1:]):
        _conv_like(p, f"down_sampler.{j + 2}.0", (co, cin, 4, 4), False, cin * 16)
        _bn(p, f"down_sampler.{j + 2}.1", co)
        cin = co
    _conv_like(p, "compress", (Nd, tem_size), True, tem_size)
    _conv_like(p, "channel_resize", (head_ch, cin + Nd, 1, 1), True, cin + Nd)
    _conv_like(p, "critic_score", (1, head_ch * 16), True, head_ch * 16)
    return p


def init_discriminator_1(tem_size=512, Nd=128):
    """discrminator_1.py:9-23."""
    return _init_critic([64, 128, 256, 512], 128, tem_size, Nd)


def init_discriminator_2(tem_size=512, Nd=128):
    """discriminator_2.py:8-25."""
    return _init_critic([16, 32, 64, 128, 256, 512], 160, tem_size, Nd)


def init_generator_2():
    """generator_2.py:45-57 (down_sampler, 4 residual blocks, up_sampler)."""
    p = OrderedDict()
    _conv_like(p, "down_sampler.0", (128, 3, 4, 4), True, 48)
    _conv_like(p, "down_sampler.2.0", (512, 128, 4, 4), False, 128 * 16)
    _bn(p, "down_sampler.2.1", 512)
    for r in range(4):
        for name, (co, ci) in (("layer1", (320, 640)), ("layer2", (320, 320)), ("layer3", (640, 320))):
            _conv_like(p, f"residual_blocks.{r}.{name}.0", (co, ci, 3, 3), False, ci * 9)
            _bn(p, f"residual_blocks.{r}.{name}.1", co)
    cin = 640
    for i, co in enumerate([320, 160, 80]):
        _conv_like(p, f"up_sampler.{i}.0", (cin, co, 4, 4), False, co * 16)
        _bn(p, f"up_sampler.{i}.1", co)
        cin = co
    _conv_like(p, "up_sampler.3", (cin, 3, 4, 4), 3, 48)
    return p


def init_all(seed=42, with_stage2=True):
    """Models in train.py:70-75 construction order under torch.manual_seed (train.py:66).
    (The reference also builds SpanBERT + Linear(768,512) first; the text side is
    synthetic here, so weights are NOT those of a real train.py run -- only the
    per-module default-init distribution and ordering is kept.)"""
    torch.manual_seed(seed)
    out = OrderedDict()
    out["con_augment_1"] = init_con_augment()
    out["critic_1"] = init_discriminator_1()
    out["gen_1"] = init_generator_1()
    if with_stage2:
        out["con_augment_2"] = init_con_augment()
        out["critic_2"] = init_discriminator_2()
        out["gen_2"] = init_generator_2()
    return out


def is_buffer(key):
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def trainable(p):
    return OrderedDict((k, v) for k, v in p.items() if not is_buffer(k))


# --------------------------------------------------------------------------- forwards
def _bn_apply(p, key, x, training):
    """nn.BatchNorm2d forward incl. buffer updates (momentum 0.1, unbiased running var)."""
    if training:
        p[key + ".num_batches_tracked"] += 1
    return F.batch_norm(x, p[key + ".running_mean"], p[key + ".running_var"],
                        p[key + ".weight"], p[key + ".bias"], training, BN_MOMENTUM, BN_EPS)


def ca_encode(p, tem):
    """con_augment.py:13-16."""
    h = F.relu(F.linear(tem, p["h.weight"], p["h.bias"]))
    return F.linear(h, p["mu.weight"], p["mu.bias"]), F.linear(h, p["sigma.weight"], p["sigma.bias"])


def ca_forward(p, tem, eps):
    """con_augment.py:18-22 with the randn_like draw (:20) supplied by the caller.
    sigma is the raw linear output used as a std-dev (no exp)."""
    mu, sigma = ca_encode(p, tem)
    return mu + sigma * eps, mu, sigma


def g1_forward(p, x, training=True):
    """generator_1.py:38-40: [B,228] -> [B,3,64,64]."""
    x = x.reshape(x.shape[0], x.shape[1], 1, 1)
    for i in range(4):
        stride, pad = (1, 0) if i == 0 else (2, 1)
        x = F.conv_transpose2d(x, p[f"upsampling.{i}.0.weight"], None, stride, pad)
        x = F.relu(_bn_apply(p, f"upsampling.{i}.1", x, training))
    x = F.conv_transpose2d(x, p["upsampling.4.weight"], p["upsampling.4.bias"], 2, 1)
    return torch.tanh(x)


def _critic_forward(p, n_bn, img, tem, training):
    x = F.leaky_relu(F.conv2d(img, p["down_sampler.0.weight"], p["down_sampler.0.bias"], 2, 1), LRELU)
    for j in range(2, 2 + n_bn):
        x = F.conv2d(x, p[f"down_sampler.{j}.0.weight"], None, 2, 1)
        x = F.leaky_relu(_bn_apply(p, f"down_sampler.{j}.1", x, training), LRELU)
    ce = F.linear(tem, p["compress.weight"], p["compress.bias"])
    rep = ce.reshape(ce.shape[0], ce.shape[1], 1, 1).repeat(1, 1, 4, 4)
    cat = torch.cat((x, rep), dim=1)
    t = F.conv2d(cat, p["channel_resize.weight"], p["channel_resize.bias"])
    return F.linear(t.flatten(1), p["critic_score.weight"], p["critic_score.bias"])


def d1_forward(p, img, tem, training=True):
    """discrminator_1.py:41-52: [B,3,64,64],[B,512] -> [B,1]."""
    return _critic_forward(p, 3, img, tem, training)


def d2_forward(p, img, tem, training=True):
    """discriminator_2.py:27-38 with the :28 fix (down_sampler(img))."""
    return _critic_forward(p, 5, img, tem, training)


def g2_forward(p, img_64, c_hat, training=True):
    """generator_2.py:59-67 (+ ResidualBlock.forward :15-26)."""
    x = F.leaky_relu(F.conv2d(img_64, p["down_sampler.0.weight"], p["down_sampler.0.bias"], 2, 1), LRELU)
    x = F.conv2d(x, p["down_sampler.2.0.weight"], None, 2, 1)
    x = F.leaky_relu(_bn_apply(p, "down_sampler.2.1", x, training), LRELU)
    rep = c_hat.reshape(c_hat.shape[0], c_hat.shape[1], 1, 1).repeat(1, 1, 16, 16)
    x = torch.cat((x, rep), dim=1)
    for r in range(4):
        k = f"residual_blocks.{r}."
        idt = x
        y = F.relu(_bn_apply(p, k + "layer1.1", F.conv2d(x, p[k + "layer1.0.weight"], None, 1, 1), training))
        y = F.relu(_bn_apply(p, k + "layer2.1", F.conv2d(y, p[k + "layer2.0.weight"], None, 1, 1), training))
        y = _bn_apply(p, k + "layer3.1", F.conv2d(y, p[k + "layer3.0.weight"], None, 1, 1), training)
        x = F.relu(y + idt)
    for i in range(3):
        x = F.conv_transpose2d(x, p[f"up_sampler.{i}.0.weight"], None, 2, 1)
        x = F.relu(_bn_apply(p, f"up_sampler.{i}.1", x, training))
    x = F.conv_transpose2d(x, p["up_sampler.3.weight"], p["up_sampler.3.bias"], 2, 1)
    return torch.tanh(x)


def sample(ca1, g1, ca2, g2, tem, z, eps_ca1, eps_ca2, g2_training=False):
    """The forward-only preview of stage_2_train_fn.py:181-195: c_hat1 -> gen_1 (eval, :59-63) -> fake_64;
    c_hat2 -> gen_2 -> fake_256.  ``g2_training=True`` keeps gen_2's BatchNorm on batch statistics, which is what
    the reference's in-loop call does (gen_2.train() at :90 is never undone); False is plain inference."""
    with torch.no_grad():
        c_hat1, _, _ = ca_forward(ca1, tem, eps_ca1)
        fake_64 = g1_forward(g1, torch.cat((c_hat1, z), dim=1), training=False)
        c_hat2, _, _ = ca_forward(ca2, tem, eps_ca2)
        fake_256 = g2_forward(g2, fake_64, c_hat2, training=g2_training)
    return fake_64, fake_256


def gradient_penalty(critic_fn, real, fake, tem, eps):
    """utils.py:8-26; ``eps`` [B] is the torch.rand((B,1,1,1)) draw of :10."""
    e = eps.reshape(-1, 1, 1, 1).to(real.dtype)
    interp = real * e + fake * (1 - e)
    mixed = critic_fn(interp, tem)
    g = torch.autograd.grad(mixed, interp, torch.ones_like(mixed), create_graph=True, retain_graph=True)[0]
    gn = g.reshape(g.shape[0], -1).norm(2, dim=1)
    return torch.mean((gn - 1) ** 2)


def kl_term(mu, sigma):
    """stage_1_train_fn.py:156-158 (no -1/2, summed over batch and features)."""
    return torch.sum(1 + torch.log(sigma.pow(2)) - mu.pow(2) - sigma.pow(2))


# --------------------------------------------------------------------------- optimiser helper
class Trainer:
    """One reference optimizer (Adam lr 1e-3 betas .9/.999, train.py:92-102) + StepLR(100,.5)
    (train.py:105-113) over the trainable entries of a parameter dict."""

    def __init__(self, p, lr=1e-3):
        self.p = p
        self.params = trainable(p)
        for v in self.params.values():
            v.requires_grad_(True)
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr, betas=(0.9, 0.999))
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=100, gamma=0.5)

    def grads(self):
        return OrderedDict((k, (v.grad.detach().clone() if v.grad is not None else None))
                           for k, v in self.params.items())


def _snap(p):
    return OrderedDict((k, v.detach().clone()) for k, v in p.items())


# --------------------------------------------------------------------------- Stage-I outer step
def _force(p, snap):
    """Overwrite a parameter dict in place with a snapshot (any dtype) -- used to re-synchronise
    a lower-precision run with the fp64 trajectory ("teacher forcing") in the parity tests."""
    with torch.no_grad():
        for k, v in snap.items():
            p[k].copy_(v.to(p[k].dtype))


def stage1_step(ca, d1, g1, real, tem, perm, z, eps_ca, eps_gp, tr, force=None, sync=None):
    """One outer step of stage_1_train_fn.py:93-196 with synthetic text embeddings.

    ca/d1/g1: parameter dicts; tr = dict(ca=Trainer, d1=Trainer, g1=Trainer);
    tem [B,512] requires_grad (leaf) so d lossG / d tem is reported;
    perm [B] (the randperm of :109), z [5,B,100] (:121), eps_ca [5,B,128]
    (con_augment.py:20), eps_gp [5,B] (utils.py:10).
    ``sync(trainer)``: optional hook run right before every optimizer step -- the data-parallel tests
    use it to average gradients over replicas (xm.optimizer_step, :149,:166-172).
    Returns losses and the gradients each optimizer saw at its step."""
    sync = sync or (lambda t: None)
    out = {"loss_critic": [], "critic_grads": [], "critic_before": [], "scores": []}
    tem_mis = tem[perm]                                        # :108-111, :127-129
    for it in range(N_CRITIC):
        if force is not None:
            _force(d1, force[it])
        out["critic_before"].append(_snap(d1))
        c_hat, mu, sigma = ca_forward(ca, tem, eps_ca[it])      # :120
        fake = g1_forward(g1, torch.cat((c_hat, z[it]), dim=1))  # :121-123 (not detached)
        s_real = d1_forward(d1, real, tem).view(-1)             # :125
        s_mis = d1_forward(d1, real, tem_mis).view(-1)          # :130
        s_fake = d1_forward(d1, fake, tem).view(-1)             # :132
        gp = gradient_penalty(lambda i, t: d1_forward(d1, i, t), real, fake, tem, eps_gp[it])  # :138
        loss_c = torch.mean(torch.cat((s_mis, s_fake))) - torch.mean(s_real) + LAMBDA_GP * gp  # :140-144
        tr["d1"].opt.zero_grad()                                # :146
        loss_c.backward(retain_graph=True)                      # :147
        out["critic_grads"].append(tr["d1"].grads())
        out["loss_critic"].append(loss_c.detach().clone())
        out["scores"].append(dict(s_real=s_real.detach().clone(), s_mis=s_mis.detach().clone(),
                                  s_fake=s_fake.detach().clone(), gp=gp.detach().clone(),
                                  fake=fake.detach().clone()))
        if it == 0:
            out["first"] = dict(fake=fake.detach().clone(), s_real=s_real.detach().clone(),
                                s_mis=s_mis.detach().clone(), s_fake=s_fake.detach().clone(),
                                gp=gp.detach().clone(), c_hat=c_hat.detach().clone(),
                                mu=mu.detach().clone(), sigma=sigma.detach().clone())
        sync(tr["d1"])
        tr["d1"].opt.step()                                     # :149
    if force is not None:
        _force(d1, force[N_CRITIC])
    out["critic_before"].append(_snap(d1))
    s = d1_forward(d1, fake, tem).view(-1)                      # :154
    lossG = -torch.mean(s) + kl_term(mu, sigma)                 # :155-159
    out["s_gen"] = s.detach().clone()
    tr["g1"].opt.zero_grad()                                    # :161-164
    tr["ca"].opt.zero_grad()
    tem.grad = None
    lossG.backward()                                            # :165
    out["lossG"] = lossG.detach().clone()
    out["g1_grads"] = tr["g1"].grads()
    out["ca_grads"] = tr["ca"].grads()
    out["dtem"] = tem.grad.detach().clone() if tem.grad is not None else None
    sync(tr["g1"])
    tr["g1"].opt.step()                                         # :166
    sync(tr["ca"])
    tr["ca"].opt.step()                                         # :172
    for k in ("d1", "g1", "ca"):                                # :187-192 (per batch)
        tr[k].sched.step()
    out["after"] = dict(ca=_snap(ca), d1=_snap(d1), g1=_snap(g1))
    return out


# --------------------------------------------------------------------------- Stage-II outer step
def stage2_step(ca1, g1, ca2, d2, g2, real, tem, perm, z, eps_ca1, eps_ca2, eps_gp, tr, sync=None, force=None):
    """One outer step of stage_2_train_fn.py:101-173 (with the :67 / discriminator_2.py:28 fixes).

    ca1/g1 are frozen and in eval mode (:52-63: running-stat BN in gen_1, CA still samples).
    G2/CA2 gradients ACCUMULATE over the five critic backward passes because
    fake_256 is not detached and opt_gen_2.zero_grad() only runs after the step
    (:131,:154,:163-168) -- reproduced here.  tr = dict(ca2=, d2=, g2=).  ``sync(trainer)``, if given, runs right
    before each optimizer step: the gradient mean over replicas of xm.optimizer_step (:155,:164,:167).
    ``force``: six critic snapshots (before each critic iteration and before the generator step) to overwrite the
    critic with -- re-synchronises a lower-precision run with a reference trajectory in the parity tests; the
    snapshots of this run are returned as ``critic_before``."""
    out = {"loss_critic": [], "critic_grads": [], "critic_before": [], "scores": []}
    tem_mis = tem[perm]
    for it in range(N_CRITIC):
        if force is not None:
            _force(d2, force[it])
        out["critic_before"].append(_snap(d2))
        with torch.no_grad():                                   # frozen params => no graph needed
            c_hat1, _, _ = ca_forward(ca1, tem, eps_ca1[it])    # :124
            fake_64 = g1_forward(g1, torch.cat((c_hat1, z[it]), dim=1), training=False)  # :126-128
        c_hat2, mu2, sigma2 = ca_forward(ca2, tem, eps_ca2[it])  # :130
        fake = g2_forward(g2, fake_64, c_hat2)                  # :131
        s_real = d2_forward(d2, real, tem).view(-1)             # :133
        s_mis = d2_forward(d2, real, tem_mis).view(-1)          # :138
        s_fake = d2_forward(d2, fake, tem).view(-1)             # :140
        gp = gradient_penalty(lambda i, t: d2_forward(d2, i, t), real, fake, tem, eps_gp[it])  # :146
        loss_c = torch.mean(torch.cat((s_mis, s_fake))) - torch.mean(s_real) + LAMBDA_GP * gp
        tr["d2"].opt.zero_grad()                                # :153
        loss_c.backward(retain_graph=True)                      # :154 (also accumulates into g2/ca2)
        out["critic_grads"].append(tr["d2"].grads())
        out["loss_critic"].append(loss_c.detach().clone())
        out["scores"].append(dict(s_real=s_real.detach().clone(), s_mis=s_mis.detach().clone(),
                                  s_fake=s_fake.detach().clone(), gp=gp.detach().clone()))
        if it == 0:
            out["g2_grads_it0"] = tr["g2"].grads()             # what one critic backward leaves in G2
            out["ca2_grads_it0"] = tr["ca2"].grads()
            out["first"] = dict(fake_64=fake_64.detach().clone(), fake=fake.detach().clone(),
                                s_real=s_real.detach().clone(), s_mis=s_mis.detach().clone(),
                                s_fake=s_fake.detach().clone(), gp=gp.detach().clone())
        if sync is not None:
            sync(tr["d2"])
        tr["d2"].opt.step()                                     # :155
    if force is not None:
        _force(d2, force[N_CRITIC])
    out["critic_before"].append(_snap(d2))
    s = d2_forward(d2, fake, tem).view(-1)                      # :157
    lossG = -torch.mean(s) + kl_term(mu2, sigma2)               # :158-162
    lossG.backward()                                            # :163 (no zero_grad before)
    out["lossG"] = lossG.detach().clone()
    out["s_gen"] = s.detach().clone()
    out["g2_grads"] = tr["g2"].grads()
    out["ca2_grads"] = tr["ca2"].grads()
    if sync is not None:
        sync(tr["g2"])
    tr["g2"].opt.step()                                         # :164
    tr["g2"].opt.zero_grad()                                    # :165
    if sync is not None:
        sync(tr["ca2"])
    tr["ca2"].opt.step()                                        # :167
    tr["ca2"].opt.zero_grad()                                   # :168
    for k in ("d2", "g2", "ca2"):                               # :170-173
        tr[k].sched.step()
    out["after"] = dict(ca2=_snap(ca2), d2=_snap(d2), g2=_snap(g2))
    return out


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_batch(B, stage, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    hw = 64 if stage == 1 else 256
    d = dict(
        tem=torch.randn(B, 512, generator=g),
        real=torch.randn(B, 3, hw, hw, generator=g).clamp_(-1, 1),
        perm=torch.randperm(B, generator=g),
        z=torch.randn(N_CRITIC, B, Z_DIM, generator=g),
        eps_ca=torch.randn(N_CRITIC, B, 128, generator=g),
        eps_gp=torch.rand(N_CRITIC, B, generator=g),
    )
    if stage == 2:
        d["eps_ca2"] = torch.randn(N_CRITIC, B, 128, generator=g)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in d.items()}


def to_dtype(p, dtype):
    return OrderedDict((k, (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()))
                       for k, v in p.items())
