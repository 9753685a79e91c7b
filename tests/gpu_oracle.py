"""The oracle (oracle/stackgan_oracle.py) run ON THE GPU as the checker of the config-size parity tests.

The oracle is plain ``torch.nn.functional`` + autograd, so the same restatement that is pinned to the
unmodified reference on CPU (tests/test_oracle_golden.py) runs on ``cuda`` tensors in fp64 (the exact
answer) or in fp32 with TF32 switched off (the precision the reference itself computes in) -- a
batch-128 Stage-I step or a batch-64 Stage-II step takes seconds there instead of minutes on the host.
It is only ever the CHECKER: nothing under imagegenerator_b200/ imports this file.
"""
import contextlib

import torch

from oracle import stackgan_oracle as O


@contextlib.contextmanager
def strict_fp32():
    """cuDNN / cuBLAS without TF32, as the judge's recipe asks (fp32 means fp32)."""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


def params_on(ps, dtype, device):
    return {name: {k: (v.detach().clone().to(device=device, dtype=dtype) if v.is_floating_point() else v.clone().to(device))
                   for k, v in p.items()} for name, p in ps.items()}


def batch_on(b, dtype, device):
    return {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device)) for k, v in b.items()}


def stage1(B, dtype=torch.float64, device="cuda", force=None, seed=0):
    """One Stage-I outer step of the oracle at batch B on ``device``.  Returns (cpu batch, result dict)."""
    ps = O.init_all(42, with_stage2=False)
    p = params_on(ps, dtype, device)
    b = O.synthetic_batch(B, 1, seed)
    bd = batch_on(b, dtype, device)
    tr = dict(ca=O.Trainer(p["con_augment_1"]), d1=O.Trainer(p["critic_1"]), g1=O.Trainer(p["gen_1"]))
    tem = bd["tem"].clone().requires_grad_(True)
    with strict_fp32():
        ref = O.stage1_step(p["con_augment_1"], p["critic_1"], p["gen_1"], bd["real"], tem, bd["perm"], bd["z"],
                            bd["eps_ca"], bd["eps_gp"], tr, force=force)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    return b, ref


def trained_stats_g1(ps, device="cuda", passes=60, B=64):
    """The state gen_1 is in when Stage-II starts after a Stage-I run (stage_2_train_fn.py:65-72 loads that checkpoint):
    its BatchNorm running statistics have converged to the batch statistics of its own activations.  Emulated by ``passes``
    train-mode forwards of the oracle's gen_1 (momentum 0.1 -> 0.9^60 = 2e-3 of the initial (0, 1) left) on seeded
    conditioning vectors; weights stay at their seeded initial values.  Returns the gen_1 state dict (fp32, CPU)."""
    p = params_on({"g": ps["gen_1"], "ca": ps["con_augment_1"]}, torch.float64, device)
    g = torch.Generator().manual_seed(777)
    with torch.no_grad():
        for _ in range(passes):
            tem = torch.randn(B, 512, generator=g).to(device, torch.float64)
            eps = torch.randn(B, 128, generator=g).to(device, torch.float64)
            z = torch.randn(B, O.Z_DIM, generator=g).to(device, torch.float64)
            c_hat, _, _ = O.ca_forward(p["ca"], tem, eps)
            O.g1_forward(p["g"], torch.cat((c_hat, z), dim=1), training=True)
    return {k: (v.float().cpu() if v.is_floating_point() else v.cpu()) for k, v in p["g"].items()}


def stage2(B, dtype=torch.float64, device="cuda", force=None, seed=0, g1_state=None):
    """One Stage-II outer step of the oracle at batch B on ``device``.  ``g1_state``: state dict to load into gen_1 first."""
    ps = O.init_all(42)
    if g1_state is not None:
        ps["gen_1"] = type(ps["gen_1"])((k, g1_state[k].clone()) for k in ps["gen_1"])
    p = params_on(ps, dtype, device)
    b = O.synthetic_batch(B, 2, seed)
    bd = batch_on(b, dtype, device)
    tr = dict(ca2=O.Trainer(p["con_augment_2"]), d2=O.Trainer(p["critic_2"]), g2=O.Trainer(p["gen_2"]))
    with strict_fp32():
        ref = O.stage2_step(p["con_augment_1"], p["gen_1"], p["con_augment_2"], p["critic_2"], p["gen_2"], bd["real"],
                            bd["tem"], bd["perm"], bd["z"], bd["eps_ca"], bd["eps_ca2"], bd["eps_gp"], tr, force=force)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    return b, ref
