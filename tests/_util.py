"""Shared helpers for the parity tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def digest(t, n=64):
    """Must match oracle/make_golden.py:digest."""
    t = t.detach().to(torch.float64).reshape(-1).cpu()
    g = torch.Generator().manual_seed(t.numel() % 9973 + 17)
    idx = torch.randint(0, t.numel(), (min(n, t.numel()),), generator=g)
    return dict(numel=t.numel(), sum=t.sum().item(), norm=t.norm().item(), idx=idx, vals=t[idx].clone())


def assert_digest(t, dg, rtol, atol, what=""):
    """Compare tensor ``t`` with a stored digest: sampled elements elementwise, L2 norm relatively."""
    t = t.detach().to(torch.float64).reshape(-1).cpu()
    assert t.numel() == dg["numel"], (what, t.numel(), dg["numel"])
    got = t[dg["idx"]]
    scale = max(dg["norm"] / max(dg["numel"], 1) ** 0.5, 1e-30)   # rms of the tensor
    ok = torch.allclose(got, dg["vals"], rtol=rtol, atol=atol)
    if not ok:
        err = (got - dg["vals"]).abs().max().item()
        raise AssertionError(f"{what}: sampled elements differ, max abs err {err:.3e} (rms {scale:.3e}, rtol {rtol}, atol {atol})")
    nerr = abs(t.norm().item() - dg["norm"])
    assert nerr <= rtol * dg["norm"] + atol * dg["numel"] ** 0.5, f"{what}: norm {t.norm().item():.6e} vs {dg['norm']:.6e}"


def assert_digest_dict(d, dgs, rtol, atol, what=""):
    for k, dg in dgs.items():
        assert k in d, (what, k)
        assert_digest(d[k], dg, rtol, atol, f"{what}[{k}]")
