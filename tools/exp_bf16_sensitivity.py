"""Which bf16-stored tensors make the Stage-II gradients noisy?  Runs the engine's dataflow on the GPU in exact
arithmetic (tests/emu_ops.py) with bf16 storage for (a) everything, (b) forward tensors only, (c) backward tensors only,
and prints the relative L2 error of a few gradient tensors against the fp64 oracle.
    python tools/exp_bf16_sensitivity.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gpu_oracle as GO  # noqa: E402
import test_parity_config_gpu as T  # noqa: E402
from emu_ops import EmuOps  # noqa: E402

BWD_CRITIC = ["da", "dy", "gda", "gdy", "v", "w", "gy", "g", "v0", "dx", "Pv"]
BWD_G2 = ["da1", "dy0", "dX", "dz", "dpre", "Pd"]


def _conv(t, dt):
    return None if t is None else torch.empty(t.shape, dtype=dt, device=t.device)


def promote(eng, which, dt=torch.float64):
    """Re-allocate the forward ('fwd') or backward ('bwd') activation buffers of the Stage-II engine in ``dt``."""
    d, g2 = eng.d, eng.g2
    bn_objs = [g2.b2] + [b for blk in g2.rb for b in blk] + list(g2.ub)
    if which == "bwd":
        for name in BWD_CRITIC:
            val = getattr(d, name)
            setattr(d, name, [_conv(t, dt) for t in val] if isinstance(val, list) else _conv(val, dt))
        for name in BWD_G2:
            val = getattr(g2, name)
            setattr(g2, name, [_conv(t, dt) for t in val] if isinstance(val, list) else _conv(val, dt))
        for b in bn_objs:
            b.dy = _conv(b.dy, dt)
            b.da = _conv(b.da, dt)
    else:
        raise ValueError(which)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    b, ref = GO.stage2(B, torch.float64)
    want = {k: (v.detach().double().cpu() if torch.is_tensor(v) else v) for k, v in T._ref_table2(ref).items()}
    force = {"critic_before": ref["critic_before"]}
    del ref
    torch.cuda.empty_cache()
    import imagegenerator_b200.engine2 as E2
    keys = ["it0 fake_256", "it0 s_fake", "it0 dD2/down_sampler.0.weight", "it0 dD2/down_sampler.4.0.weight",
            "it0 dG2/up_sampler.3.weight", "it0 dG2/up_sampler.0.0.weight", "it0 dG2/residual_blocks.3.layer3.0.weight",
            "it0 dG2/residual_blocks.0.layer1.0.weight", "it0 dG2/down_sampler.2.0.weight", "it3 dD2/down_sampler.4.0.weight",
            "G dG2/up_sampler.0.0.weight", "G dG2/residual_blocks.0.layer1.0.weight"]
    results = {}
    for label, which in (("all bf16", None), ("fwd bf16, bwd exact", "bwd")):
        orig = E2.Stage2Engine.__init__

        def patched(self, *a, **k):
            orig(self, *a, **k)
            if which:
                promote(self, which)
        E2.Stage2Engine.__init__ = patched
        try:
            got = T._run_tf2(EmuOps(torch.bfloat16, device="cuda"), b, force)
        finally:
            E2.Stage2Engine.__init__ = orig
        results[label] = {k: (got[k] - want[k]).norm().item() / max(want[k].norm().item(), 1e-30) for k in keys}
        torch.cuda.empty_cache()
    print(f"B={B}   relative L2 error against the fp64 oracle")
    print(f"{'tensor':48s} " + " ".join(f"{l:>22s}" for l in results))
    for k in keys:
        print(f"{k:48s} " + " ".join(f"{results[l][k]:22.3e}" for l in results))


if __name__ == "__main__":
    main()
