"""Why bf16 storage hurts Stage-II at initialisation: per BatchNorm'ed layer of gen_2, |batch mean| / batch std of the
pre-BN conv output -- the factor by which BatchNorm amplifies the relative rounding error of its (bf16) input.
Runs the oracle's gen_2 forward in fp64 on the GPU.    python tools/diag_g2_conditioning.py [fresh|handoff] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from oracle import stackgan_oracle as O  # noqa: E402
import gpu_oracle as GO  # noqa: E402


def main():
    state = sys.argv[1] if len(sys.argv) > 1 else "fresh"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ps = O.init_all(42)
    if state == "handoff":
        st = GO.trained_stats_g1(ps, device=dev)
        ps["gen_1"] = type(ps["gen_1"])((k, st[k].clone()) for k in ps["gen_1"])
    p = GO.params_on(ps, torch.float64, dev)
    b = GO.batch_on(O.synthetic_batch(B, 2, 0), torch.float64, dev)
    rows = []
    real_bn = F.batch_norm

    def spy(x, rm, rv, w, bias, training, mom, eps):
        if training and x.dim() == 4:
            m = x.mean((0, 2, 3))
            s = x.var((0, 2, 3), unbiased=False).sqrt()
            ratio = (m.abs() / s.clamp_min(1e-30))
            rows.append((tuple(x.shape[1:]), ratio.median().item(), ratio.max().item(), s.median().item()))
        return real_bn(x, rm, rv, w, bias, training, mom, eps)
    F.batch_norm = spy
    try:
        with torch.no_grad():
            c1, _, _ = O.ca_forward(p["con_augment_1"], b["tem"], b["eps_ca"][0])
            f64 = O.g1_forward(p["gen_1"], torch.cat((c1, b["z"][0]), 1), training=False)
            print(f"state {state}: fake_64 max|.| {f64.abs().max().item():.3f}, std over batch (mean over pixels) "
                  f"{f64.std(0).mean().item():.3e}, std over pixels {f64.std().item():.3e}")
            c2, _, _ = O.ca_forward(p["con_augment_2"], b["tem"], b["eps_ca2"][0])
            rows.clear()
            O.g2_forward(p["gen_2"], f64, c2, training=True)
    finally:
        F.batch_norm = real_bn
    print(f"{'pre-BN tensor (C,H,W)':28s} {'median |mu|/sigma':>18s} {'max |mu|/sigma':>15s} {'median sigma':>13s}")
    for shp, med, mx, s in rows:
        print(f"{str(shp):28s} {med:18.2f} {mx:15.1f} {s:13.3e}")


if __name__ == "__main__":
    main()
