"""Per-layer timing of the three conv kernels on the StackGAN shapes (dashboard for kernel work).

    python tools/bench_conv.py [filter-substring] [reps]

Directions: fprop,dgrad,wgrad (default) and wgrad_cl (channels-last accumulation, the 3x3 residual convs' path).
Each line: layer, direction, M x N x K of the implicit GEMM, mean time over ``reps`` launches with the L2
flushed between launches (CUDA events on the launch stream), algorithmic TFLOP/s (2*M*N*K) and bytes/s
of the compulsory traffic (each operand/result once)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from imagegenerator_b200.ops import CudaOps  # noqa: E402

# name, N, H, Ci, Co, k, s, p   (Conv2d orientation: x [N,H,H,Ci] -> y [N,Ho,Ho,Co])
B1, B2 = 128, 64
SHAPES = [
    ("s1.D1.ds0p", 3 * B1, 32, 48, 64, 1, 1, 0),
    ("s1.D1.ds0", 3 * B1, 64, 3, 64, 4, 2, 1),
    ("s1.G1.up4", B1, 64, 3, 24, 4, 2, 1),
    ("s1.D1.ds2", 3 * B1, 32, 64, 128, 4, 2, 1),
    ("s1.D1.ds3", 3 * B1, 16, 128, 256, 4, 2, 1),
    ("s1.D1.ds4", 3 * B1, 8, 256, 512, 4, 2, 1),
    ("s1.D1.ds3g1", B1, 16, 128, 256, 4, 2, 1),
    ("s1.G1.up1", B1, 8, 96, 192, 4, 2, 1),
    ("s1.G1.up2", B1, 16, 48, 96, 4, 2, 1),
    ("s1.G1.up3", B1, 32, 24, 48, 4, 2, 1),
    ("s2.G2.ds2", B2, 32, 128, 512, 4, 2, 1),
    ("s2.G2.res1", B2, 16, 640, 320, 3, 1, 1),
    ("s2.G2.res2", B2, 16, 320, 320, 3, 1, 1),
    ("s2.G2.res3", B2, 16, 320, 640, 3, 1, 1),
    ("s2.G2.up0", B2, 32, 320, 640, 4, 2, 1),
    ("s2.G2.up1", B2, 64, 160, 320, 4, 2, 1),
    ("s2.G2.up2", B2, 128, 80, 160, 4, 2, 1),
    ("s2.G2.up3", B2, 256, 3, 80, 4, 2, 1),
    ("s2.D2.ds0p", 3 * B2, 128, 48, 16, 1, 1, 0),
    ("s2.D2.ds0", 3 * B2, 256, 3, 16, 4, 2, 1),
    ("s2.D2.ds2", 3 * B2, 128, 16, 32, 4, 2, 1),
    ("s2.D2.ds3", 3 * B2, 64, 32, 64, 4, 2, 1),
    ("s2.D2.ds4", 3 * B2, 32, 64, 128, 4, 2, 1),
    ("s2.D2.ds5", 3 * B2, 16, 128, 256, 4, 2, 1),
    ("s2.D2.ds6", 3 * B2, 8, 256, 512, 4, 2, 1),
]


def main():
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dirs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fprop", "dgrad", "wgrad"]
    ops = CudaOps("bf16")
    for kv in filter(None, os.environ.get("SG_OPTS", "").split(",")):      # e.g. SG_OPTS=force_cg=1,force_stages=4
        key, val = kv.split("=")
        ops.set_option(key, int(val))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"{'layer':14s} {'dir':6s} {'M':>8s} {'N':>5s} {'K':>6s} {'us':>9s} {'TFLOP/s':>9s} {'GB/s':>8s}")
    for name, N, H, Ci, Co, k, s, p in SHAPES:
        if filt and filt not in name:
            continue
        Ho = (H + 2 * p - k) // s + 1
        x = (torch.randn(N, H, H, Ci, device="cuda") * 0.5).to(torch.bfloat16)
        y = (torch.randn(N, Ho, Ho, Co, device="cuda") * 0.5).to(torch.bfloat16)
        w = torch.randn(Co, Ci, k, k, device="cuda") * 0.05
        pf, pd = ops.empty((Co, k, k, Ci)), ops.empty((Ci, k, k, Co))
        ops.pack_weight(w, pf, pd)
        dw = torch.zeros_like(w)
        flops = 2.0 * N * Ho * Ho * Co * Ci * k * k
        byts = 2.0 * (x.numel() + y.numel() + w.numel())
        for d in dirs:
            if d == "fprop":
                fn, M, Nn, K = (lambda: ops.conv_fprop(x, pf, None, y, k, s, p)), N * Ho * Ho, Co, Ci * k * k
            elif d == "dgrad":
                fn, M, Nn, K = (lambda: ops.conv_dgrad(y, pd, None, x, k, s, p)), N * H * H, Ci, Co * k * k // (s * s)
            elif d == "wgrad_cl":      # channels-last accumulation buffer (what engine2 uses for the 3x3 residual convs)
                if not ops.conv_wgrad_cl_supported(x, y, k, s, p):
                    continue
                gw = torch.zeros(Co, k, k, Ci, device="cuda")
                fn, M, Nn, K = (lambda: ops.conv_wgrad_cl(x, y, gw, k, s, p)), Co, Ci * k * k, N * Ho * Ho
            else:
                fn, M, Nn, K = (lambda: ops.conv_wgrad(x, y, dw, k, s, p)), Co, Ci * k * k, N * Ho * Ho
            try:
                for _ in range(2):
                    fn()
                ts = []
                for _ in range(reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = sum(ts) / len(ts)
                print(f"{name:14s} {d:6s} {M:8d} {Nn:5d} {K:6d} {ms * 1e3:9.1f} {flops / ms / 1e9:9.1f} {byts / ms / 1e6:8.0f}",
                      flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{name:14s} {d:6s} FAILED {e}", flush=True)
        del x, y, w, pf, pd, dw


if __name__ == "__main__":
    main()
