"""Steady-state pace of the persistent conv kernel: the same GEMM (N=256, K=2048) as a k4 s2 conv, a 3x3 s1 conv-like and a
1x1 conv, at growing M, so that fixed costs (launch, prologue, pipeline fill, last drain) separate from the per-tile pace.
    python tools/exp_conv_pace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from imagegenerator_b200.ops import CudaOps  # noqa: E402


def timeit(fn, flush, reps=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def main():
    ops = CudaOps("bf16")
    for kv in filter(None, os.environ.get("SG_OPTS", "").split(",")):
        key, val = kv.split("=")
        ops.set_option(key, int(val))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"{'case':34s} {'M':>8s} {'N':>5s} {'K':>6s} {'us':>9s} {'TFLOP/s':>9s}")
    cases = []
    for mult in (1, 2, 4, 8):
        cases.append((f"k4s2 16x16x128->8x8x256 x{mult}", 384 * mult, 16, 128, 256, 4, 2, 1))
    for mult in (1, 2, 4, 8):
        cases.append((f"k1 8x8x2048->256 x{mult}", 384 * mult, 8, 2048, 256, 1, 1, 0))
    for mult in (1, 2, 4):
        cases.append((f"k3s1 16x16x256->256 x{mult}", 96 * mult, 16, 256, 256, 3, 1, 1))
    for mult in (1, 4):
        cases.append((f"k4s2 32x32x320->640 (G2 up0) x{mult}", 64 * mult, 32, 320, 640, 4, 2, 1))
    for name, N, H, Ci, Co, k, s, p in cases:
        Ho = (H + 2 * p - k) // s + 1
        x = (torch.randn(N, H, H, Ci, device="cuda") * 0.5).to(torch.bfloat16)
        y = torch.empty(N, Ho, Ho, Co, device="cuda", dtype=torch.bfloat16)
        pf = (torch.randn(Co, k, k, Ci, device="cuda") * 0.05).to(torch.bfloat16)
        us = timeit(lambda: ops.conv_fprop(x, pf, None, y, k, s, p), flush)
        M, K = N * Ho * Ho, Ci * k * k
        print(f"{name:34s} {M:8d} {Co:5d} {K:6d} {us:9.1f} {2.0 * M * Co * K / us / 1e6:9.1f}", flush=True)
        del x, y, pf


if __name__ == "__main__":
    main()
