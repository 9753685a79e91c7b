"""Epilogue-bound probe: 1x1 conv, K = 64 (one k-block per tile), N = 256: the kernel time is the accumulator drain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps
ops = CudaOps("bf16")
N, H, Ci, Co = 64, 64, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = (torch.randn(N, H, H, Ci, device="cuda") * 0.5).to(torch.bfloat16)
y = torch.empty(N, H, H, Co, device="cuda", dtype=torch.bfloat16)
w = torch.randn(Co, Ci, 1, 1, device="cuda") * 0.05
pf = ops.empty((Co, 1, 1, Ci)); ops.pack_weight(w, pf, None)
stats = torch.zeros(1, Co, 2, dtype=torch.float64, device="cuda")
M = N * H * H
for cg in (1, 2):
    for dbg in (0, 1):
        for st in (0, 1):
            ops.set_option("force_cg", cg); ops.set_option("force_bn", min(Co, 256)); ops.set_option("dbg", dbg)
            fn = (lambda: ops.conv_fprop_stats(x, pf, y, stats, 1, 1, 1, 0)) if st else (lambda: ops.conv_fprop(x, pf, None, y, 1, 1, 0))
            for _ in range(3): fn()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
            ms = min(ts)
            tiles_per_cta = (M / (128 * cg)) * max(1, Co // 256) / (148 // cg)
            print(f"cg={cg} stores={'off' if dbg else 'on '} stats={'on ' if st else 'off'}: {ms*1e3:7.1f} us  -> {ms*1e3/tiles_per_cta:5.2f} us per 128x{min(Co,256)} tile per CTA; "
                  f"out {M*Co*2/ms/1e9:6.2f} TB/s", flush=True)
