"""Experiment: direct_tc.cu with windows that start inside a swizzle atom (option dtc_diag bit 8 = two / one staged copies instead
of four / three; bit 16 = with the descriptor's base-offset field).  Compares against the shipped kernel and times both."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from imagegenerator_b200.ops import CudaOps  # noqa: E402

ops = CudaOps("bf16")
N, H, Ci, Co = 192, 128, 16, 32
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, H, Ci, generator=g).to(torch.bfloat16).cuda()
pf = (torch.randn(Co, 4, 4, Ci, generator=g) * 0.06).to(torch.bfloat16).cuda()
dy = torch.randn(N, H // 2, H // 2, Co, generator=g).to(torch.bfloat16).cuda()
pd = (torch.randn(Ci, 4, 4, Co, generator=g) * 0.09).to(torch.bfloat16).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(diag):
    ops.set_option("dtc_diag", diag)
    y, dx = ops.empty((N, H // 2, H // 2, Co)), ops.empty((N, H, H, Ci))
    ts = []
    for fn in (lambda: ops.conv_narrow_fprop(x, pf, None, y), lambda: ops.conv_narrow_dgrad(dy, pd, None, dx)):
        fn(); torch.cuda.synchronize()
        t = 0.0
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            t += a.elapsed_time(b)
        ts.append(t * 100)
    ops.set_option("dtc_diag", 0)
    return y.float(), dx.float(), ts


y0, dx0, t0 = run(0)
print(f"shipped            fprop {t0[0]:6.1f} us  dgrad {t0[1]:6.1f} us")
for diag, name in ((8, "windows in atom   "), (24, "  + base offset   ")):
    y1, dx1, t1 = run(diag)
    ey = (y1 - y0).abs().max().item() / y0.abs().max().item()
    ed = (dx1 - dx0).abs().max().item() / dx0.abs().max().item()
    print(f"{name} fprop {t1[0]:6.1f} us  dgrad {t1[1]:6.1f} us   max rel diff vs shipped: fprop {ey:.2e}  dgrad {ed:.2e}")
