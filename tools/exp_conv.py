"""Experiment driver: time one conv shape under several kernel options.  python tools/exp_conv.py name dir"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps
from tools.bench_conv import SHAPES

name, d = sys.argv[1], sys.argv[2]
cfgs = [dict(persist=0)] + [dict(persist=1, force_cg=cg, force_bn=bn, force_stages=st, dbg=dbg)
                            for cg, bn, st, dbg in eval(sys.argv[3])]
ops = CudaOps("bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nm, N, H, Ci, Co, k, s, p in SHAPES:
    if nm != name:
        continue
    Ho = (H + 2 * p - k) // s + 1
    x = (torch.randn(N, H, H, Ci, device="cuda") * 0.5).to(torch.bfloat16)
    y = (torch.randn(N, Ho, Ho, Co, device="cuda") * 0.5).to(torch.bfloat16)
    w = torch.randn(Co, Ci, k, k, device="cuda") * 0.05
    pf, pd = ops.empty((Co, k, k, Ci)), ops.empty((Ci, k, k, Co))
    ops.pack_weight(w, pf, pd)
    flops = 2.0 * N * Ho * Ho * Co * Ci * k * k
    fn = (lambda: ops.conv_fprop(x, pf, None, y, k, s, p)) if d == "fprop" else (lambda: ops.conv_dgrad(y, pd, None, x, k, s, p))
    for cfg in cfgs:
        for kk in ("persist", "force_cg", "force_bn", "force_stages", "dbg"):
            ops.set_option(kk, cfg.get(kk, 0))
        for _ in range(2):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        print(f"{name} {d} {cfg}: {ms*1e3:.1f} us  {flops/ms/1e9:.0f} TFLOP/s  (min {min(ts)*1e3:.1f})", flush=True)
