#!/usr/bin/env python
"""bench.py -- StackGAN Stage-I train step (BASELINE.json configs[1]: 64x64, batch 128/GPU, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16|fp32]

A "step" is one reference outer step (stage_1_train_fn.py:116-172): five critic updates (WGAN-GP with
matching-aware negatives, double backward through the critic) + one generator/conditioning-
augmentation update, Adam included, on synthetic text embeddings and images.  One JSON line on rank 0.

  value   images/s, whole job, inputs resident in HBM, timed with CUDA events per step (L2 flushed
          between steps, flush not timed), max over ranks
  e2e     the same metric through the public ``train_1`` call with HOST (pinned) batches: H2D copies of
          images / embeddings / noise and the D2H loss read are inside the timed region
  roofline  dominant kernel timed live with CUDA events on its launch stream
  cpu_baseline  the UNMODIFIED reference ``train_1`` (sources carried by oracle/make_ref.py; under the torch_xla / GCS
          stubs of oracle/ref_harness.py) on the box's host cores, batch 128, a bounded number of steps; the
          ``stage2`` section carries the same for ``train_2`` at batch 8

``--impl reference`` times that CPU path alone with all host threads at the same batch / steps / warm-up.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = ("StackGAN Stage-I 64x64 G+D outer step (5 critic updates w/ WGAN-GP double backward "
            "+ 1 generator/CA update, Adam), batch 128/GPU")
F_D1, F_G1 = 0.21037e9, 0.03207e9          # forward FLOPs / image
# FLOPs this implementation must execute per image and outer step (DESIGN.md "Work per step"):
# critic iteration = G fwd + 3 trunk fwd (real, fake, interp; the mismatched call reuses real's
# features) + 3 trunk bwd (dgrad+wgrad = 2 F_D each) + GP first order (1 F_D) + GP second order
# (fprop chain + wgrad = 2 F_D)  = F_G + 12 F_D ;  G step = F_D fwd + F_D dgrad + 2 F_G bwd
FLOPS_PER_IMG = 5 * (F_G1 + 12 * F_D1) + (2 * F_D1 + 2 * F_G1)
FLOPS_PER_IMG_REFERENCE_NECESSARY = 77 * F_D1 + 7 * F_G1


_SAMPLER_SRC = r"""
import sys, time, os
idx = int(sys.argv[1])
try:
    import pynvml as n
    n.nvmlInit()
    h = n.nvmlDeviceGetHandleByIndex(idx)
    mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
    while True:
        try:
            sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
            rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            sm, rs = -1, 0
        sys.stdout.write("%.6f,%d,%d,%d\n" % (time.time(), sm, mx, rs))
        sys.stdout.flush()
        time.sleep(0.02)
except Exception as e:
    sys.stdout.write("ERR %r\n" % (e,))
    sys.stdout.flush()
"""


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.  A separate process polls NVML every 20 ms from before the
    warm-up on (a thread in this process starves on the GIL while the main thread enqueues launches flat out); ``stop``
    keeps the samples whose wall-clock stamps fall inside [mark_begin, mark_end]."""
    BITS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None

    def start(self):
        idx = self.index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=2)[0]
        except Exception:
            out = ""
        rows = []
        for line in out.splitlines():
            p = line.split(",")
            if len(p) == 4:
                try:
                    rows.append((float(p[0]), int(p[1]), int(p[2]), int(p[3])))
                except ValueError:
                    pass
        t0, t1 = self.t0 or 0.0, self.t1 or 1e30
        inside = [r for r in rows if t0 - 0.02 <= r[0] <= t1 + 0.02] or rows[-3:]
        sm = sorted(r[1] for r in inside if r[1] > 0)
        bits = 0
        for r in inside:
            bits |= r[3]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "reasons": sorted(k for k, b in self.BITS.items() if bits & b), "samples": len(inside),
                "source": "NVML polled every 20 ms by a side process; samples inside the timed region"}


def build_modules(seed=42):
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.discrminator_1 import StageIDiscriminator
    from imagegenerator_b200.generator_1 import StageIGenerator
    torch.manual_seed(seed)                                   # train.py:66
    return ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)


def synthetic_host_batches(n, B, seed):
    """Pinned host batches shaped like the reference loader's (dict, real_img_64) items."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        real = torch.randn(B, 3, 64, 64, generator=g).clamp_(-1, 1).pin_memory()
        idx = torch.arange(B)
        out.append(({"idx": idx}, real))
    return out


class TableEncoder(torch.nn.Module):
    """Synthetic text side: ``encoder(idx=...)`` returns rows of a fixed embedding table as the CLS state."""

    def __init__(self, table):
        super().__init__()
        self.register_buffer("table", table)
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, idx):
        class _O:
            pass
        o = _O()
        o.last_hidden_state = self.table[idx][:, None, :]
        return o


class IdentityHead(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return x


# ------------------------------------------------------------------------------------------------ CPU reference
class _TimedLoader:
    """The loader handed to the reference's train function: yields pre-built host batches and stamps the wall clock at
    every ``__next__`` -- the time between two stamps is one whole outer step of the reference's loop body."""

    def __init__(self, batches):
        self.batches, self.stamps = batches, []

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        for b in self.batches:
            self.stamps.append(time.perf_counter())
            yield b
        self.stamps.append(time.perf_counter())

    def step_seconds(self):
        return [b - a for a, b in zip(self.stamps[:-1], self.stamps[1:])]


def cpu_reference_train(stage, B, steps, warmup, threads):
    """The reference's OWN train function on the host cores.

    kind "reference": the unmodified ``train_1`` (``train_2`` with the two one-token fixes of SURVEY.md section 0) compiled
    from the reference's sources -- /root/reference in the build container, the archive packed by oracle/make_ref.py on the
    GPU box -- under the torch_xla / google.cloud.storage / SummaryWriter stubs of oracle/ref_harness.py, synthetic text
    table, no checkpoint I/O in the timed region.
    kind "port": the oracle's restatement (oracle/stackgan_oracle.py), only when neither source is present.
    Returns dict(ips, mean_s, kind, steps_s)."""
    from oracle import ref_harness as H
    from oracle import stackgan_oracle as O
    torch.set_num_threads(threads)
    n = warmup + steps
    if not H.reference_available():
        ps = O.init_all(42, with_stage2=(stage == 2))
        times = []
        if stage == 1:
            ca, d1, g1 = ps["con_augment_1"], ps["critic_1"], ps["gen_1"]
            tr = dict(ca=O.Trainer(ca), d1=O.Trainer(d1), g1=O.Trainer(g1))
        else:
            tr = dict(ca2=O.Trainer(ps["con_augment_2"]), d2=O.Trainer(ps["critic_2"]), g2=O.Trainer(ps["gen_2"]))
        for s in range(n):
            b = O.synthetic_batch(B, stage, s)
            t0 = time.perf_counter()
            if stage == 1:
                O.stage1_step(ca, d1, g1, b["real"], b["tem"].clone().requires_grad_(True), b["perm"], b["z"], b["eps_ca"],
                              b["eps_gp"], tr)
            else:
                O.stage2_step(ps["con_augment_1"], ps["gen_1"], ps["con_augment_2"], ps["critic_2"], ps["gen_2"], b["real"],
                              b["tem"], b["perm"], b["z"], b["eps_ca"], b["eps_ca2"], b["eps_gp"], tr)
            times.append(time.perf_counter() - t0)
        ts = times[warmup:]
        mean = sum(ts) / len(ts)
        return dict(ips=B / mean, mean_s=mean, kind="port", steps_s=ts)
    # ---- the unmodified reference
    import random
    H.reset_store()
    H.seed_everything(1234)
    torch.manual_seed(42)                                                    # train.py:66
    CA = H.load("con_augment").ConditioningAugmentation
    ca1 = CA(512, 256, 128)
    d1 = H.load("discrminator_1").StageIDiscriminator(512, 128)
    g1 = H.load("generator_1").StageIGenerator(128, 100)
    gd = torch.Generator().manual_seed(0)
    hw = 64 if stage == 1 else 256
    table = torch.randn(B, 512, generator=gd)
    enc, head = H.TableEncoder(table), H.IdentityHead()
    batches = [({"idx": torch.arange(B)}, torch.randn(B, 3, hw, hw, generator=gd).clamp_(-1, 1)) for _ in range(n)]
    loader = _TimedLoader(batches)
    mk = lambda m, lr=1e-3: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.999))       # train.py:92-102
    sch = lambda o: torch.optim.lr_scheduler.StepLR(o, step_size=100, gamma=0.5)               # train.py:105-113
    if stage == 1:
        opts = [mk(enc, 0.0), mk(head, 0.0), mk(ca1), mk(d1), mk(g1)]
        fn = H.load("stage_1_train_fn").train_1
        models = [enc, head, ca1, d1, g1]
    else:
        ca2 = CA(512, 256, 128)
        d2 = H.load("discriminator_2").StageIIDiscriminator(512, 128)
        g2 = H.load("generator_2").StageIIGenerator()
        import tempfile
        with tempfile.NamedTemporaryFile() as tmp:                           # the Stage-I checkpoint train_2 loads (:65-72)
            torch.save(dict(textEncoder=enc.state_dict(), projection_head=head.state_dict(),
                            con_augment_1=ca1.state_dict(), gen_1=g1.state_dict()), tmp.name)
            with open(tmp.name, "rb") as f:
                H.store()["./checkpoint/Stage1/latest_checkpoint_stage1.pth"] = f.read()
        opts = [mk(ca2), mk(d2), mk(g2)]
        fn = H.load("stage_2_train_fn").train_2
        models = [enc, head, ca1, ca2, g1, d2, g2]
        random.seed(99)
    scheds = [sch(o) for o in opts]
    with H.quiet():
        # epoch 1 of 2: one pass over the loader, and no checkpoint upload (the reference saves when epoch % 10 == 0)
        fn(models, opts, scheds, loader, 2, "cpu", B, start_epoch=1)
    ts = loader.step_seconds()[warmup:]
    mean = sum(ts) / len(ts)
    return dict(ips=B / mean, mean_s=mean, kind="reference", steps_s=ts)


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    """``--impl reference``: the reference's own CPU path on this arm's config -- Stage-I, batch 128 per step, the same
    --steps / --warmup.  A batch-128 outer step is ~2.1 TFLOP as executed (~3 s on 16 host cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B, K, W = args.batch, args.steps, args.warmup
    r = cpu_reference_train(1, B, K, W, threads)
    ips, mean = r["ips"], r["mean_s"]
    what = ("unmodified reference train_1 (stage_1_train_fn.py:19-240 under the torch_xla / GCS stubs, synthetic text table)"
            if r["kind"] == "reference" else "oracle port (torch-CPU restatement of stage_1_train_fn.py:93-196)")
    line = {
        "impl": "reference", "metric": "stackgan_stage1_train_images_per_sec", "value": round(ips, 3),
        "unit": "images/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": round(mean * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B, "parallelism": "cpu",
                   "note": "the reference's own CPU path on the box's host cores at the SAME batch (128) and step counts as "
                           "the GPU arm; one process (the reference's data parallelism is one process per TPU core)"},
        "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": threads, "kind": r["kind"],
                         "cpu_model": _cpu_model(), "torch_threads": torch.get_num_threads(), "torch": torch.__version__,
                         "sample": f"{what}: B={B} fp32, {W} warm-up + {K} timed outer steps, mean"},
        "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ secondary workloads
F_D2, F_G2, F_CA = 0.36412e9, 15.14563e9, 0.000393e9
# executed per image and Stage-II outer step: 5 x (G1 fwd + G2 fwd/dgrad/wgrad (3 F_G2) + critic 12 F_D2 + d/d image F_D2)
# + generator step (2 F_D2 + 2 F_G2);  reference-necessary: 77 F_D2 + 17 F_G2 + 5 F_G1
FLOPS2_PER_IMG = 5 * (F_G1 + 3 * F_G2 + 13 * F_D2) + (2 * F_D2 + 2 * F_G2)
FLOPS2_PER_IMG_REFERENCE_NECESSARY = 77 * F_D2 + 17 * F_G2 + 5 * F_G1


def run_stage2(ops, comm, world, rank, dev, steps=5, warmup=3, B=64):
    """BASELINE.json configs[2]/[3]: Stage-II 256x256 outer step, batch 64 per GPU, bf16 (frozen Stage-I generator)."""
    import torch.distributed as dist
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.discriminator_2 import StageIIDiscriminator
    from imagegenerator_b200.generator_1 import StageIGenerator
    from imagegenerator_b200.generator_2 import StageIIGenerator
    from imagegenerator_b200.engine2 import Stage2Engine
    torch.manual_seed(42)
    ca1, g1 = ConditioningAugmentation(512, 256, 128), StageIGenerator(128, 100)
    ca2, d2, g2 = ConditioningAugmentation(512, 256, 128), StageIIDiscriminator(512, 128), StageIIGenerator()
    eng = Stage2Engine(ca1, g1, ca2, d2, g2, B, ops=ops, comm=comm)
    g = torch.Generator().manual_seed(3000 + rank)
    real = torch.randn(B, 3, 256, 256, generator=g).clamp_(-1, 1).to(dev)
    tem = torch.randn(B, 512, generator=g).to(dev)
    tem_mis = tem[torch.randperm(B, generator=g).to(dev)].contiguous()
    gz = torch.Generator().manual_seed(5)
    z = torch.randn(5, B, 100, generator=gz).to(dev)
    e1, e2 = torch.randn(5, B, 128, generator=gz).to(dev), torch.randn(5, B, 128, generator=gz).to(dev)
    egp = torch.rand(5, B, generator=gz).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        eng.step(real, tem, tem_mis, z, e1, e2, egp, use_graph=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.step(real, tem, tem_mis, z, e1, e2, egp, use_graph=True); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # ---- end to end through train_2 with pinned host batches (image upload + noise inside the timed region)
    from imagegenerator_b200.stage_2_train_fn import train_2
    table = torch.randn(B, 512, generator=torch.Generator().manual_seed(5000 + rank)).to(dev)
    enc, head = TableEncoder(table).to(dev), IdentityHead().to(dev)
    mk = lambda m_: torch.optim.Adam(m_.parameters(), lr=1e-3, betas=(0.9, 0.999))
    opts = [mk(ca2), mk(d2), mk(g2)]
    scheds = [torch.optim.lr_scheduler.StepLR(o, step_size=100, gamma=0.5) for o in opts]

    def host_batches(n, seed):
        gg = torch.Generator().manual_seed(seed)
        return [({"idx": torch.arange(B)}, torch.randn(B, 3, 256, 256, generator=gg).clamp_(-1, 1).pin_memory()) for _ in range(n)]
    ck_dir = f"/tmp/sgb200_bench_ckpt2_{os.getpid()}"
    quiet = lambda *a, **k: None
    args2 = dict(start_epoch=1, save_dir=ck_dir, stage1_checkpoint=None, log=quiet, use_graph=True, engine=eng)
    train_2([enc, head, ca1, ca2, g1, d2, g2], opts, scheds, host_batches(2, 1), 2, dev, B, **args2)     # warm-up pass
    hb = host_batches(steps, 2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    train_2([enc, head, ca1, ca2, g1, d2, g2], opts, scheds, hb, 2, dev, B, **args2)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    out = {"metric": "stackgan_stage2_train_images_per_sec", "value": round(B * world / (ms * 1e-3), 2), "unit": "images/s",
           "e2e": {"value": round(B * world / (e2e_ms * 1e-3), 2), "unit": "images/s", "ms_per_step": round(e2e_ms, 3),
                   "h2d_bytes_per_step": B * 3 * 256 * 256 * 4 + B * 8 + 5 * B * 100 * 4 + 5 * B * 4, "d2h_bytes_per_step": 0,
                   "api": "imagegenerator_b200.stage_2_train_fn.train_2 (loss read every 100 batches, like the reference)"},
           "ms_per_step": round(ms, 3), "batch_per_gpu": B, "steps": steps, "warmup": warmup,
           "step_tflops_per_gpu": round(FLOPS2_PER_IMG * B / (ms * 1e-3) / 1e12, 1),
           "flops_per_image_executed": FLOPS2_PER_IMG, "flops_per_image_reference_necessary": FLOPS2_PER_IMG_REFERENCE_NECESSARY,
           "gpu_launches_per_step": eng.launches_per_step, "mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)}
    del eng
    torch.cuda.empty_cache()
    return out


def run_sampling(ops, world, rank, dev, reps=20, B=64):
    """BASELINE.json configs[4]: Stage-I -> Stage-II forward-only sampling, batch sharded over the GPUs."""
    import torch.distributed as dist
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.generator_1 import StageIGenerator
    from imagegenerator_b200.generator_2 import StageIIGenerator
    from imagegenerator_b200.sampler import StackGANSampler
    torch.manual_seed(42)
    smp = StackGANSampler(ConditioningAugmentation(512, 256, 128), StageIGenerator(128, 100),
                          ConditioningAugmentation(512, 256, 128), StageIIGenerator(), B, ops=ops)
    g = torch.Generator().manual_seed(4000 + rank)
    tem = torch.randn(B, 512, generator=g).pin_memory()
    z, e1, e2 = (torch.randn(B, n, generator=g).pin_memory() for n in (100, 128, 128))
    host_out = [torch.empty(B, 3, 256, 256).pin_memory() for _ in range(2)]
    for i in range(3):
        smp.sample_to_host(tem, z, e1, e2, host_out[i & 1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        smp.graph.replay()
    b.record()
    torch.cuda.synchronize()
    ms_dev = a.elapsed_time(b) / reps
    a.record()
    # end to end: host embeddings in (pinned), fp32 NCHW images back into pinned host memory; the read-back of batch k rides
    # a copy stream under the generation of batch k+1 (StackGANSampler.sample_to_host); the closing event waits for the last copy
    torch.cuda.synchronize()
    a.record()
    for i in range(reps):
        ev = smp.sample_to_host(tem, z, e1, e2, host_out[i & 1])
    torch.cuda.current_stream().wait_event(ev)       # the copy stream is in order: the last read-back done = all done
    b.record()
    torch.cuda.synchronize()
    ms_e2e = a.elapsed_time(b) / reps
    # the same end to end with uint8 pictures (out_dtype="uint8": a quarter of the read-back)
    launches = smp.launches
    del smp
    torch.manual_seed(42)
    smp8 = StackGANSampler(ConditioningAugmentation(512, 256, 128), StageIGenerator(128, 100),
                           ConditioningAugmentation(512, 256, 128), StageIIGenerator(), B, ops=ops, out_dtype="uint8")
    host8 = [torch.empty(B, 3, 256, 256, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i in range(3):
        smp8.sample_to_host(tem, z, e1, e2, host8[i & 1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a.record()
    for i in range(reps):
        ev = smp8.sample_to_host(tem, z, e1, e2, host8[i & 1])
    torch.cuda.current_stream().wait_event(ev)
    b.record()
    torch.cuda.synchronize()
    ms_u8 = a.elapsed_time(b) / reps
    t = torch.tensor([ms_dev, ms_e2e, ms_u8], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_u8 = t.tolist()
    flops = (2 * F_CA + F_G1 + F_G2) * B
    return {"metric": "stackgan_sampling_images_per_sec", "value": round(B * world / (ms_dev * 1e-3), 1), "unit": "images/s",
            "batch_per_gpu": B, "ms_per_batch": round(ms_dev, 3), "tflops_per_gpu": round(flops / (ms_dev * 1e-3) / 1e12, 1),
            "e2e": {"value": round(B * world / (ms_e2e * 1e-3), 1), "ms_per_batch": round(ms_e2e, 3),
                    "h2d_bytes_per_step": B * (512 + 100 + 256) * 4, "d2h_bytes_per_step": B * 3 * 256 * 256 * 4,
                    "api": "StackGANSampler.sample_to_host (fp32 NCHW images into pinned host memory, read-back overlapped on a copy stream)"},
            "e2e_uint8": {"value": round(B * world / (ms_u8 * 1e-3), 1), "ms_per_batch": round(ms_u8, 3),
                          "h2d_bytes_per_step": B * (512 + 100 + 256) * 4, "d2h_bytes_per_step": B * 3 * 256 * 256,
                          "api": "StackGANSampler(out_dtype='uint8').sample_to_host (NCHW uint8 pictures, round((x + 1) * 127.5))"},
            "gpu_launches_per_batch": launches, "bn": "eval mode, folded into the packed conv weights"}


def time_hbm_kernel(ops, peaks, reps=10):
    """Dominant HBM-bound kernel of the Stage-II step: BatchNorm backward (apply) on the generator's 64x128x128x80
    activation (generator_2.py:55 `up2` block): three 168 MB tensors in, one out, each exactly once."""
    rows, C = 64 * 128 * 128, 80
    mk = lambda: torch.randn(rows, C, device=ops.device).to(ops.act_dtype)
    da, a, y, dy = mk(), mk(), mk(), mk()
    mr = torch.rand(1, C, 2, device=ops.device) + 0.5
    gamma = torch.rand(C, device=ops.device) + 0.5
    sums = torch.zeros(1, C, 2, dtype=torch.float64, device=ops.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ops.device)
    fn = lambda: ops.bn_bwd_apply(da, a, y, mr, gamma, sums, dy, 1, 1)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    byts = 4.0 * rows * C * 2
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = byts / (ms * 1e-3) / 1e9
    prof_file = "profiles/ncu_bn_r2_summary.txt"
    prof = read_profile_metrics(os.path.join(ROOT, prof_file), "bn_bwd_apply8_kernel")
    traffic = (int(prof["dram__bytes_read.sum"] + prof["dram__bytes_write.sum"])
               if "dram__bytes_read.sum" in prof and "dram__bytes_write.sum" in prof else None)
    return {"bound": "hbm", "kernel": "sg_bn_bwd_apply (bn_bwd_apply8_kernel) on the 64x128x128x80 bf16 activation of G2 up2",
            "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
            "traffic": traffic, "traffic_unit": f"B per launch (dram__bytes_read.sum + dram__bytes_write.sum parsed from {prof_file}; "
                                                "algorithmic 671.1 MB, the tail of the output is still in L2 at kernel end)",
            "algorithmic_bytes": int(byts), "kernel_ms": round(ms, 5),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"}


# ------------------------------------------------------------------------------------------------ kernel roofline
def time_dominant_kernel(ops, B, reps=20, sets=8):
    """The critic's heaviest conv (ds3: 128->256, 16x16 -> 8x8, all three image groups batched) timed alone with CUDA
    events on its launch stream; algorithmic FLOPs = 2*M*N*K.  Two figures:

    * ``ms``: average launch duration over ``reps * sets`` launches queued back to back between ONE pair of events -- the
      way the kernel runs inside the captured step.  The launches walk through ``sets`` separate (input, weights, output)
      buffer sets, 38.8 MB each: 8 sets = 310 MB > the 126 MB L2, so every launch finds its operands in HBM, not in L2;
    * ``ms_isolated``: one launch between its own pair of events after a 256 MiB L2 flush (this adds the ~2 us an event
      pair around a single short launch costs; the ncu duration of the same launch sits between the two figures)."""
    N, H, Ci, Co, k = 3 * B, 16, 128, 256, 4
    xs = [torch.randn(N, H, H, Ci, device="cuda").to(ops.act_dtype) for _ in range(sets)]
    pfs = [(torch.randn(Co, k, k, Ci, device="cuda") * 0.02).to(ops.act_dtype) for _ in range(sets)]
    ys = [torch.empty(N, H // 2, H // 2, Co, device="cuda", dtype=ops.act_dtype) for _ in range(sets)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(sets):
        ops.conv_fprop(xs[i], pfs[i], None, ys[i], k, 2, 1)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fprop(xs[0], pfs[0], None, ys[0], k, 2, 1)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms_iso = sum(ts) / len(ts)
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(sets):
            ops.conv_fprop(xs[i], pfs[i], None, ys[i], k, 2, 1)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / (reps * sets)
    flops = 2.0 * (N * (H // 2) ** 2) * Co * (k * k * Ci)
    return flops, ms, ms_iso


def read_profile_metrics(path, kernel_substr):
    """Metrics of the first launch of ``kernel_substr`` in a profiles/ncu_*_summary.txt file (tools/ncu_summary.py format:
    `  metric   value unit` lines under `-- launch N`).  Returns {} when the file is missing or holds no such launch."""
    out, active = {}, False
    try:
        with open(path) as f:
            for ln in f:
                if ln.startswith("-- launch"):
                    if out:
                        break
                    active = False
                    continue
                parts = ln.split()
                if len(parts) >= 2 and parts[0] == "Kernel" and parts[1] == "Name":
                    active = kernel_substr in ln
                    continue
                if active and len(parts) >= 2:
                    try:
                        out[parts[0]] = float(parts[1])
                    except ValueError:
                        pass
    except OSError:
        return {}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Stage-II and sampling sections of the JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from imagegenerator_b200.engine import Stage1Engine
    from imagegenerator_b200.ops import CudaOps
    from imagegenerator_b200.stage_1_train_fn import train_1
    from imagegenerator_b200.comm import make_comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    W = max(args.warmup, 3)
    K, B = args.steps, args.batch
    dev = torch.device(f"cuda:{local}")
    ops = CudaOps(args.mode, device=dev)
    ca, d1, g1 = build_modules()
    comm = make_comm(ops, device=dev) if world > 1 else None
    eng = Stage1Engine(ca, d1, g1, B, ops=ops, world_size=world, comm=comm)
    use_graph = not args.no_graph

    # ---- resident-input timing
    g = torch.Generator().manual_seed(1000 + rank)           # each replica its own images/embeddings
    real = torch.randn(B, 3, 64, 64, generator=g).clamp_(-1, 1).to(dev)
    tem = torch.randn(B, 512, generator=g).to(dev)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(7))
    tem_mis = tem[perm.to(dev)].contiguous()
    gz = torch.Generator().manual_seed(5)                     # shared noise (same z on every replica, like the reference)
    z = torch.randn(5, B, 100, generator=gz).to(dev)
    eca = torch.randn(5, B, 128, generator=gz).to(dev)
    egp = torch.rand(5, B, generator=gz).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dbg(msg):
        if os.environ.get("SG_BENCH_DEBUG"):
            print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # a side process, started before the warm-up so that it is polling when the timed region begins
    ok = 1
    if use_graph:
        try:
            eng.step(real, tem, tem_mis, z, eca, egp, use_graph=True)      # captures, then replays once
            torch.cuda.synchronize()
        except Exception as e:                                # e.g. NCCL refusing stream capture
            print(f"[bench r{rank}] CUDA-graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            ok = 0
        dbg(f"capture ok={ok}")
        if world > 1:                                         # all ranks must agree, or collectives mismatch
            t_ok = torch.tensor([ok], device=dev)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            ok = int(t_ok.item())
        if not ok:
            use_graph = False
            eng.graph = None
    for _ in range(W):
        eng.step(real, tem, tem_mis, z, eca, egp, use_graph=use_graph)
    dbg("warm-up done")
    barrier()
    sampler.mark_begin()
    n0 = ops.launch_count()
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(K):
        flush.zero_()                                         # L2 flush between timed steps (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step(real, tem, tem_mis, z, eca, egp, use_graph=use_graph)
        e1.record()
        evs.append((e0, e1))
        if os.environ.get("SG_BENCH_SYNC") == "1":
            e1.synchronize()
    barrier()
    sampler.mark_end()
    wall = time.perf_counter() - t_wall0
    step_ms = sum(a.elapsed_time(b) for a, b in evs) / K
    dbg("per-step ms: " + " ".join(f"{a.elapsed_time(b):.2f}" for a, b in evs))
    launches = (eng.launches_per_step or 0) * K if use_graph else ops.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([step_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = B * world / (step_ms * 1e-3)

    # ---- end to end through train_1 with host batches
    table = torch.randn(B, 512, generator=torch.Generator().manual_seed(2000 + rank)).to(dev)
    enc, head = TableEncoder(table).to(dev), IdentityHead().to(dev)
    mk = lambda m, lr=1e-3: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.999))
    opts = [mk(enc, 0.0), mk(head, 0.0), mk(ca), mk(d1), mk(g1)]
    scheds = [torch.optim.lr_scheduler.StepLR(o, step_size=100, gamma=0.5) for o in opts]
    quiet = lambda *a, **k: None
    ck_dir = f"/tmp/sgb200_bench_ckpt_{os.getpid()}"
    train_1([enc, head, ca, d1, g1], opts, scheds, synthetic_host_batches(2, B, 1), 2, dev, B, start_epoch=1,
            save_dir=ck_dir, log=quiet, use_graph=use_graph, engine=eng)      # warm-up pass
    batches = synthetic_host_batches(K, B, 2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    train_1([enc, head, ca, d1, g1], opts, scheds, batches, 2, dev, B, start_epoch=1, save_dir=ck_dir, log=quiet,
            use_graph=use_graph, engine=eng)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = B * 3 * 64 * 64 * 4 + B * 8 + 5 * B * 100 * 4 + 5 * B * 4       # images + idx + z + gp eps (fp32)
    d2h = 4 * 4

    # ---- data-parallel check on the hardware path just timed: after W + 2K + 2 outer steps (graph segments + NCCL) every
    # parameter must be bit-identical on all ranks (xm.optimizer_step semantics: same averaged gradient, same Adam state)
    dp_check = None
    if world > 1:
        worst = torch.zeros(1, device=dev, dtype=torch.float64)
        for fp in (eng.d.fp, eng.g.fp, eng.ca.fp):
            root = fp.flat.clone()
            dist.broadcast(root, 0)
            worst = torch.maximum(worst, (fp.flat.double() - root.double()).abs().max().reshape(1))
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        finite = torch.tensor([float(all(torch.isfinite(fp.flat).all().item() for fp in (eng.d.fp, eng.g.fp, eng.ca.fp)))], device=dev)
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
        dp_check = {"params_bit_identical_across_ranks": bool(worst.item() == 0.0), "max_abs_diff": float(worst.item()),
                    "params_finite": bool(finite.item() == 1.0), "outer_steps": W + 2 * K + 3,
                    "oracle_check": "tests/dp_nccl_worker.py (averaged-gradient oracle, 2 ranks)"}

    # gradient bytes every replica contributes per outer step: 5 critic steps + the generator's + the conditioning augmentation's
    s1_bytes_per_step = 4 * (5 * eng.d.fp.grad.numel() + eng.g.fp.grad.numel() + eng.ca.fp.grad.numel()) if comm else 0
    extras = {}
    if not args.no_extras and args.mode == "bf16":
        del eng
        torch.cuda.empty_cache()
        try:
            if rank == 0:
                pk = {}
                try:
                    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                        pk = json.load(f)
                except Exception:
                    pass
                extras["roofline_hbm"] = time_hbm_kernel(ops, pk)
                torch.cuda.empty_cache()
            extras["stage2"] = run_stage2(ops, comm, world, rank, dev)
            extras["sampling"] = run_sampling(ops, world, rank, dev)
        except Exception as e:                                    # secondary sections must not take the headline down
            extras["extras_error"] = f"{type(e).__name__}: {e}"
            if world > 1:
                raise

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        kflops, kms, kms_iso = time_dominant_kernel(ops, B)
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        ach = kflops / (kms * 1e-3) / 1e12
        prof_file = "profiles/ncu_conv_tcp_r2_summary.txt"
        prof = read_profile_metrics(os.path.join(ROOT, prof_file), "conv_tcp_kernel")
        traffic = (int(prof["dram__bytes_read.sum"] + prof["dram__bytes_write.sum"])
                   if "dram__bytes_read.sum" in prof and "dram__bytes_write.sum" in prof else None)
        line = {
            "metric": "stackgan_stage1_train_images_per_sec", "value": round(value, 2), "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(step_ms, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "cuda_graph": use_graph, "grad_allreduce": (("fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (sg_dp_adam_step), "
                                                                     "one launch per optimizer step inside the CUDA graph, " if comm.peer else
                                                                     "NCCL avg, one all-reduce of the flat gradient buffer per optimizer step, ") +
                                                                    f"{s1_bytes_per_step} B of gradients/step") if comm else "none (1 GPU)", "l2": "flushed between timed steps (256 MiB write, untimed)",
                       "flops_per_image_executed": FLOPS_PER_IMG,
                       "flops_per_image_reference_necessary": FLOPS_PER_IMG_REFERENCE_NECESSARY},
            "step_tflops": round(FLOPS_PER_IMG * B / (step_ms * 1e-3) / 1e12, 2),
            "wall_s_timed_region": round(wall, 3),
            "gpu_launches": int(launches),
            "e2e": {"value": round(B * world / (e2e_ms * 1e-3), 2), "unit": "images/s", "ms_per_step": round(e2e_ms, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": "imagegenerator_b200.stage_1_train_fn.train_1"},
            "roofline": {"bound": "tensor", "kernel": "sg_conv_fprop critic ds3 (128->256, k4 s2, 3 groups batched), conv_tcp_kernel<2>",
                         "achieved": round(ach, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4),
                         "traffic": traffic,
                         "traffic_unit": f"B per launch (dram__bytes_read.sum + dram__bytes_write.sum parsed from {prof_file}; "
                                         "algorithmic 38.8 MB incl. the 12.6 MB result, which is still in L2 at kernel end)",
                         "tensor_pipe_active_pct_of_elapsed": prof.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                         "ncu_duration_us": prof.get("gpu__time_duration.sum"),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1.59 PF",
                         "kernel_ms": round(kms, 5),
                         "timing": "average launch duration of 160 back-to-back launches between one event pair, rotating through 8 "
                                   "buffer sets (310 MB > L2: operands come from HBM)",
                         "kernel_ms_isolated": round(kms_iso, 5),
                         "frac_isolated": round(kflops / (kms_iso * 1e-3) / 1e12 / peak_tf, 4)},
            "clocks": clocks,
        }
        if dp_check is not None:
            line["dp_check"] = dp_check
        line.update(extras)
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            r = cpu_reference_train(1, B, 3, 1, threads)
            line["cpu_baseline"] = {"value": round(r["ips"], 3), "unit": "images/s", "cores": threads, "kind": r["kind"],
                                    "cpu_model": _cpu_model(),
                                    "sample": f"{'unmodified reference train_1' if r['kind'] == 'reference' else 'oracle port'} on the host "
                                              f"cores, Stage-I B={B} fp32, 1 warm-up + 3 timed outer steps, mean "
                                              f"({r['mean_s']:.2f} s/step)"}
            if "stage2" in line:
                B2 = 8
                r2 = cpu_reference_train(2, B2, 2, 1, threads)
                line["stage2"]["cpu_baseline"] = {
                    "value": round(r2["ips"], 4), "unit": "images/s", "cores": threads, "kind": r2["kind"],
                    "sample": f"{'reference train_2 (two one-token fixes)' if r2['kind'] == 'reference' else 'oracle port'} on the host "
                              f"cores, Stage-II B={B2} fp32 (the GPU figure is at B=64; a B=64 step is ~18 TFLOP), 1 warm-up + 2 "
                              f"timed outer steps, mean ({r2['mean_s']:.1f} s/step)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
