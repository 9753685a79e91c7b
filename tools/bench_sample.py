"""Stage-I -> Stage-II sampling throughput (BASELINE.json configs[4]: batch 512 sharded over the GPUs, forward-only).

    python tools/bench_sample.py [B_per_gpu] [steps]            (1 GPU)
    torchrun --nproc-per-node N ... tools/bench_sample.py ...   (N shards, no collective on the data path)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from imagegenerator_b200.con_augment import ConditioningAugmentation  # noqa: E402
from imagegenerator_b200.generator_1 import StageIGenerator  # noqa: E402
from imagegenerator_b200.generator_2 import StageIIGenerator  # noqa: E402
from imagegenerator_b200.ops import CudaOps  # noqa: E402
from imagegenerator_b200.sampler import StackGANSampler  # noqa: E402

F_CA, F_G1, F_G2 = 0.000393e9, 0.03207e9, 15.14563e9


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    ca1, g1 = ConditioningAugmentation(512, 256, 128), StageIGenerator(128, 100)
    ca2, g2 = ConditioningAugmentation(512, 256, 128), StageIIGenerator()
    smp = StackGANSampler(ca1, g1, ca2, g2, B, ops=CudaOps("bf16", device=dev))
    g = torch.Generator().manual_seed(100 + rank)
    tem = torch.randn(B, 512, generator=g).pin_memory()
    z, e1, e2 = (torch.randn(B, n, generator=g).pin_memory() for n in (100, 128, 128))
    host_out = torch.empty(B, 3, 256, 256).pin_memory()
    for _ in range(3):
        smp.sample(tem, z, e1, e2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # resident inputs: graph replay only
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        smp.graph.replay()
    e1_.record()
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1_) / steps
    # end to end: host embeddings/noise in, images back to pinned host memory
    e0.record()
    for _ in range(steps):
        _, img = smp.sample(tem, z, e1, e2)
        host_out.copy_(img, non_blocking=True)
    e1_.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1_) / steps
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    if rank == 0:
        flops = (2 * F_CA + F_G1 + F_G2) * B
        print(json.dumps({"metric": "stackgan_sampling_images_per_sec", "value": round(B * world / (ms_dev * 1e-3), 1),
                          "unit": "images/s", "n_gpus": world, "batch_per_gpu": B, "ms_per_batch": round(ms_dev, 3),
                          "tflops_per_gpu": round(flops / (ms_dev * 1e-3) / 1e12, 1),
                          "e2e": {"value": round(B * world / (ms_e2e * 1e-3), 1), "ms_per_batch": round(ms_e2e, 3),
                                  "d2h_bytes_per_step": B * 3 * 256 * 256 * 4},
                          "launches_per_batch": smp.launches, "dtype": "bf16", "bn": "eval, folded into the convs"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
