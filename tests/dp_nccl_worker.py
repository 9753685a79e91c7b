"""torchrun entry of tests/test_dp_nccl_gpu.py: the data-parallel Stage-I step over NCCL on real GPUs, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/dp_nccl_worker.py [fp32|bf16]

Each rank holds its own images / embeddings and the shared noise (stage_1_train_fn.py:98-121).  Checked on hardware:
  1. the gradient the critic's optimizer sees in iteration 0 is the MEAN over replicas of the per-replica gradients
     (xm.optimizer_step, stage_1_train_fn.py:149) -- against the oracle run on the same GPU in fp64 with an
     all-reduce hook in front of every optimizer step;
  2. after whole outer steps (CUDA-graph segments + NCCL, the path bench.py times) every parameter is BIT-IDENTICAL on
     all ranks, BatchNorm running statistics are not (no SyncBN in the reference);
  3. the post-step parameters agree with that oracle within Adam's |dw| <= lr per step amplification bound.
Prints one line ``DP_NCCL_OK ...`` on rank 0; any failure raises on its rank.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    from oracle import stackgan_oracle as O
    import gpu_oracle as GO
    from imagegenerator_b200.comm import make_comm
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.discrminator_1 import StageIDiscriminator
    from imagegenerator_b200.generator_1 import StageIGenerator
    from imagegenerator_b200.engine import Stage1Engine
    from imagegenerator_b200.ops import CudaOps

    B, dt = 16, torch.float64
    torch.manual_seed(42 + rank)          # rank-dependent init on purpose: the broadcast from rank 0 must fix it
    ca, d1, g1 = ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)
    ps = GO.params_on(O.init_all(42, with_stage2=False), dt, dev)          # what rank 0 holds
    mine = GO.batch_on(O.synthetic_batch(B, 1, 100 + rank), dt, dev)       # per-replica images / embeddings
    shared = GO.batch_on(O.synthetic_batch(B, 1, 7), dt, dev)              # same z / eps on every replica
    tr = dict(ca=O.Trainer(ps["con_augment_1"]), d1=O.Trainer(ps["critic_1"]), g1=O.Trainer(ps["gen_1"]))

    def sync(t):
        for p in t.params.values():
            dist.all_reduce(p.grad)
            p.grad.div_(world)
    tem = mine["tem"].clone().requires_grad_(True)
    ref = O.stage1_step(ps["con_augment_1"], ps["critic_1"], ps["gen_1"], mine["real"], tem, shared["perm"], shared["z"],
                        shared["eps_ca"], shared["eps_gp"], tr, sync=sync)
    # the oracle returns the per-replica gradient of iteration 0 (taken before the hook ran): average it here
    want0 = {}
    for k, v in ref["critic_grads"][0].items():
        g = v.clone()
        dist.all_reduce(g)
        want0[k] = g / world

    ops = CudaOps(mode, device=dev)
    comm = make_comm(ops, device=dev)
    comm.write_avg = True          # PeerComm: also leave the averaged gradient in every replica's .grad (this check reads it)
    eng = Stage1Engine(ca, d1, g1, B, ops=ops, world_size=world, comm=comm)
    f = lambda t: t.float().contiguous()
    real, temf, tem_mis = f(mine["real"]), f(mine["tem"]), f(mine["tem"][shared["perm"]])
    z, eca, egp = f(shared["z"]), f(shared["eps_ca"]), f(shared["eps_gp"])
    # 1. averaged gradient of iteration 0 (eager path: all-reduce, then Adam; the buffer keeps the averaged gradient)
    eng.load_batch(real, temf, tem_mis)
    eng.critic_iteration(z[0], eca[0], egp[0])
    torch.cuda.synchronize()
    worst_g = 0.0
    tol_g = 2e-3 if mode == "fp32" else 0.25
    for k, p in d1.named_parameters():
        w = want0[k]
        if w.norm().item() / w.numel() ** 0.5 < 1e-12:
            continue                                   # compress.*: exactly zero in exact arithmetic
        rel = (p.grad.double() - w).norm().item() / w.norm().item()
        worst_g = max(worst_g, rel)
        assert rel < tol_g, (mode, k, rel)
    # a single-replica gradient must NOT pass for the averaged one (the check bites)
    k0 = "down_sampler.2.0.weight"
    own = ref["critic_grads"][0][k0]
    assert (own - want0[k0]).norm().item() / want0[k0].norm().item() > 10 * tol_g or mode != "fp32", "replicas too similar to tell"

    # 2./3. whole outer steps through the graph-segment path, from a fresh state
    torch.manual_seed(42 + rank)
    ca, d1, g1 = ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)
    comm.write_avg = False
    eng = Stage1Engine(ca, d1, g1, B, ops=ops, world_size=world, comm=comm)
    eng.step(real, temf, tem_mis, z, eca, egp, use_graph=True)
    torch.cuda.synchronize()
    worst_p, lr = 0.0, 1e-3
    for m, key, steps in ((ca, "ca", 1), (d1, "d1", 5), (g1, "g1", 1)):
        sd = m.state_dict()
        for k, v in ref["after"][key].items():
            if not v.is_floating_point() or O.is_buffer(k):
                continue
            err = (sd[k].double() - v).abs().max().item()
            worst_p = max(worst_p, err / (lr * steps))
            assert err <= 2.2 * lr * steps + 5e-2 * v.abs().max().item(), (key, k, err)
    eng.step(real, temf, tem_mis, z, eca, egp, use_graph=True)         # a second step = a pure replay of the segments
    torch.cuda.synchronize()
    for fp, name in ((eng.d.fp, "critic"), (eng.g.fp, "gen"), (eng.ca.fp, "ca")):
        mineflat = fp.flat.clone()
        root = mineflat.clone()
        dist.broadcast(root, 0)
        assert torch.equal(mineflat, root), f"{name}: parameters differ between rank {rank} and rank 0 after 2 steps"
    rm = d1.down_sampler[2][1].running_mean.clone()
    rm0 = rm.clone()
    dist.broadcast(rm0, 0)
    if rank == 1:
        assert not torch.equal(rm, rm0), "BatchNorm running statistics must stay per replica"
    if comm.peer:
        comm.check()
    dist.barrier()
    if rank == 0:
        print(f"DP_NCCL_OK transport={'peer' if comm.peer else 'nccl'} mode={mode} world={world} worst_rel_l2_avg_grad={worst_g:.3e} worst_param_err_over_lr_steps={worst_p:.3f}",
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
