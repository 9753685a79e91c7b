"""Data-parallel semantics on CPU: world_size 2, gloo.  Each rank runs the engine (kernel emulator, fp64)
through ``comm.DistComm`` on ITS OWN images/embeddings with the SHARED noise, next to an oracle replica
whose optimizer steps are preceded by a gradient average -- the reference's xm.optimizer_step
(stage_1_train_fn.py:149,166-172) with per-replica BatchNorm statistics."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        from oracle import stackgan_oracle as O
        from emu_ops import EmuOps
        from imagegenerator_b200.comm import DistComm
        from imagegenerator_b200.con_augment import ConditioningAugmentation
        from imagegenerator_b200.discrminator_1 import StageIDiscriminator
        from imagegenerator_b200.generator_1 import StageIGenerator
        from imagegenerator_b200.engine import Stage1Engine
        dt, B = torch.float64, 2
        # rank-dependent init on purpose: the broadcast from rank 0 must make replicas identical
        torch.manual_seed(42 + rank)
        ca, d1, g1 = ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)
        ps = O.init_all(42, with_stage2=False)               # what rank 0 holds
        pca, pd1, pg1 = (O.to_dtype(ps[k], dt) for k in ("con_augment_1", "critic_1", "gen_1"))
        mine = O.synthetic_batch(B, 1, 100 + rank, dtype=dt)  # per-replica images / embeddings
        shared = O.synthetic_batch(B, 1, 7, dtype=dt)         # same z / eps on every replica (stage_1_train_fn.py:98-121)
        tr = dict(ca=O.Trainer(pca), d1=O.Trainer(pd1), g1=O.Trainer(pg1))

        def sync(t):
            for p in t.params.values():
                dist.all_reduce(p.grad)
                p.grad.div_(world)
        tem = mine["tem"].clone().requires_grad_(True)
        ref = O.stage1_step(pca, pd1, pg1, mine["real"], tem, shared["perm"], shared["z"], shared["eps_ca"],
                            shared["eps_gp"], tr, sync=sync)
        eng = Stage1Engine(ca, d1, g1, B, ops=EmuOps(dt), comm=DistComm())
        eng.load_batch(mine["real"], mine["tem"], mine["tem"][shared["perm"]])
        eng.outer_step(shared["z"], shared["eps_ca"], shared["eps_gp"])
        worst = 0.0
        for m, key in ((ca, "ca"), (d1, "d1"), (g1, "g1")):
            sd = m.state_dict()
            for k, v in ref["after"][key].items():
                if v.is_floating_point():
                    err = (sd[k].double() - v).abs().max().item() / max(v.abs().max().item(), 1e-12)
                    worst = max(worst, err)
                    assert err < 1e-6, (key, k, err)
        # replicas hold identical parameters after the step, different BN running stats (no SyncBN)
        w = d1.down_sampler[2][0].weight.data.clone()
        w0 = w.clone()
        dist.broadcast(w0, 0)
        assert torch.equal(w, w0)
        rm = d1.down_sampler[2][1].running_mean.clone()
        rm0 = rm.clone()
        dist.broadcast(rm0, 0)
        if rank == 1:
            assert not torch.equal(rm, rm0)
        q.put((rank, "ok", worst))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _worker2(rank, world, port, q):
    """Stage-II: the generator's gradients accumulate over the five critic backward passes and are averaged over
    replicas once, right before its step (stage_2_train_fn.py:154-168)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        from oracle import stackgan_oracle as O
        from emu_ops import EmuOps
        from imagegenerator_b200.comm import DistComm
        from imagegenerator_b200.engine2 import Stage2Engine
        from test_engine2_emulated import build_all
        dt, B = torch.float64, 1
        ms = build_all(42 + rank)                              # rank-dependent init: the broadcast must fix it
        ps = O.init_all(42)                                    # what rank 0 holds
        p = {k: O.to_dtype(ps[k], dt) for k in ps}
        mine = O.synthetic_batch(B, 2, 200 + rank, dtype=dt)
        shared = O.synthetic_batch(B, 2, 9, dtype=dt)
        tr = dict(ca2=O.Trainer(p["con_augment_2"]), d2=O.Trainer(p["critic_2"]), g2=O.Trainer(p["gen_2"]))

        def sync(t):
            for q_ in t.params.values():
                dist.all_reduce(q_.grad)
                q_.grad.div_(world)
        ref = O.stage2_step(p["con_augment_1"], p["gen_1"], p["con_augment_2"], p["critic_2"], p["gen_2"], mine["real"],
                            mine["tem"], shared["perm"], shared["z"], shared["eps_ca"], shared["eps_ca2"], shared["eps_gp"],
                            tr, sync=sync)
        eng = Stage2Engine(ms["ca1"], ms["g1"], ms["ca2"], ms["d2"], ms["g2"], B, ops=EmuOps(dt), comm=DistComm())
        eng.load_batch(mine["real"], mine["tem"], mine["tem"][shared["perm"]])
        eng.outer_step(shared["z"], shared["eps_ca"], shared["eps_ca2"], shared["eps_gp"])
        worst = 0.0
        for m, key in ((ms["ca2"], "ca2"), (ms["d2"], "d2"), (ms["g2"], "g2")):
            sd = m.state_dict()
            for k, v in ref["after"][key].items():
                if v.is_floating_point():
                    err = (sd[k].double() - v).abs().max().item() / max(v.abs().max().item(), 1e-12)
                    worst = max(worst, err)
                    assert err < 1e-5, (key, k, err)
        q.put((rank, "ok", worst))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run_world2(worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"


@pytest.mark.slow
def test_stage2_data_parallel_world2_gloo():
    _run_world2(_worker2)


@pytest.mark.slow
def test_stage1_data_parallel_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"


def _worker_loader(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tempfile
        from _util import make_coco_dir
        from imagegenerator_b200.data_loader import get_loader
        from imagegenerator_b200.train import image_transform
        # every rank writes the same (seeded) directory for itself; the captions are unique strings
        with tempfile.TemporaryDirectory() as tmp:
            root, ann, tok, rows = make_coco_dir(tmp, n_images=6, captions_per_image=2)
            loader = get_loader("b", root, ann, image_transform(64), batch_size=2, shuffle=True, tokenizer=tok, num_workers=0)
            assert len(loader) == 3                                   # 12 captions / 2 ranks / batch 2
            mine = []
            for tokenized, imgs in loader:
                assert imgs.shape == (2, 3, 64, 64)
                mine += [tuple(r.tolist()) for r in tokenized["input_ids"]]
        got = [None] * world
        dist.all_gather_object(got, mine)
        a, b = got
        enc = tok([c for c, _ in rows], padding="max_length", truncation=True, max_length=128, return_tensors="pt")["input_ids"]
        every = sorted(tuple(r.tolist()) for r in enc)
        # DistributedSampler over torch.distributed ranks: the two shards are disjoint and together cover the dataset
        assert len(a) == len(b) == 6 and sorted(a + b) == every, (len(a), len(b))
        q.put((rank, "ok", ""))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_caption_loader_shards_over_ranks_world2_gloo():
    _run_world2(_worker_loader)
