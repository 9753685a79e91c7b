"""``train_2`` -- drop-in for the reference's ``stage_2_train_fn.py:20-239`` (with its two one-token
defects fixed: ``blob`` -> ``blob_1`` at :67, ``discriminator_2.py:28`` feeds ``img``).

Same call: ``train_2(models, optimizers, schedulers, loader, num_epochs, device, batch_size,
bucket_name=..., start_epoch=0, save_dir=...)`` with ``models = [textEncoder, projection_head,
con_augment_1, con_augment_2, gen_1, critic_2, gen_2]`` and three optimizers / schedulers
(con_augment_2, critic_2, gen_2) (:40-50).  The Stage-I networks are frozen and put in eval mode
(:52-63) and their weights come from the Stage-I checkpoint (:65-72; here ``stage1_checkpoint`` on local
disk, default ``./checkpoints/Stage1/latest_checkpoint_stage1.pth`` -- the path Stage-I actually writes;
the reference reads ``./checkpoint/Stage1/...``).  The per-batch work (:120-168) is ``Stage2Engine``.
TensorBoard image logging (:175-212) is out of scope.
"""
import os

import torch
import torch.distributed as dist

from .engine import N_CRITIC, LAMBDA_GP, Z_DIM, export_optimizer_state, import_optimizer_state
from .engine2 import Stage2Engine
from .stage_1_train_fn import _world, _rank, _adam_hyper, _shared_seed_stream

n_critic = N_CRITIC
lambda_gp = LAMBDA_GP
z_dim = Z_DIM


def train_2(models, optimizers, schedulers, loader, num_epochs, device, batch_size,
            bucket_name="data-and-checkpoints-bucket", start_epoch=0, save_dir="./checkpoints/Stage2",
            stage1_checkpoint="./checkpoints/Stage1/latest_checkpoint_stage1.pth", log=print, use_graph=True, engine=None,
            preview_every=100, preview_dir=None):
    textEncoder, projection_head, con_augment_1, con_augment_2, gen_1, critic_2, gen_2 = models
    opt_con_augment_2, opt_critic_2, opt_gen_2 = optimizers
    lr_scheduler_con_augment_2, lr_scheduler_critic_2, lr_scheduler_gen_2 = schedulers
    world, rank = _world(), _rank()

    for m in (textEncoder, projection_head, con_augment_1, gen_1):      # :52-63
        m.eval()
        for p in m.parameters():
            p.requires_grad = False
    loaded_stage1 = False
    if stage1_checkpoint:                                               # :65-72 (the reference loads it unconditionally)
        if not os.path.exists(stage1_checkpoint):
            raise FileNotFoundError(
                f"Stage-I checkpoint {stage1_checkpoint!r} not found: Stage-II would train against a randomly initialised, "
                "frozen gen_1 / con_augment_1 / text side.  Pass stage1_checkpoint=None to do that on purpose.")
        loaded_stage1 = True
        ck1 = torch.load(stage1_checkpoint, map_location="cpu", weights_only=False)
        textEncoder.load_state_dict(ck1["textEncoder"])
        projection_head.load_state_dict(ck1["projection_head"])
        con_augment_1.load_state_dict(ck1["con_augment_1"])
        gen_1.load_state_dict(ck1["gen_1"])
    checkpoint_path = os.path.join(save_dir, "latest_checkpoint_stage2.pth")
    resumed = False
    if os.path.exists(checkpoint_path):                                 # :74-92
        ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        start_epoch = ck["epoch"] + 1
        con_augment_2.load_state_dict(ck["con_augment_2"])
        critic_2.load_state_dict(ck["critic_2"])
        gen_2.load_state_dict(ck["gen_2"])
        for o, k in zip(optimizers, ("opt_con_augment_2", "opt_critic_2", "opt_gen_2")):    # :84-86
            o.load_state_dict(ck[k])
        for s, k in zip(schedulers, ("lr_scheduler_con_augment_2", "lr_scheduler_critic_2", "lr_scheduler_gen_2")):
            s.load_state_dict(ck[k])
        resumed = True
        log(f"Loaded checkpoint at epoch {start_epoch - 1}")
    con_augment_2.train(); critic_2.train(); gen_2.train()              # :96-98

    if engine is None:
        comm = None
        if world > 1:
            from .comm import make_comm
            from .engine import default_ops
            comm = make_comm(default_ops(), device=torch.device(device) if not isinstance(device, torch.device) else device)
        engine = Stage2Engine(con_augment_1, gen_1, con_augment_2, critic_2, gen_2, batch_size, comm=comm)
    eng = engine
    for fp, opt in ((eng.ca2.fp, opt_con_augment_2), (eng.d.fp, opt_critic_2), (eng.g2.fp, opt_gen_2)):
        lr, b1, b2, eps = _adam_hyper(opt)
        fp.hyper[:4] = torch.tensor([lr, b1, b2, eps], dtype=torch.float32)
        fp._lr_host = lr
        if resumed:
            import_optimizer_state(opt, fp)                    # Adam moments + step count of the checkpoint
    if resumed:
        eng.refresh_all()        # a caller-supplied engine packed its operands before the weights were loaded
    if loaded_stage1:
        eng.g1.refresh_weights()
    dev = eng.ops.device
    pin = lambda t: t.pin_memory() if not t.is_cuda else t

    seed_stream = _shared_seed_stream(world, dev)
    preview_step = 0                                                    # :33 (`step`)
    for epoch in range(start_epoch, num_epochs):
        for batch_idx, (tokenized_texts, real_img_256) in enumerate(loader):
            # pageable host tensors are copied synchronously with the stream (the host would wait for the previous
            # step on every batch): go through pinned memory
            tokenized_texts = {k: pin(v).to(dev, non_blocking=True) for k, v in tokenized_texts.items()}
            seed_t = torch.randint(0, 2 ** 32 - 1, (1,), generator=seed_stream)   # :105-113
            generator = torch.Generator().manual_seed(int(seed_t.item()))
            perm = torch.randperm(batch_size, generator=generator)     # :115-118
            perm_dev = pin(perm).to(dev, non_blocking=True)
            mismatched = {k: v[perm_dev] for k, v in tokenized_texts.items()}
            with torch.no_grad():                                       # text side is frozen
                tem = projection_head(textEncoder(**tokenized_texts).last_hidden_state[:, 0, :])      # :121-123
                tem_mis = projection_head(textEncoder(**mismatched).last_hidden_state[:, 0, :])       # :135-137
            z = pin(torch.randn(n_critic, batch_size, z_dim, generator=generator))                    # :126
            eps1 = torch.randn(n_critic, batch_size, con_augment_1.c_dim, device=dev)                 # con_augment.py:20
            eps2 = torch.randn(n_critic, batch_size, con_augment_2.c_dim, device=dev)
            eps_gp = pin(torch.rand(n_critic, batch_size))                                            # utils.py:10
            eng.step(real_img_256, tem.float(), tem_mis.float(), z, eps1, eps2, eps_gp, use_graph=use_graph)
            for s in schedulers:                                        # :170-173
                s.step()
            eng.d.fp.set_lr(opt_critic_2.param_groups[0]["lr"])
            eng.g2.fp.set_lr(opt_gen_2.param_groups[0]["lr"])
            eng.ca2.fp.set_lr(opt_con_augment_2.param_groups[0]["lr"])
            if rank == 0 and batch_idx % preview_every == 0 and batch_idx > 0:    # :175-212
                losses = eng.losses.tolist()
                log(f"Epoch [{epoch}/{num_epochs}] Batch {batch_idx}/{len(loader)} "
                    f"Loss D: {losses[0]:.4f}, loss G: {losses[2]:.4f}")
                # the fixed-noise preview of :181-195 (gen_2 in train mode, like the reference) and the two scalars of
                # :208-210; instead of TensorBoard event files on GCS (:36-38) they go to save_dir/previews
                fixed_generator = torch.Generator().manual_seed(456)                                   # :186
                fixed_noise = torch.randn(batch_size, z_dim, generator=fixed_generator)                # :187-189
                e1 = torch.randn(batch_size, con_augment_1.c_dim, device=dev)                          # con_augment.py:20
                e2 = torch.randn(batch_size, con_augment_2.c_dim, device=dev)
                fake_64, fake_256 = eng.preview(tem.float(), fixed_noise, e1, e2)
                pdir = preview_dir or os.path.join(save_dir, "previews")
                os.makedirs(pdir, exist_ok=True)
                img = fake_256[0]                                                                       # :200 (first image)
                lo, hi = img.min(), img.max()
                img = ((img - lo) / (hi - lo).clamp_min(1e-5)).cpu()                                    # make_grid(normalize=True)
                torch.save({"epoch": epoch, "batch": batch_idx, "step": preview_step, "fake_256": img,
                            "critic_2_loss": losses[0], "generator_2_loss": losses[2]},
                           os.path.join(pdir, f"preview_{preview_step:06d}.pt"))
                with open(os.path.join(pdir, "scalars.csv"), "a") as f:
                    f.write(f"{preview_step},{epoch},{batch_idx},{losses[0]},{losses[2]}\n")
                preview_step += 1                                                                       # :211
        if epoch % 10 == 0:
            eng.gather_optimizer_state()                                 # every rank: the Adam moments are sharded over replicas
        if rank == 0 and epoch % 10 == 0:                               # :214-235
            for opt, fp in ((opt_con_augment_2, eng.ca2.fp), (opt_critic_2, eng.d.fp), (opt_gen_2, eng.g2.fp)):
                export_optimizer_state(opt, fp)                # exp_avg / exp_avg_sq / step of the fused Adam
            checkpoint = {
                "con_augment_2": con_augment_2.state_dict(), "critic_2": critic_2.state_dict(), "gen_2": gen_2.state_dict(),
                "opt_con_augment_2": opt_con_augment_2.state_dict(), "opt_critic_2": opt_critic_2.state_dict(),
                "opt_gen_2": opt_gen_2.state_dict(),
                "lr_scheduler_con_augment_2": lr_scheduler_con_augment_2.state_dict(),
                "lr_scheduler_critic_2": lr_scheduler_critic_2.state_dict(),
                "lr_scheduler_gen_2": lr_scheduler_gen_2.state_dict(), "epoch": epoch,
            }
            os.makedirs(os.path.join(save_dir, "epochs"), exist_ok=True)
            torch.save(checkpoint, f"{save_dir}/epochs/checkpoint_epoch_{epoch}.pth")
            torch.save(checkpoint, checkpoint_path)
    return eng
