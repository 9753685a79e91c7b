"""``train_1`` -- drop-in for the reference's ``stage_1_train_fn.py:19-238``.

Same call: ``train_1(models, optimizers, schedulers, loader, num_epochs, device, batch_size,
start_epoch=0, bucket_name=..., save_dir=...)`` with ``models = [textEncoder, projection_head,
con_augment_1, critic_1, gen_1]`` and five optimizers / schedulers in that order (:37-53); the loader
yields ``(dict_of_tensors, real_img_64)`` (:93).  What changes underneath:

  * the five critic updates + generator/CA update of a batch (:116-172) run as ONE CUDA-graph replay
    of hand-written sm_100a kernels (``Stage1Engine``); autograd is only used for the caller's text
    encoder / projection head, which receive d lossG / d tem from the kernels (:162-171);
  * ``xm.optimizer_step`` (gradient all-reduce over replicas + step) becomes an NCCL all-reduce of the
    flat gradient buffer + a fused Adam kernel; the per-batch seed all-reduce (:98-105) becomes ONE broadcast of
    the master's base seed per call, from which every rank derives the same per-batch seeds (``_shared_seed_stream``);
  * checkpoints (:211-238) keep the reference's dictionary keys but go to ``save_dir`` on local disk
    instead of a GCS bucket (``bucket_name`` is accepted and ignored);
  * LR schedulers are stepped on every rank (the reference steps them on rank 0 only, :187-192, which
    lets replicas' learning rates diverge -- identical for world size 1).
"""
import os

import torch
import torch.distributed as dist

from .engine import Stage1Engine, N_CRITIC, LAMBDA_GP, Z_DIM

n_critic = N_CRITIC     # stage_1_train_fn.py:14
lambda_gp = LAMBDA_GP   # :15
z_dim = Z_DIM           # :16


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _adam_hyper(opt):
    g = opt.param_groups[0]
    return g["lr"], g["betas"][0], g["betas"][1], g["eps"]


def _shared_seed_stream(world, dev):
    """Where the per-batch seeds come from.  The reference draws one on the master and all-reduces it to every replica on
    EVERY batch (stage_1_train_fn.py:98-105) -- a collective plus a device-to-host read in the middle of the step.  Here the
    master's RNG is shared ONCE per call: rank 0 draws a base seed, it is broadcast, and every rank then draws the per-batch
    seeds from its own copy of that generator -- identical on all ranks with no further communication.  World size 1 keeps
    the reference's draw from the global RNG (None), so seeded single-process runs consume the RNG exactly like the
    reference does."""
    if world == 1:
        return None
    base = torch.randint(0, 2 ** 32 - 1, (1,)).to(dev)
    dist.broadcast(base, 0)
    return torch.Generator().manual_seed(int(base.item()))


def make_allreduce(world):
    """Gradient mean over replicas right before each optimizer step (xm.optimizer_step semantics)."""
    if world == 1:
        return None

    def allreduce(flat_grad):
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.mul_(1.0 / world)
    return allreduce


def _report(pending, log, rank, num_epochs, n_batches):
    epoch, batch_idx, host, ev = pending
    ev.synchronize()
    if rank == 0:
        log(f"Epoch [{epoch}/{num_epochs}] Batch {batch_idx}/{n_batches} "
            f"Loss D: {float(host[0]):.4f}, loss G: {float(host[2]):.4f}")


def train_1(models, optimizers, schedulers, loader, num_epochs, device, batch_size, start_epoch=0,
            bucket_name="data-and-checkpoints-bucket", save_dir="./checkpoints/Stage1",
            log=print, use_graph=True, engine=None):
    textEncoder, projection_head, con_augment_1, critic_1, gen_1 = models
    opt_encoder, opt_projection_head, opt_con_augment_1, opt_critic_1, opt_gen_1 = optimizers
    (lr_scheduler_encoder, lr_scheduler_projection_head, lr_scheduler_con_augment_1,
     lr_scheduler_critic_1, lr_scheduler_gen_1) = schedulers
    world, rank = _world(), _rank()

    checkpoint_path = os.path.join(save_dir, "latest_checkpoint_stage1.pth")
    resumed = False
    if os.path.exists(checkpoint_path):                       # :55-82
        ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        start_epoch = ck["epoch"] + 1
        for m, k in ((textEncoder, "textEncoder"), (projection_head, "projection_head"),
                     (con_augment_1, "con_augment_1"), (critic_1, "critic_1"), (gen_1, "gen_1")):
            m.load_state_dict(ck[k])
        for o, k in zip(optimizers, ("opt_encoder", "opt_projection_head", "opt_con_augment_1", "opt_critic_1",
                                     "opt_gen_1")):        # :69-73; the fused-Adam moments are taken over below
            o.load_state_dict(ck[k])
        for s, k in zip(schedulers, ("lr_scheduler_encoder", "lr_scheduler_projection_head",
                                     "lr_scheduler_con_augment_1", "lr_scheduler_critic_1", "lr_scheduler_gen_1")):
            s.load_state_dict(ck[k])
        resumed = True
        log(f"Loaded checkpoint at epoch {start_epoch - 1}")

    for m in models:
        m.train()                                              # :86-90
    if engine is None:
        from .comm import make_comm
        from .engine import default_ops
        comm = make_comm(default_ops(), device=torch.device(device) if not isinstance(device, torch.device) else device)
        engine = Stage1Engine(con_augment_1, critic_1, gen_1, batch_size, world_size=world, comm=comm)
    eng = engine
    for fp, opt in ((eng.ca.fp, opt_con_augment_1), (eng.d.fp, opt_critic_1), (eng.g.fp, opt_gen_1)):
        lr, b1, b2, eps = _adam_hyper(opt)
        fp.hyper[:4] = torch.tensor([lr, b1, b2, eps], dtype=torch.float32)
        fp._lr_host = lr
        if resumed:
            eng.import_optimizer_state(opt, fp)                # Adam moments + step count of the checkpoint
    if resumed:
        eng.refresh_all()        # a caller-supplied engine packed its bf16 operands / collapsed head before the weights were loaded
    dev = eng.ops.device
    pin = lambda t: t.pin_memory() if not t.is_cuda else t

    seed_stream = _shared_seed_stream(world, dev)
    loss_bufs = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = None
    for epoch in range(start_epoch, num_epochs):
        for batch_idx, (tokenized_texts, real_img_64) in enumerate(loader):
            # pageable host tensors are copied synchronously with the stream (the host would wait for the previous
            # step on every batch): go through pinned memory
            tokenized_texts = {k: pin(v).to(dev, non_blocking=True) for k, v in tokenized_texts.items()}
            seed_t = torch.randint(0, 2 ** 32 - 1, (1,), generator=seed_stream)   # :98-105: the master's seed for every replica
            generator = torch.Generator().manual_seed(int(seed_t.item()))
            perm = torch.randperm(batch_size, generator=generator)        # :108-111
            perm_dev = pin(perm).to(dev, non_blocking=True)
            mismatched = {k: v[perm_dev] for k, v in tokenized_texts.items()}

            # DELIBERATE DEVIATION (documented): the reference re-encodes the captions inside each of the five critic
            # iterations (:117-129) with the encoder in train mode, so an encoder WITH dropout gives five different `tem`; here the
            # text is encoded once per outer step and the same `tem` / `tem_mis` feed all five iterations and the generator
            # update (d lossG / d tem flows through that one dropout mask).  Identical for a dropout-free encoder, which is
            # what the parity tests pin; the encoder is outside SURVEY section 8's hot path (it runs on stock PyTorch).
            tem = projection_head(textEncoder(**tokenized_texts).last_hidden_state[:, 0, :])     # :117-119
            with torch.no_grad():
                tem_mis = projection_head(textEncoder(**mismatched).last_hidden_state[:, 0, :])  # :127-129
            # the noise the reference draws inside the loop (:121, con_augment.py:20, utils.py:10)
            z = pin(torch.randn(n_critic, batch_size, z_dim, generator=generator))
            eps_ca = torch.randn(n_critic, batch_size, con_augment_1.c_dim, device=dev)
            eps_gp = pin(torch.rand(n_critic, batch_size))

            eng.step(real_img_64, tem.detach().float(), tem_mis.float(), z, eps_ca, eps_gp, use_graph=use_graph)

            if tem.requires_grad:                              # :161-171 encoder / projection-head update
                opt_encoder.zero_grad()
                opt_projection_head.zero_grad()
                tem.backward(eng.d.dtem.to(tem.dtype))
                for opt in (opt_encoder, opt_projection_head):
                    if world > 1:
                        for g in opt.param_groups:
                            for p in g["params"]:
                                if p.grad is not None:
                                    dist.all_reduce(p.grad)
                                    p.grad.div_(world)
                    opt.step()

            # the reference prints both losses after every batch (:178-181), a device -> host read.  Here the read is a
            # non-blocking copy into pinned memory; the line for batch k is printed while batch k+1 is already running
            # on the GPU (and after the loop for the last one), so the host never stalls the stream
            if pending is not None:
                _report(pending, log, rank, num_epochs, len(loader))
            host = loss_bufs[batch_idx % 2]
            host.copy_(eng.losses, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            pending = (epoch, batch_idx, host, ev)
            for s in schedulers:                               # :187-192 (per batch)
                s.step()
            eng.d.fp.set_lr(opt_critic_1.param_groups[0]["lr"])
            eng.g.fp.set_lr(opt_gen_1.param_groups[0]["lr"])
            eng.ca.fp.set_lr(opt_con_augment_1.param_groups[0]["lr"])

        if pending is not None:                                # last batch of the epoch
            _report(pending, log, rank, num_epochs, len(loader))
            pending = None
        if epoch % 10 == 0:
            eng.gather_optimizer_state()                       # every rank: the Adam moments are sharded over replicas
        if rank == 0 and epoch % 10 == 0:                     # :211-238
            eng.export_optimizer_state(opt_con_augment_1, eng.ca.fp)
            eng.export_optimizer_state(opt_critic_1, eng.d.fp)
            eng.export_optimizer_state(opt_gen_1, eng.g.fp)
            checkpoint = {
                "textEncoder": textEncoder.state_dict(), "projection_head": projection_head.state_dict(),
                "con_augment_1": con_augment_1.state_dict(), "critic_1": critic_1.state_dict(),
                "gen_1": gen_1.state_dict(),
                "opt_encoder": opt_encoder.state_dict(), "opt_projection_head": opt_projection_head.state_dict(),
                "opt_con_augment_1": opt_con_augment_1.state_dict(), "opt_critic_1": opt_critic_1.state_dict(),
                "opt_gen_1": opt_gen_1.state_dict(),
                "lr_scheduler_encoder": lr_scheduler_encoder.state_dict(),
                "lr_scheduler_projection_head": lr_scheduler_projection_head.state_dict(),
                "lr_scheduler_con_augment_1": lr_scheduler_con_augment_1.state_dict(),
                "lr_scheduler_critic_1": lr_scheduler_critic_1.state_dict(),
                "lr_scheduler_gen_1": lr_scheduler_gen_1.state_dict(),
                "epoch": epoch,
            }
            os.makedirs(os.path.join(save_dir, "epochs"), exist_ok=True)
            torch.save(checkpoint, f"{save_dir}/epochs/checkpoint_epoch_{epoch}.pth")
            torch.save(checkpoint, checkpoint_path)
    return eng
