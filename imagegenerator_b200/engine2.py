"""Stage-II of the StackGAN train step on the C-ABI kernels (reference stage_2_train_fn.py:120-168,
generator_2.py:59-67, discriminator_2.py:27-38).

Differences from Stage-I that matter for parity:
  * ``gen_1`` / ``con_augment_1`` are frozen and in eval mode (:52-63): gen_1's BatchNorm uses running
    statistics, con_augment_1 still samples;
  * ``fake_256`` is not detached and ``opt_gen_2.zero_grad()`` runs only after the step (:131,:154,
    :163-168), so the gradients G2 / CA2 are stepped with are  d lossG + sum_5 d loss_critic_i  --
    every critic iteration back-propagates through the critic INTO the generator, including the
    gradient-penalty's second-order term through the interpolated images;
  * there is no zero_grad before ``lossG.backward()`` (:163).
"""
from __future__ import annotations

import torch

from .engine import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, LAMBDA_GP, N_CRITIC, Z_DIM, CART, CriticRT, GenRT,
                     SideStream, _LayerRT, _alloc_ctx, _capture_stream, _conv_out, _prio, _side_run, default_ops)
from .layers import FlatParams


class _BN:
    """Buffers of one BatchNorm'ed tensor: pre-BN conv output y, activation a, their gradients."""

    def __init__(self, ops, shape, C, alloc_a=True):
        self.y, self.dy = ops.empty(shape), ops.empty(shape)
        self.a = ops.empty(shape) if alloc_a else None
        self.da = ops.empty(shape) if alloc_a else None
        self.mr = ops.empty((1, C, 2), ops.f32)
        self.stats = ops.zeros((1, C, 2), ops.f64)
        self.sums = ops.zeros((1, C, 2), ops.f64)


def _conv_bn_forward(ops, direction, x, L, buf, out, act, training, residual=None):
    """conv (fprop 'f' / transposed 'd') -> train- or eval-mode BN -> activation (+ residual).  In training
    mode the batch statistics come out of the conv kernel's epilogue (sg_conv_*_stats)."""
    bn = L.bn
    C = buf.y.shape[-1]
    if training:
        ops.zero(buf.stats)
        if direction == "f":
            ops.conv_fprop_stats(x, L.pf, buf.y, buf.stats, 1, L.k, L.s, L.p)
        else:
            ops.conv_dgrad_stats(x, L.pd, buf.y, buf.stats, 1, L.k, L.s, L.p)
        ops.bn_finalize_act(buf.stats, buf.y.numel() // C, buf.mr, bn.running_mean, bn.running_var, bn.num_batches_tracked, 1,
                            buf.y, bn.weight.data, bn.bias.data, out, act, residual=residual)
    else:
        if direction == "f":
            ops.conv_fprop(x, L.pf, None, buf.y, L.k, L.s, L.p)
        else:
            ops.conv_dgrad(x, L.pd, None, buf.y, L.k, L.s, L.p)
        ops.bn_eval_mr(bn.running_mean, bn.running_var, buf.mr)
        ops.bn_act(buf.y, buf.mr, bn.weight.data, bn.bias.data, out, 1, act, residual=residual)


def _bn_backward(ops, bn, buf, da, a_out, act, side=None, from_y=True, bn_items=None, reduced=False, zeroed=False):
    """dy = BN-backward of (da masked by act'(a_out)); accumulates gamma/beta grads.  ``from_y``: the activation follows
    the BN directly, so the kernels take its sign from y and do not stream a_out (False for the layer that closes a
    residual block, whose ReLU sees bn(y) + identity).  ``bn_items``: list collecting (sums, gamma.grad, beta.grad) instead of
    launching the per-layer parameter-gradient kernel."""
    # reduced: buf.sums came out of the epilogue of the conv that produced da (_conv_bstats); else reduce + apply (one launch
    # for the layers that fit the SMs' shared memory)
    # zeroed: the caller zeroed buf.sums (ops.zero_multi at the start of its pass)
    kw = {"beta": bn.bias.data} if from_y else {}
    if reduced:
        ops.bn_bwd_apply(da, a_out, buf.y, buf.mr, bn.weight.data, buf.sums, buf.dy, 1, act, **kw)
    else:
        ops.bn_bwd(da, a_out, buf.y, buf.mr, bn.weight.data, buf.sums, buf.dy, 1, act, zeroed=zeroed, **kw)
    if bn_items is not None:
        bn_items.append((buf.sums, bn.weight.grad, bn.bias.grad))       # the caller ends its pass with ONE launch for all layers
    else:
        _side_run(side, lambda: ops.bn_param_grad(buf.sums, bn.weight.grad, bn.bias.grad))
    return buf.dy


def _conv_bstats(ops, direction, dy, L, da_out, below_bn, below_buf, act):
    """Data-gradient conv of layer L whose result ``da_out`` is d loss / d a of the BatchNorm'ed layer below: the launch also
    reduces that layer's BatchNorm-backward statistics (below_buf.sums) in its epilogue.  Returns whether it did (False: only
    the conv ran; the BatchNorm backward of the layer below reduces + applies in one call)."""
    return ops.conv_bstats_opt(direction, dy, L.pf if direction == "f" else L.pd, da_out, below_buf.y, below_buf.mr,
                               below_bn.weight.data, below_bn.bias.data, below_buf.sums, 1, act, L.k, L.s, L.p)


class Gen2RT:
    """Stage-II generator runtime."""

    def __init__(self, ops, module, B, x_in=None, out=None):
        self.ops, self.m, self.B = ops, module, B
        self.fp = FlatParams.of(module, ops.device, dtype=ops.f32)
        m = module
        f32 = ops.f32
        self.ds0 = _LayerRT(ops, m.down_sampler[0], None)
        self.ds2 = _LayerRT(ops, m.down_sampler[2][0], m.down_sampler[2][1])
        self.res = [[_LayerRT(ops, c, bn) for c, bn in blk.conv_layers()] for blk in m.residual_blocks]
        self.ups = [_LayerRT(ops, m.up_sampler[i][0], m.up_sampler[i][1]) for i in range(3)]
        self.up3 = _LayerRT(ops, m.up_sampler[3], None)
        # thin (3-channel) operators: forward / data gradient are direct kernels (thin_conv.cu); their weight gradients
        # are 1x1 GEMMs over a patch matrix in the PyTorch weight order
        self.K0 = 3 * 16
        self.x_in = x_in if x_in is not None else ops.empty((B, 64, 64, 3))
        self.P0 = ops.empty((B, 32, 32, self.K0))
        self.a1, self.da1, self.dy0 = ops.empty((B, 32, 32, 128)), ops.empty((B, 32, 32, 128)), ops.empty((B, 32, 32, 128))
        self.b2 = _BN(ops, (B, 16, 16, 512), 512)
        self.c_hat = None
        self.dc_hat = ops.empty((B, m.C_TEXT), f32)
        self.X = [ops.empty((B, 16, 16, 640)) for _ in range(5)]       # residual-block boundaries
        self.dX = [ops.empty((B, 16, 16, 640)) for _ in range(5)]
        self.dz = ops.empty((B, 16, 16, 640))
        self.rb = [[_BN(ops, (B, 16, 16, 320), 320), _BN(ops, (B, 16, 16, 320), 320), _BN(ops, (B, 16, 16, 640), 640, alloc_a=False)]
                   for _ in range(4)]
        self.ub = [_BN(ops, (B, 32, 32, 320), 320), _BN(ops, (B, 64, 64, 160), 160), _BN(ops, (B, 128, 128, 80), 80)]
        self.out = out if out is not None else ops.empty((B, 256, 256, 3))
        self.dpre = ops.empty((B, 256, 256, 3))
        self.Pd = ops.empty((B, 128, 128, self.K0))
        self.ones = torch.ones(B, dtype=f32).to(ops.device)

    def _wgrad(self, L, x, dy):
        """Weight gradient of one residual-block conv: 3x3 taps are not a multiple of 4, so in the PyTorch layout
        the tcgen05 kernel would leave through 4-byte atomics; accumulate channels-last and fold once per step."""
        ops = self.ops
        if getattr(L, "gw", None) is None:
            L.gw = ops.zeros((L.co, L.k, L.k, L.ci), ops.f32) if ops.conv_wgrad_cl_supported(x, dy, L.k, L.s, L.p) else False
        if not torch.is_tensor(L.gw):
            ops.conv_wgrad(x, dy, L.conv.weight.grad, L.k, L.s, L.p)
        else:
            ops.conv_wgrad_cl(x, dy, L.gw, L.k, L.s, L.p)

    def fold_grads(self):
        """Add the channels-last accumulation buffers into the parameter gradients and clear them."""
        for blk in self.res:
            for L in blk:
                if torch.is_tensor(getattr(L, "gw", None)):
                    self.ops.fold_grad_cl(L.gw, L.conv.weight.grad)

    def all_layers(self):
        out = [self.ds0, self.ds2]
        for blk in self.res:
            out += blk
        return out + self.ups + [self.up3]

    def refresh_weights(self):
        ops = self.ops
        for L in self.all_layers():
            L.pack(ops)

    def forward(self, c_hat, training=True):
        """x_in (NHWC T) and c_hat [B,128] fp32 -> self.out [B,256,256,3] (generator_2.py:59-67)."""
        ops = self.ops
        self.c_hat = c_hat
        L = self.ds0
        ops.conv_fprop(self.x_in, L.pf, L.conv.bias.data, self.a1, L.k, L.s, L.p, act=ACT_LRELU)
        L = self.ds2
        _conv_bn_forward(ops, "f", self.a1, L, self.b2, self.b2.a, ACT_LRELU, training)
        ops.concat_rep(self.b2.a, c_hat, self.X[0])
        for r in range(4):
            l1, l2, l3 = self.res[r]
            b1, b2, b3 = self.rb[r]
            _conv_bn_forward(ops, "f", self.X[r], l1, b1, b1.a, ACT_RELU, training)
            _conv_bn_forward(ops, "f", b1.a, l2, b2, b2.a, ACT_RELU, training)
            _conv_bn_forward(ops, "f", b2.a, l3, b3, self.X[r + 1], ACT_RELU, training, residual=self.X[r])   # x += identity; relu
        x = self.X[4]
        for i in range(3):
            L, b = self.ups[i], self.ub[i]
            _conv_bn_forward(ops, "d", x, L, b, b.a, ACT_RELU, training)
            x = b.a
        L = self.up3
        # ConvT(80 -> 3) + bias + Tanh (generator_2.py:55-57): one direct kernel, the col matrix never leaves the SM
        ops.conv_dgrad(x, L.pd, L.conv.bias.data, self.out, L.k, L.s, L.p, act=ACT_TANH)
        return self.out

    def backward(self, dout, side=None):
        """Accumulates parameter gradients; leaves d/d c_hat in self.dc_hat (fp32).  ``side``: optional SideStream for
        the parameter-gradient kernels (the caller joins before the optimizer step)."""
        ops = self.ops
        sr = lambda fn: _side_run(side, fn)
        # every BatchNorm layer's backward sums zeroed by one kernel node (a memset per layer is a graph node in the chain each)
        ops.zero_multi([b.sums for b in self.ub] + [b.sums for rb in self.rb for b in rb] + [self.b2.sums])
        L = self.up3
        bn_items = []             # (sums, gamma.grad, beta.grad) of the 16 BatchNorm layers: one launch at the end of the pass
        ops.act_bwd(dout, self.out, self.dpre, ACT_TANH)

        def pgrad_up3(L=L):
            ops.patchify(self.dpre, self.Pd, L.k, L.s, L.p)
            ops.colsum(self.dpre, L.conv.bias.grad)
            ops.conv_wgrad(self.Pd, self.ub[2].a, L.conv.weight.grad.view(L.co, self.K0, 1, 1), 1, 1, 0)
        sr(pgrad_up3)
        ops.conv_fprop(self.dpre, L.pf, None, self.ub[2].da, L.k, L.s, L.p)
        reduced = False           # the next layer's BN-backward statistics already came out of a conv epilogue
        for i in range(2, -1, -1):
            L, b = self.ups[i], self.ub[i]
            dy = _bn_backward(ops, L.bn, b, b.da, b.a, ACT_RELU, side, bn_items=bn_items, zeroed=True, reduced=reduced)
            x_in = self.ub[i - 1].a if i > 0 else self.X[4]
            sr(lambda L=L, dy=dy, x_in=x_in: ops.conv_wgrad(dy, x_in, L.conv.weight.grad, L.k, L.s, L.p))
            if i > 0:
                reduced = _conv_bstats(ops, "f", dy, L, self.ub[i - 1].da, self.ups[i - 1].bn, self.ub[i - 1], ACT_RELU)
            else:
                ops.conv_fprop(dy, L.pf, None, self.dX[4], L.k, L.s, L.p)
        for r in range(3, -1, -1):
            l1, l2, l3 = self.res[r]
            b1, b2, b3 = self.rb[r]
            dy3 = _bn_backward(ops, l3.bn, b3, self.dX[r + 1], self.X[r + 1], ACT_RELU, side, from_y=False, bn_items=bn_items, zeroed=True)
            ops.act_bwd(self.dX[r + 1], self.X[r + 1], self.dz, ACT_RELU)            # identity branch
            sr(lambda l3=l3, b2=b2, dy3=dy3: self._wgrad(l3, b2.a, dy3))
            red = _conv_bstats(ops, "d", dy3, l3, b2.da, l2.bn, b2, ACT_RELU)
            dy2 = _bn_backward(ops, l2.bn, b2, b2.da, b2.a, ACT_RELU, side, bn_items=bn_items, zeroed=True, reduced=red)
            sr(lambda l2=l2, b1=b1, dy2=dy2: self._wgrad(l2, b1.a, dy2))
            red = _conv_bstats(ops, "d", dy2, l2, b1.da, l1.bn, b1, ACT_RELU)
            dy1 = _bn_backward(ops, l1.bn, b1, b1.da, b1.a, ACT_RELU, side, bn_items=bn_items, zeroed=True, reduced=red)
            sr(lambda l1=l1, r=r, dy1=dy1: self._wgrad(l1, self.X[r], dy1))
            ops.conv_dgrad(dy1, l1.pd, None, self.dX[r], 3, 1, 1)
            ops.scale_rows_add(self.dz, self.ones, self.dX[r], True)
        ops.split_rep_bwd(self.dX[0], self.b2.da, self.dc_hat)
        L = self.ds2
        dy2 = _bn_backward(ops, L.bn, self.b2, self.b2.da, self.b2.a, ACT_LRELU, side, bn_items=bn_items, zeroed=True)
        sr(lambda L=L, dy2=dy2: ops.conv_wgrad(self.a1, dy2, L.conv.weight.grad, L.k, L.s, L.p))
        ops.conv_dgrad(dy2, L.pd, None, self.da1, L.k, L.s, L.p)
        L = self.ds0
        ops.act_bwd(self.da1, self.a1, self.dy0, ACT_LRELU)

        def pgrad_ds0(L=L):
            ops.bn_param_grad_multi(bn_items)
            ops.patchify(self.x_in, self.P0, L.k, L.s, L.p)
            ops.conv_wgrad(self.P0, self.dy0, L.conv.weight.grad.view(L.co, self.K0, 1, 1), 1, 1, 0)
            ops.colsum(self.dy0, L.conv.bias.grad)
        sr(pgrad_ds0)
        return self.dc_hat


class Stage2Engine:
    """One reference Stage-II outer step with caller-supplied noise."""

    def __init__(self, ca1, gen1, ca2, critic2, gen2, batch_size, ops=None, lr=1e-3, comm=None):
        ops = ops or default_ops()
        self.ops, self.B = ops, batch_size
        B = batch_size
        with _alloc_ctx(comm):
            self.d = CriticRT(ops, critic2, B)
            self.ca1, self.ca2 = CART(ops, ca1), CART(ops, ca2)
            self.ca1.ensure(B)
            self.ca2.ensure(B)
            self.g1 = GenRT(ops, gen1, B)                               # frozen, eval mode
            self.g2 = Gen2RT(ops, gen2, B, x_in=self.g1.out, out=self.d.group_view(self.d.a[0], 1, 1))
        for fp in (self.d.fp, self.g2.fp, self.ca2.fp):
            fp.set_lr(lr)
        self.losses = ops.zeros((4,), ops.f32)
        self.side = SideStream(ops)
        self.pack_side = SideStream(ops, priority=_prio("SG_PACK_PRIO", 0))   # critic weight re-packing, overlapped with the next generator forward
        self._ce_ready = False                        # compressed text valid for the current weights + batch
        self.comm = comm
        self.one_minus_eps = ops.empty((B,), ops.f32)
        self.dcg2 = ops.zeros((B, 1, 1, ca2.c_dim), ops.f32)
        if comm is not None:
            for fp in (self.d.fp, self.g2.fp, self.ca2.fp, self.g1.fp, self.ca1.fp):
                comm.broadcast_params(fp.flat)
        self.g1.refresh_weights()
        self.refresh_all()
        ops.zero(self.g2.fp.grad)
        ops.zero(self.ca2.fp.grad)

    def refresh_all(self):
        self.d.refresh_weights()
        self.g2.refresh_weights()

    def gather_optimizer_state(self):
        """COLLECTIVE (every rank): make the sharded Adam moments whole before a checkpoint is written."""
        if self.comm is not None and self.comm.peer:
            for fp in (self.d.fp, self.g2.fp, self.ca2.fp):
                self.comm.gather_state(fp)

    def sync_grads(self):
        self.side.join()
        self._sync_grads()

    def _sync_grads(self):
        """Make every ``.grad`` current (the residual-block weight gradients accumulate in channels-last side
        buffers during the step and are folded in here; ``optimizer_step`` does it for you)."""
        self.g2.fold_grads()

    def optimizer_step(self, fp):
        """xm.optimizer_step (stage_2_train_fn.py:155,164,167): gradient mean over replicas, then Adam.  Under
        multi-GPU graph capture the NCCL call is issued eagerly between two graph segments (engine._SegmentedGraph)."""
        self.side.join()
        if fp is self.g2.fp:
            self.g2.fold_grads()
        if self.comm is not None and self.comm.peer:
            self.comm.step(fp)               # one kernel over peer memory, inside the captured graph
            return
        if self.comm is not None:
            seg = getattr(self, "_seg", None)
            if seg is not None and seg.capturing:
                self.pack_side.join()                               # a graph segment must end with every fork joined
                seg.cut(lambda: (self.comm.allreduce_async(fp.grad), self.comm.wait_all()))
            else:
                self.comm.allreduce_async(fp.grad)
                self.comm.wait_all()
        self.ops.adam_step(fp.flat, fp.grad, fp.m, fp.v, fp.hyper)

    def load_batch(self, real_nchw, tem, tem_mis):
        d = self.d
        self.ops.nchw_to_nhwc(real_nchw, d.group_view(d.a[0], 0, 1))
        d.set_text(tem, tem_mis)
        self._ce_ready = False

    def _generate(self, z, eps_ca1, eps_ca2):
        d, B = self.d, self.B
        tem = d.tem_all[:B]
        self.ca1.forward(tem, eps_ca1, z, cg=self.g1.cg)           # stage_2_train_fn.py:124-127 (frozen)
        self.g1.forward(training=False)                             # :128, eval-mode BN
        st2 = self.ca2.forward(tem, eps_ca2, None)                  # :130
        self.g2.forward(st2.c_hat, training=True)                   # :131 -> critic group 1

    def preview(self, tem, z, eps_ca1, eps_ca2):
        """The in-loop sample of stage_2_train_fn.py:181-195: con_augment_1 -> gen_1 (eval) -> con_augment_2 -> gen_2, with
        gen_2 still in TRAIN mode as in the reference (batch statistics, running statistics updated, :90 is never undone).
        Returns (fake_64, fake_256) as fp32 NCHW tensors.  Overwrites the generator's activations and the critic's fake
        image buffer, which the next train step recomputes anyway."""
        ops, B = self.ops, self.B
        tem_d = self.d.tem_all[:B]
        tem_d.copy_(tem, non_blocking=True)
        f32 = ops.f32
        dv = lambda t: t.to(device=ops.device, dtype=f32, non_blocking=True).contiguous()
        self.ca1.forward(tem_d, dv(eps_ca1), dv(z), cg=self.g1.cg)
        self.g1.forward(training=False)
        st2 = self.ca2.forward(tem_d, dv(eps_ca2), None)
        self.g2.forward(st2.c_hat, training=True)
        out64 = ops.empty((B, 3, self.g1.out.shape[1], self.g1.out.shape[2]), f32)
        out256 = ops.empty((B, 3, 256, 256), f32)
        ops.nhwc_to_nchw(self.g1.out, out64)
        ops.nhwc_to_nchw(self.g2.out, out256)
        return out64, out256

    def _generator_backward(self, dfake, kl_scale):
        """Back-propagate d loss / d fake_256 into G2 and CA2 (accumulating)."""
        ops = self.ops
        dc = self.g2.backward(dfake, side=self.side)                # fp32 [B,128]
        self.ca2.backward_from_dc(dc, kl_scale)

    def critic_iteration(self, z, eps_ca1, eps_ca2, eps_gp):
        ops, d, B = self.ops, self.d, self.B
        self._generate(z, eps_ca1, eps_ca2)
        X = d.a[0]
        ops.interp(d.group_view(X, 0, 1), d.group_view(X, 1, 1), eps_gp, d.group_view(X, 2, 1))   # utils.py:10-11
        d.forward(0, 3, dup_first=2, training=True, with_mismatched=True,        # :133-140 + utils.py:13
                  before_weights=self.pack_side.join, ce_ready=self._ce_ready, patches_on=self.side)
        ops.zero(d.fp.grad)                                         # :153
        ops.zero(d.head_grads)                                      # dA, dBv
        d.gp_first_order()
        # :148-152; only the host reads the loss values: off the main stream
        self.side.run(lambda: ops.critic_loss(d.score[0], d.score[1], d.score[2], d.sq, LAMBDA_GP, self.losses[0:2]))
        d.gp_second_order(2.0 * LAMBDA_GP / B, side=self.side, defer_wgrad=True)
        d.backward(0, 3, d.coef_critic, inject=True, param_grads=True, need_input_grad=True,      # :154
                   input_grad_from=1, side=self.side, merge_gp=True)   # d/d real images is never used
        self.side.run(lambda: d.text_backward(d.coef_text, 2 * B, 0.0, True, None))
        # d loss_critic / d fake_256 = d/d(fake group) + (1 - eps) * d/d(interpolated group)  (utils.py:11, not detached)
        ops.affine_f32(eps_gp, -1.0, 1.0, self.one_minus_eps)
        dfake = d.group_view(d.dx, 1, 1)
        ops.scale_rows_add(d.group_view(d.dx, 2, 1), self.one_minus_eps, dfake, True)
        self._generator_backward(dfake, 0.0)                        # accumulates into G2 / CA2 (:154, no zero_grad)
        self.optimizer_step(d.fp)                                   # :155
        self.pack_side.run(lambda: d.refresh_weights(with_text=True, events=True))   # the next critic forward waits layer by layer
        self._ce_ready = True                                       # until the text changes (load_batch / next outer step)

    def generator_step(self):
        ops, d, B = self.ops, self.d, self.B
        d.forward(1, 1, dup_first=1, training=True, before_weights=self.pack_side.join,     # :157
                  ce_ready=self._ce_ready)
        st = self.ca2.st
        ops.gen_loss(d.score[2], st.mu, st.sigma, self.losses[2:4])  # :158-162
        d.backward(1, 1, d.coef_gen, inject=False, param_grads=False, need_input_grad=True)
        self._generator_backward(d.group_view(d.dx, 1, 1), 1.0)     # :163 (no zero_grad before)
        self.optimizer_step(self.g2.fp)                             # :164
        self.optimizer_step(self.ca2.fp)                            # :167
        ops.zero(self.g2.fp.grad)                                   # :165
        ops.zero(self.ca2.fp.grad)                                  # :168
        self.g2.refresh_weights()

    def outer_step(self, z, eps_ca1, eps_ca2, eps_gp):
        self._ce_ready = False                                      # a new batch: its text has not been compressed yet
        for it in range(N_CRITIC):
            self.critic_iteration(z[it], eps_ca1[it], eps_ca2[it], eps_gp[it])
        self.generator_step()

    # -- whole step behind static buffers, replayed as one CUDA graph (single GPU)
    def step(self, real_nchw, tem, tem_mis, z, eps_ca1, eps_ca2, eps_gp, use_graph=True):
        ops, B = self.ops, self.B
        if getattr(self, "s_real", None) is None:
            f = ops.f32
            self.s_real = ops.empty((B, 3, 256, 256), f)
            self.s_z = ops.empty((N_CRITIC, B, Z_DIM), f)
            self.s_e1 = ops.empty((N_CRITIC, B, self.ca1.m.c_dim), f)
            self.s_e2 = ops.empty((N_CRITIC, B, self.ca2.m.c_dim), f)
            self.s_egp = ops.empty((N_CRITIC, B), f)
            self.graph, self.launches_per_step = None, None
        if (not real_nchw.is_cuda) and real_nchw.is_pinned() and not getattr(ops, "is_emulator", False):
            # host batch (50 MB of fp32 images): upload over a copy stream into one of two staging buffers, so that batch
            # k+1 crosses PCIe while step k computes; the compute stream only does a device-to-device copy
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=ops.device)
                self._stage = [torch.empty_like(self.s_real) for _ in range(2)]
                self._stage_free, self._stage_i = [None, None], 0
            i = self._stage_i
            self._stage_i ^= 1
            cur = torch.cuda.current_stream(ops.device)
            with torch.cuda.stream(self._copy_stream):
                if self._stage_free[i] is not None:
                    self._copy_stream.wait_event(self._stage_free[i])
                self._stage[i].copy_(real_nchw, non_blocking=True)
                up = torch.cuda.Event()
                up.record(self._copy_stream)
            cur.wait_event(up)
            self.s_real.copy_(self._stage[i], non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            self._stage_free[i] = done
        else:
            self.s_real.copy_(real_nchw, non_blocking=True)
        for dst, src in ((self.d.tem_all[:B], tem), (self.d.tem_all[B:], tem_mis), (self.s_z, z),
                         (self.s_e1, eps_ca1), (self.s_e2, eps_ca2), (self.s_egp, eps_gp)):
            dst.copy_(src, non_blocking=True)

        def body():
            if hasattr(ops, "set_option"):
                ops.set_option("pdl", 1)          # programmatic dependent launch: +2 % on this step (see engine.Stage1Engine._body)
            ops.nchw_to_nhwc(self.s_real, self.d.group_view(self.d.a[0], 0, 1))
            self.outer_step(self.s_z, self.s_e1, self.s_e2, self.s_egp)
        if not use_graph or getattr(ops, "is_emulator", False):
            n0 = ops.launch_count() if hasattr(ops, "launch_count") else 0
            body()
            if hasattr(ops, "launch_count"):
                self.launches_per_step = ops.launch_count() - n0
            return
        if self.comm is not None and self.comm.world > 1 and not self.comm.peer:
            # all-reduce transport: graph segments on a private stream, the collective eager in between (engine.Stage1Engine.step)
            from .engine import _SegmentedGraph
            if getattr(self, "gstream", None) is None:
                self.gstream = torch.cuda.Stream(device=ops.device)
            cur = torch.cuda.current_stream(ops.device)
            self.gstream.wait_stream(cur)
            with torch.cuda.stream(self.gstream):
                if self.graph is None:
                    torch.cuda.synchronize()
                    n0 = ops.launch_count()
                    self._seg = _SegmentedGraph(ops)
                    self._seg.begin()
                    try:
                        body()
                    finally:
                        self._seg.end()
                    self.launches_per_step = ops.launch_count() - n0
                    self.graph = self._seg
                self.graph.replay()
            cur.wait_stream(self.gstream)
            return
        if self.graph is None:
            torch.cuda.synchronize()
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=_capture_stream(self.ops.device, 0)):
                body()
            self.launches_per_step = ops.launch_count() - n0
            self.graph = g
        self.graph.replay()
