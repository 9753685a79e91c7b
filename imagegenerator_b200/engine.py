"""Host-side orchestration of the StackGAN train step on top of the C-ABI kernels.

This file contains no arithmetic: every tensor operation is a call into ``ops`` (the ctypes
binding of libsgb200.so, ``imagegenerator_b200.ops.CudaOps``).  Autograd is not used -- the
backward passes, including the second-order path of the WGAN-GP gradient penalty
(reference utils.py:8-26 differentiated by stage_1_train_fn.py:147), are written out by hand
(DESIGN.md "Hand-derived backward").

Layout: activations are NHWC in the storage type T of the ops object (bf16 or fp32);
parameters/gradients/statistics are fp32 (sums in fp64).  The critic processes up to three
*groups* of B images in one batched pass -- (real, fake, interpolated) -- each group with its own
BatchNorm statistics, exactly like the reference's separate critic calls
(stage_1_train_fn.py:125-138); the "mismatched text" call (:130) shares the real group's
feature maps and differs only in the affine head.
"""
from __future__ import annotations

import contextlib
import os
from types import SimpleNamespace

import torch

from .layers import FlatParams, flat_allocator


def _alloc_ctx(comm):
    """Parameter / gradient buffers in symmetric memory when the transport addresses peers directly (comm.PeerComm)."""
    return flat_allocator(comm.alloc) if (comm is not None and getattr(comm, "peer", False)) else contextlib.nullcontext()

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
N_CRITIC = 5       # stage_1_train_fn.py:14
LAMBDA_GP = 10.0   # stage_1_train_fn.py:15
Z_DIM = 100        # stage_1_train_fn.py:16

_DEFAULT_OPS = None


def default_ops():
    """The CUDA ops object.  Fails loudly when the extension or a GPU is missing: there is no
    CPU fallback in this package."""
    global _DEFAULT_OPS
    if _DEFAULT_OPS is None:
        from .ops import CudaOps
        _DEFAULT_OPS = CudaOps()
    return _DEFAULT_OPS


def _conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


class SideStream:
    """Fork/join helper: ``run(fn)`` queues ``fn``'s kernels on a second CUDA stream behind everything already queued
    on the current one, ``join()`` makes the current stream wait for them.  Used for the parameter-gradient kernels
    (wgrad, BN gamma/beta, bias sums), which nothing reads before the optimizer step: the data-gradient chain keeps
    the main stream, and the weight-gradient kernels fill the SMs its tile waves leave idle.  Works under CUDA-graph
    capture (the fork/join become graph edges).  No-op for the CPU emulator or with SG_NO_SIDE_STREAM=1."""

    def __init__(self, ops, priority=0):
        self.ops = ops
        self.enabled = (not getattr(ops, "is_emulator", False)) and os.environ.get("SG_NO_SIDE_STREAM") != "1"
        self.stream, self.pending, self.priority = None, False, priority
        # timing experiment only (results are WRONG): drop this stream's work to see what the main chain costs alone
        self.skip = os.environ.get("SG_EXPERIMENT_SKIP_SIDE") == "1"

    def run(self, fn):
        if self.skip:
            return
        if not self.enabled:
            fn()
            return
        dev = self.ops.device
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev, priority=self.priority)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            fn()
        self.pending = True

    def join(self):
        if self.enabled and self.pending:
            torch.cuda.current_stream(self.ops.device).wait_stream(self.stream)
            self.pending = False


def _prio(name, default):
    v = os.environ.get(name)
    return int(v) if v not in (None, "") else default


def _capture_stream(device, default_priority=0):
    """Stream the step's graph is captured on.  Kernel nodes inherit the priority of the stream they were captured from
    (CUDA: lower number = higher priority, 0 is the lowest): capturing the main chain on a HIGH-priority stream while the
    parameter-gradient side stream keeps priority 0 makes the block scheduler serve the critical path first whenever
    main-chain CTAs and wgrad CTAs are both waiting for an SM.  Measured on B200 (bench.py): Stage-I 5.57 -> 5.35 ms;
    Stage-II 34.9 -> 35.4 ms (its side stream carries 13 ms of wgrad per step: starved, it arrives late at the joins), so
    Stage-II keeps equal priorities.  SG_MAIN_PRIO overrides."""
    pr = _prio("SG_MAIN_PRIO", default_priority)
    return torch.cuda.Stream(device=device, priority=pr) if pr != 0 else None


# A/B switch (default off, measured): the data-gradient conv of the critic's second layer stores dy0 = da1 * act'(a1) itself and
# leaves the first conv's bias gradient in its epilogue sums (ops.conv_dgrad_masked) instead of conv + activation-backward pass.
# Correct (tests/test_kernels_gpu.py::test_conv_dgrad_masked, step parity green) and SLOWER: Stage-I 4.86 -> 5.34 ms -- the
# statistics epilogue of conv_tcp_kernel costs more on this 512-deep reduction than the 21 us pass it removes.
MASKED_DGRAD = os.environ.get("SG_MASKED_DGRAD") == "1"
# The one-launch BatchNorm backward (option bn_fused) for the gradient penalty's first-order pass only (one image group: 2-8 MB
# tensors, and the side streams are nearly idle there): Stage-I 4.89 -> 4.85 ms.  SG_BN_FUSED_GP1=0 switches it off.
BN_FUSED_GP1 = os.environ.get("SG_BN_FUSED_GP1", "1") == "1"
# Where the next critic iteration's fake batch is produced on the generator's side stream: 0 = right after the critic forward
# (next to the penalty's first-order pass), 1 = after that pass (next to the second-order pass), 2 = next to the plain backward.
# Measured (Stage-I ms per step, two runs each): 4.766 / 4.783 / 4.741 -- the two penalty passes run one-launch BatchNorm kernels
# that want every SM, the plain backward's kernels share SMs with other streams anyway.
GEN_AHEAD_AT = int(os.environ.get("SG_GEN_AHEAD_AT", "2"))
# A/B switch (default off, measured neutral: Stage-I 4.757 -> 4.753 ms): the critic forward's BatchNorm finalize + apply with
# bulk-copy staged ranges (option bn_act_bulk, one CTA per SM) -- y was just written by the conv and is L2-resident, so the deeper
# prefetch buys nothing there.
BN_ACT_BULK = os.environ.get("SG_BN_ACT_BULK", "0") == "1"


def _side_run(side, fn):
    if side is None:
        fn()
    else:
        side.run(fn)


class _LayerRT:
    """One conv operator (Conv2d orientation: ``co`` x ``ci`` x k x k weight) + optional BN."""

    def __init__(self, ops, conv, bn, packs=True):
        self.conv, self.bn = conv, bn
        w = conv.weight
        self.co, self.ci, self.k = w.shape[0], w.shape[1], w.shape[2]
        self.s, self.p = conv.stride, conv.pad
        self.pf = ops.empty((self.co, self.k, self.k, self.ci)) if packs else None
        self.pd = ops.empty((self.ci, self.k, self.k, self.co)) if packs else None

    def pack(self, ops):
        ops.pack_weight(self.conv.weight.data, self.pf, self.pd)


class Up0Gemm:
    """The Stage-I generator's first layer, ConvTranspose2d(228 -> 192, k4, s1, p0) on a 1x1 input (generator_1.py:9-13,
    :38-39), as what it is: a GEMM  [B, 228] x [228, 16*192]  whose output row IS the NHWC [4][4][192] feature map.

    The conv kernels read 64-channel blocks through TMA, and 228 is not a multiple of 8 -- as a convolution this layer
    (and only this layer) used to fall back to the CUDA-core kernel.  Here the reduction is zero-padded to Kp = 256
    channels ([c_hat, z, 0...] rows, sg_ca_reparam) and all three directions run as 1x1 convolutions on the tensor-core
    kernels:
        forward      y[B, (h,w,ci)]   = cg[B, Kp]        . wf[(h,w,ci)][Kp]^T
        input grad   dcg[B, Kp]       = dy[B, (h,w,ci)]  . wb[Kp][(h,w,ci)]^T
        weight grad  gw[Kp][(h,w,ci)] += cg^T dy   -> folded into the PyTorch-layout .grad [228][192][4][4]
    """

    def __init__(self, ops, conv, B):
        w = conv.weight
        self.ops, self.conv, self.B = ops, conv, B
        self.co, self.ci, self.k = w.shape[0], w.shape[1], w.shape[2]
        self.Kp = (self.co + 63) // 64 * 64
        self.NO = self.ci * self.k * self.k
        self.wf = ops.empty((self.NO, 1, 1, self.Kp))
        self.wb = ops.zeros((self.Kp, 1, 1, self.NO))                          # rows >= co stay zero
        self.gw = ops.zeros((self.Kp, self.NO, 1, 1), ops.f32)

    def pack(self):
        ops, W = self.ops, self.conv.weight.data
        ops.pack_gemm_t(W, self.wf)                                                            # [(h,w,ci)][Kp]
        ops.pack_weight(W, self.wb.view(self.Kp, self.k, self.k, self.ci)[:self.co], None)      # [co][(h,w,ci)]

    def forward(self, cg, y):
        self.ops.conv_fprop(cg, self.wf, None, y.view(y.shape[0], 1, 1, self.NO), 1, 1, 0)

    def input_grad(self, dy, dcg):
        self.ops.conv_fprop(dy.view(dy.shape[0], 1, 1, self.NO), self.wb, None, dcg, 1, 1, 0)

    def weight_grad(self, cg, dy):
        """``conv.weight.grad`` += the gradient (accumulated in gw, then folded and cleared)."""
        ops = self.ops
        ops.conv_wgrad(dy.view(dy.shape[0], 1, 1, self.NO), cg, self.gw, 1, 1, 0)
        ops.fold_grad_cl(self.gw.view(self.Kp, self.k, self.k, self.ci)[:self.co], self.conv.weight.grad)


# ============================================================================================ CA
class CART:
    """Conditioning augmentation runtime (reference con_augment.py:13-22)."""

    def __init__(self, ops, module):
        self.ops, self.m = ops, module
        self.fp = None
        self.st = None

    def ensure(self, B):
        ops, m = self.ops, self.m
        if self.fp is None:
            self.fp = FlatParams.of(m, ops.device, dtype=ops.f32)
        if self.st is None or self.st.B != B:
            f = ops.f32
            self.st = SimpleNamespace(
                B=B, h=ops.empty((B, m.h_dim), f), mu=ops.empty((B, m.c_dim), f), sigma=ops.empty((B, m.c_dim), f),
                c_hat=ops.empty((B, m.c_dim), f), dmu=ops.empty((B, m.c_dim), f), dsigma=ops.empty((B, m.c_dim), f),
                dh=ops.empty((B, m.h_dim), f), tem=None, eps=None)
        return self.st

    def forward(self, tem, eps, z, cg=None):
        """tem [B,512] fp32; eps [B,128] or None (encode only); z [B,nz] or None; cg: optional
        [B,1,1,128+nz] T buffer receiving [c_hat, z] (stage_1_train_fn.py:120-122)."""
        ops, m = self.ops, self.m
        st = self.ensure(tem.shape[0])
        st.tem, st.eps = tem, eps
        ops.ca_forward(tem, m.h.weight.data, m.h.bias.data, m.mu.weight.data, m.mu.bias.data, m.sigma.weight.data,
                       m.sigma.bias.data, eps, z, st.h, st.mu, st.sigma, st.c_hat, cg)      # one launch
        return st

    def backward(self, dcg, kl_scale, dtem, dtem_acc):
        """Accumulates parameter grads; dtem (+)= d/d tem.  dcg: grad of the [c_hat, z] buffer
        (T) or None; kl_scale multiplies d/d(mu,sigma) of sum(1+log s^2-mu^2-s^2)
        (stage_1_train_fn.py:156-159).  Two launches: per-sample data gradients, then all six parameter gradients."""
        ops, m, st = self.ops, self.m, self.st
        ops.ca_backward(dcg, st.eps, st.mu, st.sigma, kl_scale, st.h, st.tem, m.mu.weight.data, m.sigma.weight.data,
                        m.h.weight.data, st.dmu, st.dsigma, st.dh, m.mu.weight.grad, m.mu.bias.grad, m.sigma.weight.grad,
                        m.sigma.bias.grad, m.h.weight.grad, m.h.bias.grad, dtem, dtem_acc)

    def backward_from_dc(self, dc, kl_scale):
        """Same with d loss / d c_hat given directly as an fp32 [B,c_dim] tensor and no d/d tem wanted
        (Stage-II: the text side is frozen, stage_2_train_fn.py:52-57)."""
        self.backward(dc, kl_scale, None, False)


# ============================================================================================ generator (Stage-I)
class GenRT:
    """Stage-I generator runtime (reference generator_1.py:38-40).  Every layer is a
    ConvTranspose2d = the data-gradient direction of a Conv2d operator."""

    def __init__(self, ops, module, B, out=None):
        self.ops, self.m, self.B = ops, module, B
        self.fp = FlatParams.of(module, ops.device, dtype=ops.f32)
        cl = module.conv_layers()
        self.layers = [_LayerRT(ops, c, bn, packs=(i > 0)) for i, (c, bn) in enumerate(cl)]
        self.up0 = Up0Gemm(ops, cl[0][0], B)             # first layer: a GEMM over the zero-padded [c_hat, z] rows
        self.cg = ops.zeros((B, 1, 1, self.up0.Kp))
        self.dcg = ops.empty((B, 1, 1, self.up0.Kp))
        h = 1
        self.y, self.a, self.dy, self.da, self.mr, self.stats, self.sums = [], [], [], [], [], [], []
        # BN statistics accumulators of all layers in one buffer: one memset per forward instead of one per layer
        self.stats_flat = ops.zeros((2 * sum(L.ci for L in self.layers if L.bn is not None),), ops.f64)
        soff = 0
        for L in self.layers:
            h = (h - 1) * L.s - 2 * L.p + L.k
            shp = (B, h, h, L.ci)
            if L.bn is not None:
                self.y.append(ops.empty(shp)); self.a.append(ops.empty(shp))
                self.dy.append(ops.empty(shp)); self.da.append(ops.empty(shp))
                self.mr.append(ops.empty((1, L.ci, 2), ops.f32))
                self.stats.append(self.stats_flat[soff:soff + 2 * L.ci].view(1, L.ci, 2))
                soff += 2 * L.ci
                self.sums.append(ops.zeros((1, L.ci, 2), ops.f64))
            else:
                self.out = out if out is not None else ops.empty(shp)
                self.dpre = ops.empty(shp)
                hin = (h + 2 * L.p - L.k) // L.s + 1
                self.K_last = L.ci * L.k * L.k
                # weight gradient of the 3-channel layer: a 1x1 GEMM over the patch matrix of d/d(pre-tanh)
                self.Pd = ops.empty((B, hin, hin, self.K_last))
        self.packed = False

    def refresh_weights(self):
        self.up0.pack()
        for L in self.layers[1:]:
            L.pack(self.ops)
        self.packed = True

    def set_input(self, x):
        """x [B, c_dim+z_dim] fp32 -> cg buffer (module-level API only; the engine writes cg directly)."""
        xp = torch.zeros(self.B, self.up0.Kp, 1, 1, dtype=self.ops.f32, device=self.cg.device)
        xp[:, :x.shape[1], 0, 0] = x.to(device=xp.device, dtype=xp.dtype)
        self.ops.nchw_to_nhwc(xp, self.cg)

    def forward(self, training=True):
        ops = self.ops
        x = self.cg
        if training:
            ops.zero(self.stats_flat)
        for i, L in enumerate(self.layers):
            if L.bn is None:
                # ConvT(C -> 3) + bias + Tanh: one direct kernel (thin_conv.cu), the col matrix stays on the SM
                ops.conv_dgrad(x, L.pd, L.conv.bias.data, self.out, L.k, L.s, L.p, act=ACT_TANH)
                break
            bn = L.bn
            if i == 0:
                self.up0.forward(x, self.y[0])
                if training:
                    ops.col_stats(self.y[0], self.stats[0], 1)
            elif training:
                # conv + BN batch statistics in one kernel (reduced in the tcgen05 epilogue when the shape allows)
                ops.conv_dgrad_stats(x, L.pd, self.y[i], self.stats[i], 1, L.k, L.s, L.p)
            else:
                ops.conv_dgrad(x, L.pd, None, self.y[i], L.k, L.s, L.p)
            if training:
                n = self.y[i].numel() // L.ci
                ops.bn_finalize_act(self.stats[i], n, self.mr[i], bn.running_mean, bn.running_var,
                                    bn.num_batches_tracked, 1, self.y[i], bn.weight.data, bn.bias.data, self.a[i], ACT_RELU)
            else:
                ops.bn_eval_mr(bn.running_mean, bn.running_var, self.mr[i])
                ops.bn_act(self.y[i], self.mr[i], bn.weight.data, bn.bias.data, self.a[i], 1, ACT_RELU)
            x = self.a[i]
        return self.out

    def backward(self, dout, side=None):
        """dout: d loss / d out (T, NHWC).  Accumulates parameter grads, leaves d/d cg in self.dcg.  ``side``: optional
        SideStream for the parameter-gradient kernels (join before the optimizer step)."""
        ops = self.ops
        last = self.layers[-1]
        ops.zero_multi(self.sums)          # every layer's BatchNorm-backward sums in one kernel node (no memset per layer)
        ops.act_bwd(dout, self.out, self.dpre, ACT_TANH)

        def pgrad_last():
            ops.patchify(self.dpre, self.Pd, last.k, last.s, last.p)
            ops.colsum(self.dpre, last.conv.bias.grad)
            ops.conv_wgrad(self.Pd, self.a[-1], last.conv.weight.grad.view(last.co, self.K_last, 1, 1), 1, 1, 0)
        _side_run(side, pgrad_last)
        ops.conv_fprop(self.dpre, last.pf, None, self.da[-1], last.k, last.s, last.p)
        bn_items = []             # (sums, gamma.grad, beta.grad) of every BatchNorm layer: ONE launch at the end
        reduced = False           # sums[i] already came out of the epilogue of the conv that produced da[i]
        for i in range(len(self.layers) - 2, -1, -1):
            L, bn = self.layers[i], self.layers[i].bn
            if reduced:
                ops.bn_bwd_apply(self.da[i], self.a[i], self.y[i], self.mr[i], bn.weight.data, self.sums[i], self.dy[i], 1,
                                 ACT_RELU, beta=bn.bias.data)
            else:
                ops.bn_bwd(self.da[i], self.a[i], self.y[i], self.mr[i], bn.weight.data, self.sums[i], self.dy[i], 1,
                           ACT_RELU, beta=bn.bias.data, zeroed=True)
            x_in = self.a[i - 1] if i > 0 else self.cg

            bn_items.append((self.sums[i], bn.weight.grad, bn.bias.grad))

            def pgrad(i=i, L=L, bn=bn, x_in=x_in):
                if i == 0:
                    self.up0.weight_grad(x_in, self.dy[0])
                else:
                    ops.conv_wgrad(self.dy[i], x_in, L.conv.weight.grad, L.k, L.s, L.p)
            _side_run(side, pgrad)
            if i == 0:
                self.up0.input_grad(self.dy[0], self.dcg)
            else:
                # d/d a of the layer below + that layer's BatchNorm-backward statistics in the same kernel
                bnb = self.layers[i - 1].bn
                reduced = ops.conv_bstats_opt("f", self.dy[i], L.pf, self.da[i - 1], self.y[i - 1], self.mr[i - 1], bnb.weight.data,
                                              bnb.bias.data, self.sums[i - 1], 1, ACT_RELU, L.k, L.s, L.p)
        _side_run(side, lambda: ops.bn_param_grad_multi(bn_items))
        return self.dcg


# ============================================================================================ critic
class CriticRT:
    """Critic runtime for both stages (reference discrminator_1.py:41-52 / discriminator_2.py:27-38)."""

    NG = 3   # groups: 0 real, 1 fake, 2 interpolated

    def __init__(self, ops, module, B):
        self.ops, self.m, self.B = ops, module, B
        self.fp = FlatParams.of(module, ops.device, dtype=ops.f32)
        self.layers = [_LayerRT(ops, c, bn) for c, bn in module.conv_layers()]
        G, f32, f64 = self.NG, ops.f32, ops.f64
        h = module.in_hw
        self.a = [ops.empty((G * B, h, h, 3))]
        self.y, self.dy, self.da = [None], [], [None]
        self.mr, self.stats, self.sums = [None], [None], [None]
        # gradient-penalty chain buffers (one group)
        self.gda, self.gdy, self.gsums, self.tsums, self.v, self.w, self.gy = [None], [], [None], [None], [], [None], [None]
        # BN statistics accumulators of all layers in one buffer (consumed by bn_finalize right after each conv): one
        # memset per forward instead of one per layer
        self.stats_flat = ops.zeros((2 * G * sum(L.co for L in self.layers[1:]),), f64)
        soff = 0
        # The weight gradient of layer l has two terms with the same weights: wgrad(a[l], dy[l]) over the three image groups of
        # the plain backward and wgrad(w[l], gdy[l]) of the gradient penalty's second-order pass (one group, a third of the
        # rows: a launch at half the efficiency).  w[l] and gdy[l] therefore live as a FOURTH group right behind a[l] and
        # dy[l] (a_full / dy_full), so that one launch over 4B images does both (backward(merge_gp=True)).
        self.a_full, self.dy_full = [None], []
        for l, L in enumerate(self.layers):
            h = _conv_out(h, L.k, L.s, L.p)
            shp, shp1 = (G * B, h, h, L.co), (B, h, h, L.co)
            af, dyf = ops.empty(((G + 1) * B, h, h, L.co)), ops.empty(((G + 1) * B, h, h, L.co))
            self.a_full.append(af)
            self.dy_full.append(dyf)
            self.a.append(af[:G * B])
            self.da.append(ops.empty(shp))
            self.dy.append(dyf[:G * B])
            self.gda.append(ops.empty(shp1))
            self.gdy.append(dyf[G * B:])
            self.v.append(ops.empty(shp1))
            self.w.append(af[G * B:])
            if l > 0:
                self.y.append(ops.empty(shp))
                self.mr.append(ops.empty((G, L.co, 2), f32))
                self.stats.append(self.stats_flat[soff:soff + 2 * G * L.co].view(G, L.co, 2))
                soff += 2 * G * L.co
                self.sums.append(ops.zeros((G, L.co, 2), f64))
                self.gsums.append(ops.zeros((1, L.co, 2), f64))
                self.tsums.append(ops.zeros((L.co, 3), f64))
                self.gy.append(ops.empty(shp1))
        assert h == 4, h
        self.nl = len(self.layers)
        # identity "BatchNorm" table for the first layer (conv + bias + LeakyReLU, no BN): with it the masked data-gradient conv
        # (ops.conv_dgrad_masked) writes dy0 = da1 * act'(a1) directly and leaves the bias gradient in sums0[.][.][0]
        C0 = self.layers[0].co
        self.sums0, self.gsums0 = ops.zeros((G, C0, 2), f64), ops.zeros((1, C0, 2), f64)
        self.id_mr = ops.zeros((G, C0, 2), f32)
        self.id_mr[:, :, 1] = 1.0
        self.id_gamma, self.id_beta = ops.zeros((C0,), f32), ops.zeros((C0,), f32)
        self.id_gamma.fill_(1.0)
        self.bias_scratch = ops.zeros((C0,), f32)                    # receives the meaningless "gamma gradient" of that table
        # first layer (3-channel image): forward and input gradient are direct kernels (thin_conv.cu); its WEIGHT gradient
        # is a 1x1 GEMM over the patch matrix P[pix][ci*16+tap] -- the PyTorch weight order, so the gradient lands in place
        L0 = self.layers[0]
        h1 = _conv_out(module.in_hw, L0.k, L0.s, L0.p)
        self.K0 = L0.ci * L0.k * L0.k
        self.P_full = ops.empty(((G + 1) * B, h1, h1, self.K0))        # patches of the images, then (4th group) of v0
        self.P, self.Pv = self.P_full[:G * B], self.P_full[G * B:]
        cl, Nd = self.layers[-1].co, module.Nd
        self.head_grads = ops.zeros((16 * cl + Nd,), f32)            # dA and dBv: zeroed together every iteration
        self.A, self.dA = ops.empty((16, cl), f32), self.head_grads[:16 * cl].view(16, cl)
        self.Bv, self.dBv = ops.empty((Nd,), f32), self.head_grads[16 * cl:]
        self.c0, self.dc0 = ops.empty((1,), f32), ops.zeros((1,), f32)
        self.tem_all = ops.empty((2 * B, module.tem_size), f32)     # rows [0,B) tem, [B,2B) mismatched
        self.ce = ops.empty((2 * B, Nd), f32)
        self.dce = ops.empty((2 * B, Nd), f32)
        self.score = ops.empty((4, B), f32)                          # real, mismatched, fake, interpolated
        self.g = ops.empty((B, module.in_hw, module.in_hw, 3))       # d score_interp / d interp
        self.v0 = ops.empty((B, module.in_hw, module.in_hw, 3))
        self.dx = ops.empty((G * B, module.in_hw, module.in_hw, 3))  # d loss / d images (when requested)
        self.sq = ops.empty((B,), f32)
        self.dtem = ops.zeros((B, module.tem_size), f32)
        # per-sample head coefficients (d loss / d score)
        cc = torch.zeros(G * B, dtype=f32)
        cc[:B] = -1.0 / (2 * B)      # real (-1/B) and mismatched (+1/2B) share features
        cc[B:2 * B] = 1.0 / (2 * B)  # fake
        self.coef_critic = cc.to(ops.device)
        ct = torch.zeros(2 * B, dtype=f32)
        ct[:B] = -1.0 / (2 * B)      # text rows: real(-1/B)+fake(+1/2B) use tem; mismatched rows +1/2B
        ct[B:] = 1.0 / (2 * B)
        self.coef_text = ct.to(ops.device)
        self.coef_one = torch.ones(B, dtype=f32).to(ops.device)
        self.coef_gen = torch.full((B,), -1.0 / B, dtype=f32).to(ops.device)

    # ---------------------------------------------------------------- helpers
    def group_view(self, t, g0, ng):
        B = self.B
        return t[g0 * B:(g0 + ng) * B]

    def refresh_weights(self, with_text=False, events=False):
        """bf16 operand packs + the collapsed head from the fp32 masters.  ``with_text``: also the compressed text of the
        current batch (it depends on the compress weights only), so that the next forward finds it ready.  ``events``
        (when this runs on a side stream): record one CUDA event per layer and one for the head, so that the next forward
        waits for layer l's operands right before conv l instead of for the whole re-pack before its first conv (the main
        stream used to stall ~60 us per critic iteration behind head_prepare and the text compression)."""
        ops, m = self.ops, self.m
        evs = [] if (events and not getattr(ops, "is_emulator", False)) else None
        cur = torch.cuda.current_stream(ops.device) if evs is not None else None

        def mark():
            if evs is not None:
                e = torch.cuda.Event()
                e.record(cur)
                evs.append(e)
        for L in self.layers:
            L.pack(ops)
            mark()
        ops.head_prepare(m.channel_resize.weight.data, m.channel_resize.bias.data, m.critic_score.weight.data,
                         m.critic_score.bias.data, self.A, self.Bv, self.c0)
        if with_text:
            ops.linear_fwd(self.tem_all, m.compress.weight.data, m.compress.bias.data, self.ce)
        mark()
        self._pack_events = evs

    def _wait_pack(self, i):
        """Make the current stream wait for the i-th event of the last side-stream re-pack (layer i's operands; -1: the head)."""
        evs = getattr(self, "_pack_events", None)
        if evs:
            torch.cuda.current_stream(self.ops.device).wait_event(evs[i])

    def set_text(self, tem, tem_mis):
        self.tem_all[:self.B].copy_(tem)
        if tem_mis is not None:
            self.tem_all[self.B:].copy_(tem_mis)

    # ---------------------------------------------------------------- forward
    def forward(self, g0, ng, dup_first, training=True, with_mismatched=False, before_weights=None, ce_ready=False,
                patches_on=None):
        """Trunk + head on groups [g0, g0+ng).  BN running statistics are updated once per group,
        group g0 ``dup_first`` times (the mismatched-text call sees the real images again).
        ``before_weights()`` is called right before the first kernel that reads packed weights (the engines re-pack
        them on a side stream after each optimizer step and join here).  ``ce_ready``: the compressed text was already
        computed by ``refresh_weights(with_text=True)`` for the current weights and batch.  ``patches_on``: a SideStream (or
        False for "this stream") on which to build the patch matrix of the input images that the first layer's weight
        gradient will read -- it must be built NOW when the image buffer is overwritten before the backward pass runs
        (Stage-I generates the next fake batch early); None: no weight gradient will be asked for."""
        ops, m, B = self.ops, self.m, self.B
        gv = lambda t: self.group_view(t, g0, ng)
        L0 = self.layers[0]
        if patches_on is not None:
            _side_run(patches_on or None, lambda: ops.patchify(gv(self.a[0]), gv(self.P), L0.k, L0.s, L0.p))
        if training:
            ops.zero(self.stats_flat)
        per_layer = bool(getattr(self, "_pack_events", None))
        if before_weights is not None and not per_layer:
            before_weights()
        self._wait_pack(0)
        ops.conv_fprop(gv(self.a[0]), L0.pf, L0.conv.bias.data, gv(self.a[1]), L0.k, L0.s, L0.p, act=ACT_LRELU)
        bulk = BN_ACT_BULK and training and hasattr(ops, "set_option")
        if bulk:
            ops.set_option("bn_act_bulk", 1)     # launch-time switch: the critic forward runs (nearly) alone on the GPU
        for l in range(1, self.nl):
            L, bn = self.layers[l], self.layers[l].bn
            self._wait_pack(l)
            y = gv(self.y[l])
            mr = self.mr[l][g0:g0 + ng]
            if training:
                st = self.stats[l][g0:g0 + ng]
                ops.conv_fprop_stats(gv(self.a[l]), L.pf, y, st, ng, L.k, L.s, L.p)
                ops.bn_finalize_act(st, y.numel() // (ng * L.co), mr, bn.running_mean, bn.running_var,
                                    bn.num_batches_tracked, dup_first, y, bn.weight.data, bn.bias.data,
                                    gv(self.a[l + 1]), ACT_LRELU)
            else:
                ops.conv_fprop(gv(self.a[l]), L.pf, None, y, L.k, L.s, L.p)
                for g in range(ng):
                    ops.bn_eval_mr(bn.running_mean, bn.running_var, mr[g:g + 1])
                ops.bn_act(y, mr, bn.weight.data, bn.bias.data, gv(self.a[l + 1]), ng, ACT_LRELU)
        if bulk:
            ops.set_option("bn_act_bulk", 0)
        # head: compressed text, then the collapsed affine score
        self._wait_pack(-1)
        if per_layer:
            if before_weights is not None:
                before_weights()                 # the side stream's last event has been waited for: this join is free
            self._pack_events = None             # consumed: later forwards of the same weights need no waits
        nt = 2 * B if with_mismatched else B
        if not ce_ready:
            ops.linear_fwd(self.tem_all[:nt], m.compress.weight.data, m.compress.bias.data, self.ce[:nt])
        # one launch for all score rows: (first activation row, first text row, first score element) per job
        jobs = [(g * B, 0, {0: 0, 1: 2, 2: 3}[g] * B) for g in range(g0, g0 + ng)]
        if with_mismatched:
            jobs.append((0, B, B))
        ops.head_fwd_multi(self.a[self.nl], self.ce, self.A, self.Bv, self.c0, self.score, jobs, B)

    # ---------------------------------------------------------------- first-order backward
    def input_grad(self, dy0, dx):
        """d/d image of the first conv: ConvTranspose2d(C0 -> 3) as one direct kernel (thin_conv.cu)."""
        ops, L0 = self.ops, self.layers[0]
        ops.conv_dgrad(dy0, L0.pd, None, dx, L0.k, L0.s, L0.p)

    def seed_heads(self, coef):
        """d score / d (last activation) = coef (x) A for the plain backward over all three groups and 1 (x) A for the
        penalty's first-order pass: both depend on the head weights only, so the engine issues them behind the re-pack of
        the weights (re-pack stream) and runs the passes with ``seeded=True`` -- two launches less on the main chain."""
        self.ops.head_bwd_data(self.coef_one, self.A, self.gda[self.nl])
        self.ops.head_bwd_data(coef, self.A, self.da[self.nl])

    def zero_pass_buffers(self):
        """The per-channel sums of ALL three passes of a critic update (gp_first_order, gp_second_order, backward over the three
        groups) + the per-sample squared norms, zeroed by one kernel node; the passes then run with ``prezeroed=True``.  The
        engine issues it behind the previous optimizer step on the re-pack stream -- off the main chain."""
        nl = self.nl
        self.ops.zero_multi([self.sums[l] for l in range(1, nl)] + [self.gsums[l] for l in range(1, nl)] +
                            [self.tsums[l] for l in range(1, nl)] + [self.sq, self.sums0, self.gsums0])

    def backward(self, g0, ng, coef, inject, param_grads, need_input_grad, head_reduce=True, input_grad_from=None, side=None,
                 merge_gp=False, prezeroed=False, seeded=False):
        """Backward of sum_n coef[n]*score[n] over groups [g0,g0+ng) (+ ``inject``: extra
        d loss / d y_l on the interpolated group from the gradient-penalty second-order pass).  ``coef`` holds the ng*B
        coefficients of these groups.  Groups are independent (own BN statistics, disjoint buffer slices; parameter gradients
        accumulate with atomic adds), so two calls on disjoint groups may run on different streams.  ``merge_gp`` (with all
        three groups): every conv weight gradient also covers the gradient penalty's second-order term -- the fourth group of
        a_full / dy_full, filled by gp_first_order / gp_second_order(defer_wgrad=True) -- in the same launch."""
        merge_gp = merge_gp and param_grads and g0 == 0 and ng == self.NG
        ops, nl = self.ops, self.nl
        gv = lambda t: self.group_view(t, g0, ng)
        a4 = gv(self.a[nl])
        if not prezeroed:
            ops.zero_multi([self.sums[l][g0:g0 + ng] for l in range(1, nl)])  # this pass's BatchNorm-backward sums, one node
        if not seeded:
            ops.head_bwd_data(coef, self.A, gv(self.da[nl]))
        if param_grads and head_reduce:                  # (same stream as gp_second_order's term: both ADD to dA)
            _side_run(side, lambda: ops.head_bwd_reduce(coef, a4, self.dA))
        bn_items = []             # (sums, gamma.grad, beta.grad) of every BatchNorm layer: ONE launch at the end
        reduced = False           # sums[l] already came out of the epilogue of the conv that produced da[l + 1]
        L0, dy0, masked0 = self.layers[0], gv(self.dy[0]), False
        fuse = BN_FUSED_GP1 and not param_grads and hasattr(ops, "set_option")     # no wgrad stream next to this pass
        if fuse:
            ops.set_option("bn_fused", 1)
        for l in range(nl - 1, 0, -1):
            L, bn = self.layers[l], self.layers[l].bn
            mr, sums = self.mr[l][g0:g0 + ng], self.sums[l][g0:g0 + ng]
            da, a_out, y, dy = gv(self.da[l + 1]), gv(self.a[l + 1]), gv(self.y[l]), gv(self.dy[l])
            if reduced:
                ops.bn_bwd_apply(da, a_out, y, mr, bn.weight.data, sums, dy, ng, ACT_LRELU,
                                 inject=self.gy[l] if inject else None, inject_group=2 - g0, beta=bn.bias.data)
            else:
                ops.bn_bwd(da, a_out, y, mr, bn.weight.data, sums, dy, ng, ACT_LRELU,
                           inject=self.gy[l] if inject else None, inject_group=2 - g0, beta=bn.bias.data, zeroed=True)
            if param_grads:
                bn_items.append((sums, bn.weight.grad, bn.bias.grad))
                if merge_gp:
                    _side_run(side, lambda l=l, L=L: ops.conv_wgrad(self.a_full[l], self.dy_full[l], L.conv.weight.grad, L.k, L.s, L.p))
                else:
                    _side_run(side, lambda l=l, L=L, dy=dy: ops.conv_wgrad(gv(self.a[l]), dy, L.conv.weight.grad, L.k, L.s, L.p))
            if l >= 2:
                # d/d a of layer l-1 + that layer's BatchNorm-backward statistics in the same kernel
                bnb = self.layers[l - 1].bn
                reduced = ops.conv_bstats_opt("d", dy, L.pd, gv(self.da[l]), gv(self.y[l - 1]), self.mr[l - 1][g0:g0 + ng],
                                              bnb.weight.data, bnb.bias.data, self.sums[l - 1][g0:g0 + ng], ng, ACT_LRELU,
                                              L.k, L.s, L.p)
            elif MASKED_DGRAD and ops.conv_dgrad_masked_supported(dy, dy0, L.k, L.s, L.p, ng):
                # l == 1: the data-gradient conv stores dy0 = da1 * act'(a1) itself (no activation-backward pass over the
                # critic's largest tensor) and its epilogue leaves the first conv's bias gradient in sums0[.][.][0]
                s0 = self.sums0[g0:g0 + ng]
                ops.conv_dgrad_masked(dy, L.pd, dy0, gv(self.a[1]), self.id_mr[g0:g0 + ng], self.id_gamma, self.id_beta, s0, ng,
                                      ACT_LRELU, L.k, L.s, L.p, zeroed=prezeroed)
                if param_grads:
                    bn_items.append((s0, self.bias_scratch, L0.conv.bias.grad))
                masked0 = True
            else:
                ops.conv_dgrad(dy, L.pd, None, gv(self.da[l]), L.k, L.s, L.p)
        if fuse:
            ops.set_option("bn_fused", 0)
        if not masked0:
            # the first conv's bias gradient (column sums of dy0) rides in the activation-backward pass
            ops.act_bwd(gv(self.da[1]), gv(self.a[1]), dy0, ACT_LRELU, colsum=L0.conv.bias.grad if param_grads else None)
        if param_grads:
            def pgrad0():
                ops.bn_param_grad_multi(bn_items)
                if merge_gp:
                    ops.conv_wgrad(self.P_full, self.dy_full[0], L0.conv.weight.grad.view(L0.co, self.K0, 1, 1), 1, 1, 0)
                else:
                    ops.conv_wgrad(gv(self.P), dy0, L0.conv.weight.grad.view(L0.co, self.K0, 1, 1), 1, 1, 0)
            _side_run(side, pgrad0)
        if need_input_grad:
            # only for the groups [input_grad_from, g0+ng): the real images need no gradient
            gi = g0 if input_grad_from is None else input_grad_from
            sl = lambda t: self.group_view(t, gi, g0 + ng - gi)
            self.input_grad(sl(self.dy[0]), sl(self.dx))

    def text_backward(self, coef_text, nt, dc0, param_grads, dtem):
        """Head/text parameter grads for d loss/d score coefficients; dtem (=) d/d tem rows [0,B)."""
        ops, m, B = self.ops, self.m, self.B
        ops.head_bwd_data(coef_text[:nt], self.Bv, self.dce[:nt])            # dce = coef (x) Bv
        if param_grads:
            ops.head_bwd_reduce(coef_text[:nt], self.ce[:nt], self.dBv)
            ops.fill(self.dc0, dc0)
            ops.linear_bwd(self.tem_all[:nt], m.compress.weight.data, self.dce[:nt], m.compress.weight.grad,
                           m.compress.bias.grad, None, dx_acc=False)
            ops.head_param_grads(self.dA, self.dBv, self.dc0, m.channel_resize.weight.data,
                                 m.channel_resize.bias.data, m.critic_score.weight.data,
                                 m.channel_resize.weight.grad, m.channel_resize.bias.grad,
                                 m.critic_score.weight.grad, m.critic_score.bias.grad)
        if dtem is not None:
            ops.linear_bwd(self.tem_all[:B], m.compress.weight.data, self.dce[:B], None, None, dtem, dx_acc=False)

    # ---------------------------------------------------------------- gradient penalty
    def gp_first_order(self, prezeroed=False, seeded=False):
        """g = d sum_b score_interp[b] / d interp through train-mode BN (utils.py:15-21)."""
        ops, nl = self.ops, self.nl
        i2 = lambda t: self.group_view(t, 2, 1)
        if not prezeroed:
            ops.zero_multi([self.gsums[l] for l in range(1, nl)] + [self.sq])
        if not seeded:
            ops.head_bwd_data(self.coef_one, self.A, self.gda[nl])
        reduced, masked0 = False, False
        fuse = BN_FUSED_GP1 and hasattr(ops, "set_option")
        if fuse:
            ops.set_option("bn_fused", 1)      # (launch-time switch: baked into the captured graph)
        for l in range(nl - 1, 0, -1):
            L, bn = self.layers[l], self.layers[l].bn
            mr = self.mr[l][2:3]
            if reduced:
                ops.bn_bwd_apply(self.gda[l + 1], i2(self.a[l + 1]), i2(self.y[l]), mr, bn.weight.data, self.gsums[l],
                                 self.gdy[l], 1, ACT_LRELU, beta=bn.bias.data)
            else:
                ops.bn_bwd(self.gda[l + 1], i2(self.a[l + 1]), i2(self.y[l]), mr, bn.weight.data, self.gsums[l],
                           self.gdy[l], 1, ACT_LRELU, beta=bn.bias.data, zeroed=True)
            if l >= 2:
                bnb = self.layers[l - 1].bn
                reduced = ops.conv_bstats_opt("d", self.gdy[l], L.pd, self.gda[l], i2(self.y[l - 1]), self.mr[l - 1][2:3],
                                              bnb.weight.data, bnb.bias.data, self.gsums[l - 1], 1, ACT_LRELU, L.k, L.s, L.p)
            elif MASKED_DGRAD and ops.conv_dgrad_masked_supported(self.gdy[l], self.gdy[0], L.k, L.s, L.p, 1):
                ops.conv_dgrad_masked(self.gdy[l], L.pd, self.gdy[0], i2(self.a[1]), self.id_mr[2:3], self.id_gamma, self.id_beta,
                                      self.gsums0, 1, ACT_LRELU, L.k, L.s, L.p, zeroed=prezeroed)
                masked0 = True
            else:
                ops.conv_dgrad(self.gdy[l], L.pd, None, self.gda[l], L.k, L.s, L.p)
        if fuse:
            ops.set_option("bn_fused", 0)
        if not masked0:
            ops.act_bwd(self.gda[1], i2(self.a[1]), self.gdy[0], ACT_LRELU)
        self.input_grad(self.gdy[0], self.g)
        ops.sample_sqnorm(self.g, self.sq, zeroed=True)

    def gp_second_order(self, coef, side=None, defer_wgrad=False, prezeroed=False):
        """Backward of coef/2 * sum_b (||g_b||-1)^2 through the first-order graph: parameter grads
        via wgrad / gamma / head, and gy[l] = d/d y_l for the plain backward to pick up.  ``defer_wgrad``: leave the conv
        weight-gradient terms to the plain backward that follows (backward(merge_gp=True) covers w[l] / gdy[l] as a fourth
        group of its own launches); only the patch matrix of v0 is built here."""
        ops, nl = self.ops, self.nl
        i2 = lambda t: self.group_view(t, 2, 1)
        if not prezeroed:
            ops.zero_multi([self.tsums[l] for l in range(1, nl)])
        ops.gp_seed(self.g, self.sq, coef, self.v0)
        L0 = self.layers[0]
        ops.conv_fprop(self.v0, L0.pf, None, self.v[0], L0.k, L0.s, L0.p)

        def pgrad_v0():
            ops.patchify(self.v0, self.Pv, L0.k, L0.s, L0.p)
            if not defer_wgrad:
                ops.conv_wgrad(self.Pv, self.gdy[0], L0.conv.weight.grad.view(L0.co, self.K0, 1, 1), 1, 1, 0)
        _side_run(side, pgrad_v0)
        ops.act_bwd(self.v[0], i2(self.a[1]), self.w[1], ACT_LRELU)
        for l in range(1, nl):
            L, bn = self.layers[l], self.layers[l].bn
            ops.conv_fprop(self.w[l], L.pf, None, self.v[l], L.k, L.s, L.p)
            if not defer_wgrad:
                _side_run(side, lambda l=l, L=L: ops.conv_wgrad(self.w[l], self.gdy[l], L.conv.weight.grad, L.k, L.s, L.p))
            mr = self.mr[l][2:3]
            ops.gp_bn(self.v[l], self.gda[l + 1], i2(self.a[l + 1]), i2(self.y[l]), mr, bn.weight.data, self.gsums[l],
                      self.tsums[l], self.w[l + 1], self.gy[l], bn.weight.grad, ACT_LRELU, zeroed=True)
        # dA is read by the head / text parameter gradients only (side stream): keep the reduction off the data-gradient chain
        _side_run(side, lambda: ops.head_bwd_reduce(self.coef_one, self.w[nl], self.dA))


def export_optimizer_state(opt, fp):
    """Mirror the fused Adam state (flat moments + device step counter) into a torch.optim.Adam so that
    ``opt.state_dict()`` checkpoints carry exp_avg / exp_avg_sq / step like the reference's
    (stage_1_train_fn.py:218-222, stage_2_train_fn.py:219-221)."""
    step = float(fp.step_count())
    off = 0
    for p in fp.params:
        k = p.numel()
        opt.state[p] = {"step": torch.tensor(step), "exp_avg": fp.m[off:off + k].view(p.shape).clone(),
                        "exp_avg_sq": fp.v[off:off + k].view(p.shape).clone()}
        off += k


def import_optimizer_state(opt, fp):
    """The inverse, for resuming (the reference restores its optimizers, stage_1_train_fn.py:69-73,
    stage_2_train_fn.py:84-86): after ``opt.load_state_dict(checkpoint[...])`` copy exp_avg / exp_avg_sq / step of every
    parameter into the flat moment buffers and the device step counter the fused Adam kernel uses.  Returns the number of
    parameters that had state."""
    off, step, found = 0, None, 0
    for p in fp.params:
        k = p.numel()
        st = opt.state.get(p)
        if st is not None and "exp_avg" in st:
            fp.m[off:off + k].copy_(st["exp_avg"].detach().reshape(-1))
            fp.v[off:off + k].copy_(st["exp_avg_sq"].detach().reshape(-1))
            t = float(st["step"])
            assert step is None or step == t, "parameters of one optimizer carry different step counts"
            step, found = t, found + 1
        off += k
    if step is not None:
        fp.set_step(step)
    return found


class _SegmentedGraph:
    """A train step captured as a SEQUENCE of CUDA graphs with the NCCL calls issued eagerly between them.

    Capturing NCCL collectives inside a CUDA graph hung on this stack (torch 2.11 / NCCL 2.28, see
    DESIGN.md); cutting the graph at every hand-off to the communicator keeps the ~650 kernel launches
    of a step off the host's critical path and still lets the side-stream all-reduce overlap the next
    graph segment."""

    def __init__(self, ops):
        self.ops, self.items, self.g, self.n0 = ops, [], None, 0
        self.capturing = False

    def begin(self):
        self.g = torch.cuda.CUDAGraph()
        self.n0 = self.ops.launch_count()
        self.g.capture_begin(capture_error_mode="thread_local")
        self.capturing = True

    def _close(self):
        self.g.capture_end()
        self.capturing = False
        if self.ops.launch_count() > self.n0:
            self.items.append(self.g)
        self.g = None

    def cut(self, action):
        self._close()
        self.items.append(action)
        self.begin()

    def end(self):
        self._close()

    def replay(self):
        for it in self.items:
            if isinstance(it, torch.cuda.CUDAGraph):
                it.replay()
            else:
                it()


# ============================================================================================ Stage-I engine
class Stage1Engine:
    """One reference outer step (stage_1_train_fn.py:116-172): five critic updates + one
    generator/CA update, with all noise supplied by the caller."""

    def __init__(self, ca, critic, gen, batch_size, ops=None, lr=1e-3, world_size=1, allreduce=None, comm=None):
        ops = ops or default_ops()
        self.ops, self.B = ops, batch_size
        self.ca_m, self.d_m, self.g_m = ca, critic, gen
        with _alloc_ctx(comm):
            self.d = CriticRT(ops, critic, batch_size)
            self.ca = CART(ops, ca)
            self.ca.ensure(batch_size)
            # the generator writes its tanh output straight into the critic's "fake" group
            self.g = GenRT(ops, gen, batch_size, out=self.d.group_view(self.d.a[0], 1, 1))
        for fp in (self.d.fp, self.g.fp, self.ca.fp):
            fp.set_lr(lr)
        self.losses = ops.zeros((4,), ops.f32)       # [loss_critic, gp, lossG, kl]
        self.side = SideStream(ops)
        self.gen_side = SideStream(ops)               # next iteration's generator forward (critic_iteration)
        # weight re-packing after the critic's optimizer step (its priority measured irrelevant: 5.34-5.40 ms at 0, -1, -2)
        self.pack_side = SideStream(ops, priority=_prio("SG_PACK_PRIO", 0))
        self._ce_ready = False                        # compressed text valid for the current weights + batch
        self._fake_ready = False
        self._interp_ready = False       # the next iteration's interpolated images were produced ahead as well
        self.allreduce = allreduce                   # callable(flat_grad) or None (legacy, unbucketed)
        self.comm = comm                             # comm.PeerComm / comm.DistComm or None
        self.world = world_size
        if comm is not None:
            for fp in (self.d.fp, self.g.fp, self.ca.fp):
                comm.broadcast_params(fp.flat)       # train.py:78-85
        self.real_nchw = None
        self.refresh_all()

    def refresh_all(self):
        self.d.refresh_weights()
        self.g.refresh_weights()

    def optimizer_step(self, fp):
        """xm.optimizer_step (stage_1_train_fn.py:149,166,172): average gradients over replicas, then Adam."""
        self.side.join()
        if self.comm is not None and self.comm.peer:
            self.comm.step(fp)               # reduce-scatter + Adam + all-gather over peer memory, one kernel, in the graph
            return
        if self.comm is not None:
            self._comm_allreduce(fp.grad)
            self._comm_wait()
        elif self.allreduce is not None:
            self.allreduce(fp.grad)
        self.ops.adam_step(fp.flat, fp.grad, fp.m, fp.v, fp.hyper)

    def gather_optimizer_state(self):
        """COLLECTIVE (every rank): make the sharded Adam moments whole before a checkpoint is written."""
        if self.comm is not None and self.comm.peer:
            for fp in (self.d.fp, self.g.fp, self.ca.fp):
                self.comm.gather_state(fp)

    def export_optimizer_state(self, opt, fp):
        export_optimizer_state(opt, fp)

    def import_optimizer_state(self, opt, fp):
        import_optimizer_state(opt, fp)

    def _comm_allreduce(self, t):
        seg = getattr(self, "_seg", None)
        if seg is not None and seg.capturing:
            self.gen_side.join()                                # a graph segment must end with every fork joined
            self.pack_side.join()
            seg.cut(lambda: self.comm.allreduce_async(t))       # eager NCCL between two graph segments
        else:
            self.comm.allreduce_async(t)

    def _comm_wait(self):
        seg = getattr(self, "_seg", None)
        if seg is not None and seg.capturing:
            seg.cut(self.comm.wait_all)
        else:
            self.comm.wait_all()

    # -- inputs
    def load_batch(self, real_nchw, tem, tem_mis):
        d = self.d
        self.ops.nchw_to_nhwc(real_nchw, d.group_view(d.a[0], 0, 1))
        d.set_text(tem, tem_mis)
        self._ce_ready = False

    def _generate(self, z, eps_ca):
        self.ca.forward(self.d.tem_all[:self.B], eps_ca, z, cg=self.g.cg)   # stage_1_train_fn.py:120-122
        self.g.forward(training=True)                            # :123 -> critic group 1

    def critic_iteration(self, z, eps_ca, eps_gp, next_noise=None, grads_zeroed=False, zero_after=False):
        """One critic update (stage_1_train_fn.py:120-149).  ``next_noise = (z, eps_ca)`` of the FOLLOWING iteration,
        if given, lets its fake batch be generated on a second stream while this iteration's gradient penalty and
        backward run: the generator's weights do not change between critic updates, and once the critic forward has
        consumed the image buffer nothing reads it until the next iteration.  ``grads_zeroed``: the critic's gradient
        buffers are already zero (the previous iteration ran with ``zero_after``: they were cleared behind its optimizer
        step on the re-pack stream, off the main chain, together with the per-channel sums of this iteration's three passes and
        the head seeds of its two backward passes -- seven graph nodes less between forward and optimizer step).  A third
        element of ``next_noise`` (the next eps_gp) moves the next interpolation onto the generator's side stream too."""
        ops, d, B = self.ops, self.d, self.B
        X = d.a[0]
        interp = lambda e: ops.interp(d.group_view(X, 0, 1), d.group_view(X, 1, 1), e, d.group_view(X, 2, 1))   # utils.py:10-11
        if self._fake_ready:
            self.gen_side.join()
            self._fake_ready = False
            if not self._interp_ready:
                interp(eps_gp)
        else:
            self._generate(z, eps_ca)
            interp(eps_gp)
        self._interp_ready = False
        d.forward(0, 3, dup_first=2, training=True, with_mismatched=True,    # :125-132 + utils.py:13
                  before_weights=self.pack_side.join, ce_ready=self._ce_ready, patches_on=self.side)
        def generate_ahead():
            if next_noise is not None and self.gen_side.enabled:
                self.side.join()             # the patch matrix of this iteration's images is built: group 1 may be overwritten
                if len(next_noise) > 2:      # + the next interpolation (its eps given): one launch less at the head of the chain
                    self.gen_side.run(lambda: (self._generate(*next_noise[:2]), interp(next_noise[2])))
                    self._interp_ready = True
                else:
                    self.gen_side.run(lambda: self._generate(*next_noise))
                self._fake_ready = True
        if GEN_AHEAD_AT == 0:
            generate_ahead()
        if not grads_zeroed:
            ops.zero(d.fp.grad)                                  # :146
            ops.zero(d.head_grads)                               # dA, dBv
            d.zero_pass_buffers()
        d.gp_first_order(prezeroed=True, seeded=grads_zeroed)    # utils.py:15-24
        if GEN_AHEAD_AT == 1:
            generate_ahead()
        # :140-144; only the host reads the loss values: off the main stream
        self.side.run(lambda: ops.critic_loss(d.score[0], d.score[1], d.score[2], d.sq, LAMBDA_GP, self.losses[0:2]))
        d.gp_second_order(2.0 * LAMBDA_GP / B, side=self.side, defer_wgrad=True, prezeroed=True)
        if GEN_AHEAD_AT == 2:
            generate_ahead()
        # head/text gradients on the side stream, in order: dA is complete once the plain backward's head term is added
        self.side.run(lambda: (ops.head_bwd_reduce(d.coef_critic, d.a[d.nl], d.dA),
                               d.text_backward(d.coef_text, 2 * B, 0.0, True, None)))
        # One batched backward over all three groups.  Running the gradient-penalty chain (one group) and the interpolated
        # group's backward on a second stream next to a two-group backward was measured and REJECTED (B200, round 2:
        # Stage-I 5.56 -> 5.98 ms, Stage-II 34.6 -> 36.1 ms, 508 -> 603 launches): the persistent conv kernels take every SM
        # they can get, so the two chains do not really overlap, and every split launch pays its fixed ~10 us again.
        d.backward(0, 3, d.coef_critic, inject=True, param_grads=True, need_input_grad=False,   # :147
                   head_reduce=False, side=self.side, merge_gp=True, prezeroed=True, seeded=grads_zeroed)
        self.optimizer_step(d.fp)                                # :149
        # re-pack the bf16 operands on a side stream: the next forward's interpolation / patch matrix need no weights
        def after_step():
            d.refresh_weights(with_text=True, events=True)
            if zero_after:                                       # the gradients are dead once the optimizer has read them (:146 of
                ops.zero(d.fp.grad)                              # the NEXT iteration, issued here: nothing writes them before its
                ops.zero(d.head_grads)                           # gradient-penalty pass, which waits for this stream's weights)
                d.zero_pass_buffers()                            # and the per-channel sums of its three passes
                d.seed_heads(d.coef_critic)                      # its backward seeds: functions of the new head weights only
        self.pack_side.run(after_step)
        self._ce_ready = True                                    # until the text changes (load_batch / next outer step)

    def generator_step(self, grads_zeroed=False):
        ops, d, B = self.ops, self.d, self.B
        d.forward(1, 1, dup_first=1, training=True, before_weights=self.pack_side.join,   # :154 (updated critic, last fake)
                  ce_ready=self._ce_ready)
        st = self.ca.st
        ops.gen_loss(d.score[2], st.mu, st.sigma, self.losses[2:4])          # :155-159
        if not grads_zeroed:
            ops.zero(self.g.fp.grad); ops.zero(self.ca.fp.grad)  # :161-164
        d.backward(1, 1, d.coef_gen, inject=False, param_grads=False, need_input_grad=True)
        d.text_backward(d.coef_gen, B, -1.0, False, d.dtem)      # d lossG/d tem through the critic head
        self.g.backward(d.group_view(d.dx, 1, 1), side=self.side)
        self.ca.backward(self.g.dcg, 1.0, d.dtem, True)
        self.optimizer_step(self.g.fp)                           # :166
        self.optimizer_step(self.ca.fp)                          # :172
        self.g.refresh_weights()

    def outer_step(self, z, eps_ca, eps_gp):
        """z [5,B,100], eps_ca [5,B,128], eps_gp [5,B] (fp32, device)."""
        self._ce_ready = False                                   # a new batch: its text has not been compressed yet
        ops = self.ops
        # :161-164 early and off the main chain: nothing touches the generator's / CA's gradients during the critic updates,
        # and every optimizer_step joins this stream
        self.side.run(lambda: (ops.zero(self.g.fp.grad), ops.zero(self.ca.fp.grad)))
        for it in range(N_CRITIC):
            nxt = (z[it + 1], eps_ca[it + 1], eps_gp[it + 1]) if it + 1 < N_CRITIC else None
            self.critic_iteration(z[it], eps_ca[it], eps_gp[it], next_noise=nxt, grads_zeroed=it > 0,
                                  zero_after=it + 1 < N_CRITIC)
        self.generator_step(grads_zeroed=True)

    # -- whole step behind static buffers, replayed as one CUDA graph
    def _ensure_static(self):
        if getattr(self, "s_real", None) is None:
            ops, B, f = self.ops, self.B, self.ops.f32
            hw = self.d_m.in_hw
            self.s_real = ops.empty((B, 3, hw, hw), f)
            self.s_z = ops.empty((N_CRITIC, B, Z_DIM), f)
            self.s_eca = ops.empty((N_CRITIC, B, self.ca_m.c_dim), f)
            self.s_egp = ops.empty((N_CRITIC, B), f)
            self.graph = None
            self.launches_per_step = None

    def _body(self):
        d = self.d
        # programmatic dependent launch measured -1.5 % on this step (many short kernels on three streams) and +2 % on
        # the Stage-II step; the attribute is baked into the launches (and the captured graph) issued below
        if hasattr(self.ops, "set_option"):
            self.ops.set_option("pdl", int(os.environ.get("SG_PDL_S1", "0")))
        self.ops.nchw_to_nhwc(self.s_real, d.group_view(d.a[0], 0, 1))
        self.outer_step(self.s_z, self.s_eca, self.s_egp)

    def step(self, real_nchw, tem, tem_mis, z, eps_ca, eps_gp, use_graph=True):
        """One outer step.  Inputs may live on the host (pinned) or the device; they are copied into
        static device buffers, then the ~650 kernels of the step run as one CUDA-graph replay
        (Stage-I kernels are microseconds long: launch-bound otherwise, SURVEY.md section 7 hard part 3)."""
        self._ensure_static()
        if (not real_nchw.is_cuda) and real_nchw.is_pinned() and not getattr(self.ops, "is_emulator", False):
            # host batch: the 6 MB image upload goes over a copy stream into one of two staging buffers, so that batch
            # k+1 crosses PCIe while step k still computes; the compute stream only does a device-to-device copy
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.ops.device)
                self._stage = [torch.empty_like(self.s_real) for _ in range(2)]
                self._stage_free = [None, None]
                self._stage_i = 0
            i = self._stage_i
            self._stage_i ^= 1
            cur = torch.cuda.current_stream(self.ops.device)
            with torch.cuda.stream(self._copy_stream):
                if self._stage_free[i] is not None:
                    self._copy_stream.wait_event(self._stage_free[i])      # the step that last read this buffer
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
        self.d.tem_all[:self.B].copy_(tem, non_blocking=True)
        self.d.tem_all[self.B:].copy_(tem_mis, non_blocking=True)
        self.s_z.copy_(z, non_blocking=True)
        self.s_eca.copy_(eps_ca, non_blocking=True)
        self.s_egp.copy_(eps_gp, non_blocking=True)
        if not use_graph or getattr(self.ops, "is_emulator", False):
            n0 = self.ops.launch_count() if hasattr(self.ops, "launch_count") else 0
            self._body()
            if hasattr(self.ops, "launch_count"):
                self.launches_per_step = self.ops.launch_count() - n0
            return
        if self.comm is not None and self.comm.world > 1 and not self.comm.peer:
            # multi-GPU over NCCL / gloo all-reduce: graph segments on a private stream, the collective eager in between
            if getattr(self, "gstream", None) is None:
                self.gstream = torch.cuda.Stream(device=self.ops.device)
            cur = torch.cuda.current_stream(self.ops.device)
            self.gstream.wait_stream(cur)
            with torch.cuda.stream(self.gstream):
                if self.graph is None:
                    torch.cuda.synchronize()
                    n0 = self.ops.launch_count()
                    self._seg = _SegmentedGraph(self.ops)
                    self._seg.begin()
                    try:
                        self._body()
                    finally:
                        self._seg.end()
                    self.launches_per_step = self.ops.launch_count() - n0
                    self.graph = self._seg
                self.graph.replay()
            cur.wait_stream(self.gstream)
            return
        if self.graph is None:
            torch.cuda.synchronize()
            n0 = self.ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=_capture_stream(self.ops.device, -1)):
                self._body()
            self.launches_per_step = self.ops.launch_count() - n0
            self.graph = g
        self.graph.replay()
