"""Stage-I -> Stage-II forward-only sampling (BASELINE.json configs[4]; reference stage_2_train_fn.py:181-195):

    tem -> con_augment_1 -> [c_hat1, z] -> gen_1 -> fake_64 -> (con_augment_2(tem), fake_64) -> gen_2 -> fake_256

B200 design of the forward-only conv path: every BatchNorm is in eval mode, so it is FOLDED into the conv in front
of it -- the packed bf16 weights carry gamma/sqrt(var+eps) per output channel and the conv epilogue adds the shift
and applies the activation (the residual add of ``ResidualBlock`` too).  No BN kernel, no pre-activation tensor:
one tcgen05 kernel per layer, replayed as one CUDA graph.  Eval-mode BN has no batch coupling, so a large batch is
sharded over GPUs with no collective.

``bn_batch_stats=True`` reproduces the reference's in-loop preview instead, where ``gen_2`` is still in train mode
(:194 under ``gen_2.train()``, :90): batch statistics, running statistics updated -- that path runs the training
forward kernels of ``engine2.Gen2RT``.
"""
from __future__ import annotations

import torch

from .engine import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, CART, GenRT, Up0Gemm, Z_DIM, default_ops


class _Folded:
    """One conv operator with the following eval-mode BN folded in: packed weights + per-channel shift."""

    def __init__(self, ops, conv, bn, direction):
        w = conv.weight
        self.conv, self.bn, self.dir = conv, bn, direction
        self.co, self.ci, self.k, self.s, self.p = w.shape[0], w.shape[1], w.shape[2], conv.stride, conv.pad
        n_out = self.co if direction == "f" else self.ci
        self.pack = ops.empty((self.co, self.k, self.k, self.ci) if direction == "f" else (self.ci, self.k, self.k, self.co))
        self.scale, self.shift = ops.empty((n_out,), ops.f32), ops.empty((n_out,), ops.f32)

    def refresh(self, ops):
        bn, w = self.bn, self.conv.weight.data
        if bn is None:
            ops.pack_weight(w, self.pack if self.dir == "f" else None, self.pack if self.dir == "d" else None)
            return
        ops.bn_fold(bn.running_mean, bn.running_var, bn.weight.data, bn.bias.data, self.scale, self.shift)
        if self.dir == "f":
            ops.pack_weight_scaled(w, self.scale, 0, self.pack, None)
        else:
            ops.pack_weight_scaled(w, self.scale, 1, None, self.pack)

    def run(self, ops, x, out, act, residual=None):
        bias = self.shift if self.bn is not None else (self.conv.bias.data if self.conv.bias is not None else None)
        if self.dir == "d":
            ops.conv_dgrad(x, self.pack, bias, out, self.k, self.s, self.p, act=act)
        elif residual is not None:
            ops.conv_fprop_res(x, self.pack, bias, residual, out, self.k, self.s, self.p, act=act)
        else:
            ops.conv_fprop(x, self.pack, bias, out, self.k, self.s, self.p, act=act)


class StackGANSampler:
    def __init__(self, con_augment_1, gen_1, con_augment_2, gen_2, batch_size, ops=None, bn_batch_stats=False, out_dtype="fp32"):
        """``out_dtype``: "fp32" -- NCHW float images in [-1, 1], what the reference's generators return -- or "uint8": NCHW
        pictures round((x + 1) * 127.5), a quarter of the bytes to read back (a host that collects 8 GPUs' 50 MB fp32
        batches is bound by its own memory / PCIe bandwidth at ~90 GB/s: 116 k of the 437 k images/s the GPUs generate)."""
        ops = ops or default_ops()
        assert out_dtype in ("fp32", "uint8")
        self.ops, self.B, self.bn_batch_stats, self.out_dtype = ops, batch_size, bn_batch_stats, out_dtype
        B, f32 = batch_size, ops.f32
        self.ca1, self.ca2 = CART(ops, con_augment_1), CART(ops, con_augment_2)
        self.ca1.ensure(B)
        self.ca2.ensure(B)
        self.g1_m, self.g2_m = gen_1, gen_2
        for m in (gen_1, gen_2):                         # parameters and BN buffers onto the device (fp32 masters)
            for p in m.parameters():
                p.data = p.data.to(device=ops.device, dtype=f32)
            for mod in m.modules():
                for name, buf in list(mod._buffers.items()):
                    if buf is not None:
                        mod._buffers[name] = buf.to(device=ops.device, dtype=f32 if buf.is_floating_point() else buf.dtype)
        # ---- Stage-I generator (generator_1.py:38-40), always eval here (stage_2_train_fn.py:59-63)
        cl1 = gen_1.conv_layers()
        # first layer: the GEMM of engine.Up0Gemm over the zero-padded [c_hat, z] rows + eval-mode BN as its own small pass
        # (its BN cannot be folded into per-output-column weights without a second transposed pack; the tensor is 4x4x192)
        self.up0, self.bn0 = Up0Gemm(ops, cl1[0][0], B), cl1[0][1]
        self.g1 = [None] + [_Folded(ops, c, bn, "d") for c, bn in cl1[1:]]
        self.cg = ops.zeros((B, 1, 1, self.up0.Kp))
        self.y0 = ops.empty((B, self.up0.k, self.up0.k, self.up0.ci))
        self.mr0 = ops.empty((1, self.up0.ci, 2), f32)
        self.a1, h = [ops.empty((B, self.up0.k, self.up0.k, self.up0.ci))], self.up0.k
        for L in self.g1[1:-1]:
            h = (h - 1) * L.s - 2 * L.p + L.k
            self.a1.append(ops.empty((B, h, h, L.ci)))
        self.fake_64 = ops.empty((B, 2 * h, 2 * h, 3))
        # ---- Stage-II generator (generator_2.py:59-67)
        m = gen_2
        if bn_batch_stats:
            from .engine2 import Gen2RT
            self.g2rt = Gen2RT(ops, gen_2, B, x_in=self.fake_64)
            self.fake_256 = self.g2rt.out
        else:
            self.ds0 = _Folded(ops, m.down_sampler[0], None, "f")
            self.ds2 = _Folded(ops, m.down_sampler[2][0], m.down_sampler[2][1], "f")
            self.res = [[_Folded(ops, c, bn, "f") for c, bn in blk.conv_layers()] for blk in m.residual_blocks]
            self.ups = [_Folded(ops, m.up_sampler[i][0], m.up_sampler[i][1], "d") for i in range(3)]
            self.up3 = _Folded(ops, m.up_sampler[3], None, "d")
            self.x1 = ops.empty((B, 32, 32, 128))
            self.x2 = ops.empty((B, 16, 16, 512))
            self.X = [ops.empty((B, 16, 16, 640)) for _ in range(2)]       # residual-block boundaries, ping-pong
            self.r1, self.r2 = ops.empty((B, 16, 16, 320)), ops.empty((B, 16, 16, 320))
            self.u = [ops.empty((B, 32, 32, 320)), ops.empty((B, 64, 64, 160)), ops.empty((B, 128, 128, 80))]
            self.fake_256 = ops.empty((B, 256, 256, 3))
        self.s_tem = ops.empty((B, con_augment_1.h.weight.shape[1]), f32)
        self.s_z = ops.empty((B, Z_DIM), f32)
        self.s_e1 = ops.empty((B, con_augment_1.c_dim), f32)
        self.s_e2 = ops.empty((B, con_augment_2.c_dim), f32)
        odt = f32 if out_dtype == "fp32" else torch.uint8
        self.out_64 = ops.empty((B, 3, self.fake_64.shape[1], self.fake_64.shape[2]), odt)
        self.out_256 = ops.empty((B, 3, 256, 256), odt)
        self.graph, self.launches = None, None
        self.refresh_weights()

    def refresh_weights(self):
        """Re-pack after the parameters / running statistics changed (e.g. a checkpoint was loaded)."""
        ops = self.ops
        self.up0.pack()
        for L in self.g1[1:]:
            L.refresh(ops)
        if self.bn_batch_stats:
            self.g2rt.refresh_weights()
            return
        self.ds0.refresh(ops)
        self.ds2.refresh(ops)
        for blk in self.res:
            for L in blk:
                L.refresh(ops)
        for L in self.ups + [self.up3]:
            L.refresh(ops)
        self.graph = None

    def _body(self):
        ops = self.ops
        self.ca1.forward(self.s_tem, self.s_e1, self.s_z, cg=self.cg)                 # :184-189
        bn0 = self.bn0                                                                 # :190 gen_1 (eval)
        self.up0.forward(self.cg, self.y0)
        ops.bn_eval_mr(bn0.running_mean, bn0.running_var, self.mr0)
        ops.bn_act(self.y0, self.mr0, bn0.weight.data, bn0.bias.data, self.a1[0], 1, ACT_RELU)
        x = self.a1[0]
        for L, a in zip(self.g1[1:-1], self.a1[1:]):
            L.run(ops, x, a, ACT_RELU)
            x = a
        self.g1[-1].run(ops, x, self.fake_64, ACT_TANH)                                # ConvT(C -> 3) + Tanh, direct kernel
        st2 = self.ca2.forward(self.s_tem, self.s_e2, None)                            # :192
        if self.bn_batch_stats:
            self.g2rt.forward(st2.c_hat, training=True)                                # :193, gen_2 in train mode
        else:
            self.ds0.run(ops, self.fake_64, self.x1, ACT_LRELU)                        # Conv2d(3 -> 128), direct kernel
            self.ds2.run(ops, self.x1, self.x2, ACT_LRELU)
            ops.concat_rep(self.x2, st2.c_hat, self.X[0])
            cur = 0
            for l1, l2, l3 in self.res:
                l1.run(ops, self.X[cur], self.r1, ACT_RELU)
                l2.run(ops, self.r1, self.r2, ACT_RELU)
                l3.run(ops, self.r2, self.X[1 - cur], ACT_RELU, residual=self.X[cur])
                cur = 1 - cur
            x = self.X[cur]
            for L, u in zip(self.ups, self.u):
                L.run(ops, x, u, ACT_RELU)
                x = u
            self.up3.run(ops, x, self.fake_256, ACT_TANH)
        to_out = ops.nhwc_to_nchw if self.out_dtype == "fp32" else ops.nhwc_to_nchw_u8
        to_out(self.fake_64, self.out_64)
        to_out(self.fake_256, self.out_256)

    def sample_to_host(self, tem, z, eps_ca1, eps_ca2, host_256, host_64=None):
        """``sample`` + the device-to-host copy of the images, off the compute stream: the batch is generated by the graph
        replay, parked in one of two device staging buffers (a 50 MB device-to-device copy, ~35 us) and sent to the
        caller's PINNED host tensors over a copy stream while the next batch is already being generated.  Returns a CUDA
        event; ``event.synchronize()`` (or a stream wait) before the host tensors are read.  Without this the 50 MB
        fp32 read-back of a 64-image batch sat on the compute stream and halved the end-to-end rate."""
        dev = self.ops.device
        if getattr(self, "_d2h_stream", None) is None:
            self._d2h_stream = torch.cuda.Stream(device=dev)
            self._park = [(torch.empty_like(self.out_64), torch.empty_like(self.out_256)) for _ in range(2)]
            self._park_free = [None, None]
            self._park_i = 0
        assert host_256.is_pinned() and (host_64 is None or host_64.is_pinned()), "host buffers must be pinned"
        out_64, out_256 = self.sample(tem, z, eps_ca1, eps_ca2)
        i = self._park_i
        self._park_i ^= 1
        cur = torch.cuda.current_stream(dev)
        if self._park_free[i] is not None:
            cur.wait_event(self._park_free[i])                  # the copy that last read this staging pair has finished
        p64, p256 = self._park[i]
        p256.copy_(out_256, non_blocking=True)
        if host_64 is not None:
            p64.copy_(out_64, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._d2h_stream):
            self._d2h_stream.wait_event(ready)
            host_256.copy_(p256, non_blocking=True)
            if host_64 is not None:
                host_64.copy_(p64, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._d2h_stream)
        self._park_free[i] = done
        return done

    def sample(self, tem, z, eps_ca1, eps_ca2, use_graph=True):
        """tem [B,512], z [B,100], eps_ca1/eps_ca2 [B,128] (host or device, fp32) -> (fake_64 [B,3,64,64],
        fake_256 [B,3,256,256]) fp32 NCHW device tensors (static buffers, overwritten by the next call)."""
        for dst, src in ((self.s_tem, tem), (self.s_z, z), (self.s_e1, eps_ca1), (self.s_e2, eps_ca2)):
            dst.copy_(src, non_blocking=True)
        ops = self.ops
        if not use_graph or getattr(ops, "is_emulator", False) or self.bn_batch_stats:
            n0 = ops.launch_count() if hasattr(ops, "launch_count") else 0
            self._body()
            if hasattr(ops, "launch_count"):
                self.launches = ops.launch_count() - n0
            return self.out_64, self.out_256
        if self.graph is None:
            torch.cuda.synchronize()
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.launches = ops.launch_count() - n0
            self.graph = g
        self.graph.replay()
        return self.out_64, self.out_256
