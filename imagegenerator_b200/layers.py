"""Parameter containers that reproduce the reference modules' ``state_dict`` layout and default
initialisation without using torch.nn compute layers.

The reference builds its networks from ``nn.Linear`` / ``nn.Conv2d`` / ``nn.ConvTranspose2d`` /
``nn.BatchNorm2d`` (e.g. generator_1.py:9-36, discrminator_1.py:9-39) and never overrides their
initialisation, so checkpoints interchange iff the key names, shapes and (for seeded construction)
the order of RNG draws agree.  The holders below only own tensors; all
arithmetic is done by the CUDA kernels behind ``imagegenerator_b200.ops``.
"""
from __future__ import annotations

import contextlib
import math

import torch
from torch import nn


def _default_init_(weight, bias, fan_in):
    # what torch's Linear/_ConvNd.reset_parameters do: kaiming_uniform(a=sqrt(5)) then U(+-1/sqrt(fan_in))
    nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    if bias is not None:
        bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
        nn.init.uniform_(bias, -bound, bound)


class DenseParams(nn.Module):
    """weight [out, in] (+ bias [out])  -- state_dict twin of nn.Linear."""

    def __init__(self, n_in, n_out):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n_out, n_in))
        self.bias = nn.Parameter(torch.empty(n_out))
        _default_init_(self.weight, self.bias, n_in)


class ConvParams(nn.Module):
    """Conv2d-layout weight [Cout, Cin, k, k]; ``transposed`` stores ConvTranspose2d's [Cin, Cout, k, k].

    In both cases dim 0 is what this package calls the layer's *wide* side ``co`` and dim 1 its
    *narrow* side ``ci`` of the underlying Conv2d operator (a ConvTranspose2d is that operator's
    data-gradient), so one set of kernels serves both ("one operator, two directions")."""

    def __init__(self, c_in, c_out, k, stride, pad, bias=False, transposed=False):
        super().__init__()
        self.c_in, self.c_out, self.k, self.stride, self.pad, self.transposed = c_in, c_out, k, stride, pad, transposed
        shape = (c_in, c_out, k, k) if transposed else (c_out, c_in, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        if bias:
            self.bias = nn.Parameter(torch.empty(c_out))
        else:
            self.register_parameter("bias", None)
        _default_init_(self.weight, self.bias, shape[1] * k * k)   # torch: fan_in = size(1) * receptive field


class BNParams(nn.Module):
    """state_dict twin of nn.BatchNorm2d (eps 1e-5, momentum 0.1)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class Slot(nn.Module):
    """Parameter-less placeholder keeping nn.Sequential indices equal to the reference's
    (activation modules occupy an index there)."""


def block(conv, c_bn):
    """(conv, bn, activation-slot) triple -> keys '<i>.0.weight', '<i>.1.*' like the reference's
    upsampling_block / downsampling_block helpers."""
    return nn.Sequential(conv, BNParams(c_bn), Slot())


class _NoAutograd(torch.autograd.Function):
    """Identity whose backward explains itself.  The module-level forwards run hand-written kernels and carry no autograd
    graph; the reference's callers differentiate through them (``loss.backward(retain_graph=True)``, and the critic twice
    with ``create_graph=True``, stage_1_train_fn.py:147, utils.py:15-21).  Instead of torch's generic "element 0 of tensors
    does not require grad", a caller who tries gets told where the gradients are produced."""

    @staticmethod
    def forward(ctx, out, anchor, name):
        ctx.name = name
        return out.view_as(out)

    @staticmethod
    def backward(ctx, grad):
        raise RuntimeError(
            f"{ctx.name}.forward() runs hand-written CUDA kernels and does not record an autograd graph: "
            "backward()/autograd.grad through it is not supported.  The gradients of the StackGAN step (including the "
            "WGAN-GP double backward) are produced by imagegenerator_b200.stage_1_train_fn.train_1 / "
            "stage_2_train_fn.train_2 (engine.Stage1Engine / engine2.Stage2Engine), which write .grad of every parameter.")


def no_autograd(out, module):
    """Tag ``out`` so that differentiating through ``module.forward`` raises a clear error (only when grad mode is on and the
    module has trainable parameters; otherwise ``out`` is returned as is)."""
    if not torch.is_grad_enabled():
        return out
    anchor = next((p for p in module.parameters() if p.requires_grad), None)
    if anchor is None:
        return out
    return _NoAutograd.apply(out, anchor, type(module).__name__)


_FLAT_ALLOC = None      # callable(numel) -> zeroed fp32 CUDA tensor, or None (torch.zeros)


@contextlib.contextmanager
def flat_allocator(fn):
    """Within this context ``FlatParams`` takes its parameter and gradient buffers from ``fn(numel)`` -- the data-parallel
    engines pass ``comm.PeerComm.alloc`` so that the buffers live in symmetric memory every peer GPU can address."""
    global _FLAT_ALLOC
    prev, _FLAT_ALLOC = _FLAT_ALLOC, fn
    try:
        yield
    finally:
        _FLAT_ALLOC = prev


class FlatParams:
    """All trainable parameters of a module re-pointed into ONE contiguous fp32 buffer, with a
    matching flat gradient buffer and Adam moments: one fused Adam launch and one all-reduce per
    optimizer, and ``module.state_dict()`` keeps working because the Parameters are views."""

    @classmethod
    def of(cls, module, device, dtype=torch.float32):
        """THE flat buffer of ``module``: created on first use, then shared by every runtime built on that module (the
        train engine, the module-level ``forward`` / ``gradient_penalty`` API, the sampler).  A second ``FlatParams`` would
        re-point ``p.data`` / ``p.grad`` away from the buffers a live engine's Adam, all-reduce and captured CUDA graph
        keep using -- training would silently stop while ``state_dict()`` reads the new copy."""
        fp = module.__dict__.get("_flat")
        if fp is not None:
            if fp.flat.device != torch.device(device) or fp.flat.dtype != dtype:
                raise RuntimeError(
                    f"{type(module).__name__} already lives in a flat parameter buffer on {fp.flat.device} ({fp.flat.dtype}); "
                    f"it cannot also be driven by an ops object on {torch.device(device)} ({dtype}) -- build a separate module")
            if any(p.data.data_ptr() != v.data_ptr() for p, v in zip(fp.params, fp.views)):
                raise RuntimeError(f"the parameters of {type(module).__name__} were re-pointed behind its flat buffer "
                                   "(module.to(...) / p.data = ... after an engine was built)")
            return fp
        fp = cls(module, device, dtype=dtype)
        module.__dict__["_flat"] = fp              # not a submodule / buffer: invisible to state_dict() and .to()
        return fp

    def __init__(self, module, device, opt_hyper=None, dtype=torch.float32):
        ps = [p for p in module.parameters()]
        self.params = ps
        n = sum(p.numel() for p in ps)
        # pad to a multiple of 4 floats so the vectorised Adam kernel needs no tail
        self.n = n
        npad = (n + 3) // 4 * 4
        if _FLAT_ALLOC is not None and dtype == torch.float32:
            self.flat, self.grad = _FLAT_ALLOC(npad), _FLAT_ALLOC(npad)
        else:
            self.flat = torch.zeros(npad, dtype=dtype, device=device)
            self.grad = torch.zeros(npad, dtype=dtype, device=device)
        self.m = torch.zeros(npad, dtype=dtype, device=device)
        self.v = torch.zeros(npad, dtype=dtype, device=device)
        off = 0
        self.views = []
        for p in ps:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1).to(device=device, dtype=dtype))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self.views.append(p.data)
            off += k
        lr, b1, b2, eps = opt_hyper or (1e-3, 0.9, 0.999, 1e-8)
        # [lr, beta1, beta2, eps, step, step_lo, step_hi, -]: lives on the device so a captured CUDA graph sees updates.
        # ``step`` (what the kernel's bias correction reads) is a float and stops counting exactly at 2^24; the exact count
        # for checkpoints is step_hi * 2^23 + step_lo (``step_count`` / ``set_step``)
        self.hyper = torch.tensor([lr, b1, b2, eps, 0.0, 0.0, 0.0, 0.0], dtype=dtype, device=device)
        for mod in module.modules():               # buffers follow to the device
            for name, buf in list(mod._buffers.items()):
                if buf is not None:
                    mod._buffers[name] = buf.to(device=device, dtype=dtype if buf.is_floating_point() else buf.dtype)

    def step_count(self):
        h = self.hyper[4:7].tolist()
        return int(h[2]) * (1 << 23) + int(h[1]) if (h[1] or h[2]) else int(h[0])

    def set_step(self, t):
        t = int(t)
        self.hyper[4] = float(t)
        self.hyper[5] = float(t % (1 << 23))
        self.hyper[6] = float(t // (1 << 23))

    def set_lr(self, lr):
        # a scalar write from pageable host memory synchronises the stream: only touch the device when the value
        # changes (StepLR: every 100 batches), or the train loop would stall on every batch
        if getattr(self, "_lr_host", None) == lr:
            return
        self._lr_host = lr
        self.hyper[0] = lr
