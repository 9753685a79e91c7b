"""ctypes binding of libsgb200.so (include/sgb200.h) -- the only compute backend of this package.

``CudaOps`` exposes one method per C-ABI entry point, taking torch CUDA tensors (used purely as
device-memory handles) and launching on torch's current stream.  There is NO CPU fallback:
constructing ``CudaOps`` without the built library or without an sm_100 device raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SG_LIB") or os.path.join(_HERE, "libsgb200.so")      # SG_LIB: A/B builds of the same library (tools/)

SG_F32, SG_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3

_c = ctypes
_P, _I, _L, _F = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

# name -> argtypes (stream is always the trailing void*)
_SIGS = {
    "sg_zero": [_P, _L, _P],
    "sg_fill_f32": [_P, _F, _L, _P],
    "sg_affine_f32": [_P, _F, _F, _P, _L, _P],
    "sg_nchw_to_nhwc": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sg_nhwc_to_nchw": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sg_pack_weight": [_P, _P, _P, _I, _I, _I, _I, _P],
    "sg_pack_gemm_t": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sg_patchify": [_P, _P] + [_I] * 10 + [_P],
    "sg_bn_fold": [_P, _P, _P, _P, _F, _P, _P, _I, _P],
    "sg_pack_weight_scaled": [_P, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "sg_conv_fprop_res": [_P, _P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_fprop_tc_res": [_P, _P, _P, _P, _P] + [_I] * 11 + [_P],
    "sg_add_act": [_P, _P, _P, _L, _I, _I, _P],
    "sg_unpatchify": [_P, _P, _P] + [_I] * 11 + [_P],
    "sg_conv_fprop_f32out": [_P, _P, _P] + [_I] * 11 + [_P],
    "sg_conv_fprop_tc_f32out": [_P, _P, _P] + [_I] * 10 + [_P],
    "sg_concat_rep": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "sg_split_rep_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "sg_conv_fprop": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_dgrad": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_wgrad": [_P, _P, _P] + [_I] * 11 + [_P],
    "sg_conv_wgrad_cl": [_P, _P, _P] + [_I] * 11 + [_P],
    "sg_fold_grad_cl": [_P, _P, _I, _I, _I, _P],
    "sg_conv_fprop_ffma": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_dgrad_ffma": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_wgrad_ffma": [_P, _P, _P] + [_I] * 11 + [_P],
    "sg_conv_fprop_tc": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_dgrad_tc": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_wgrad_tc": [_P, _P, _P] + [_I] * 11 + [_P],
    "sg_conv_fprop_stats": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_dgrad_stats": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_fprop_tc_stats": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_conv_dgrad_tc_stats": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "sg_colsum": [_P, _P, _L, _I, _I, _P],
    "sg_col_stats": [_P, _P, _L, _I, _I, _I, _P],
    "sg_bn_finalize": [_P, _L, _P, _P, _P, _P, _I, _I, _F, _F, _I, _I, _P],
    "sg_bn_finalize_act": [_P, _L, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "sg_bn_eval_mr": [_P, _P, _P, _F, _I, _P],
    "sg_bn_act": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "sg_bn_bwd_reduce": [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "sg_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _L, _I, _I, _I, _I, _P],
    "sg_bn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _L, _I, _I, _I, _I, _I, _P, _P],
    "sg_zero_multi": [_P, _P, _I, _P],
    "sg_bn_bwd_reduce_y": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "sg_bn_bwd_apply_y": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _L, _I, _I, _I, _I, _P],
    "sg_bn_param_grad": [_P, _P, _P, _I, _I, _P],
    "sg_act_bwd": [_P, _P, _P, _L, _I, _I, _P],
    "sg_act_bwd_colsum": [_P, _P, _P, _P, _L, _I, _I, _I, _P],
    "sg_gp_bn_reduce": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "sg_gp_bn_reduce_acc": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "sg_gp_bn_apply": [_P] * 11 + [_L, _I, _I, _I, _P],
    "sg_gp_bn": [_P] * 11 + [_L, _I, _I, _I, _I, _P, _P],
    "sg_linear_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "sg_linear_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "sg_head_prepare": [_P] * 7 + [_I, _I, _I, _P],
    "sg_head_fwd": [_P] * 6 + [_I, _I, _I, _I, _P],
    "sg_head_fwd_multi": [_P] * 6 + [_I, _P, _P, _P, _I, _I, _I, _I, _P],
    "sg_outer": [_P, _P, _P, _I, _I, _I, _P],
    "sg_wsum_rows": [_P, _P, _P, _I, _I, _I, _P],
    "sg_head_param_grads": [_P] * 10 + [_I, _I, _I, _P],
    "sg_ca_reparam": [_P] * 6 + [_I, _I, _I, _I, _I, _P],
    "sg_ca_bwd_seed": [_P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _P],
    "sg_bn_param_grad_multi": [_P, _P, _P, _P, _P, _I, _P],
    "sg_nhwc_to_nchw_u8": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sg_conv_fprop_bstats": [_P] * 8 + [_I] * 13 + [_P],
    "sg_conv_dgrad_bstats": [_P] * 8 + [_I] * 13 + [_P],
    "sg_conv_bstats_in_epilogue": [_I] * 13,
    "sg_conv_dgrad_tc_bstats_masked": [_P] * 8 + [_I] * 13 + [_P],
    "sg_conv_dgrad_tc_bstats_masked_supported": [_I] * 11,
    "sg_conv_fprop_tc_bstats": [_P] * 8 + [_I] * 12 + [_P],
    "sg_conv_dgrad_tc_bstats": [_P] * 8 + [_I] * 12 + [_P],
    "sg_ca_forward": [_P] * 14 + [_I] * 7 + [_P],
    "sg_ca_backward": [_P] * 4 + [_F] + [_P] * 15 + [_I] * 7 + [_P],
    "sg_interp": [_P, _P, _P, _P, _I, _L, _I, _P],
    "sg_sample_sqnorm": [_P, _P, _I, _L, _I, _P],
    "sg_sample_sqnorm_acc": [_P, _P, _I, _L, _I, _P],
    "sg_gp_seed": [_P, _P, _F, _P, _I, _L, _I, _P],
    "sg_critic_loss": [_P, _P, _P, _P, _F, _P, _I, _P],
    "sg_gen_loss": [_P, _P, _P, _P, _I, _I, _P],
    "sg_scale_rows_add": [_P, _P, _P, _I, _I, _L, _I, _P],
    "sg_adam_step": [_P, _P, _P, _P, _P, _L, _P],
    "sg_conv_narrow_fprop": [_P, _P, _P, _P, _P] + [_I] * 7 + [_P],
    "sg_conv_narrow_dgrad": [_P, _P, _P, _P] + [_I] * 6 + [_P],
    "sg_conv_thin_fprop": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "sg_conv_thin_dgrad": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
}

EXPORTS = sorted(list(_SIGS) + ["sg_version", "sg_last_error", "sg_check_device", "sg_launch_count", "sg_conv_tc_supported", "sg_conv_wgrad_tc_supported", "sg_set_option",
                                 "sg_conv_tc_stats_supported", "sg_conv_wgrad_cl_supported", "sg_debug_conv_trace", "sg_conv_thin_supported", "sg_conv_narrow_supported", "sg_conv_narrow_routed", "sg_init_workspace",
                                 "sg_dp_max_world", "sg_dp_flag_ints", "sg_dp_sync_ints", "sg_dp_adam_step"])


def load_library(path=LIB_PATH):
    """dlopen libsgb200.so and attach argtypes.  Raises if it has not been built."""
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -m imagegenerator_b200.build` "
            "(there is no CPU or PyTorch fallback for the StackGAN kernels)")
    lib = ctypes.CDLL(path)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _I
    lib.sg_conv_tc_supported.argtypes = [_I] * 11
    lib.sg_conv_tc_supported.restype = _I
    lib.sg_conv_wgrad_tc_supported.argtypes = [_I] * 10
    lib.sg_conv_wgrad_tc_supported.restype = _I
    lib.sg_conv_wgrad_cl_supported.argtypes = [_I] * 11
    lib.sg_conv_wgrad_cl_supported.restype = _I
    lib.sg_set_option.argtypes = [_c.c_char_p, _I]
    lib.sg_set_option.restype = _I
    lib.sg_debug_conv_trace.argtypes = [_P]
    lib.sg_debug_conv_trace.restype = _I
    for name in ("sg_dp_max_world", "sg_dp_flag_ints", "sg_dp_sync_ints"):
        getattr(lib, name).argtypes = []
        getattr(lib, name).restype = _I
    lib.sg_dp_adam_step.argtypes = [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P]
    lib.sg_dp_adam_step.restype = _I
    lib.sg_version.restype = _I
    lib.sg_last_error.restype = _c.c_char_p
    lib.sg_check_device.restype = _I
    lib.sg_init_workspace.restype = _I
    lib.sg_launch_count.restype = _L
    return lib


def _ptr(t):
    return None if t is None else t.data_ptr()


class CudaOps:
    """The CUDA backend.  ``dtype``: 'bf16' (tensor-core mode) or 'fp32' (validation mode)."""

    is_emulator = False

    def __init__(self, dtype="bf16", device=None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("imagegenerator_b200 needs a CUDA device (sm_100a); none is visible")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        with torch.cuda.device(self.device):
            if self.lib.sg_check_device() != 0:
                raise RuntimeError(self.lib.sg_last_error().decode())
            if self.lib.sg_init_workspace() != 0:         # one-time allocations (never inside a launch / stream capture)
                raise RuntimeError(self.lib.sg_last_error().decode())
        if dtype in ("bf16", torch.bfloat16):
            self.act_dtype, self.dt = torch.bfloat16, SG_BF16
        elif dtype in ("fp32", torch.float32):
            self.act_dtype, self.dt = torch.float32, SG_F32
        else:
            raise ValueError(dtype)
        self.f32, self.f64 = torch.float32, torch.float64
        # work words of the one-launch BatchNorm backward (sg_bn_bwd): 1 KB per call site, allocated up front
        self._bn_work = torch.zeros(256 * 1024, dtype=torch.int32, device=self.device)
        self._bn_slot = {}
        # A/B measurements without code edits: SG_OPTS="name=value,..." sets library options (include/sgb200.h, sg_set_option)
        for kv in filter(None, os.environ.get("SG_OPTS", "").split(",")):
            self.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    # ---- plumbing
    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"libsgb200 error {rc}: {self.lib.sg_last_error().decode()}")

    def set_option(self, name, value):
        self._ck(self.lib.sg_set_option(name.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.sg_launch_count())

    def conv_trace(self, buf):
        """Profiling hook: ``buf`` = int64 CUDA tensor of >= 296*16 elements receiving %globaltimer stamps, or None."""
        self._ck(self.lib.sg_debug_conv_trace(_ptr(buf)))

    def empty(self, shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.act_dtype, device=self.device)

    def zeros(self, shape, dtype=None):
        return torch.zeros(shape, dtype=dtype or self.act_dtype, device=self.device)

    def _dt_of(self, t):
        if t.dtype == torch.float32:
            return SG_F32
        if t.dtype == torch.bfloat16:
            return SG_BF16
        raise TypeError(f"unsupported tensor dtype {t.dtype}")

    @staticmethod
    def _c(*ts):
        for t in ts:
            if t is not None:
                assert t.is_cuda and t.is_contiguous(), "libsgb200 needs contiguous CUDA tensors"

    # ---- memory
    def zero_multi(self, ts):
        """Up to 32 small buffers (per-channel sums of a backward pass) zeroed by ONE kernel node instead of a memset node in
        front of every reduction (each a graph node between two dependent kernels: ~3 us on the latency-bound main chain)."""
        ts = [t for t in ts if t is not None]
        for i in range(0, len(ts), 32):
            part = ts[i:i + 32]
            self._c(*part)
            n = len(part)
            ptrs = (_c.c_void_p * n)(*[t.data_ptr() for t in part])
            nbytes = (_c.c_int64 * n)(*[t.numel() * t.element_size() for t in part])
            self._ck(self.lib.sg_zero_multi(ptrs, nbytes, n, self._st()))

    def zero(self, t):
        self._c(t)
        self._ck(self.lib.sg_zero(_ptr(t), t.numel() * t.element_size(), self._st()))

    def fill(self, t, value):
        self._c(t)
        assert t.dtype == torch.float32
        self._ck(self.lib.sg_fill_f32(_ptr(t), float(value), t.numel(), self._st()))

    def affine_f32(self, x, a, b, out):
        self._c(x, out)
        self._ck(self.lib.sg_affine_f32(_ptr(x), float(a), float(b), _ptr(out), x.numel(), self._st()))

    # ---- layout
    def nchw_to_nhwc(self, src, dst):
        self._c(src, dst)
        assert src.dtype == torch.float32
        N, C, H, W = src.shape
        assert dst.numel() == src.numel()
        self._ck(self.lib.sg_nchw_to_nhwc(_ptr(src), _ptr(dst), N, C, H, W, self._dt_of(dst), self._st()))

    def nhwc_to_nchw(self, src, dst):
        self._c(src, dst)
        N, C, H, W = dst.shape
        assert dst.dtype == torch.float32 and dst.numel() == src.numel()
        self._ck(self.lib.sg_nhwc_to_nchw(_ptr(src), _ptr(dst), N, C, H, W, self._dt_of(src), self._st()))

    def nhwc_to_nchw_u8(self, src, dst):
        """tanh output in [-1, 1] -> NCHW uint8 image, round((x + 1) * 127.5)."""
        self._c(src, dst)
        N, C, H, W = dst.shape
        assert dst.dtype == torch.uint8 and dst.numel() == src.numel()
        self._ck(self.lib.sg_nhwc_to_nchw_u8(_ptr(src), _ptr(dst), N, C, H, W, self._dt_of(src), self._st()))

    def pack_weight(self, w, pf, pd):
        self._c(w, pf, pd)
        Co, Ci, k, _ = w.shape
        ref = pf if pf is not None else pd
        self._ck(self.lib.sg_pack_weight(_ptr(w), _ptr(pf), _ptr(pd), Co, Ci, k * k, self._dt_of(ref), self._st()))

    def pack_gemm_t(self, w, wt):
        """wt[(t, ci)][Kp] = w[co][ci][t] (columns co >= Co zero)."""
        self._c(w, wt)
        Co, Ci, k, _ = w.shape
        Kp = wt.shape[-1]
        assert wt.numel() == Ci * k * k * Kp
        self._ck(self.lib.sg_pack_gemm_t(_ptr(w), _ptr(wt), Co, Ci, k * k, Kp, self._dt_of(wt), self._st()))

    def bn_fold(self, running_mean, running_var, gamma, beta, scale, shift, eps=1e-5):
        self._c(running_mean, running_var, gamma, beta, scale, shift)
        self._ck(self.lib.sg_bn_fold(_ptr(running_mean), _ptr(running_var), _ptr(gamma), _ptr(beta), eps, _ptr(scale),
                                     _ptr(shift), gamma.numel(), self._st()))

    def pack_weight_scaled(self, w, scale, axis, pf, pd):
        self._c(w, scale, pf, pd)
        Co, Ci, k, _ = w.shape
        ref = pf if pf is not None else pd
        self._ck(self.lib.sg_pack_weight_scaled(_ptr(w), _ptr(scale), axis, _ptr(pf), _ptr(pd), Co, Ci, k * k,
                                                self._dt_of(ref), self._st()))

    def conv_fprop_res(self, x, pf, bias, residual, y, k, s, p, act=ACT_NONE):
        self._c(x, pf, bias, residual, y)
        d = self._conv_dims(x, y)
        self._ck(self.lib.sg_conv_fprop_res(_ptr(x), _ptr(pf), _ptr(bias), _ptr(residual), _ptr(y), *d, k, s, p, act,
                                            self._dt_of(x), self._st()))

    def add_act(self, a, b, out, act):
        self._c(a, b, out)
        self._ck(self.lib.sg_add_act(_ptr(a), _ptr(b), _ptr(out), a.numel(), act, self._dt_of(a), self._st()))

    def concat_rep(self, x, c, out):
        self._c(x, c, out)
        N, H, W, Cx = x.shape
        Cc = c.shape[1]
        assert out.shape[-1] == Cx + Cc and c.dtype == torch.float32
        self._ck(self.lib.sg_concat_rep(_ptr(x), _ptr(c), _ptr(out), N, H * W, Cx, Cc, self._dt_of(x), self._st()))

    def split_rep_bwd(self, dout, dx, dc):
        self._c(dout, dx, dc)
        N, H, W, Cx = dx.shape
        Cc = dc.shape[1]
        self._ck(self.lib.sg_split_rep_bwd(_ptr(dout), _ptr(dx), _ptr(dc), N, H * W, Cx, Cc, self._dt_of(dout), self._st()))

    def patchify(self, x, P, k, s, p):
        self._c(x, P)
        N, H, W, C = x.shape
        _, Ho, Wo, K = P.shape
        assert K == C * k * k
        self._ck(self.lib.sg_patchify(_ptr(x), _ptr(P), N, H, W, C, Ho, Wo, k, s, p, self._dt_of(x), self._st()))

    def unpatchify(self, col, bias, out, k, s, p, act=ACT_NONE):
        """col2im: out[n,oh,ow,c] = act(bias[c] + sum of the taps of col[n,ih,iw, c*k*k+kh*k+kw] landing on (oh,ow))."""
        self._c(col, bias, out)
        N, Hi, Wi, K = col.shape
        _, Ho, Wo, C = out.shape
        assert K == C * k * k and col.dtype == torch.float32
        self._ck(self.lib.sg_unpatchify(_ptr(col), _ptr(bias), _ptr(out), N, Hi, Wi, C, Ho, Wo, k, s, p, act,
                                        self._dt_of(out), self._st()))

    def conv_fprop_f32out(self, x, pf, y, k, s, p):
        """y (fp32) = conv(x, W) without rounding the accumulators to the storage type."""
        self._c(x, pf, y)
        d = self._conv_dims(x, y)
        assert y.dtype == torch.float32
        self._ck(self.lib.sg_conv_fprop_f32out(_ptr(x), _ptr(pf), _ptr(y), *d, k, s, p, self._dt_of(x), self._st()))

    # ---- convolution operator
    def _conv_dims(self, x, y):
        N, H, W, Ci = x.shape
        N2, Ho, Wo, Co = y.shape
        assert N == N2
        return N, H, W, Ci, Ho, Wo, Co

    def conv_fprop(self, x, pf, bias, y, k, s, p, act=ACT_NONE, impl=""):
        self._c(x, pf, bias, y)
        d = self._conv_dims(x, y)
        fn = getattr(self.lib, "sg_conv_fprop" + impl)
        self._ck(fn(_ptr(x), _ptr(pf), _ptr(bias), _ptr(y), *d, k, s, p, act, self._dt_of(x), self._st()))

    def conv_dgrad(self, dy, pd, bias, dx, k, s, p, act=ACT_NONE, impl=""):
        self._c(dy, pd, bias, dx)
        d = self._conv_dims(dx, dy)
        fn = getattr(self.lib, "sg_conv_dgrad" + impl)
        self._ck(fn(_ptr(dy), _ptr(pd), _ptr(bias), _ptr(dx), *d, k, s, p, act, self._dt_of(dy), self._st()))

    def conv_narrow_fprop(self, x, pf, bias, y, act=ACT_NONE, stats=None, groups=1):
        """The direct k4 s2 p1 kernel of narrow_conv.cu by name (conv_fprop / conv_fprop_stats route to it where it is faster)."""
        self._c(x, pf, bias, y, stats)
        N, H, W, Ci, Ho, Wo, Co = self._conv_dims(x, y)
        self._ck(self.lib.sg_conv_narrow_fprop(_ptr(x), _ptr(pf), _ptr(bias), _ptr(y), _ptr(stats), groups, N, H, W, Ci, Co, act, self._st()))

    def conv_narrow_dgrad(self, dy, pd, bias, dx, act=ACT_NONE):
        self._c(dy, pd, bias, dx)
        N, H, W, Ci, Ho, Wo, Co = self._conv_dims(dx, dy)
        self._ck(self.lib.sg_conv_narrow_dgrad(_ptr(dy), _ptr(pd), _ptr(bias), _ptr(dx), N, Ho, Wo, Ci, Co, act, self._st()))

    def conv_fprop_stats(self, x, pf, y, stats, groups, k, s, p):
        """y = conv(x); stats[groups][Co][2] += (sum, sum^2) of y per image group (BN batch statistics)."""
        self._c(x, pf, y, stats)
        d = self._conv_dims(x, y)
        assert stats.dtype == torch.float64 and tuple(stats.shape) == (groups, d[6], 2)
        self._ck(self.lib.sg_conv_fprop_stats(_ptr(x), _ptr(pf), _ptr(y), _ptr(stats), groups, *d, k, s, p,
                                              self._dt_of(x), self._st()))

    def conv_dgrad_stats(self, dy, pd, dx, stats, groups, k, s, p):
        self._c(dy, pd, dx, stats)
        d = self._conv_dims(dx, dy)
        assert stats.dtype == torch.float64 and tuple(stats.shape) == (groups, d[3], 2)
        self._ck(self.lib.sg_conv_dgrad_stats(_ptr(dy), _ptr(pd), _ptr(dx), _ptr(stats), groups, *d, k, s, p,
                                              self._dt_of(dy), self._st()))

    def conv_fprop_bstats(self, x, pf, y, ybn, mr, gamma, beta, sums, groups, act, k, s, p):
        """y = conv(x) is d loss / d a of the layer below (a = act(bn(ybn))); sums[groups][Co][2] = that layer's BN-backward
        statistics, reduced in the conv epilogue (what bn_bwd_reduce(y, ., ybn, mr, sums, gamma=, beta=) would compute)."""
        self._c(x, pf, y, ybn, mr, gamma, beta, sums)
        d = self._conv_dims(x, y)
        assert sums.dtype == torch.float64 and tuple(sums.shape) == (groups, d[6], 2) and ybn.shape == y.shape
        self._ck(self.lib.sg_conv_fprop_bstats(_ptr(x), _ptr(pf), _ptr(y), _ptr(ybn), _ptr(mr), _ptr(gamma), _ptr(beta), _ptr(sums),
                                               groups, act, *d, k, s, p, self._dt_of(x), self._st()))
        return True

    def conv_dgrad_bstats(self, dy, pd, dx, ybn, mr, gamma, beta, sums, groups, act, k, s, p):
        self._c(dy, pd, dx, ybn, mr, gamma, beta, sums)
        d = self._conv_dims(dx, dy)
        assert sums.dtype == torch.float64 and tuple(sums.shape) == (groups, d[3], 2) and ybn.shape == dx.shape
        self._ck(self.lib.sg_conv_dgrad_bstats(_ptr(dy), _ptr(pd), _ptr(dx), _ptr(ybn), _ptr(mr), _ptr(gamma), _ptr(beta), _ptr(sums),
                                               groups, act, *d, k, s, p, self._dt_of(dy), self._st()))
        return True

    def conv_dgrad_masked_supported(self, dy, dx, k, s, p, groups):
        d = self._conv_dims(dx, dy)
        return dy.dtype == torch.bfloat16 and bool(self.lib.sg_conv_dgrad_tc_bstats_masked_supported(*d, k, s, p, groups))

    def conv_dgrad_masked(self, dy, pd, dx, ybn, mr, gamma, beta, sums, groups, act, k, s, p, zeroed=False):
        """dx = conv_dgrad(dy) * act'(gamma * xhat(ybn) + beta) -- the data-gradient conv and the activation backward of the layer
        below in one kernel; sums[groups][C][0] = column sums of dx (with the identity table and ybn = that layer's stored
        activation: its bias gradient)."""
        self._c(dy, pd, dx, ybn, mr, gamma, beta, sums)
        d = self._conv_dims(dx, dy)
        assert sums.dtype == torch.float64 and tuple(sums.shape) == (groups, d[3], 2) and ybn.shape == dx.shape
        self._ck(self.lib.sg_conv_dgrad_tc_bstats_masked(_ptr(dy), _ptr(pd), _ptr(dx), _ptr(ybn), _ptr(mr), _ptr(gamma), _ptr(beta),
                                                         _ptr(sums), groups, act, *d, k, s, p, 1 if zeroed else 0, self._st()))

    def conv_bstats_opt(self, direction, src, pw, dst, ybn, mr, gamma, beta, sums, groups, act, k, s, p):
        """The engines' entry: ``direction`` 'f' / 'd'.  Shapes whose statistics come out of the tcgen05 epilogue run
        conv_{fprop,dgrad}_bstats and return True (``sums`` written); for the others only the conv runs and the return is
        False -- the caller's ``bn_bwd`` then reduces + applies in one launch instead of conv, reduce, apply."""
        d = self._conv_dims(src, dst) if direction == "f" else self._conv_dims(dst, src)
        if self.lib.sg_conv_bstats_in_epilogue(0 if direction == "f" else 1, *d, k, s, p, groups, self._dt_of(src)):
            (self.conv_fprop_bstats if direction == "f" else self.conv_dgrad_bstats)(src, pw, dst, ybn, mr, gamma, beta, sums,
                                                                                    groups, act, k, s, p)
            return True
        (self.conv_fprop if direction == "f" else self.conv_dgrad)(src, pw, None, dst, k, s, p)
        return False

    def conv_wgrad(self, x, dy, dw, k, s, p, impl=""):
        self._c(x, dy, dw)
        d = self._conv_dims(x, dy)
        assert dw.dtype == torch.float32 and tuple(dw.shape) == (d[6], d[3], k, k), (dw.shape, d)
        fn = getattr(self.lib, "sg_conv_wgrad" + impl)
        self._ck(fn(_ptr(x), _ptr(dy), _ptr(dw), *d, k, s, p, self._dt_of(x), self._st()))

    def conv_wgrad_cl_supported(self, x, dy, k, s, p):
        d = self._conv_dims(x, dy)
        return bool(self.lib.sg_conv_wgrad_cl_supported(*d, k, s, p, self._dt_of(x)))

    def conv_wgrad_cl(self, x, dy, gw, k, s, p):
        """gw[Co][k][k][Ci] (fp32, channels-last accumulation buffer) += weight gradient."""
        self._c(x, dy, gw)
        d = self._conv_dims(x, dy)
        assert gw.dtype == torch.float32 and tuple(gw.shape) == (d[6], k, k, d[3]), (gw.shape, d)
        self._ck(self.lib.sg_conv_wgrad_cl(_ptr(x), _ptr(dy), _ptr(gw), *d, k, s, p, self._dt_of(x), self._st()))

    def fold_grad_cl(self, gw, dw):
        """dw[Co][Ci][k][k] += gw[Co][k][k][Ci]; gw = 0."""
        self._c(gw, dw)
        Co, k, _, Ci = gw.shape
        assert tuple(dw.shape) == (Co, Ci, k, k)
        self._ck(self.lib.sg_fold_grad_cl(_ptr(gw), _ptr(dw), Co, Ci, k * k, self._st()))

    def colsum(self, x, out):
        self._c(x, out)
        C = x.shape[-1]
        self._ck(self.lib.sg_colsum(_ptr(x), _ptr(out), x.numel() // C, C, self._dt_of(x), self._st()))

    # ---- batch norm
    def col_stats(self, y, stats, groups):
        self._c(y, stats)
        C = y.shape[-1]
        assert stats.dtype == torch.float64
        self._ck(self.lib.sg_col_stats(_ptr(y), _ptr(stats), y.numel() // (C * groups), C, groups, self._dt_of(y), self._st()))

    def bn_finalize(self, stats, count, mr, running_mean, running_var, nbt, dup_first, update_running=True,
                    momentum=0.1, eps=1e-5):
        self._c(stats, mr, running_mean, running_var, nbt)
        G, C, _ = stats.shape
        self._ck(self.lib.sg_bn_finalize(_ptr(stats), int(count), _ptr(mr), _ptr(running_mean), _ptr(running_var),
                                         _ptr(nbt), dup_first, int(update_running), momentum, eps, G, C, self._st()))

    def bn_finalize_act(self, stats, count, mr, running_mean, running_var, nbt, dup_first, y, gamma, beta, out, act,
                        residual=None, update_running=True, momentum=0.1, eps=1e-5):
        """``bn_finalize`` + ``bn_act`` in one launch (training-mode BatchNorm forward from the conv epilogue's sums)."""
        self._c(stats, mr, running_mean, running_var, nbt, y, gamma, beta, out, residual)
        G, C, _ = stats.shape
        self._ck(self.lib.sg_bn_finalize_act(_ptr(stats), int(count), _ptr(mr), _ptr(running_mean), _ptr(running_var),
                                             _ptr(nbt), dup_first, int(update_running), momentum, eps, _ptr(y), _ptr(gamma),
                                             _ptr(beta), _ptr(residual), _ptr(out), y.numel() // (C * G), C, G, act,
                                             self._dt_of(y), self._st()))

    def bn_eval_mr(self, running_mean, running_var, mr, eps=1e-5):
        self._c(running_mean, running_var, mr)
        self._ck(self.lib.sg_bn_eval_mr(_ptr(running_mean), _ptr(running_var), _ptr(mr), eps, running_mean.numel(), self._st()))

    def bn_act(self, y, mr, gamma, beta, out, groups, act, residual=None):
        self._c(y, mr, gamma, beta, out, residual)
        C = y.shape[-1]
        self._ck(self.lib.sg_bn_act(_ptr(y), _ptr(mr), _ptr(gamma), _ptr(beta), _ptr(residual), _ptr(out),
                                    y.numel() // (C * groups), C, groups, act, self._dt_of(y), self._st()))

    def bn_bwd_reduce(self, da, a_out, y, mr, sums, groups, act, gamma=None, beta=None):
        """``gamma``/``beta`` given (BN directly followed by the activation): the activation tensor is not streamed, its
        sign is recomputed from y."""
        self._c(da, a_out, y, mr, sums, gamma, beta)
        C = y.shape[-1]
        if gamma is not None and C % 8 == 0 and act != ACT_TANH:
            self._ck(self.lib.sg_bn_bwd_reduce_y(_ptr(da), _ptr(y), _ptr(mr), _ptr(gamma), _ptr(beta), _ptr(sums),
                                                 y.numel() // (C * groups), C, groups, act, self._dt_of(y), self._st()))
            return
        self._ck(self.lib.sg_bn_bwd_reduce(_ptr(da), _ptr(a_out), _ptr(y), _ptr(mr), _ptr(sums),
                                           y.numel() // (C * groups), C, groups, act, self._dt_of(y), self._st()))

    def bn_bwd_apply(self, da, a_out, y, mr, gamma, sums, dy, groups, act, inject=None, inject_group=0, beta=None):
        self._c(da, a_out, y, mr, gamma, sums, dy, inject, beta)
        C = y.shape[-1]
        if beta is not None and C % 8 == 0 and act != ACT_TANH:
            self._ck(self.lib.sg_bn_bwd_apply_y(_ptr(da), _ptr(y), _ptr(mr), _ptr(gamma), _ptr(beta), _ptr(sums),
                                                _ptr(inject), inject_group, _ptr(dy), y.numel() // (C * groups), C, groups,
                                                act, self._dt_of(y), self._st()))
            return
        self._ck(self.lib.sg_bn_bwd_apply(_ptr(da), _ptr(a_out), _ptr(y), _ptr(mr), _ptr(gamma), _ptr(sums),
                                          _ptr(inject), inject_group, _ptr(dy), y.numel() // (C * groups), C, groups,
                                          act, self._dt_of(y), self._st()))

    def bn_bwd(self, da, a_out, y, mr, gamma, sums, dy, groups, act, inject=None, inject_group=0, beta=None, zeroed=False):
        """bn_bwd_reduce + bn_bwd_apply as one call (``beta`` given: the activation's sign is recomputed from y).  Tensors that
        fit the SMs' shared memory run as ONE launch (sg_bn_bwd); each call site -- identified by its ``sums`` buffer -- owns
        1 KB of work words in a pool allocated up front (nothing is allocated under graph capture).  ``zeroed``: the caller
        zeroed ``sums`` with ``zero_multi`` at the start of its pass -- no memset node in front of the reduction."""
        C = y.shape[-1]
        if C % 8 != 0 or act == ACT_TANH:        # (these zero ``sums`` themselves, whatever ``zeroed`` says)
            self.bn_bwd_reduce(da, a_out, y, mr, sums, groups, act, gamma=gamma if beta is not None else None, beta=beta)
            self.bn_bwd_apply(da, a_out, y, mr, gamma, sums, dy, groups, act, inject=inject, inject_group=inject_group, beta=beta)
            return
        self._c(da, a_out, y, mr, gamma, sums, dy, inject, beta)
        work = self._work_slot(sums.data_ptr())
        self._ck(self.lib.sg_bn_bwd(_ptr(da), None if beta is not None else _ptr(a_out), _ptr(y), _ptr(mr), _ptr(gamma),
                                    _ptr(beta), _ptr(sums), _ptr(inject), inject_group, _ptr(dy), y.numel() // (C * groups), C,
                                    groups, act, self._dt_of(y), 1 if zeroed else 0, work, self._st()))

    def bn_param_grad(self, sums, dgamma, dbeta):
        self._c(sums, dgamma, dbeta)
        G, C, _ = sums.shape
        self._ck(self.lib.sg_bn_param_grad(_ptr(sums), _ptr(dgamma), _ptr(dbeta), G, C, self._st()))

    def bn_param_grad_multi(self, items):
        """``items``: [(sums [G,C,2] fp64, dgamma [C], dbeta [C]), ...] -- every BatchNorm layer of a backward pass in one launch."""
        for i in range(0, len(items), 24):
            part = items[i:i + 24]
            n = len(part)
            for t in part:
                self._c(*t)
            P = ctypes.c_void_p * n
            I = ctypes.c_int * n
            self._ck(self.lib.sg_bn_param_grad_multi(P(*[_ptr(t[0]) for t in part]), P(*[_ptr(t[1]) for t in part]),
                                                     P(*[_ptr(t[2]) for t in part]), I(*[t[0].shape[0] for t in part]),
                                                     I(*[t[0].shape[1] for t in part]), n, self._st()))

    def act_bwd(self, da, a_out, out, act, colsum=None):
        """out = da * act'(a_out); ``colsum`` (fp32 [C]) += the column sums of out (the layer's bias gradient) in the same pass."""
        self._c(da, a_out, out, colsum)
        if colsum is not None:
            C = da.shape[-1]
            self._ck(self.lib.sg_act_bwd_colsum(_ptr(da), _ptr(a_out), _ptr(out), _ptr(colsum), da.numel() // C, C, act,
                                                self._dt_of(da), self._st()))
            return
        self._ck(self.lib.sg_act_bwd(_ptr(da), _ptr(a_out), _ptr(out), da.numel(), act, self._dt_of(da), self._st()))

    def gp_bn_reduce(self, v, da, a_out, y, mr, tsums, act, zeroed=False):
        self._c(v, da, a_out, y, mr, tsums)
        C = y.shape[-1]
        fn = self.lib.sg_gp_bn_reduce_acc if zeroed else self.lib.sg_gp_bn_reduce
        self._ck(fn(_ptr(v), _ptr(da), _ptr(a_out), _ptr(y), _ptr(mr), _ptr(tsums),
                                          y.numel() // C, C, act, self._dt_of(y), self._st()))

    def _work_slot(self, key):
        """1 KB of work words for the call site ``key`` (the address of its sums buffer); None once the pool is used up --
        the library then runs the two-kernel path for that site."""
        slot = self._bn_slot.get(key)
        if slot is None:
            if len(self._bn_slot) >= 1024:
                return None
            slot = self._bn_slot[key] = len(self._bn_slot)
        return self._bn_work.data_ptr() + 1024 * slot

    def gp_bn(self, v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, act, zeroed=False):
        """gp_bn_reduce + gp_bn_apply as one call (one launch where the four tensors fit the SMs' shared memory)."""
        self._c(v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma)
        C = y.shape[-1]
        self._ck(self.lib.sg_gp_bn(_ptr(v), _ptr(da), _ptr(a_out), _ptr(y), _ptr(mr), _ptr(gamma), _ptr(sums), _ptr(tsums),
                                   _ptr(w_out), _ptr(gy_out), _ptr(dgamma), y.numel() // C, C, act, self._dt_of(y),
                                   1 if zeroed else 0, self._work_slot(tsums.data_ptr()), self._st()))

    def gp_bn_apply(self, v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, act):
        self._c(v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma)
        C = y.shape[-1]
        self._ck(self.lib.sg_gp_bn_apply(_ptr(v), _ptr(da), _ptr(a_out), _ptr(y), _ptr(mr), _ptr(gamma), _ptr(sums),
                                         _ptr(tsums), _ptr(w_out), _ptr(gy_out), _ptr(dgamma), y.numel() // C, C, act,
                                         self._dt_of(y), self._st()))

    # ---- dense
    def linear_fwd(self, x, w, b, out, relu=False):
        self._c(x, w, b, out)
        N, K = x.shape
        M = w.shape[0]
        self._ck(self.lib.sg_linear_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(out), N, K, M, int(relu), self._st()))

    def linear_bwd(self, x, w, dout, dw, db, dx, dx_acc=False, relu_out=None):
        self._c(x, w, dout, dw, db, dx, relu_out)
        N, K = x.shape
        M = w.shape[0]
        self._ck(self.lib.sg_linear_bwd(_ptr(x), _ptr(w), _ptr(dout), _ptr(relu_out), _ptr(dw), _ptr(db), _ptr(dx),
                                        int(dx_acc), N, K, M, self._st()))

    # ---- critic head
    def head_prepare(self, wcr, bcr, wcs, bcs, A, Bv, c0):
        self._c(wcr, bcr, wcs, bcs, A, Bv, c0)
        K = wcr.shape[0]
        Cx, Nd = A.shape[1], Bv.shape[0]
        self._ck(self.lib.sg_head_prepare(_ptr(wcr), _ptr(bcr), _ptr(wcs), _ptr(bcs), _ptr(A), _ptr(Bv), _ptr(c0),
                                          K, Cx, Nd, self._st()))

    def head_fwd(self, a4, ce, A, Bv, c0, score):
        self._c(a4, ce, A, Bv, c0, score)
        N = a4.shape[0]
        self._ck(self.lib.sg_head_fwd(_ptr(a4), _ptr(ce), _ptr(A), _ptr(Bv), _ptr(c0), _ptr(score), N,
                                      a4.numel() // N, Bv.shape[0], self._dt_of(a4), self._st()))

    def head_fwd_multi(self, a4, ce, A, Bv, c0, score, jobs, N):
        """``jobs``: up to four ``(first row of a4, first row of ce, first element of score)``; each scores N rows."""
        self._c(a4, ce, A, Bv, c0, score)
        n = len(jobs)
        arr = lambda k: (_c.c_int * n)(*[int(j[k]) for j in jobs])
        a0, c0r, s0 = arr(0), arr(1), arr(2)
        self._ck(self.lib.sg_head_fwd_multi(_ptr(a4), _ptr(ce), _ptr(A), _ptr(Bv), _ptr(c0), _ptr(score), n, a0, c0r, s0, N,
                                            a4.numel() // a4.shape[0], Bv.shape[0], self._dt_of(a4), self._st()))

    def head_bwd_data(self, coef, A, da4):
        self._c(coef, A, da4)
        N = coef.shape[0]
        self._ck(self.lib.sg_outer(_ptr(coef), _ptr(A), _ptr(da4), N, A.numel(), self._dt_of(da4), self._st()))

    def head_bwd_reduce(self, coef, a4, dA):
        self._c(coef, a4, dA)
        N = coef.shape[0]
        assert a4.numel() == N * dA.numel()
        self._ck(self.lib.sg_wsum_rows(_ptr(coef), _ptr(a4), _ptr(dA), N, dA.numel(), self._dt_of(a4), self._st()))

    def head_param_grads(self, dA, dBv, dc0, wcr, bcr, wcs, dwcr, dbcr, dwcs, dbcs):
        self._c(dA, dBv, dc0, wcr, bcr, wcs, dwcr, dbcr, dwcs, dbcs)
        K = wcr.shape[0]
        self._ck(self.lib.sg_head_param_grads(_ptr(dA), _ptr(dBv), _ptr(dc0), _ptr(wcr), _ptr(bcr), _ptr(wcs),
                                              _ptr(dwcr), _ptr(dbcr), _ptr(dwcs), _ptr(dbcs), K, dA.shape[1],
                                              dBv.shape[0], self._st()))

    # ---- conditioning augmentation
    def ca_reparam(self, mu, sigma, eps, z, c_hat, cg):
        """c_hat = mu + sigma*eps; cg rows = [c_hat, z, zero padding up to cg's row length]."""
        self._c(mu, sigma, eps, z, c_hat, cg)
        N, C = mu.shape
        ld = cg.numel() // N if cg is not None else C
        nz = z.shape[1] if (z is not None and cg is not None) else 0
        self._ck(self.lib.sg_ca_reparam(_ptr(mu), _ptr(sigma), _ptr(eps), _ptr(z), _ptr(c_hat), _ptr(cg), N, C, nz, ld,
                                        self._dt_of(cg) if cg is not None else SG_F32, self._st()))

    def ca_bwd_seed(self, dcg, eps, mu, sigma, kl_scale, dmu, dsigma):
        self._c(dcg, eps, mu, sigma, dmu, dsigma)
        N, C = mu.shape
        ld = dcg.numel() // N if dcg is not None else 0
        self._ck(self.lib.sg_ca_bwd_seed(_ptr(dcg), _ptr(eps), _ptr(mu), _ptr(sigma), float(kl_scale), _ptr(dmu),
                                         _ptr(dsigma), N, C, ld, self._dt_of(dcg) if dcg is not None else SG_F32,
                                         self._st()))

    def ca_forward(self, tem, Wh, bh, Wmu, bmu, Wsg, bsg, eps, z, h, mu, sigma, c_hat, cg):
        """The whole conditioning-augmentation forward in one launch (con_augment.py:13-22 + the [c_hat, z] row)."""
        self._c(tem, Wh, bh, Wmu, bmu, Wsg, bsg, eps, z, h, mu, sigma, c_hat, cg)
        N, Tm = tem.shape
        Hd, C = Wh.shape[0], Wmu.shape[0]
        ld = cg.numel() // N if cg is not None else C
        nz = z.shape[1] if (z is not None and cg is not None) else 0
        self._ck(self.lib.sg_ca_forward(_ptr(tem), _ptr(Wh), _ptr(bh), _ptr(Wmu), _ptr(bmu), _ptr(Wsg), _ptr(bsg), _ptr(eps),
                                        _ptr(z if cg is not None else None), _ptr(h), _ptr(mu), _ptr(sigma), _ptr(c_hat), _ptr(cg),
                                        N, Tm, Hd, C, nz, ld, self._dt_of(cg) if cg is not None else SG_F32, self._st()))

    def ca_backward(self, dcg, eps, mu, sigma, kl_scale, h, tem, Wmu, Wsg, Wh, dmu, dsigma, dh, gWmu, gbmu, gWsg, gbsg, gWh,
                    gbh, dtem, dtem_acc):
        """Backward of ca_forward in two launches; parameter gradients accumulate, dtem (may be None) (+)=."""
        self._c(dcg, eps, mu, sigma, h, tem, Wmu, Wsg, Wh, dmu, dsigma, dh, gWmu, gbmu, gWsg, gbsg, gWh, gbh, dtem)
        N, Tm = tem.shape
        Hd, C = Wh.shape[0], Wmu.shape[0]
        ld = dcg.numel() // N if dcg is not None else 0
        self._ck(self.lib.sg_ca_backward(_ptr(dcg), _ptr(eps), _ptr(mu), _ptr(sigma), float(kl_scale), _ptr(h), _ptr(tem),
                                         _ptr(Wmu), _ptr(Wsg), _ptr(Wh), _ptr(dmu), _ptr(dsigma), _ptr(dh), _ptr(gWmu),
                                         _ptr(gbmu), _ptr(gWsg), _ptr(gbsg), _ptr(gWh), _ptr(gbh), _ptr(dtem),
                                         1 if dtem_acc else 0, N, Tm, Hd, C, ld,
                                         self._dt_of(dcg) if dcg is not None else SG_F32, self._st()))

    # ---- losses
    def interp(self, real, fake, eps, out):
        self._c(real, fake, eps, out)
        N = real.shape[0]
        self._ck(self.lib.sg_interp(_ptr(real), _ptr(fake), _ptr(eps), _ptr(out), N, real.numel() // N,
                                    self._dt_of(real), self._st()))

    def sample_sqnorm(self, g, out, zeroed=False):
        self._c(g, out)
        N = g.shape[0]
        fn = self.lib.sg_sample_sqnorm_acc if zeroed else self.lib.sg_sample_sqnorm
        self._ck(fn(_ptr(g), _ptr(out), N, g.numel() // N, self._dt_of(g), self._st()))

    def gp_seed(self, g, sq, coef, v):
        self._c(g, sq, v)
        N = g.shape[0]
        self._ck(self.lib.sg_gp_seed(_ptr(g), _ptr(sq), float(coef), _ptr(v), N, g.numel() // N, self._dt_of(g), self._st()))

    def critic_loss(self, s_real, s_mis, s_fake, sq, lam, out):
        self._c(s_real, s_mis, s_fake, sq, out)
        self._ck(self.lib.sg_critic_loss(_ptr(s_real), _ptr(s_mis), _ptr(s_fake), _ptr(sq), float(lam), _ptr(out),
                                         s_real.numel(), self._st()))

    def gen_loss(self, s_fake, mu, sigma, out):
        self._c(s_fake, mu, sigma, out)
        N, C = mu.shape
        self._ck(self.lib.sg_gen_loss(_ptr(s_fake), _ptr(mu), _ptr(sigma), _ptr(out), N, C, self._st()))

    def scale_rows_add(self, x, scale, out, accumulate):
        self._c(x, scale, out)
        N = x.shape[0]
        self._ck(self.lib.sg_scale_rows_add(_ptr(x), _ptr(scale), _ptr(out), int(accumulate), N, x.numel() // N,
                                            self._dt_of(x), self._st()))

    # ---- optimiser
    def dp_adam_step(self, grad_ptrs, param_ptrs, flag_ptrs, m, v, hyper, sync, n, rank, world, slot, write_avg=False):
        """Data-parallel optimizer step over peer memory (csrc/dp_adam.cu).  ``*_ptrs``: ctypes arrays of ``world`` device
        pointers (comm.PeerComm keeps them alive)."""
        self._c(m, v, hyper, sync)
        self._ck(self.lib.sg_dp_adam_step(grad_ptrs, param_ptrs, flag_ptrs, _ptr(m), _ptr(v), _ptr(hyper), _ptr(sync), int(n),
                                          rank, world, slot, int(write_avg), self._st()))

    def adam_step(self, p, g, m, v, hyper):
        self._c(p, g, m, v, hyper)
        self._ck(self.lib.sg_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(hyper), p.numel(), self._st()))
