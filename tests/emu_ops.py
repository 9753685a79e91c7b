"""Torch-CPU statement of what every C-ABI kernel computes -- TEST DOUBLE, lives in tests/ only.

Same method names/arguments as ``imagegenerator_b200.ops.CudaOps``.  Two uses:
  * host-logic tests: run the engine (imagegenerator_b200/engine.py) on CPU in fp64 and
    compare a whole train step with the oracle (no GPU needed);
  * kernel tests (-m gpu): each CUDA kernel is compared against the method of the
    same name here on random inputs.
The product package never imports this file.
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
SLOPE = 0.1


def _act(x, act):
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_LRELU:
        return F.leaky_relu(x, SLOPE)
    if act == ACT_TANH:
        return torch.tanh(x)
    return x


def _mask(a_out, act):
    """d act / d pre-activation, from the stored activation OUTPUT."""
    if act == ACT_RELU:
        return (a_out > 0).to(a_out.dtype)
    if act == ACT_LRELU:
        return torch.where(a_out > 0, torch.ones_like(a_out), torch.full_like(a_out, SLOPE))
    if act == ACT_TANH:
        return 1 - a_out * a_out
    return torch.ones_like(a_out)


def nchw(x):
    return x.permute(0, 3, 1, 2)


def nhwc(x):
    return x.permute(0, 2, 3, 1)


class EmuOps:
    is_emulator = True

    def __init__(self, dtype=torch.float64, device="cpu"):
        self.act_dtype = dtype          # storage type of activations ("T")
        self.device = torch.device(device)
        self.launches = 0

    # ---- allocation helpers the engine uses
    def empty(self, shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.act_dtype, device=self.device)

    def zeros(self, shape, dtype=None):
        return torch.zeros(shape, dtype=dtype or self.act_dtype, device=self.device)

    @property
    def f32(self):
        # "fp32" side tensors (params, grads, stats outputs) follow the emulation precision
        return torch.float64 if self.act_dtype == torch.float64 else torch.float32

    @property
    def f64(self):
        return torch.float64

    # ---- layout
    def nchw_to_nhwc(self, src, dst):
        dst.copy_(nhwc(src).to(dst.dtype))

    def nhwc_to_nchw(self, src, dst):
        dst.copy_(nchw(src).to(dst.dtype))

    def nhwc_to_nchw_u8(self, src, dst):
        dst.copy_(torch.clamp(torch.round((nchw(src).to(torch.float32) + 1.0) * 127.5), 0, 255).to(torch.uint8))

    def pack_weight(self, w, pf, pd):
        """w [Co,Ci,kh,kw] fp32 -> pf [Co,kh,kw,Ci], pd [Ci,kh,kw,Co] in T."""
        if pf is not None:
            pf.copy_(w.permute(0, 2, 3, 1).to(pf.dtype))
        if pd is not None:
            pd.copy_(w.permute(1, 2, 3, 0).to(pd.dtype))

    def pack_gemm_t(self, w, wt):
        """wt[(t, ci)][Kp] = w[co][ci][t], columns co >= Co zero."""
        Co, Ci, k, _ = w.shape
        Kp = wt.shape[-1]
        out = torch.zeros(k * k * Ci, Kp, dtype=wt.dtype, device=wt.device)
        out[:, :Co] = w.reshape(Co, Ci, k * k).permute(2, 1, 0).reshape(k * k * Ci, Co).to(wt.dtype)
        wt.copy_(out.reshape(wt.shape))

    def bn_fold(self, running_mean, running_var, gamma, beta, scale, shift, eps=1e-5):
        sc = gamma.double() / torch.sqrt(running_var.double() + eps)
        scale.copy_(sc.to(scale.dtype))
        shift.copy_((beta.double() - running_mean.double() * sc).to(shift.dtype))

    def pack_weight_scaled(self, w, scale, axis, pf, pd):
        shp = (-1, 1, 1, 1) if axis == 0 else (1, -1, 1, 1)
        self.pack_weight(w * scale.to(w.dtype).view(shp), pf, pd)

    def conv_fprop_res(self, x, pf, bias, residual, y, k, s, p, act=ACT_NONE):
        w = pf.permute(0, 3, 1, 2).double()
        out = F.conv2d(nchw(x).double(), w, None if bias is None else bias.double(), s, p) + nchw(residual).double()
        y.copy_(nhwc(_act(out, act)).to(y.dtype))

    def add_act(self, a, b, out, act):
        out.copy_(_act(a.double() + b.double(), act).to(out.dtype))

    def patchify(self, x, P, k, s, p):
        """P[n,oh,ow, ci*k*k + kh*k + kw] = x[n, oh*s-p+kh, ow*s-p+kw, ci]  (F.unfold order)."""
        N, Ho, Wo, K = P.shape
        u = F.unfold(nchw(x).double(), k, padding=p, stride=s)          # [N, C*k*k, Ho*Wo]
        P.copy_(u.transpose(1, 2).reshape(N, Ho, Wo, K).to(P.dtype))

    def unpatchify(self, col, bias, out, k, s, p, act=ACT_NONE):
        """col2im (F.fold) of col[n,ih,iw, c*k*k + kh*k + kw] onto out[n,oh,ow,c], + bias, activation."""
        N, Hi, Wi, K = col.shape
        _, Ho, Wo, C = out.shape
        u = col.reshape(N, Hi * Wi, K).transpose(1, 2).double()
        o = F.fold(u, (Ho, Wo), k, padding=p, stride=s)                  # [N, C, Ho, Wo]
        if bias is not None:
            o = o + bias.double().view(1, C, 1, 1)
        out.copy_(nhwc(_act(o, act)).to(out.dtype))

    # ---- convolutions (Conv2d-layout semantics; ConvTranspose2d layers use them mirrored)
    def conv_fprop(self, x, pf, bias, y, k, s, p, act=ACT_NONE, impl=""):
        # arithmetic is always fp64; only STORAGE follows the emulated mode (ideal-rounding model)
        w = pf.permute(0, 3, 1, 2).double()                         # [Co,Ci,kh,kw]
        out = F.conv2d(nchw(x).double(), w, None if bias is None else bias.double(), s, p)
        y.copy_(nhwc(_act(out, act)).to(y.dtype))

    def conv_fprop_f32out(self, x, pf, y, k, s, p):
        self.conv_fprop(x, pf, None, y, k, s, p)

    def conv_dgrad(self, dy, pd, bias, dx, k, s, p, act=ACT_NONE, impl=""):
        w = pd.permute(3, 0, 1, 2).double()                         # [Co,Ci,kh,kw] (= convT weight [in,out,kh,kw])
        out = F.conv_transpose2d(nchw(dy).double(), w, None if bias is None else bias.double(), s, p)
        assert out.shape[2] == dx.shape[1], (out.shape, dx.shape)
        dx.copy_(nhwc(_act(out, act)).to(dx.dtype))

    def conv_narrow_fprop(self, x, pf, bias, y, act=ACT_NONE, stats=None, groups=1):
        self.conv_fprop(x, pf, bias, y, 4, 2, 1, act=act)
        if stats is not None:
            self.col_stats(y, stats, groups)

    def conv_narrow_dgrad(self, dy, pd, bias, dx, act=ACT_NONE):
        self.conv_dgrad(dy, pd, bias, dx, 4, 2, 1, act=act)

    def conv_fprop_stats(self, x, pf, y, stats, groups, k, s, p):
        """conv + (sum, sum^2) of the STORED result per image group (the statistics of the BN that follows)."""
        self.conv_fprop(x, pf, None, y, k, s, p)
        self.col_stats(y, stats, groups)

    def conv_dgrad_stats(self, dy, pd, dx, stats, groups, k, s, p):
        self.conv_dgrad(dy, pd, None, dx, k, s, p)
        self.col_stats(dx, stats, groups)

    def conv_fprop_bstats(self, x, pf, y, ybn, mr, gamma, beta, sums, groups, act, k, s, p):
        """conv, then the BatchNorm-backward statistics of the layer below from the STORED result."""
        self.conv_fprop(x, pf, None, y, k, s, p)
        self.bn_bwd_reduce(y, None, ybn, mr, sums, groups, act, gamma=gamma, beta=beta)

    def conv_dgrad_bstats(self, dy, pd, dx, ybn, mr, gamma, beta, sums, groups, act, k, s, p):
        self.conv_dgrad(dy, pd, None, dx, k, s, p)
        self.bn_bwd_reduce(dx, None, ybn, mr, sums, groups, act, gamma=gamma, beta=beta)

    def conv_dgrad_masked_supported(self, dy, dx, k, s, p, groups):
        return True

    def conv_dgrad_masked(self, dy, pd, dx, ybn, mr, gamma, beta, sums, groups, act, k, s, p, zeroed=False):
        """dx = conv_dgrad(dy) * act'(gamma * xhat(ybn) + beta); sums (+)= (sum dx, sum dx * xhat) per (group, channel) of the stored dx."""
        da = torch.zeros_like(dx)
        self.conv_dgrad(dy, pd, None, da, k, s, p)
        C = dx.shape[-1]
        sign = self._act_sign_from_y(ybn, mr, gamma, beta, groups)
        dx.copy_((da.to(torch.float64) * _mask(sign, act).to(torch.float64)).to(dx.dtype))
        dz = dx.to(torch.float64).reshape(groups, -1, C)
        xh = self._xhat(ybn, mr, groups).to(torch.float64)
        part = torch.stack([dz.sum(1), (dz * xh).sum(1)], dim=-1)
        if zeroed:
            sums.add_(part)
        else:
            sums.copy_(part)

    def conv_bstats_opt(self, direction, src, pw, dst, ybn, mr, gamma, beta, sums, groups, act, k, s, p):
        """CudaOps.conv_bstats_opt: True = the statistics were reduced with the conv (always, here)."""
        (self.conv_fprop_bstats if direction == "f" else self.conv_dgrad_bstats)(src, pw, dst, ybn, mr, gamma, beta, sums,
                                                                                groups, act, k, s, p)
        return True

    def conv_wgrad(self, x, dy, dw, k, s, p, impl=""):
        """dw[Co,Ci,kh,kw] (fp32) += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh*s-p+kh,ow*s-p+kw,ci]."""
        g = torch.nn.grad.conv2d_weight(nchw(x).double(), dw.shape, nchw(dy).double(), stride=s, padding=p)
        dw.add_(g.to(dw.dtype))

    def conv_wgrad_cl_supported(self, x, dy, k, s, p):
        return True

    def conv_wgrad_cl(self, x, dy, gw, k, s, p):
        """channels-last accumulation buffer gw[Co,kh,kw,Ci] += weight gradient."""
        Co, _, _, Ci = gw.shape
        g = torch.nn.grad.conv2d_weight(nchw(x).double(), (Co, Ci, k, k), nchw(dy).double(), stride=s, padding=p)
        gw.add_(g.permute(0, 2, 3, 1).to(gw.dtype))

    def fold_grad_cl(self, gw, dw):
        dw.add_(gw.permute(0, 3, 1, 2))
        gw.zero_()

    def colsum(self, x, out):
        """out[C] (fp32) += sum over all leading dims of x[..., C]."""
        out.add_(x.reshape(-1, x.shape[-1]).to(out.dtype).sum(0))

    # ---- batch norm
    def col_stats(self, y, stats, groups):
        """stats[G,C,2] (f64) += (sum, sum of squares) per group; rows split evenly into groups."""
        C = y.shape[-1]
        v = y.reshape(groups, -1, C).to(torch.float64)
        stats[:, :, 0] += v.sum(1)
        stats[:, :, 1] += (v * v).sum(1)

    def bn_finalize(self, stats, count, mr, running_mean, running_var, nbt, dup_first, update_running=True,
                    momentum=0.1, eps=1e-5):
        """mr[G,C,2] = (mean, rstd) from biased variance; running stats EMA-updated once per group
        in group order, group 0 ``dup_first`` times (real + mismatched share statistics)."""
        mean = stats[:, :, 0] / count
        var = (stats[:, :, 1] / count - mean * mean).clamp_min(0)
        mr[:, :, 0] = mean.to(mr.dtype)
        mr[:, :, 1] = (1.0 / torch.sqrt(var + eps)).to(mr.dtype)
        if update_running:
            G = stats.shape[0]
            order = [0] * dup_first + list(range(1, G))
            for g in order:
                running_mean.mul_(1 - momentum).add_(momentum * mean[g].to(running_mean.dtype))
                running_var.mul_(1 - momentum).add_(momentum * (var[g] * count / max(count - 1, 1)).to(running_var.dtype))
            nbt += len(order)

    def bn_finalize_act(self, stats, count, mr, running_mean, running_var, nbt, dup_first, y, gamma, beta, out, act,
                        residual=None, update_running=True, momentum=0.1, eps=1e-5):
        self.bn_finalize(stats, count, mr, running_mean, running_var, nbt, dup_first, update_running, momentum, eps)
        self.bn_act(y, mr, gamma, beta, out, stats.shape[0], act, residual)

    def bn_eval_mr(self, running_mean, running_var, mr, eps=1e-5):
        mr[0, :, 0] = running_mean.to(mr.dtype)
        mr[0, :, 1] = (1.0 / torch.sqrt(running_var.to(torch.float64) + eps)).to(mr.dtype)

    def _xhat(self, y, mr, groups):
        C = y.shape[-1]
        v = y.reshape(groups, -1, C).double()
        mr = mr.double()
        return (v - mr[:, None, :, 0]) * mr[:, None, :, 1]

    def bn_act(self, y, mr, gamma, beta, out, groups, act, residual=None):
        xh = self._xhat(y, mr, groups)
        z = xh * gamma.double() + beta.double()
        z = z.reshape(y.shape)
        if residual is not None:
            z = z + residual.double()
        out.copy_(_act(z, act).to(out.dtype))

    def _act_sign_from_y(self, y, mr, gamma, beta, groups):
        """what the *_y kernels use instead of the stored activation: gamma*xhat+beta (same sign as act(...))."""
        C = y.shape[-1]
        z = self._xhat(y, mr, groups).to(torch.float64) * gamma.double()[None, None, :] + beta.double()[None, None, :]
        return z.reshape(y.shape)

    def bn_bwd_reduce(self, da, a_out, y, mr, sums, groups, act, gamma=None, beta=None):
        """sums[G,C,2] (f64) = (S1, S2) = (sum dz, sum dz*xhat), dz = da * act'(.) from a_out -- or, when gamma/beta
        are given (BN directly followed by the activation), from the sign of gamma*xhat+beta recomputed from y."""
        C = y.shape[-1]
        if gamma is not None and C % 8 == 0 and act != ACT_TANH:
            a_out = self._act_sign_from_y(y, mr, gamma, beta, groups)
        dz = (da.to(torch.float64) * _mask(a_out, act).to(torch.float64)).reshape(groups, -1, C)
        xh = self._xhat(y, mr, groups).to(torch.float64)
        sums[:, :, 0] = dz.sum(1)
        sums[:, :, 1] = (dz * xh).sum(1)

    def bn_bwd_apply(self, da, a_out, y, mr, gamma, sums, dy, groups, act, inject=None, inject_group=0, beta=None):
        """dy = gamma*rstd/N * (N dz - S1 - xhat S2)  [+ inject on rows of group inject_group]."""
        C = y.shape[-1]
        ft = torch.float64
        if beta is not None and C % 8 == 0 and act != ACT_TANH:
            a_out = self._act_sign_from_y(y, mr, gamma, beta, groups)
        dz = (da.to(ft) * _mask(a_out, act).to(ft)).reshape(groups, -1, C)
        n = dz.shape[1]
        xh = self._xhat(y, mr, groups)
        a = gamma.to(ft)[None, None, :] * mr.to(ft)[:, None, :, 1] / n
        out = a * (n * dz - sums[:, None, :, 0].to(ft) - xh * sums[:, None, :, 1].to(ft))
        if inject is not None:
            out[inject_group] += inject.reshape(-1, C).to(ft)
        dy.copy_(out.reshape(dy.shape).to(dy.dtype))

    def bn_bwd(self, da, a_out, y, mr, gamma, sums, dy, groups, act, inject=None, inject_group=0, beta=None, zeroed=False):
        """reduce + apply in one call (CudaOps.bn_bwd: one launch when the tensor fits the SMs' shared memory).  ``zeroed``:
        the caller promises zeroed sums and the kernels ADD to them -- so does this statement (a missing zero_multi shows)."""
        part = torch.zeros_like(sums)
        self.bn_bwd_reduce(da, a_out, y, mr, part, groups, act, gamma=gamma if beta is not None else None, beta=beta)
        if zeroed:
            sums.add_(part)
        else:
            sums.copy_(part)
        self.bn_bwd_apply(da, a_out, y, mr, gamma, sums, dy, groups, act, inject=inject, inject_group=inject_group, beta=beta)

    def bn_param_grad(self, sums, dgamma, dbeta):
        dgamma.add_(sums[:, :, 1].sum(0).to(dgamma.dtype))
        dbeta.add_(sums[:, :, 0].sum(0).to(dbeta.dtype))

    def bn_param_grad_multi(self, items):
        for sums, dgamma, dbeta in items:
            self.bn_param_grad(sums, dgamma, dbeta)

    def act_bwd(self, da, a_out, out, act, colsum=None):
        out.copy_((da.to(torch.float64) * _mask(a_out, act).to(torch.float64)).to(out.dtype))
        if colsum is not None:
            self.colsum(out, colsum)

    # ---- gradient-penalty second order through a train-mode BN (one group)
    def gp_bn_reduce(self, v, da, a_out, y, mr, tsums, act, zeroed=False):
        """tsums[C,3] (f64) = (sum v, sum v*xhat, sum v*dz)."""
        C = y.shape[-1]
        vv = v.reshape(-1, C).to(torch.float64)
        dz = (da.to(torch.float64) * _mask(a_out, act).to(torch.float64)).reshape(-1, C)
        xh = self._xhat(y, mr, 1)[0].to(torch.float64)
        part = torch.stack([vv.sum(0), (vv * xh).sum(0), (vv * dz).sum(0)], dim=1).to(tsums.dtype)
        if zeroed:                                   # the caller zeroed tsums (zero_multi); the kernel adds
            tsums.add_(part)
        else:
            tsums.copy_(part)

    def gp_bn(self, v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, act, zeroed=False):
        self.gp_bn_reduce(v, da, a_out, y, mr, tsums, act, zeroed=zeroed)
        self.gp_bn_apply(v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, act)

    def gp_bn_apply(self, v, da, a_out, y, mr, gamma, sums, tsums, w_out, gy_out, dgamma, act):
        """Backward of dy = BNbwd(dz; y, gamma) given v = dL/d dy  (SURVEY section 7):
             w_out  = mask * a (N v - T1 - xhat T2)                      (dL/d da: feeds the next conv_fprop)
             gy_out = r (G - sum G/N - xhat sum(G xhat)/N) - P r xhat/N   (dL/d y: injected into the plain backward)
             dgamma += P / gamma,  P = sum v*dy = a (N T3 - S1 T1 - S2 T2),  G = -a (v S2 + dz T2)."""
        C = y.shape[-1]
        ft = torch.float64
        vv = v.reshape(-1, C).to(ft)
        m = _mask(a_out, act).reshape(-1, C).to(ft)
        dz = da.reshape(-1, C).to(ft) * m
        n = vv.shape[0]
        r = mr[0, :, 1].to(ft)
        xh = self._xhat(y, mr, 1)[0].to(ft)
        g = gamma.to(ft)
        a = g * r / n
        S1, S2 = sums[0, :, 0].to(ft), sums[0, :, 1].to(ft)
        T1, T2, T3 = tsums[:, 0].to(ft), tsums[:, 1].to(ft), tsums[:, 2].to(ft)
        u = a * (n * vv - T1 - xh * T2)
        w_out.copy_((u * m).reshape(w_out.shape).to(w_out.dtype))
        P = a * (n * T3 - S1 * T1 - S2 * T2)
        G = -a * (vv * S2 + dz * T2)
        sG = -a * (S2 * T1 + S1 * T2)
        sGx = -2 * a * S2 * T2
        gy = r * (G - sG / n - xh * sGx / n) - P * r * xh / n
        gy_out.copy_(gy.reshape(gy_out.shape).to(gy_out.dtype))
        dgamma.add_((r / n * (n * T3 - S1 * T1 - S2 * T2)).to(dgamma.dtype))

    # ---- small dense layers (fp32)
    def linear_fwd(self, x, w, b, out, relu=False):
        o = F.linear(x.double(), w.double(), None if b is None else b.double())
        out.copy_((F.relu(o) if relu else o).to(out.dtype))

    def linear_bwd(self, x, w, dout, dw, db, dx, dx_acc=False, relu_out=None):
        """dw += dout^T x ; db += sum dout ; dx (may be None) (+)= dout w.  If relu_out is given,
        dout is first masked by relu_out > 0."""
        dout, x, w = dout.double(), x.double(), w.double()
        if relu_out is not None:
            dout = dout * (relu_out > 0).to(dout.dtype)
        if dw is not None:
            dw.add_((dout.t() @ x).to(dw.dtype))
        if db is not None:
            db.add_(dout.sum(0).to(db.dtype))
        if dx is not None:
            if dx_acc:
                dx.add_((dout @ w).to(dx.dtype))
            else:
                dx.copy_((dout @ w).to(dx.dtype))

    # ---- critic head (text replicate + concat + 1x1 conv + linear collapse into A, Bv, c0)
    def head_prepare(self, wcr, bcr, wcs, bcs, A, Bv, c0):
        """score = <A, a4> + <Bv, ce> + c0 with A[hw,c] = sum_k wcs[k,hw] wcr[k,c] (c < 512),
        Bv[j] = sum_k (sum_hw wcs[k,hw]) wcr[k,512+j], c0 = sum_k bcr[k] sum_hw wcs[k,hw] + bcs."""
        K = wcr.shape[0]
        ws = wcs.reshape(K, 16).double()
        wr = wcr.reshape(K, -1).double()
        nx = A.shape[1]
        A.copy_((ws.t() @ wr[:, :nx]).to(A.dtype))
        sw = ws.sum(1)
        Bv.copy_((sw @ wr[:, nx:]).to(Bv.dtype))
        c0.copy_(((bcr.double() * sw).sum().reshape(1) + bcs.double()).to(c0.dtype))

    def head_fwd(self, a4, ce, A, Bv, c0, score):
        n = a4.shape[0]
        s = (a4.reshape(n, -1).double() * A.double().reshape(1, -1)).sum(1) + ce.double() @ Bv.double() + c0.double()
        score.copy_(s.to(score.dtype))

    def head_fwd_multi(self, a4, ce, A, Bv, c0, score, jobs, N):
        flat = score.reshape(-1)
        for a0, c0r, s0 in jobs:
            self.head_fwd(a4[a0:a0 + N], ce[c0r:c0r + N], A, Bv, c0, flat[s0:s0 + N])

    def head_bwd_data(self, coef, A, da4):
        """da4[n] = coef[n] * A."""
        da4.copy_((coef.double()[:, None, None] * A.double()[None]).reshape(da4.shape).to(da4.dtype))

    def head_bwd_reduce(self, coef, a4, dA):
        """dA[hw,c] += sum_n coef[n] a4[n,hw,c]."""
        n = a4.shape[0]
        dA.add_((coef[:, None].double() * a4.reshape(n, -1).double()).sum(0).reshape(dA.shape).to(dA.dtype))

    def head_param_grads(self, dA, dBv, dc0, wcr, bcr, wcs, dwcr, dbcr, dwcs, dbcs):
        K = wcr.shape[0]
        ws = wcs.reshape(K, 16).double()
        wr = wcr.reshape(K, -1).double()
        dA, dBv, dc0, bcr = dA.double(), dBv.double(), dc0.double(), bcr.double()
        nx = dA.shape[1]
        sw = ws.sum(1)
        g = torch.zeros_like(wr)
        g[:, :nx] = ws @ dA
        g[:, nx:] = sw[:, None] * dBv[None, :]
        dwcr.add_(g.reshape(dwcr.shape).to(dwcr.dtype))
        dbcr.add_((sw * dc0).to(dbcr.dtype))
        gs = wr[:, :nx] @ dA.t() + (wr[:, nx:] @ dBv)[:, None] + (bcr * dc0)[:, None]
        dwcs.add_(gs.reshape(dwcs.shape).to(dwcs.dtype))
        dbcs.add_(dc0.to(dbcs.dtype))

    # ---- conditioning augmentation pieces
    def ca_reparam(self, mu, sigma, eps, z, c_hat, cg):
        """c_hat = mu + sigma*eps (fp32); cg[N,1,1,128+nz] (T) = [c_hat, z] when cg is given."""
        c = mu.double() + sigma.double() * eps.double()
        c_hat.copy_(c.to(c_hat.dtype))
        if cg is not None:
            n = c.shape[0]
            row = cg.reshape(n, -1)
            row[:, :c.shape[1]] = c.to(cg.dtype)
            nz = z.shape[1] if z is not None else 0
            if z is not None:
                row[:, c.shape[1]:c.shape[1] + nz] = z.to(cg.dtype)
            row[:, c.shape[1] + nz:] = 0          # zero padding up to the row length

    def ca_bwd_seed(self, dcg, eps, mu, sigma, kl_scale, dmu, dsigma):
        """dmu = dc + kl_scale*(-2 mu); dsigma = dc*eps + kl_scale*(2/sigma - 2 sigma);
        dc = first 128 columns of dcg (T) or 0 when dcg is None."""
        nc = mu.shape[1]
        mu, sigma, eps = mu.double(), sigma.double(), eps.double()
        dc = 0 if dcg is None else dcg.reshape(mu.shape[0], -1)[:, :nc].double()
        dmu.copy_((dc + kl_scale * (-2 * mu)).to(dmu.dtype))
        dsigma.copy_((dc * eps + kl_scale * (2 / sigma - 2 * sigma)).to(dsigma.dtype))

    def ca_forward(self, tem, Wh, bh, Wmu, bmu, Wsg, bsg, eps, z, h, mu, sigma, c_hat, cg):
        """con_augment.py:13-22 in one call: the composition of the statements above."""
        self.linear_fwd(tem, Wh, bh, h, relu=True)
        self.linear_fwd(h, Wmu, bmu, mu)
        self.linear_fwd(h, Wsg, bsg, sigma)
        if eps is not None:
            self.ca_reparam(mu, sigma, eps, z, c_hat, cg)

    def ca_backward(self, dcg, eps, mu, sigma, kl_scale, h, tem, Wmu, Wsg, Wh, dmu, dsigma, dh, gWmu, gbmu, gWsg, gbsg, gWh,
                    gbh, dtem, dtem_acc):
        self.ca_bwd_seed(dcg, eps, mu, sigma, kl_scale, dmu, dsigma)
        self.linear_bwd(h, Wmu, dmu, gWmu, gbmu, dh, dx_acc=False)
        self.linear_bwd(h, Wsg, dsigma, gWsg, gbsg, dh, dx_acc=True)
        dh.mul_((h > 0).to(dh.dtype))                               # stored masked, like the kernel
        self.linear_bwd(tem, Wh, dh, gWh, gbh, dtem, dx_acc=dtem_acc)

    # ---- losses / gradient penalty
    def interp(self, real, fake, eps, out):
        e = eps.to(torch.float64)[:, None, None, None]
        out.copy_((real.to(torch.float64) * e + fake.to(torch.float64) * (1 - e)).to(out.dtype))

    def sample_sqnorm(self, g, out, zeroed=False):
        sq = (g.reshape(g.shape[0], -1).to(torch.float64) ** 2).sum(1).to(out.dtype)
        if zeroed:
            out.add_(sq)
        else:
            out.copy_(sq)

    def gp_seed(self, g, sq, coef, v):
        """v = coef * (1 - 1/||g||) g  (= d[coef/2 * sum (||g||-1)^2] / dg)."""
        nrm = torch.sqrt(sq.to(torch.float64))
        f = coef * (1 - 1 / nrm)
        v.copy_((g.to(torch.float64) * f[:, None, None, None]).to(v.dtype))

    def critic_loss(self, s_real, s_mis, s_fake, sq, lam, out):
        """out[0] = mean(cat(mis,fake)) - mean(real) + lam*gp ; out[1] = gp."""
        gp = ((torch.sqrt(sq.double()) - 1) ** 2).mean()
        out[0] = torch.cat((s_mis, s_fake)).double().mean() - s_real.double().mean() + lam * gp
        out[1] = gp

    def gen_loss(self, s_fake, mu, sigma, out):
        """out[0] = -mean(s) + sum(1 + log sigma^2 - mu^2 - sigma^2) ; out[1] = kl term."""
        mu, sigma = mu.double(), sigma.double()
        kl = (1 + torch.log(sigma * sigma) - mu * mu - sigma * sigma).sum()
        out[0] = -s_fake.double().mean() + kl
        out[1] = kl

    # ---- optimiser
    def adam_step(self, p, g, m, v, hyper):
        """hyper (fp32, device) = [lr, beta1, beta2, eps, step]; step is incremented first (torch Adam)."""
        hyper[4] += 1
        hyper[5] += 1            # exact count (lo word; the kernel carries into hyper[6] at 2^23)
        lr, b1, b2, eps, t = (float(x) for x in hyper[:5])
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)

    def concat_rep(self, x, c, out):
        """out[n,h,w,:Cx] = x ; out[n,h,w,Cx:] = c[n]  (generator_2.py:61-63: reshape/repeat/cat)."""
        Cx = x.shape[-1]
        out[..., :Cx] = x.to(out.dtype)
        out[..., Cx:] = c[:, None, None, :].to(out.dtype)

    def split_rep_bwd(self, dout, dx, dc):
        """dx = dout[..., :Cx] ; dc[n] (fp32, =) = sum_hw dout[n,h,w,Cx:]."""
        Cx = dx.shape[-1]
        dx.copy_(dout[..., :Cx])
        dc.copy_(dout[..., Cx:].double().sum((1, 2)).to(dc.dtype))

    def affine_f32(self, x, a, b, out):
        """out = a*x + b  (fp32 vectors)."""
        out.copy_(a * x + b)

    def fill(self, t, value):
        t.fill_(value)

    def zero(self, t):
        t.zero_()

    def zero_multi(self, ts):
        for t in ts:
            if t is not None:
                t.zero_()

    def scale_rows_add(self, x, scale, out, accumulate):
        """out (+)= scale[n] * x[n, ...]   (per-sample scale)."""
        s = scale.to(torch.float64).reshape(-1, *([1] * (x.dim() - 1)))
        r = x.to(torch.float64) * s
        if accumulate:
            out.copy_((out.to(torch.float64) + r).to(out.dtype))
        else:
            out.copy_(r.to(out.dtype))
