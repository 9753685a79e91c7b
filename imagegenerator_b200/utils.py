"""``gradient_penalty`` -- drop-in for the reference's ``utils.py:8-26``.

Same signature and value: the WGAN-GP term mean_b (||d critic(x_b)/d x_b||_2 - 1)^2 at
x = eps*real + (1-eps)*fake with one uniform eps per sample.  The input gradient is produced by the
hand-written first-order backward chain of the critic (dgrad kernels + train-mode BN backward), not
by autograd; the returned 0-d tensor therefore carries no autograd graph -- the parameter gradients
of this term (the double backward) are produced inside ``stage_*_train_fn.train_*`` by
``CriticRT.gp_second_order``.
"""
import torch


def gradient_penalty(critic, real, fake, tem, device, eps=None):
    if not critic.training:
        # the reference's penalty differentiates whatever mode the critic is in; only the train-mode BatchNorm backward
        # (batch statistics) is implemented by the kernels, which is the mode both train functions call it in
        raise RuntimeError("gradient_penalty: the critic must be in train() mode (eval-mode BatchNorm backward is not implemented)")
    B = real.shape[0]
    rt = critic.runtime(B)
    ops = rt.ops
    if eps is None:
        eps = torch.rand((B, 1, 1, 1))                       # utils.py:10 (global CPU RNG)
    eps = eps.reshape(B).to(device=ops.device, dtype=torch.float32).contiguous()
    rt.refresh_weights()
    X = rt.a[0]
    ops.nchw_to_nhwc(real.contiguous().float(), rt.group_view(X, 0, 1))
    ops.nchw_to_nhwc(fake.detach().contiguous().float(), rt.group_view(X, 1, 1))
    ops.interp(rt.group_view(X, 0, 1), rt.group_view(X, 1, 1), eps, rt.group_view(X, 2, 1))
    rt.set_text(tem.detach().contiguous().float(), None)
    rt.forward(2, 1, dup_first=1, training=critic.training)
    rt.gp_first_order()
    out = torch.zeros(2, device=ops.device, dtype=torch.float32)
    zero = torch.zeros(B, device=ops.device, dtype=torch.float32)
    ops.critic_loss(zero, zero, zero, rt.sq, 1.0, out)
    from .layers import no_autograd
    return no_autograd(out[1].clone(), critic)
