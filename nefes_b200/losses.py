"""Mirror of script/models/losses.py:96-132 (NerfWLoss) -- same constructor, same `inputs` / `targets` convention --
with the transient-head case evaluated by two kernels (forward, backward) instead of ~45 elementwise launches."""
import torch
from torch import nn
from torch.autograd import Function

from . import _lib as L


class _NerfW(Function):
    @staticmethod
    def forward(ctx, rgb0, rgb, beta, tsig, target, coef, lambda_u):
        L.need_cuda(rgb0, rgb, beta, tsig, target)
        rgb0_c, rgb_c, beta_c, tsig_c, tgt_c = (L.f32c(t) for t in (rgb0, rgb, beta, tsig, target))
        N, S = tsig_c.shape
        dev = rgb_c.device
        scratch = torch.empty(8, device=dev)
        loss = torch.empty((), device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().nefes_nerfw_loss_fwd(L.ptr(rgb0_c), L.ptr(rgb_c), L.ptr(beta_c), L.ptr(tsig_c), L.ptr(tgt_c), N, S,
                                                 float(coef), float(lambda_u), L.ptr(scratch), L.ptr(loss), L.stream_of(rgb_c)),
                    "nefes_nerfw_loss_fwd")
        ctx.save_for_backward(rgb0_c, rgb_c, beta_c, tgt_c)
        ctx.meta = (N, S, float(coef), float(lambda_u), tuple(rgb0.shape), tuple(rgb.shape), tuple(beta.shape), tuple(tsig.shape))
        return loss

    @staticmethod
    def backward(ctx, g):
        rgb0_c, rgb_c, beta_c, tgt_c = ctx.saved_tensors
        N, S, coef, lambda_u, s0, s1, s2, s3 = ctx.meta
        dev = rgb_c.device
        g = L.f32c(g.reshape(1))
        d0, d1, db = torch.empty_like(rgb0_c), torch.empty_like(rgb_c), torch.empty_like(beta_c)
        dts = torch.empty(N, S, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().nefes_nerfw_loss_bwd(L.ptr(rgb0_c), L.ptr(rgb_c), L.ptr(beta_c), L.ptr(tgt_c), L.ptr(g), N, S, coef,
                                                 lambda_u, L.ptr(d0), L.ptr(d1), L.ptr(db), L.ptr(dts), L.stream_of(rgb_c)),
                    "nefes_nerfw_loss_bwd")
        return d0.reshape(s0), d1.reshape(s1), db.reshape(s2), dts.reshape(s3), None, None, None


class NerfWLoss(nn.Module):
    """Drop-in for losses.py:96.  inputs: 'rgb_coarse' [N,3], 'rgb_fine' [N,3], 'beta' [N], 'transient_sigmas' [N,S]."""

    def __init__(self, coef=1, lambda_u=0.01):
        super().__init__()
        self.coef = coef
        self.lambda_u = lambda_u

    def forward(self, inputs, targets, loss_mode=0):
        if 'rgb_fine' not in inputs or 'beta' not in inputs:
            raise RuntimeError("nefes_b200: NerfWLoss is built for the coarse+fine NeRF-W case (every reference config)")
        return _NerfW.apply(inputs['rgb_coarse'], inputs['rgb_fine'], inputs['beta'], inputs['transient_sigmas'], targets,
                            self.coef, self.lambda_u)
