"""Mirror of script/models/losses.py:96-173 (NerfWLoss, ColorFeatureFusionNerfWLoss) -- same constructors, same `inputs` /
`targets` conventions -- with the transient-head colour loss and the feature losses evaluated by two kernels each (forward,
backward) instead of ~45 / ~10 elementwise launches."""
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


class _FeatLoss(Function):
    """mean |a - t| (+ mean |b - t|) or the squared version: nefes_feat_loss_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, a, b, target, mode):
        ts = [t for t in (a, b, target) if t is not None]
        L.need_cuda(*ts)
        a_c, t_c = L.f32c(a), L.f32c(target)
        b_c = None if b is None else L.f32c(b)
        if a_c.shape != t_c.shape or (b_c is not None and b_c.shape != t_c.shape):
            raise RuntimeError(f"nefes_b200: feature loss shapes differ: {tuple(a.shape)} / {tuple(target.shape)}")
        n = a_c.numel()
        dev = a_c.device
        scratch, loss = torch.empty(2, device=dev), torch.empty((), device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().nefes_feat_loss_fwd(L.ptr(a_c), L.ptr(b_c) if b_c is not None else None, L.ptr(t_c), n, int(mode),
                                                L.ptr(scratch), L.ptr(loss), L.stream_of(a_c)), "nefes_feat_loss_fwd")
        ctx.save_for_backward(*( [a_c, t_c] + ([b_c] if b_c is not None else []) ))
        ctx.meta = (n, int(mode), tuple(a.shape), b is not None)
        return loss

    @staticmethod
    def backward(ctx, g):
        n, mode, shape, has_b = ctx.meta
        a_c, t_c = ctx.saved_tensors[:2]
        b_c = ctx.saved_tensors[2] if has_b else None
        g = L.f32c(g.reshape(1))
        d_a = torch.empty_like(a_c)
        d_b = torch.empty_like(b_c) if has_b else None
        with torch.cuda.device(a_c.device):
            L.check(L.lib().nefes_feat_loss_bwd(L.ptr(a_c), L.ptr(b_c) if has_b else None, L.ptr(t_c), L.ptr(g), n, mode, L.ptr(d_a),
                                                L.ptr(d_b) if has_b else None, L.stream_of(a_c)), "nefes_feat_loss_bwd")
        return d_a.reshape(shape), (d_b.reshape(shape) if has_b else None), None, None


class ColorFeatureFusionNerfWLoss(nn.Module):
    """Drop-in for losses.py:134-173.  forward(inputs, targets, switch_on, color_only_switch):
    colour-only -> loss; stage 2 (switch_on=False) -> (loss, loss_f); stage 3 -> (loss, loss_f, loss_fusion), where
    loss_f = f(feat_fine, t) [+ f(feat_coarse, t)] and loss_fusion = f(feat_fusion, t), f = L1 or MSE mean."""

    def __init__(self, coef=1, L1_loss=False, lambda_u=0.01):
        super().__init__()
        self.coef = coef
        self.lambda_u = lambda_u
        self.loss = NerfWLoss(coef=coef, lambda_u=lambda_u)
        self.mode = 0 if L1_loss else 1

    def forward(self, inputs, targets, switch_on=True, color_only_switch=False):
        loss = self.loss(inputs, targets['rgb'])
        if color_only_switch:
            return loss
        loss_f = _FeatLoss.apply(inputs['feat_fine'], inputs.get('feat_coarse'), targets['feat'], self.mode)
        if switch_on:
            return loss, loss_f, _FeatLoss.apply(inputs['feat_fusion'], None, targets['feat'], self.mode)
        return loss, loss_f
