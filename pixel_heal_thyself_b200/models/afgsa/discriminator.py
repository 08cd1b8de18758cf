"""VGG-style WGAN critic (reference: pht/models/afgsa/model.py:264-344).

Adjacent to the hot path.  Its convolutions and linear layers stay stock PyTorch / cuDNN; the
BatchNorm2d + LeakyReLU of every conv block -- whose second-order backward (the gradient penalty differentiates the
critic's input gradient) PyTorch unrolls into hundreds of small launches, ~45 % of the critic step's GPU time -- runs on
hand-written kernels (``pht_bn_act_fwd / _bwd / _bwd_bwd``) through a twice-differentiable autograd function.  Same
parameter names, buffers and init order as the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


# Set by GradientPenaltyLoss around its ``torch.autograd.grad(pred, x_hat, create_graph=True)``: that pass only wants the
# INPUT gradient, but a Python autograd function cannot see which of its outputs the engine needs, so without the hint
# every convolution would also compute (and throw away) its weight and bias gradients there.
_input_grad_only = False


class input_grad_only:
    def __enter__(self):
        global _input_grad_only
        self.prev, _input_grad_only = _input_grad_only, True

    def __exit__(self, *exc):
        global _input_grad_only
        _input_grad_only = self.prev


_colsum_ws: dict = {}


def _bias_grad(g: torch.Tensor) -> torch.Tensor:
    """sum over (batch, rows, cols); outside a differentiable pass on the deterministic column-sum kernel"""
    C = g.shape[1]
    if (not torch.is_grad_enabled() and g.is_cuda and g.dtype == torch.float32 and C % 4 == 0 and C <= 1024
            and (C // 4) & (C // 4 - 1) == 0 and g.is_contiguous(memory_format=torch.channels_last)):
        from ... import ops
        key = (C, str(g.device))
        ws = _colsum_ws.get(key)
        if ws is None:
            ws = _colsum_ws[key] = ops.bn_act_ws(C, g.device)
        out = torch.empty(C, dtype=torch.float32, device=g.device)
        ops.colsum_nhwc(g, out, ws)
        return out
    return g.sum((0, 2, 3))


class _Conv2dForGP(torch.autograd.Function):
    """``F.conv2d`` whose backward is written with differentiable torch ops (conv_transpose2d / conv2d_weight).

    The WGAN-GP term differentiates THROUGH the critic's input gradient (losses.py:12-57).  With stock ``nn.Conv2d`` the
    second-order step runs ``aten::_convolution_double_backward``, which expresses the weight term as a convolution
    whose "kernel" is a 128x128 gradient map: cuDNN falls back to a generic indexed kernel (3-6 ms per layer at
    8 x 128 x 128, ~14 ms of a 36 ms critic step on B200).  Recording the first-order input gradient as a plain
    ``conv_transpose2d`` lets the second-order pass use the ordinary cuDNN data / weight gradient kernels.  Same arithmetic.
    """

    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.padding, ctx.has_bias = stride, padding, b is not None
        return F.conv2d(x, w, b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = F.conv_transpose2d(g, w, None, stride=ctx.stride, padding=ctx.padding)
            if gx.shape[2:] != x.shape[2:]:     # (sizes the stride does not divide: pad the tail like conv2d's backward)
                gx = F.pad(gx, (0, x.shape[3] - gx.shape[3], 0, x.shape[2] - gx.shape[2]))
        if ctx.needs_input_grad[1] and not _input_grad_only:
            gw = torch.nn.grad.conv2d_weight(x, w.shape, g, stride=ctx.stride, padding=ctx.padding)
        if ctx.has_bias and ctx.needs_input_grad[2] and not _input_grad_only:
            gb = _bias_grad(g)
        return gx, gw, gb, None, None


class _BNActBwdFn(torch.autograd.Function):
    """First-order backward of BatchNorm2d(train) + LeakyReLU as a differentiable op: (gz, x, gamma) -> (gx, g_gamma,
    g_beta); its own backward (the gradient-penalty pass) is ``pht_bn_act_bwd_bwd``."""

    @staticmethod
    def forward(ctx, gz, x, gamma, beta, stat, slope, ws):
        from ... import ops
        gz = gz.contiguous(memory_format=torch.channels_last)
        gx = torch.empty_like(x, memory_format=torch.channels_last)
        gg, gb = torch.empty_like(gamma), torch.empty_like(beta)
        ops.bn_act_bwd(x, gz, gamma, beta, stat, gx, gg, gb, ws, slope=slope)
        ctx.save_for_backward(gz, x, gamma, beta, stat)
        ctx.slope, ctx.ws = slope, ws
        ctx.set_materialize_grads(False)
        return gx, gg, gb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, h_gx, h_gg, h_gb):
        from ... import ops
        if h_gg is not None or h_gb is not None:
            raise NotImplementedError("second-order gradients through d/d gamma, d/d beta of the critic's BatchNorm are "
                                      "not needed by WGAN-GP (the penalty differentiates the INPUT gradient) and not built")
        gz, x, gamma, beta, stat = ctx.saved_tensors
        if h_gx is None:
            return None, None, None, None, None, None, None
        h = h_gx.contiguous(memory_format=torch.channels_last)
        h_gz, h_x = torch.empty_like(h), torch.empty_like(h)
        h_gamma = torch.empty_like(gamma)
        ops.bn_act_bwd_bwd(x, gz, h, gamma, beta, stat, h_gz, h_x, h_gamma, ctx.ws, slope=ctx.slope)
        return h_gz, h_x, h_gamma, None, None, None, None


class _BNActFn(torch.autograd.Function):
    """z = LeakyReLU(BatchNorm2d_train(x + conv_bias)) on channels-last fp32 CUDA tensors; running statistics updated in
    place.  ``conv_bias`` (may be None) is the bias of the convolution in front, which the caller did NOT add: batch
    normalisation cancels a per-channel constant exactly, so z and every gradient are those of x alone; only the running
    mean sees the bias (``pht_bn_act_fwd``'s ``pre_bias``).  Its gradient -- analytically zero, rounding noise in the
    reference -- is the column sum of the input gradient, as autograd would compute it."""

    @staticmethod
    def forward(ctx, x, gamma, beta, conv_bias, run_mean, run_var, eps, momentum, slope, ws):
        from ... import ops
        # (the caller passes a channels-last tensor: a copy made here would cut x out of the double-backward graph)
        assert x.is_contiguous(memory_format=torch.channels_last)
        z = torch.empty_like(x, memory_format=torch.channels_last)
        stat = torch.empty(2, x.shape[1], dtype=torch.float32, device=x.device)
        ops.bn_act_fwd(x, gamma, beta, run_mean, run_var, stat, z, ws, eps=eps, momentum=momentum, slope=slope,
                       pre_bias=None if conv_bias is None else conv_bias.detach())
        ctx.save_for_backward(x, gamma, beta, stat)
        ctx.slope, ctx.ws = slope, ws
        return z

    @staticmethod
    def backward(ctx, gz):
        x, gamma, beta, stat = ctx.saved_tensors
        gx, gg, gb = _BNActBwdFn.apply(gz, x, gamma, beta, stat, ctx.slope, ctx.ws)
        g_cb = _bias_grad(gx) if ctx.needs_input_grad[3] and not _input_grad_only else None
        return gx, gg, gb, g_cb, None, None, None, None, None, None


def _block(cin, cout, k, stride, bn):
    mods: list[nn.Module] = [nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=1)]
    if bn:
        mods.append(nn.BatchNorm2d(cout, affine=True))
    mods.append(nn.LeakyReLU(0.2, True))
    return nn.Sequential(*mods)


class DiscriminatorVGG(nn.Module):
    def __init__(self, in_nc: int, base_nf: int, input_size: int) -> None:
        super().__init__()
        n_down = int(math.log2(input_size / 4))
        feats = [_block(in_nc, base_nf, 3, 1, bn=False)]
        nc = base_nf
        for i in range(n_down):
            nxt = min(base_nf * 2 ** (i + 1), base_nf * 8)
            feats.append(_block(nc, nxt, 3, 1, bn=True))
            feats.append(_block(nxt, nxt, 4, 2, bn=True))
            nc = nxt
        self.features = nn.Sequential(*feats)
        side = input_size // 2 ** n_down
        self.classifier = nn.Sequential(nn.Linear(nc * side * side, 100), nn.LeakyReLU(0.2, True), nn.Linear(100, 1))
        import os
        self.fused_bn_act = os.environ.get("PHT_CRITIC_FUSED_BN", "1") != "0"
        self.fold_conv_bias = os.environ.get("PHT_CRITIC_FOLD_BIAS", "1") != "0"
        self._bn_workspaces: dict = {}
        # The convolution weights LIVE in channels-last memory (same shapes, names and values: a stride property only):
        # cuDNN's NHWC kernels take them as they are, instead of one layout-conversion kernel per weight per forward and
        # another per weight gradient (~100 small launches per critic step).
        for block in self.features:
            block[0].weight.data = block[0].weight.data.contiguous(memory_format=torch.channels_last)

    def _bn_ws(self, C, device):
        """one reduction workspace per channel count (launches of a step are stream-ordered)"""
        from ... import ops
        key = (C, str(device))
        ws = self._bn_workspaces.get(key)
        if ws is None:
            ws = self._bn_workspaces[key] = ops.bn_act_ws(C, device)
        return ws

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.is_cuda:
            # Same arithmetic as the reference module, laid out for cuDNN's tensor-op kernels: channels-last activations
            # (no NCHW<->NHWC conversion kernels around every conv), the 3-channel input of the first conv zero-padded to
            # 8 channels (the padded channels multiply zero weights), and convolutions whose second-order backward (the
            # gradient penalty) stays on cuDNN's regular kernels (_Conv2dForGP).  Parameters, names, state dict: untouched.
            nbt = []
            for bi, block in enumerate(self.features):
                conv = block[0]
                w = conv.weight
                if bi == 0:
                    pad_c = (-x.shape[1]) % 8
                    x = F.pad(x, (0, 0, 0, 0, 0, pad_c))
                    w = F.pad(w, (0, 0, 0, 0, 0, pad_c))
                x = x.contiguous(memory_format=torch.channels_last)
                bn = block[1] if isinstance(block[1], nn.BatchNorm2d) else None
                fused = bn is not None and self.training and x.dtype == torch.float32 and self.fused_bn_act
                fold = fused and self.fold_conv_bias
                # (in front of the fused BatchNorm the convolution's bias is not added: see _BNActFn)
                x = _Conv2dForGP.apply(x, w.contiguous(memory_format=torch.channels_last), None if fold else conv.bias,
                                       conv.stride, conv.padding)
                if fused:
                    # BatchNorm2d (batch statistics) + LeakyReLU(0.2) on the hand-written kernels, twice differentiable
                    x = _BNActFn.apply(x.contiguous(memory_format=torch.channels_last), bn.weight, bn.bias,
                                       conv.bias if fold else None,
                                       bn.running_mean, bn.running_var, bn.eps, bn.momentum, 0.2,
                                       self._bn_ws(bn.num_features, x.device))
                    nbt.append(bn.num_batches_tracked)
                else:
                    x = block[1:](x)
            if nbt:
                torch._foreach_add_(nbt, 1)      # nn.BatchNorm2d's num_batches_tracked += 1, one launch for all layers
        else:
            x = self.features(x)
        return self.classifier(x.reshape(x.size(0), -1))
