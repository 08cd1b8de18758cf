"""VGG-style WGAN critic (reference: pht/models/afgsa/model.py:264-344).

Adjacent to the hot path: it stays stock PyTorch / cuDNN in
this round and exists so a full GAN training step (base_trainer.py:388-457)
can run and be timed.  Same parameter names and init order as the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


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
        if ctx.needs_input_grad[1]:
            gw = torch.nn.grad.conv2d_weight(x, w.shape, g, stride=ctx.stride, padding=ctx.padding)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.sum((0, 2, 3))
        return gx, gw, gb, None, None


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

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.is_cuda:
            # Same arithmetic as the reference module, laid out for cuDNN's tensor-op kernels: channels-last activations
            # (no NCHW<->NHWC conversion kernels around every conv), the 3-channel input of the first conv zero-padded to
            # 8 channels (the padded channels multiply zero weights), and convolutions whose second-order backward (the
            # gradient penalty) stays on cuDNN's regular kernels (_Conv2dForGP).  Parameters, names, state dict: untouched.
            for bi, block in enumerate(self.features):
                conv = block[0]
                w = conv.weight
                if bi == 0:
                    pad_c = (-x.shape[1]) % 8
                    x = F.pad(x, (0, 0, 0, 0, 0, pad_c))
                    w = F.pad(w, (0, 0, 0, 0, 0, pad_c))
                x = x.contiguous(memory_format=torch.channels_last)
                x = _Conv2dForGP.apply(x, w.contiguous(memory_format=torch.channels_last), conv.bias, conv.stride, conv.padding)
                x = block[1:](x)
        else:
            x = self.features(x)
        return self.classifier(x.reshape(x.size(0), -1))
