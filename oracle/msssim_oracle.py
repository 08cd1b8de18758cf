"""CPU oracle of the optional MS-SSIM + L1 image loss (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

PARITY UNPINNED.  The reference's ``SSIMLoss`` (pht/models/losses.py:248-263) wraps
``kornia.losses.MS_SSIMLoss(reduction="mean")`` from kornia 0.8.0 (pyproject.toml:13, uv.lock).  kornia is neither
vendored under the reference checkout nor installed in this image and there is no network, so its source cannot be
imported and the reference holds no test or golden vector for this loss.  This module restates

  * the reference's own wrapper (losses.py:256-263): per-pixel scale = max(channel-max of the TARGET, 1) and
    ``ms_ssim(input / scale, target / scale)``;
  * kornia 0.8.0's published ``MS_SSIMLoss`` algorithm (the MS-SSIM + gaussian-weighted L1 mix of Zhao et al., "Loss
    Functions for Image Restoration with Neural Networks", in the widely copied pytorch-msssim-l1 formulation kornia
    adopted), with its documented defaults: sigmas (0.5, 1, 2, 4, 8), data_range 1, K = (0.01, 0.03), alpha 0.025,
    compensation 200, 33 x 33 normalised gaussian windows (size 4 * sigma_max + 1), zero padding 2 * sigma_max, and one
    grouped convolution with 3 * 5 windows ordered [sigma0 x3, sigma1 x3, ...] over 3 channel groups -- which pairs
    output channel c with input channel c // 5 and sigma c // 3 (so e.g. the red channel is filtered with sigma 0.5 three
    times and sigma 1 twice), takes the luminance term from the last three outputs and multiplies ALL 15
    contrast-structure maps.

Autograd through this function is the backward oracle.  If kornia's source becomes importable, pin this file against it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SIGMAS = (0.5, 1.0, 2.0, 4.0, 8.0)
K1, K2 = 0.01, 0.03
ALPHA = 0.025
COMPENSATION = 200.0
DATA_RANGE = 1.0
FILTER = int(4 * SIGMAS[-1] + 1)      # 33
PAD = int(2 * SIGMAS[-1])             # 16


def gauss_1d(sigma: float, size: int = FILTER, dtype=torch.float32) -> torch.Tensor:
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def window_bank(dtype=torch.float32) -> torch.Tensor:
    """[15, 1, 33, 33]: windows 3*i .. 3*i+2 are the sigma_i gaussian."""
    m = torch.zeros(3 * len(SIGMAS), 1, FILTER, FILTER, dtype=dtype)
    for i, s in enumerate(SIGMAS):
        g = gauss_1d(s, dtype=dtype)
        m[3 * i:3 * i + 3, 0] = torch.outer(g, g)
    return m


def ms_ssim_l1(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """kornia 0.8.0 MS_SSIMLoss(reduction="mean").forward for [B, 3, H, W] inputs."""
    g = window_bank(img1.dtype).to(img1.device)
    ch = img1.shape[-3]
    c1, c2 = (K1 * DATA_RANGE) ** 2, (K2 * DATA_RANGE) ** 2
    conv = lambda t: F.conv2d(t, g, groups=ch, padding=PAD)
    mux, muy = conv(img1), conv(img2)
    mux2, muy2, muxy = mux * mux, muy * muy, mux * muy
    sigmax2 = conv(img1 * img1) - mux2
    sigmay2 = conv(img2 * img2) - muy2
    sigmaxy = conv(img1 * img2) - muxy
    lc = (2 * muxy + c1) / (mux2 + muy2 + c1)
    cs = (2 * sigmaxy + c2) / (sigmax2 + sigmay2 + c2)
    l_m = lc[:, -1] * lc[:, -2] * lc[:, -3]
    loss_ms_ssim = 1 - l_m * cs.prod(dim=1)                              # [B, H, W]
    loss_l1 = (img1 - img2).abs()                                        # [B, C, H, W]
    gaussian_l1 = F.conv2d(loss_l1, g[-ch:], groups=ch, padding=PAD).mean(1)
    loss = ALPHA * loss_ms_ssim + (1 - ALPHA) * gaussian_l1 / DATA_RANGE
    return (COMPENSATION * loss).mean()


def ssim_loss(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """SSIMLoss.forward, pht/models/losses.py:256-263."""
    scale = torch.maximum(target.max(dim=1, keepdim=True)[0], torch.tensor(1.0, dtype=target.dtype, device=target.device))
    return ms_ssim_l1(output / scale, target / scale)
