"""Validation metrics of the reference's validation loop on the GPU (no CPU fallback).

Mirrors pht/models/afgsa/metric.py (``calculate_psnr`` / ``calculate_ssim`` / ``calculate_rmse``: batch inputs return the
SUM over the batch, like the reference's ndim == 4 branches) and pht/models/afgsa/util.py:77-119 (``tensor2img``) on
CUDA tensors through the C ABI (``pht_tonemap_u8``, ``pht_image_metrics_u8``, ``pht_mrse``).
"""
from __future__ import annotations

import math

import torch

from . import _lib as L

lib = L.lib


def tensor2img(x: torch.Tensor, post_spec: bool = False) -> torch.Tensor:
    """[B,C,H,W] fp32 (log-space radiance if ``post_spec``) -> tone-mapped uint8 [B,H,W,C] (util.py:77-119)."""
    L.require_cuda(x)
    x = x.contiguous().float()
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, dtype=torch.uint8, device=x.device)
    L.check(lib.pht_tonemap_u8(x.data_ptr(), out.data_ptr(), B, C, H, W, int(post_spec), L.stream_ptr()), "pht_tonemap_u8")
    return out


def _ws(B: int, device) -> torch.Tensor:
    return torch.empty(int(lib.pht_image_metrics_ws_bytes(B)) // 8 + 1, dtype=torch.float64, device=device)


def image_metrics(img1: torch.Tensor, img2: torch.Tensor) -> tuple[float, float]:
    """(sum of per-image PSNR, sum of per-image SSIM) of two uint8 [B,H,W,C] batches (metric.py:9-73)."""
    L.require_cuda(img1, img2)
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    assert img1.dtype == torch.uint8 and img2.dtype == torch.uint8 and img1.dim() == 4
    img1, img2 = img1.contiguous(), img2.contiguous()
    B, H, W, C = img1.shape
    sq = torch.empty(B, dtype=torch.int64, device=img1.device)
    ss = torch.empty(B, dtype=torch.float64, device=img1.device)
    ws = _ws(B, img1.device)
    L.check(lib.pht_image_metrics_u8(img1.data_ptr(), img2.data_ptr(), B, H, W, C, sq.data_ptr(), ss.data_ptr(), ws.data_ptr(),
                                     ws.numel() * 8, L.stream_ptr()), "pht_image_metrics_u8")
    psnr = 0.0
    for s in sq.tolist():
        mse = s / float(H * W * C)
        psnr += 0.0 if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))      # metric.py:20-24
    ssim = float(ss.sum()) / float((H - 10) * (W - 10) * C)
    return psnr, ssim


def calculate_psnr(img1: torch.Tensor, img2: torch.Tensor) -> float:
    return image_metrics(img1, img2)[0]


def calculate_ssim(img1: torch.Tensor, img2: torch.Tensor) -> float:
    return image_metrics(img1, img2)[1]


def calculate_rmse(output: torch.Tensor, gt: torch.Tensor, output_is_log: bool = False) -> float:
    """Sum over the batch of 0.5 * mean((a - b)^2 / (b^2 + 0.01)) (metric.py:76-94); ``output_is_log`` applies
    postprocess_specular (exp(x) - 1) to ``output`` on the fly, as the validation loop does before the call."""
    L.require_cuda(output, gt)
    if output.shape != gt.shape:
        raise ValueError("Input images must have the same dimensions!")
    a, b = output.contiguous().float(), gt.contiguous().float()
    B = a.shape[0]
    n = a[0].numel()
    out = torch.empty(B, dtype=torch.float64, device=a.device)
    ws = _ws(B, a.device)
    L.check(lib.pht_mrse(a.data_ptr(), b.data_ptr(), B, n, int(output_is_log), out.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                         L.stream_ptr()), "pht_mrse")
    return 0.5 * float(out.sum()) / n
