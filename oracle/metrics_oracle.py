"""CPU oracle of the validation metrics (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Plain numpy restatement of pht/models/afgsa/metric.py:9-94 and pht/models/afgsa/util.py:71-119 with the gaussian
filtering written out (no cv2): pinned against the real reference functions by tests/golden/make_golden_metrics.py.
"""
from __future__ import annotations

import math

import numpy as np


def tensor2img(x: np.ndarray, post_spec: bool = False) -> np.ndarray:
    """util.py:77-119 for [B,C,H,W] float32 -> uint8 [B,H,W,C] (postprocess_specular = exp(x) - 1, tone mapping gamma 2.2)."""
    x = np.transpose(x.astype(np.float32), (0, 2, 3, 1))
    if post_spec:
        x = np.exp(x) - 1
    with np.errstate(invalid="ignore"):
        t = np.clip(x ** (1.0 / 2.2), 0, 1) * 255.0
        return np.nan_to_num(np.clip(t, 0, 255), nan=0.0).astype(np.uint8)


def psnr(img1: np.ndarray, img2: np.ndarray) -> float:
    """metric.py:9-24 (batch: sum over images)."""
    if img1.ndim == 4:
        return sum(psnr(a, b) for a, b in zip(img1, img2))
    mse = np.mean((img1.astype(np.float64) - img2.astype(np.float64)) ** 2)
    return 0.0 if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def _gauss11() -> np.ndarray:
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 * 1.5))
    return g / g.sum()


def _filter_valid(img: np.ndarray) -> np.ndarray:
    """11x11 gaussian correlation, valid region only (== cv2.filter2D(...)[5:-5, 5:-5], metric.py:36-43)."""
    g = _gauss11()
    H, W = img.shape[:2]
    tmp = sum(g[i] * img[i:H - 10 + i] for i in range(11))
    return sum(g[j] * tmp[:, j:W - 10 + j] for j in range(11))


def ssim(img1: np.ndarray, img2: np.ndarray) -> float:
    """metric.py:27-73 (batch: sum over images; HWC images: mean over all channels)."""
    if img1.ndim == 4:
        return sum(ssim(a, b) for a, b in zip(img1, img2))
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    mu1, mu2 = _filter_valid(a), _filter_valid(b)
    s1 = _filter_valid(a * a) - mu1 ** 2
    s2 = _filter_valid(b * b) - mu2 ** 2
    s12 = _filter_valid(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))
    return float(m.mean())


def rmse(img1: np.ndarray, img2: np.ndarray) -> float:
    """metric.py:76-94 (batch: sum over images)."""
    if img1.ndim == 4:
        return sum(rmse(a, b) for a, b in zip(img1, img2))
    return float(0.5 * np.mean((img1 - img2) ** 2 / (img2 ** 2 + 1.0e-2)))
