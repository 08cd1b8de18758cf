"""CPU oracle for the patch-index sampler.  TEST INFRASTRUCTURE ONLY.

Restates (a) the third-party RNG the reference depends on -- CPython's
``random.Random`` (MT19937, stdlib; reference needs Python >= 3.11,
pyproject.toml:6) -- and (b) ``sample_patches_dart_throwing``
(pht/models/afgsa/preprocessing.py:171-213).

The MT19937 restatement follows the published Matsumoto-Nishimura algorithm and
CPython's documented derivations:
  seed(int)       -> init_by_array(32-bit little-endian limbs of abs(seed))
  getrandbits(k)  -> genrand_uint32() >> (32 - k)            (k <= 32)
  _randbelow(n)   -> k = n.bit_length(); r = getrandbits(k) until r < n
  randint(a, b)   -> a + _randbelow(b - a + 1)
  random()        -> ((u32 >> 5) * 2**26 + (u32 >> 6)) / 2**53
It is pinned against ``random.Random`` itself in tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import math

import numpy as np

_N, _M = 624, 397
_MASK = 0xFFFFFFFF


class MT19937:
    """Bit-exact restatement of CPython's random.Random core."""

    def __init__(self, seed: int) -> None:
        key = []
        s = abs(int(seed))
        if s == 0:
            key = [0]
        while s:
            key.append(s & _MASK)
            s >>= 32
        self._init_by_array(key)

    def _init_genrand(self, s: int) -> None:
        mt = [0] * _N
        mt[0] = s & _MASK
        for i in range(1, _N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & _MASK
        self.mt, self.idx = mt, _N

    def _init_by_array(self, key: list[int]) -> None:
        self._init_genrand(19650218)
        mt = self.mt
        i, j = 1, 0
        for _ in range(max(_N, len(key))):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & _MASK
            i += 1
            j += 1
            if i >= _N:
                mt[0] = mt[_N - 1]
                i = 1
            if j >= len(key):
                j = 0
        for _ in range(_N - 1):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & _MASK
            i += 1
            if i >= _N:
                mt[0] = mt[_N - 1]
                i = 1
        mt[0] = 0x80000000

    def _twist(self) -> None:
        mt = self.mt
        for k in range(_N):
            y = (mt[k] & 0x80000000) | (mt[(k + 1) % _N] & 0x7FFFFFFF)
            mt[k] = mt[(k + _M) % _N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.idx = 0

    def u32(self) -> int:
        if self.idx >= _N:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & _MASK

    def getrandbits(self, k: int) -> int:
        assert 0 < k <= 32
        return self.u32() >> (32 - k)

    def randbelow(self, n: int) -> int:
        k = n.bit_length()
        r = self.getrandbits(k)
        while r >= n:
            r = self.getrandbits(k)
        return r

    def randint(self, a: int, b: int) -> int:
        return a + self.randbelow(b - a + 1)

    def random(self) -> float:
        a = self.u32() >> 5
        b = self.u32() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


def dart_throwing(exr_shape, patch_size, num_patches, rng, max_iter=5000):
    """sample_patches_dart_throwing, preprocessing.py:179-213.

    ``rng`` needs ``randint(a, b)`` (MT19937 above or random.Random).
    Returns int64 (num_patches, 2) rows [x, y] (top-left corners).
    """
    h, w = exr_shape
    radius = math.sqrt((float(h * w) / num_patches) / math.pi)
    min_sq = (2 * radius) ** 2
    pts = np.zeros((num_patches, 2), dtype=np.int64)
    x_max = w - patch_size - 1
    y_max = h - patch_size - 1
    for n in range(num_patches):
        placed = False
        while not placed:
            for _ in range(max_iter):
                x = rng.randint(0, x_max)
                y = rng.randint(0, y_max)
                if n == 0:
                    ok = True  # reference: distance to an empty set is +inf
                else:
                    dx = pts[:n, 0] - x
                    dy = pts[:n, 1] - y
                    ok = int((dx * dx + dy * dy).min()) > min_sq
                if ok:
                    pts[n] = (x, y)
                    placed = True
                    break
            if not placed:
                radius *= 0.96
                min_sq = (2 * radius) ** 2
    return pts


def crop_patches(frame_nhwc: np.ndarray, centres: np.ndarray, patch_size: int) -> np.ndarray:
    """crop, preprocessing.py:325-344 applied to every centre [x, y]:
    rows py-P/2 : py+P/2, cols px-P/2 : px+P/2 (even P)."""
    half = patch_size // 2
    out = np.empty((len(centres), patch_size, patch_size, frame_nhwc.shape[-1]), frame_nhwc.dtype)
    for i, (px, py) in enumerate(centres):
        out[i] = frame_nhwc[py - half:py + half, px - half:px + half, :]
    return out


# ---------------------------------------------------------------------------------------------------------------
# importance sampling (preprocessing.py:119-168, 223-322)
# ---------------------------------------------------------------------------------------------------------------
def _reflect_index(i: np.ndarray, n: int) -> np.ndarray:
    """scipy.ndimage mode='reflect' (half-sample symmetric: d c b a | a b c d | d c b a)."""
    i = np.where(i < 0, -i - 1, i)
    return np.where(i >= n, 2 * n - 1 - i, i)


def uniform_filter1d(a: np.ndarray, size: int, axis: int) -> np.ndarray:
    """scipy.ndimage.uniform_filter1d (scipy 1.15.2 pinned by the reference's uv.lock; third-party, restated):
    out[i] = mean(a[i - size//2 : i - size//2 + size]) with 'reflect' borders, accumulated in double, result in
    the input dtype."""
    n = a.shape[axis]
    idx = np.arange(n)[:, None] - size // 2 + np.arange(size)[None, :]           # [n, size]
    g = np.take(a.astype(np.float64), _reflect_index(idx, n).reshape(-1), axis=axis)
    shape = list(a.shape)
    shape[axis:axis + 1] = [n, size]
    return (g.reshape(shape).sum(axis=axis + 1) / size).astype(a.dtype)


def uniform_filter_pp1(buffer: np.ndarray, patch_size: int) -> np.ndarray:
    """ndimage.uniform_filter(buffer, size=(P, P, 1)): separable, axis 0 then axis 1, intermediate in buffer.dtype."""
    return uniform_filter1d(uniform_filter1d(buffer, patch_size, 0), patch_size, 1)


def variance_map(buffer: np.ndarray, patch_size: int, relative: bool) -> np.ndarray:
    """get_variance_map, preprocessing.py:119-140 (float32 arithmetic like the reference's arrays)."""
    mean = uniform_filter_pp1(buffer, patch_size)
    square_mean = uniform_filter_pp1(buffer ** 2, patch_size)
    variance = np.maximum(square_mean - mean ** 2, 0)
    if relative:
        variance = variance / np.maximum(mean ** 2, 1e-4)
    variance = variance.max(axis=2)
    variance = np.minimum(variance ** (1.0 / 2.2), 1.0)
    return variance / np.maximum(variance.max(), 1e-4)


def importance_map(noisy: np.ndarray, normal: np.ndarray, patch_size: int) -> np.ndarray:
    """get_importance_map as called by importance_sampling (preprocessing.py:143-168, 293-300):
    relative variance of the noisy radiance + variance of the normals, weights 1, normalised by the maximum."""
    imp = variance_map(noisy, patch_size, True) * 1.0
    imp += variance_map(normal, patch_size, False) * 1.0
    return imp / np.max(imp)


def region_list(shape, step):
    """get_region_list, preprocessing.py:223-238: serpentine scan of step x step regions as (x0, x1, y0, y1)."""
    regions = []
    for y in range(0, shape[0], step):
        xs = list(range(0, shape[1], step))
        if (y // step) % 2 == 1:
            xs.reverse()
        regions.extend((x, x + step, y, y + step) for x in xs)
    return regions


def prune_patches(shape, centres: np.ndarray, patch_size: int, imp: np.ndarray, rng) -> np.ndarray:
    """prune_patches + split_patches, preprocessing.py:241-281: error-diffusion thinning.  Region bounds are
    inclusive on both sides; the arithmetic is numpy float32 scalar arithmetic (NEP 50, numpy 2.2.4 pinned): the error
    accumulator is float32 and rng.random() is rounded to float32 for the comparison."""
    remain = [tuple(int(v) for v in p) for p in centres]
    kept = []
    error = np.float32(0.0)
    for (x0, x1, y0, y1) in region_list(shape, 4 * patch_size):
        current = [p for p in remain if x0 <= p[0] <= x1 and y0 <= p[1] <= y1]
        remain = [p for p in remain if not (x0 <= p[0] <= x1 and y0 <= p[1] <= y1)]
        for (x, y) in current:
            v = np.float32(imp[y, x])
            if np.float32(v - error) > np.float32(rng.random()):
                kept.append((x, y))
                error = np.float32(error + np.float32(np.float32(1.0) - v))
            else:
                error = np.float32(error + np.float32(np.float32(0.0) - v))
    return np.array(kept, dtype=np.int64).reshape(-1, 2)


def importance_sampling(noisy: np.ndarray, normal: np.ndarray, patch_size: int, num_patches: int, rng,
                        imp: np.ndarray | None = None) -> np.ndarray:
    """importance_sampling, preprocessing.py:284-322 -> kept patch CENTRES [x, y] in processing order.
    ``rng`` needs randint and random (MT19937 above or random.Random); the map can be passed in precomputed."""
    if imp is None:
        imp = importance_map(noisy, normal, patch_size)
    corners = dart_throwing(noisy.shape[:2], patch_size, num_patches, rng)
    pad = patch_size // 2
    pruned = np.maximum(0, prune_patches(noisy.shape[:2], corners + pad, patch_size, imp, rng) - pad)
    return pruned + pad
