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
