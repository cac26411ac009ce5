"""Deterministic synthetic Y-plane generators shared by tests, goldens and bench."""
from __future__ import annotations

import numpy as np


def _box_blur(a: np.ndarray, k: int) -> np.ndarray:
    """k x k box filter with edge replication (pure numpy, cumulative sums; deterministic)."""
    if k <= 1:
        return a
    pad = k // 2
    p = np.pad(a, ((pad, k - 1 - pad), (pad, k - 1 - pad)), mode="edge")
    c = np.cumsum(np.cumsum(p, axis=0), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    H, W = a.shape
    s = c[k:k + H, k:k + W] - c[:H, k:k + W] - c[k:k + H, :W] + c[:H, :W]
    return s / float(k * k)


def texture(seed: int, height: int, width: int, blur: int = 9) -> np.ndarray:
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 256, size=(height, width)).astype(np.float64)
    f = _box_blur(f, blur)
    lo, hi = f.min(), f.max()
    return ((f - lo) / (hi - lo) * 255.0).astype(np.uint8)


def moving_clip(seed: int, height: int, width: int, nframes: int, *, step: int = 6, clamp: int = 96,
                noise: int = 2, blur: int = 9) -> np.ndarray:
    """Base texture cropped along a seeded random walk plus i.i.d. noise (S-1080 recipe, §8(d))."""
    margin = clamp + 32
    T = texture(seed, height + 2 * margin, width + 2 * margin, blur).astype(np.int16)
    rng = np.random.default_rng(seed + 7919)
    dx = dy = 0
    out = np.empty((nframes, height, width), dtype=np.uint8)
    crop_buf = np.empty((height, width), dtype=np.int16)
    for t in range(nframes):
        dx = int(np.clip(dx + rng.integers(-step, step + 1), -clamp, clamp))
        dy = int(np.clip(dy + rng.integers(-step, step + 1), -clamp, clamp))
        crop = T[margin + dy: margin + dy + height, margin + dx: margin + dx + width]
        nz = np.random.default_rng(seed + 1 + t).integers(-noise, noise + 1, size=(height, width), dtype=np.int8)
        np.clip(crop + nz, 0, 255, out=crop_buf)
        out[t] = crop_buf
    return out


def posterised_clip(seed: int, height: int, width: int, nframes: int, levels: int = 4) -> np.ndarray:
    """Tie-heavy content: large flat regions so many candidates share the minimum SAD."""
    c = moving_clip(seed, height, width, nframes, step=3, clamp=16, noise=0, blur=15)
    q = 256 // levels
    return ((c // q) * q).astype(np.uint8)
