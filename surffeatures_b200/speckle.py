"""Deterministic synthetic ultrasound speckle -- numpy only.

Fully-developed speckle: a complex Gaussian scatterer field blurred by an anisotropic Gaussian PSF
(sigma_axial(y) = 1.2 px, sigma_lateral(x) = 2.0 px), envelope-detected (Rayleigh), log-compressed
over 50 dB and quantised to 8 bit.  The reference's own inputs are tracked ultrasound .mha sweeps
(SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:542-594) that are not in its tree.
"""
from __future__ import annotations

import hashlib

import numpy as np

SIGMA_Y = 1.2
SIGMA_X = 2.0


def _gauss_taps(sigma: float) -> np.ndarray:
    r = int(np.ceil(4.0 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur_axis(a: np.ndarray, taps: np.ndarray, axis: int) -> np.ndarray:
    r = len(taps) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    n = a.shape[axis]
    for i, t in enumerate(taps):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(i, i + n)
        out += t * p[tuple(sl)]
    return out


def _psf(z: np.ndarray) -> np.ndarray:
    return _blur_axis(_blur_axis(z, _gauss_taps(SIGMA_Y), -2), _gauss_taps(SIGMA_X), -1)


def _quantise(z: np.ndarray) -> np.ndarray:
    e = np.abs(z)
    e = e / e.mean()
    db = 20.0 * np.log10(e + 1e-6)
    g = (db + 40.0) * (255.0 / 50.0)
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)


def _field(rng: np.random.Generator, h: int, w: int) -> np.ndarray:
    return rng.standard_normal((h, w)) + 1j * rng.standard_normal((h, w))


def _blobs(rng: np.random.Generator, h: int, w: int, n: int) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    gain = np.ones((h, w))
    for _ in range(n):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        s = rng.uniform(3.0, 12.0)
        gain += 3.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    return gain


def speckle_frame(h: int, w: int, seed: int, blobs: int = 0) -> np.ndarray:
    """One h x w uint8 speckle frame."""
    rng = np.random.default_rng(seed)
    z = _field(rng, h, w)
    if blobs:
        z = z * _blobs(rng, h, w, blobs)
    return _quantise(_psf(z))


def speckle_stream(n: int, h: int, w: int, seed: int, shift=(0.7, 0.3), fresh: float = 0.02) -> np.ndarray:
    """n correlated frames: one scatterer field translated by shift*t px plus `fresh` new field."""
    rng = np.random.default_rng(seed)
    pad_x = int(np.ceil(abs(shift[0]) * n)) + 2
    pad_y = int(np.ceil(abs(shift[1]) * n)) + 2
    big = _field(rng, h + pad_y, w + pad_x)
    out = np.empty((n, h, w), np.uint8)
    for t in range(n):
        dx, dy = shift[0] * t, shift[1] * t
        ix, iy = int(np.floor(dx)), int(np.floor(dy))
        fx, fy = dx - ix, dy - iy
        a = big[iy:iy + h + 1, ix:ix + w + 1]
        z = ((1 - fy) * (1 - fx) * a[:-1, :-1] + (1 - fy) * fx * a[:-1, 1:]
             + fy * (1 - fx) * a[1:, :-1] + fy * fx * a[1:, 1:])
        z = np.sqrt(1.0 - fresh) * z + np.sqrt(fresh) * _field(rng, h, w)
        out[t] = _quantise(_psf(z))
    return out


def speckle_batch(n: int, h: int, w: int, seed0: int) -> np.ndarray:
    """n independent frames with seeds seed0 .. seed0+n-1."""
    return np.stack([speckle_frame(h, w, seed0 + i) for i in range(n)])


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
