"""Restatement of the reference's custom SSIM and PSNR (CPU NumPy).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (TensorFlow ops; see oracle/__init__.py).

  window ................ ops/image_ops_impl.py:131-151  (11 taps, sigma 1.5, softmax-normalised,
                          full (non-separable) 11^ndim window)
  reducer (VALID conv) .. ops/image_ops_impl.py:206-222
  luminance / cs ........ ops/image_ops_impl.py:106-128  (K1=.01, K2=.03, compensation 1.0)
  per-channel mean ...... ops/image_ops_impl.py:227-233, 293
  caller ................ smoe.py:993-1010 (SYMMETRIC pad 5, [6,1,1]/8 or mean)
  psnr .................. plotter.py:14-15
"""
from __future__ import annotations

from itertools import product

import numpy as np


def gauss_window(ndim, size=11, sigma=1.5, dtype=np.float32):
    c = np.arange(size, dtype=dtype) - dtype(size - 1) / dtype(2.0)
    g = (c * c) * dtype(-0.5 / (sigma * sigma))
    if ndim == 2:
        full = g[None, :] + g[:, None]
    else:
        full = g[None, None, :] + g[None, :, None] + g[:, None, None]
    flat = full.reshape(-1)
    e = np.exp(flat - flat.max())
    return (e / e.sum()).astype(dtype).reshape(full.shape)


def _reduce_valid(x, win):
    """VALID correlation of x (spatial..., C) with win (spatial...)."""
    nd = win.ndim
    size = win.shape[0]
    out_shape = tuple(s - size + 1 for s in x.shape[:nd]) + x.shape[nd:]
    out = np.zeros(out_shape, dtype=x.dtype)
    for off in product(range(size), repeat=nd):
        sl = tuple(slice(o, o + n) for o, n in zip(off, out_shape[:nd]))
        out += win[off] * x[sl]
    return out


def custom_ssim(img1, img2, max_val=1.0, ndim=2, dtype=np.float32):
    """Per-channel SSIM of two (spatial..., C) arrays (already padded by the caller)."""
    a = np.asarray(img1, dtype)
    b = np.asarray(img2, dtype)
    win = gauss_window(ndim, dtype=dtype)
    c1 = dtype((0.01 * max_val) ** 2)
    c2 = dtype((0.03 * max_val) ** 2)
    mean0 = _reduce_valid(a, win)
    mean1 = _reduce_valid(b, win)
    num0 = mean0 * mean1 * dtype(2.0)
    den0 = mean0 * mean0 + mean1 * mean1
    lum = (num0 + c1) / (den0 + c1)
    num1 = _reduce_valid(a * b, win) * dtype(2.0)
    den1 = _reduce_valid(a * a + b * b, win)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(axis=tuple(range(ndim)))


def smoe_ssim(res, target, use_yuv, dtype=np.float32):
    """SSIM as the reference's loss graph evaluates it (smoe.py:993-1010): SYMMETRIC pad by 5
    along every domain axis, per-channel custom_ssim, then [6,1,1]/8 (YUV) or channel mean."""
    nd = res.ndim - 1
    pad = ((5, 5),) * nd + ((0, 0),)
    a = np.pad(np.asarray(res, dtype), pad, mode="symmetric")
    b = np.pad(np.asarray(target, dtype), pad, mode="symmetric")
    per_ch = custom_ssim(a, b, 1.0, nd, dtype)
    if use_yuv:
        return float(np.sum(per_ch * np.array([6, 1, 1], dtype)) / 8), per_ch
    return float(np.mean(per_ch)), per_ch


def psnr(mse, precision):
    return 10 * np.log10((2 ** precision) ** 2 / mse)
