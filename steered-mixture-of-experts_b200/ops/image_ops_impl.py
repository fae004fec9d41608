"""`custom_ssim` with the reference's signature (ops/image_ops_impl.py:235-293) and the PSNR / MSE
metrics, as GPU kernels (smoe_ssim / smoe_sqerr of include/smoe_b200.h)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _ffi
from .._ffi import check, lib, ptr, stream_ptr
import ctypes as C


def _to_dev(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(device)


def ssim_per_channel_padded(res, target, device=None):
    """Per-channel SSIM of two UNPADDED (spatial..., C) arrays, evaluated as the loss graph does
    (smoe.py:993-1004): SYMMETRIC pad 5 on every domain axis, then custom_ssim (VALID).  The
    padding is folded into the kernel's index reflection."""
    _ffi.require_cuda()
    device = device or torch.device("cuda", torch.cuda.current_device())
    a, b = _to_dev(res, device), _to_dev(target, device)
    assert a.shape == b.shape and a.dim() in (3, 4)
    d, Cc = a.dim() - 1, a.shape[-1]
    dims = (C.c_int32 * 3)(*(list(a.shape[:d]) + [1] * (3 - d)))
    ws = torch.empty((lib().smoe_ssim_workspace_bytes(d, dims, Cc) + 7) // 8, dtype=torch.float64, device=device)
    out = torch.zeros((4,), dtype=torch.float64, device=device)
    check(lib().smoe_ssim(d, dims, Cc, ptr(a), ptr(b), ptr(out), ptr(ws), stream_ptr()), "smoe_ssim")
    return out[:Cc].cpu().numpy()


def custom_ssim(img1, img2, max_val=1.0, ndim=2, device=None):
    """Reference signature: inputs are ALREADY padded by the caller; returns SSIM per channel over the
    VALID region.  Implemented by un-padding (the kernel re-creates the symmetric halo), which is
    exact when the inputs were padded symmetrically by 5 as at the reference's only call site."""
    if max_val != 1 and max_val != 1.0:
        raise NotImplementedError("max_val != 1")
    sl = (slice(5, -5),) * ndim
    a = img1[sl] if not isinstance(img1, torch.Tensor) else img1[sl]
    b = img2[sl] if not isinstance(img2, torch.Tensor) else img2[sl]
    return ssim_per_channel_padded(a, b, device)


def smoe_ssim(res, target, use_yuv=True, device=None):
    """Channel combination of smoe.py:1006-1009."""
    per = ssim_per_channel_padded(res, target, device)
    if use_yuv:
        return float(np.sum(per * np.array([6, 1, 1], dtype=np.float64)[:len(per)] if len(per) == 3 else per * 8) / 8), per
    return float(np.mean(per)), per


def mse_gpu(a, b, device=None):
    """mean((a-b)^2) with a fixed-order double accumulation on the GPU."""
    _ffi.require_cuda()
    device = device or torch.device("cuda", torch.cuda.current_device())
    x, y = _to_dev(a, device), _to_dev(b, device)
    ws = torch.empty((1024,), dtype=torch.float64, device=device)
    out = torch.zeros((1,), dtype=torch.float64, device=device)
    check(lib().smoe_sqerr(ptr(x), ptr(y), C.c_size_t(x.numel()), ptr(out), ptr(ws), stream_ptr()), "smoe_sqerr")
    return float(out.item()) / x.numel()
