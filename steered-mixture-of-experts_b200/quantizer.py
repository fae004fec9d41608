"""Parameter quantiser round trip with the reference's interface (quantizer.py:4-145):
`quantize_params(smoe, params) -> qparams`, `rescaler(smoe, qparams) -> rparams`.

The arithmetic runs on the GPU (smoe_colminmax / smoe_quantize / smoe_rescale of
include/smoe_b200.h): separately rounded IEEE operations in the precision NumPy uses at each
call site (float32 for data-dependent min/max bounds, float64 for the fixed bounds that the
reference builds as `np.ones(...) * python_float`), so codes are bit-exact with the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi
from ._ffi import check, lib, ptr, stream_ptr
from .utils import reduce_params

_ORDER = ("A_diagonal", "A_corr", "musX", "nu_e", "pis", "gamma_e")
_STEP_KEY = {"A_diagonal": "A", "A_corr": "A", "musX": "musX", "nu_e": "nu_e", "pis": "pis", "gamma_e": "gamma_e"}
_BOUND_SLOT = {"A_diagonal": 0, "A_corr": 0, "musX": 1, "nu_e": 2, "pis": 3, "gamma_e": 4}


def _device(smoe):
    return getattr(smoe, "device", None) or torch.device("cuda", torch.cuda.current_device())


def _bounds_for(smoe, name, x_dev, rows, cols, shape_tail):
    """(lb, ub) as float64 device vectors of `cols` entries + the NumPy arrays stored in qparams
    + whether NumPy would compute this tensor in float64."""
    qm = smoe.quantization_mode
    fixed = (qm == 2) or (name == "pis" and (qm == 2 or smoe.quantize_pis or qm > 1))
    if name == "pis" and qm <= 1 and not smoe.quantize_pis:
        fixed = False
    if name == "pis" and qm == 3 and not smoe.quantize_pis:
        raise UnboundLocalError("lb_pis (quantization_mode 3 needs quantize_pis, quantizer.py:36-41)")
    dev = x_dev.device
    if fixed:
        slot = _BOUND_SLOT[name]
        lb = torch.full((cols,), float(smoe.lower_bounds[slot]), dtype=torch.float64, device=dev)
        ub = torch.full((cols,), float(smoe.upper_bounds[slot]), dtype=torch.float64, device=dev)
        lb_np = np.ones((1,) + shape_tail) * smoe.lower_bounds[slot]
        ub_np = np.ones((1,) + shape_tail) * smoe.upper_bounds[slot]
        if name == "pis":
            lb_np, ub_np = lb_np.reshape(1), ub_np.reshape(1)
        return lb, ub, lb_np, ub_np, True
    lb = torch.empty((cols,), dtype=torch.float64, device=dev)
    ub = torch.empty((cols,), dtype=torch.float64, device=dev)
    check(lib().smoe_colminmax(ptr(x_dev), rows, cols, ptr(lb), ptr(ub), stream_ptr()), "smoe_colminmax")
    lb_np = lb.cpu().numpy().astype(np.float32).reshape((1,) + shape_tail)
    ub_np = ub.cpu().numpy().astype(np.float32).reshape((1,) + shape_tail)
    if name == "pis":
        lb_np, ub_np = lb_np.reshape(1), ub_np.reshape(1)
    return lb, ub, lb_np, ub_np, False


def quantize_params(smoe, params):
    _ffi.require_cuda()
    params, _ = reduce_params(params)
    radial = bool(getattr(smoe, "radial_as", False))       # A_diagonal is (K,), A_corr is not quantised
    dev = _device(smoe)
    bd = smoe.bit_depths
    steps = {"A": 2 ** bd[0] - 1, "musX": 2 ** bd[1] - 1, "nu_e": 2 ** bd[2] - 1, "pis": 2 ** bd[3] - 1,
             "gamma_e": 2 ** bd[4] - 1}
    lower, upper, out = {}, {}, {}
    for name in _ORDER:
        if radial and name == "A_corr":                     # quantizer.py:11, 45, 61, 80
            continue
        x = np.ascontiguousarray(np.asarray(params[name], dtype=np.float32))
        rows = x.shape[0]
        tail = tuple(x.shape[1:])
        cols = int(np.prod(tail)) if tail else 1
        if rows == 0:
            raise ValueError("no kernel with pi > 0 left to quantise")
        xd = torch.from_numpy(x.reshape(rows, cols)).to(dev)
        lb, ub, lb_np, ub_np, f64 = _bounds_for(smoe, name, xd, rows, cols, tail)
        codes = torch.empty((rows, cols), dtype=torch.float64 if f64 else torch.float32, device=dev)
        check(lib().smoe_quantize(ptr(xd), ptr(lb), ptr(ub), rows, cols, C.c_double(float(steps[_STEP_KEY[name]])),
                                  int(f64), ptr(codes), stream_ptr()), "smoe_quantize")
        out[name] = codes.cpu().numpy().reshape(x.shape)
        lower[name], upper[name] = lb_np, ub_np
    qparams = {"lower_bounds": lower, "upper_bounds": upper, "steps": steps}
    qparams.update(out)
    return qparams


def rescaler(smoe, qparams):
    _ffi.require_cuda()
    dev = _device(smoe)
    steps, lower, upper = qparams["steps"], qparams["lower_bounds"], qparams["upper_bounds"]
    r = {}
    radial = bool(getattr(smoe, "radial_as", False))
    for name in _ORDER:
        if radial and name == "A_corr":
            continue
        q = np.ascontiguousarray(qparams[name])
        lb_np, ub_np = np.asarray(lower[name]), np.asarray(upper[name])
        f64 = (q.dtype == np.float64) or (lb_np.dtype == np.float64)
        rows = q.shape[0]
        cols = int(np.prod(q.shape[1:])) if q.ndim > 1 else 1
        qd = torch.from_numpy(q.astype(np.float64 if f64 else np.float32).reshape(rows, cols)).to(dev)
        lb = torch.from_numpy(np.broadcast_to(lb_np.astype(np.float64).reshape(-1), (cols,)).copy()).to(dev)
        ub = torch.from_numpy(np.broadcast_to(ub_np.astype(np.float64).reshape(-1), (cols,)).copy()).to(dev)
        outd = torch.empty_like(qd)
        check(lib().smoe_rescale(ptr(qd), ptr(lb), ptr(ub), rows, cols, C.c_double(float(steps[_STEP_KEY[name]])),
                                 int(f64), ptr(outd), stream_ptr()), "smoe_rescale")
        r[name] = outd.cpu().numpy().reshape(q.shape)
    if radial:                                              # quantizer.py:132-136
        d = smoe.dim_domain
        rA = np.zeros((len(r["A_diagonal"]), d, d))
        rA[:, np.arange(d), np.arange(d)] = np.asarray(r["A_diagonal"]).reshape(-1, 1)
    else:
        rA = r["A_diagonal"] + r["A_corr"]                  # quantizer.py:138
    rmusX = r["musX"]
    if getattr(smoe, "use_diff_center", False):
        rmusX = rmusX + smoe.musX_init                        # quantizer.py:140-141
    return {"A": rA, "musX": rmusX, "nu_e": r["nu_e"], "pis": r["pis"], "gamma_e": r["gamma_e"]}
