"""Restatement of the reference quantiser round trip (host NumPy).

TEST INFRASTRUCTURE ONLY.  PINNED against the reference's own quantizer.py /
utils.reduce_params through tests/golden/quant_*.npz (oracle/make_golden.py).

  reduce_params ...... utils.py:7-15       (drops pi<=0 rows, MUTATES its input)
  quantize_params .... quantizer.py:4-88
  rescaler ........... quantizer.py:90-145

`smoe` is any object exposing quantization_mode, quantize_pis, radial_as, dim_domain,
image, lower_bounds, upper_bounds, bit_depths, use_diff_center, musX_init.
"""
from __future__ import annotations

import numpy as np

_TENSORS = ("A_diagonal", "A_corr", "musX", "nu_e", "gamma_e")
_STEP_OF = {"A_diagonal": "A", "A_corr": "A", "musX": "musX", "nu_e": "nu_e", "pis": "pis", "gamma_e": "gamma_e"}
_BOUND_IDX = {"A_diagonal": 0, "A_corr": 0, "musX": 1, "nu_e": 2, "pis": 3, "gamma_e": 4}


def reduce_params(params):
    idx = params["pis"] > 0
    for k in ("pis", "A_diagonal", "A_corr", "nu_e", "gamma_e", "musX"):
        params[k] = params[k][idx]
    return params, idx


def quantize_params(smoe, params):
    params, _ = reduce_params(params)
    qm = smoe.quantization_mode
    names = [n for n in _TENSORS if not (smoe.radial_as and n == "A_corr")]
    lb, ub = {}, {}
    if qm <= 1 or qm == 3:                                   # quantizer.py:8-19 per-entry min/max
        for n in names:
            lb[n] = np.amin(params[n], axis=0, keepdims=True)
            ub[n] = np.amax(params[n], axis=0, keepdims=True)
    elif qm == 2:                                            # quantizer.py:20-34 fixed bounds (float64)
        d, C = smoe.dim_domain, smoe.image.shape[-1]
        shapes = {"A_diagonal": (1,) if smoe.radial_as else (1, d, d), "A_corr": (1, d, d),
                  "musX": (1, d), "nu_e": (1, C), "gamma_e": (1, d, C)}
        for n in names:
            lb[n] = np.ones(shapes[n]) * smoe.lower_bounds[_BOUND_IDX[n]]
            ub[n] = np.ones(shapes[n]) * smoe.upper_bounds[_BOUND_IDX[n]]
    if qm <= 1 and not smoe.quantize_pis:                    # quantizer.py:36-41
        lb["pis"] = np.amin(params["pis"], axis=0, keepdims=True)
        ub["pis"] = np.amax(params["pis"], axis=0, keepdims=True)
    elif qm == 2 or smoe.quantize_pis:
        lb["pis"] = np.ones((1,)) * smoe.lower_bounds[3]
        ub["pis"] = np.ones((1,)) * smoe.upper_bounds[3]
    # (qm == 3 and not quantize_pis leaves the pi bounds unbound in the reference -> NameError there)

    bd = smoe.bit_depths
    steps = {"A": 2 ** bd[0] - 1, "musX": 2 ** bd[1] - 1, "nu_e": 2 ** bd[2] - 1,
             "pis": 2 ** bd[3] - 1, "gamma_e": 2 ** bd[4] - 1}
    q = {}
    for n in names + ["pis"]:                                # quantizer.py:58-75
        normalized = (params[n] - lb[n]) / (ub[n] - lb[n] + 10e-12)
        q[n] = np.round(normalized * steps[_STEP_OF[n]])
    out = {"lower_bounds": lb, "upper_bounds": ub, "steps": steps}
    out.update(q)
    return out


def rescaler(smoe, qparams):
    steps, lb, ub = qparams["steps"], qparams["lower_bounds"], qparams["upper_bounds"]
    names = [n for n in _TENSORS if not (smoe.radial_as and n == "A_corr")] + ["pis"]
    r = {n: qparams[n] / steps[_STEP_OF[n]] * (ub[n] - lb[n]) + lb[n] for n in names}   # quantizer.py:124-130
    if smoe.radial_as:
        rA = np.zeros((len(r["A_diagonal"]), smoe.dim_domain, smoe.dim_domain))
        for i in range(rA.shape[0]):
            np.fill_diagonal(rA[i], r["A_diagonal"][i])
    else:
        rA = r["A_diagonal"] + r["A_corr"]
    musX = r["musX"]
    if smoe.use_diff_center:
        musX += smoe.musX_init
    return {"A": rA, "musX": musX, "nu_e": r["nu_e"], "pis": r["pis"], "gamma_e": r["gamma_e"]}
