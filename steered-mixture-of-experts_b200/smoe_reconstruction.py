"""Reconstruction from a trained-parameter pickle: the reference's `smoe_reconstruction.py`
(:15-105) on the B200 engine.  Same CLI flags, same outputs (reconstruction image + the quantised
parameter side-car pickle).

HEAD defects of the reference script that are not reproduced: `read_image` returns
a 3-tuple, `run_batched` returns 4 values, `write_image` needs `precision`; and the model is built
with the flags stored in the checkpoint (`use_determinant`, `use_yuv`) and with the
||A^T(x-mu)||^2 gate form (decision D2) instead of the class defaults.
"""
from __future__ import annotations

import argparse
import os
import pickle
import re

from .smoe import Smoe
from .utils import load_params, read_image, write_image
from .quantizer import quantize_params, rescaler


def main(image_path, results_path, params_file, batches=1, bit_depths=(20, 18, 6, 10, 10), quant_params=True):
    bit_depths = list(bit_depths)
    if len(bit_depths) != 5:
        raise ValueError("Number of bit depths must be five!")
    orig, precision, _ = read_image(image_path)
    init_params = load_params(params_file)
    if results_path is not None and not os.path.exists(results_path):
        os.mkdir(results_path)
    with open(params_file, "rb") as fd:
        cp = pickle.load(fd)
    qm = cp.get("quantization_mode")
    smoe = Smoe(orig, init_params=init_params, start_batches=batches, bit_depths=bit_depths, precision=precision,
                use_determinant=bool(cp.get("use_determinant", False)),
                use_yuv=bool(cp.get("use_yuv")) if qm is not None else False, train_inverse_cov=False)
    smoe.quantization_mode = qm if qm is not None else 0
    smoe.quantize_pis = bool(cp.get("quantized_pis")) if qm is not None else False
    smoe.lower_bounds, smoe.upper_bounds = cp.get("lower_bounds"), cp.get("upper_bounds")
    with_quantized_params = smoe.quantization_mode <= 0 and quant_params
    if with_quantized_params:
        smoe.quantize_pis = False               # min/max bounds, as the reference's mode-0 branch
        smoe.qparams = quantize_params(smoe, smoe.get_params())
        smoe.rparams = rescaler(smoe, smoe.qparams)
    loss, mse, _, _ = smoe.run_batched(train=False, update_reconstruction=True,
                                       with_quantized_params=with_quantized_params)
    iter_str = re.findall(r"\d+", os.path.basename(params_file))[-1] if re.findall(r"\d+", os.path.basename(params_file)) else "0"
    reconstruction_path = results_path + "/" + iter_str + "_reconstruction"
    if with_quantized_params:
        reconstruction = smoe.get_qreconstruction()
        add = "_{0:1d}_{1:1d}_{2:1d}_{3:1d}_{4:1d}".format(*bit_depths)
        reconstruction_path += add
        qparams = smoe.qparams
        qparams.update({"dim_of_domain": smoe.dim_domain, "dim_of_output": smoe.image.shape[-1],
                        "shape_of_img": smoe.image.shape[:-1], "used_ranges": False, "quantized_tria_params": True,
                        "trained_gamma": smoe.train_gammas, "radial_as": smoe.radial_as,
                        "trained_pis": smoe.train_pis})
        with open(results_path + "/" + iter_str + "_params" + add + ".pkl", "wb") as fd:
            pickle.dump(qparams, fd)
    else:
        reconstruction = smoe.get_reconstruction()
    write_image(reconstruction, reconstruction_path, smoe.dim_domain, smoe.use_yuv, precision)
    return smoe, loss, mse, reconstruction_path


def str2bool(v):
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--image_path", type=str, required=True, help="input image")
    parser.add_argument("-r", "--results_path", type=str, required=True, help="results path")
    parser.add_argument("-p", "--params_file", type=str, required=True, help="parameter file for model initialization.")
    parser.add_argument("-b", "--batches", type=int, default=1)
    parser.add_argument("-bd", "--bit_depths", type=int, default=[20, 18, 6, 10, 10], nargs="+")
    parser.add_argument("-qp", "--quant_params", type=str2bool, nargs="?", const=True, default=True)
    main(**vars(parser.parse_args()))
