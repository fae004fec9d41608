"""Minimal training driver with the core flags of the reference's `smoe_test.py` (which, despite its name,
is the training CLI; :260-356).  Only what the hot path needs: build a model on a regular kernel grid, three
Adam optimizers at lr, lr/lr_div and lr*lr_mult (smoe_test.py:84-88), train with optional pi-sparsification,
save the parameter pickle (utils.save_model schema) and the reconstruction.  Plotters, the kernel-adding loop
and the global-motion model are out of scope.

    python -m smoe_b200.smoe_test -i img.png -r out -k 32 32 -n 1000 -reg 1
"""
from __future__ import annotations

import argparse
import os

from .smoe import AdamOptimizer, Smoe
from .utils import read_image, save_model, write_image

_B = lambda v: str(v).lower() in ("yes", "true", "t", "y", "1")
_FLAGS = [  # (short, long, type, default, nargs)
    ("-i", "--image_path", str, None, None), ("-r", "--results_path", str, None, None),
    ("-n", "--iterations", int, 10000, None), ("-v", "--validation_iterations", int, 100, None),
    ("-k", "--kernels_per_dim", int, [12], "+"), ("-p", "--params_file", str, None, None),
    ("-reg", "--l1reg", float, 0.0, None), ("-lr", "--base_lr", float, 0.001, None),
    ("-b", "--batches", int, 1, None), ("-d", "--lr_div", float, 100.0, None), ("-m", "--lr_mult", float, 1000.0, None),
    ("-ud", "--use_determinant", _B, True, None), ("-np", "--normalize_pis", _B, True, None),
    ("-qm", "--quantization_mode", int, 0, None), ("-bd", "--bit_depths", int, [20, 18, 6, 10, 10], "+"),
    ("-qp", "--quantize_pis", _B, False, None), ("-lb", "--lower_bounds", float, [-2500, -.3, -5, 0, -32], "+"),
    ("-ub", "--upper_bounds", float, [2500, 1.3, 5, 2, 32], "+"), ("-yuv", "--use_yuv", _B, True, None),
    ("-tiv", "--train_inverse_cov", _B, False, None), ("-ukl", "--update_kernel_list_iterations", int, None, None),
]


def main(image_path, results_path, iterations, validation_iterations, kernels_per_dim, params_file, l1reg, base_lr,
         batches, lr_div, lr_mult, use_determinant, normalize_pis, quantization_mode, bit_depths, quantize_pis,
         lower_bounds, upper_bounds, use_yuv, train_inverse_cov, update_kernel_list_iterations, callbacks=()):
    if len(bit_depths) != 5 or len(lower_bounds) != 5 or len(upper_bounds) != 5:
        raise ValueError("Number of bit depths / bounds must be five!")
    orig, precision, _ = read_image(image_path, use_yuv)
    os.makedirs(results_path, exist_ok=True)
    init = None
    if params_file is not None:
        from .utils import load_params
        init = load_params(params_file)
    smoe = Smoe(orig, kernels_per_dim, init_params=init, start_batches=batches, use_determinant=use_determinant,
                normalize_pis=normalize_pis, quantization_mode=quantization_mode, bit_depths=bit_depths,
                quantize_pis=quantize_pis, lower_bounds=lower_bounds, upper_bounds=upper_bounds, use_yuv=use_yuv,
                precision=precision, train_inverse_cov=train_inverse_cov)
    smoe.set_optimizer(AdamOptimizer(base_lr), AdamOptimizer(base_lr / lr_div), AdamOptimizer(base_lr * lr_mult))
    smoe.train(iterations, val_iter=validation_iterations, ukl_iter=update_kernel_list_iterations, pis_l1=l1reg,
               callbacks=list(callbacks))
    save_model(smoe, os.path.join(results_path, "params_best.pkl"), best=True, quantize=quantization_mode >= 1)
    save_model(smoe, os.path.join(results_path, "params_last.pkl"), best=False, quantize=quantization_mode >= 1)
    write_image(smoe.get_reconstruction(), os.path.join(results_path, "reconstruction"), smoe.dim_domain, use_yuv, precision)
    return smoe


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    for short, long_, typ, default, nargs in _FLAGS:
        kw = dict(type=typ, default=default, required=default is None and long_ in ("--image_path", "--results_path"))
        if nargs:
            kw["nargs"] = nargs
        ap.add_argument(short, long_, **kw)
    main(**vars(ap.parse_args()))
