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
    # the rows widened after the first slice (same names as the reference, smoe_test.py:276-352)
    ("-bz", "--batch_size", int, [None], "+"), ("-dp", "--disable_train_pis", _B, False, None),
    ("-dg", "--disable_train_gammas", _B, False, None), ("-dm", "--disable_train_musx", _B, False, None),
    ("-udc", "--use_diff_center", _B, False, None), ("-ra", "--radial_as", _B, False, None),
    ("-oyg", "--only_y_gamma", _B, False, None), ("-ssim", "--ssim_opt", _B, False, None),
    ("-sp", "--sampling_percentage", int, 100, None), ("-ovl", "--overlap_of_batches", int, 0, None),
    ("-kcn", "--kernel_count_norm_l1", _B, False, None), ("-mask", "--loss_mask_path", str, None, None),
]


def main(image_path, results_path, iterations, validation_iterations, kernels_per_dim, params_file, l1reg, base_lr,
         batches, lr_div, lr_mult, use_determinant, normalize_pis, quantization_mode, bit_depths, quantize_pis,
         lower_bounds, upper_bounds, use_yuv, train_inverse_cov, update_kernel_list_iterations, callbacks=(),
         batch_size=(None,), disable_train_pis=False, disable_train_gammas=False, disable_train_musx=False,
         use_diff_center=False, radial_as=False, only_y_gamma=False, ssim_opt=False, sampling_percentage=100,
         overlap_of_batches=0, kernel_count_norm_l1=False, loss_mask_path=None):
    if len(bit_depths) != 5 or len(lower_bounds) != 5 or len(upper_bounds) != 5:
        raise ValueError("Number of bit depths / bounds must be five!")
    if ssim_opt:                                             # smoe_test.py:30-31
        sampling_percentage = 100
    if sampling_percentage <= 0 or sampling_percentage > 100:
        raise ValueError("Value of Sampling Percentage must be in range (0,100]")
    if quantization_mode >= 2:                               # smoe_test.py:36-37
        quantize_pis = True
    orig, precision, _ = read_image(image_path, use_yuv)
    if not orig.shape[-1] == 3:                              # smoe_test.py:41-44
        use_yuv = False
    if not use_yuv:
        only_y_gamma = False
    loss_mask = None
    if loss_mask_path is not None:                           # smoe_test.py:55-59
        import numpy as np
        loss_mask = np.load(loss_mask_path)["loss_mask"]
    os.makedirs(results_path, exist_ok=True)
    init = None
    if params_file is not None:
        from .utils import load_params
        init = load_params(params_file)
    smoe = Smoe(orig, kernels_per_dim, init_params=init, start_batches=batches, batch_size=list(batch_size),
                use_determinant=use_determinant,
                normalize_pis=normalize_pis, quantization_mode=quantization_mode, bit_depths=bit_depths,
                quantize_pis=quantize_pis, lower_bounds=lower_bounds, upper_bounds=upper_bounds, use_yuv=use_yuv,
                precision=precision, train_inverse_cov=train_inverse_cov, train_pis=not disable_train_pis,
                train_gammas=not disable_train_gammas, train_musx=not disable_train_musx,
                use_diff_center=use_diff_center, radial_as=radial_as, only_y_gamma=only_y_gamma, ssim_opt=ssim_opt,
                overlap_of_batches=overlap_of_batches, kernel_count_as_norm_l1=kernel_count_norm_l1,
                loss_mask=loss_mask)
    smoe.set_optimizer(AdamOptimizer(base_lr), AdamOptimizer(base_lr / lr_div), AdamOptimizer(base_lr * lr_mult))
    smoe.train(iterations, val_iter=validation_iterations, ukl_iter=update_kernel_list_iterations, pis_l1=l1reg,
               sampling_percentage=sampling_percentage, callbacks=list(callbacks), use_loss_mask=loss_mask is not None)
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
