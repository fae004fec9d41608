"""Decoder: decoded-bitstream pickle -> reconstruction, the reference's
`smoe_reconstruction_decoded.py` (:16-81) on the B200 engine (BASELINE config 5).

The pickle carries, for the K' surviving kernels of a (H/4 x W/4) grid: `pis[0]`, `gamma_e`,
`musX` (offsets from the grid centres of `used_kernels[0]`), `nu_e`, `A_diagonal` (K',2),
`A_corr` (K',1), `shape_of_img[0]`, `dim_of_output[0]`, `used_determinants`.  A is rebuilt as
[[d0,0],[c,d1]] (:36-39) and fed over the compacted tensors (run_batched
with_quantized_params=True).  Decision D2: the gate form is ||A^T(x-mu)||^2 (the reference
script relies on a class default that contradicts the Cholesky-style A it feeds).
"""
from __future__ import annotations

import argparse
import os
import pickle

import numpy as np

from .smoe import Smoe
from .utils import read_image, write_image


def decode_params(cp, musX_init):
    """rparams dict from the decoded-bitstream dict (smoe_reconstruction_decoded.py:32-39)."""
    used = np.asarray(cp["used_kernels"][0]).astype(bool)
    rA_diagonal, rA_corr = np.asarray(cp["A_diagonal"]), np.asarray(cp["A_corr"])
    rA = np.concatenate((rA_diagonal, rA_corr, np.zeros_like(rA_corr)), axis=1)
    rA = rA[:, [0, 3, 2, 1]].reshape((rA_corr.shape[0], 2, 2))
    return {"A": rA, "musX": np.asarray(cp["musX"]) + musX_init[used, :], "nu_e": np.asarray(cp["nu_e"]),
            "pis": np.asarray(cp["pis"][0]), "gamma_e": np.asarray(cp["gamma_e"])}


def main(image_path=None, results_path=None, params_file=None, batches=1, cp=None, write=True):
    if cp is None:
        with open(params_file, "rb") as fd:
            cp = pickle.load(fd)
    k = [int(v) for v in np.int32(np.asarray(cp["shape_of_img"][0][:]) / 4)]
    precision = 8
    if image_path is not None:
        orig, precision, _ = read_image(image_path)
    else:
        orig = np.zeros((*cp["shape_of_img"][0][:], *cp["dim_of_output"][0][:]), dtype=np.float32)
    smoe = Smoe(orig, kernels_per_dim=k, start_batches=batches, use_determinant=bool(cp["used_determinants"]),
                use_yuv=True, train_inverse_cov=False, precision=precision, _decoder_only=True)
    smoe.rparams = decode_params(cp, smoe.musX_init)
    loss, mse, _, _ = smoe.run_batched(train=False, update_reconstruction=True, with_quantized_params=True)
    reconstruction = smoe.get_qreconstruction()
    if write:
        if results_path is None:
            results_path = "/tmp"
        elif not os.path.exists(results_path):
            os.mkdir(results_path)
        write_image(reconstruction, results_path + "/output", smoe.dim_domain, smoe.use_yuv, precision)
    return smoe, reconstruction, loss, mse


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--image_path", type=str, required=False, help="input image")
    parser.add_argument("-r", "--results_path", type=str, required=False, help="results path")
    parser.add_argument("-p", "--params_file", type=str, required=True, help="decoded parameter file")
    parser.add_argument("-b", "--batches", type=int, default=1)
    a = parser.parse_args()
    main(a.image_path, a.results_path, a.params_file, a.batches)
