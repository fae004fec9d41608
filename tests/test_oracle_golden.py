"""The oracle's NumPy restatements vs vectors produced by the REAL reference code
(oracle/make_golden.py imported /root/reference/{smoe,quantizer,utils}.py under stubs)."""
import copy
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import init_ref, quant, ssim


def test_init_helpers_bit_exact_vs_reference():
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    for name in ("c1", "rgb", "vid", "one"):
        img = z[f"{name}_image"]
        d = img.ndim - 1
        k = [int(v) for v in z[f"{name}_k"]]
        np.testing.assert_array_equal(init_ref.gen_domain(img, d), z[f"{name}_joint_domain"])
        mus, A = init_ref.kernel_grid(k, d, bool(z[f"{name}_tic"]))
        np.testing.assert_array_equal(mus, z[f"{name}_musX"])
        np.testing.assert_array_equal(A, z[f"{name}_A"])
        nu, ga = init_ref.experts(img, mus)
        np.testing.assert_array_equal(nu, z[f"{name}_nu_e"])
        assert nu.dtype == np.float32
        np.testing.assert_array_equal(ga, z[f"{name}_gamma_e"])
        np.testing.assert_array_equal(init_ref.pis(mus.shape[0], bool(z[f"{name}_norm"])), z[f"{name}_pis"])
    # golden values quoted in SURVEY.md 8c
    np.testing.assert_allclose(z["c1_musX"][:3], [(.03125, .03125), (.03125, .09375), (.03125, .15625)])
    np.testing.assert_array_equal(z["c1_A"][0], np.diag([34., 34.]))
    assert z["c1_pis"][0] == np.float32(1 / 256)


def test_batch_shape_and_sliding_window_vs_reference():
    z = np.load(os.path.join(GOLDEN, "batch_shapes.npz"))
    for key in z.files:
        nb, shp = key.split("_")
        shape = tuple(int(v) for v in shp.split("x"))
        assert init_ref.get_batch_shape(int(nb), shape) == tuple(z[key]), key
    assert init_ref.get_batch_shape(4, (512, 512, 3)) == (256, 256, 3)
    assert init_ref.get_batch_shape(16, (1080, 1920, 5)) == (270, 480, 5)
    assert init_ref.get_batch_shape(8, (720, 1280, 32, 6)) == (360, 640, 16, 6)
    s = np.load(os.path.join(GOLDEN, "sliding_window.npz"))
    for img, bs, ov, ck, wk in ((s["img2"], (3, 4), 0, "coords2", "wins2"), (s["img3"], (2, 3, 2), 0, "coords3", "wins3"),
                                (s["img2"], (3, 4), 1, "coords2_ov", "wins2_ov")):
        got = list(init_ref.sliding_window(img, ov, bs))
        np.testing.assert_array_equal(np.array([c for c, _ in got]), s[ck])
        np.testing.assert_array_equal(np.array([w for _, w in got]), s[wk])


class _Shim:
    pass


def _quant_cases():
    z = np.load(os.path.join(GOLDEN, "quant_cases.npz"))
    for ci in range(int(z["num_cases"])):
        pre = f"case{ci}_"
        qm, qp, d, C, K = [int(v) for v in z[pre + "meta"][:5]]
        s = _Shim()
        s.quantization_mode, s.quantize_pis, s.radial_as, s.dim_domain = qm, bool(qp), False, d
        s.image = np.zeros((4,) * d + (C,), np.float32)
        s.lower_bounds, s.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
        s.bit_depths = [int(v) for v in z[pre + "meta"][5:]]
        s.use_diff_center, s.musX_init = False, None
        p = {k: z[pre + "in_" + k] for k in ("pis", "musX", "A_diagonal", "A_corr", "gamma_e", "nu_e")}
        yield ci, s, p, z, pre


def test_quantizer_round_trip_bit_exact_vs_reference():
    n = 0
    for ci, s, p, z, pre in _quant_cases():
        q = quant.quantize_params(s, copy.deepcopy(p))
        r = quant.rescaler(s, q)
        for k in ("A_diagonal", "A_corr", "musX", "nu_e", "pis", "gamma_e"):
            np.testing.assert_array_equal(q[k], z[pre + "q_" + k], err_msg=f"case {ci} codes {k}")
            assert q[k].dtype == z[pre + "q_" + k].dtype
            np.testing.assert_array_equal(np.asarray(q["lower_bounds"][k]), z[pre + "lb_" + k])
            np.testing.assert_array_equal(np.asarray(q["upper_bounds"][k]), z[pre + "ub_" + k])
        for k in ("A", "musX", "nu_e", "pis", "gamma_e"):
            np.testing.assert_array_equal(r[k], z[pre + "r_" + k], err_msg=f"case {ci} rescaled {k}")
        n += 1
    assert n >= 20


def test_quantizer_properties():
    for ci, s, p, z, pre in _quant_cases():
        if s.quantization_mode == 2:
            continue
        q = quant.quantize_params(s, copy.deepcopy(p))
        r = quant.rescaler(s, q)
        keep = p["pis"] > 0
        step = (np.asarray(q["upper_bounds"]["musX"]) - np.asarray(q["lower_bounds"]["musX"])) / q["steps"]["musX"]
        assert np.all(np.abs(r["musX"] - p["musX"][keep]) <= 0.5 * step * (1 + 1e-4) + 1e-7)   # error <= half a step
        assert q["pis"].shape[0] == int(keep.sum())
        # codes are integers in [0, step]
        for k, sk in (("musX", "musX"), ("nu_e", "nu_e"), ("gamma_e", "gamma_e"), ("A_diagonal", "A")):
            assert np.all(q[k] == np.round(q[k])) and q[k].min() >= 0 and q[k].max() <= q["steps"][sk]


def test_ssim_properties():
    rs = np.random.RandomState(3)
    a = rs.uniform(0, 1, (24, 28, 3)).astype(np.float32)
    val, per = ssim.smoe_ssim(a, a, use_yuv=False)
    np.testing.assert_allclose(per, 1.0, atol=1e-6)
    b = np.clip(a + 0.1 * rs.standard_normal(a.shape), 0, 1).astype(np.float32)
    v2, p2 = ssim.smoe_ssim(a, b, use_yuv=True)
    assert 0 < v2 < 1 and abs(v2 - float((p2 * [6, 1, 1]).sum() / 8)) < 1e-6
    w = ssim.gauss_window(2)
    w1 = np.exp(-0.5 * (np.arange(11) - 5.0) ** 2 / 1.5 ** 2)
    w1 /= w1.sum()
    np.testing.assert_allclose(w, np.outer(w1, w1), rtol=2e-6)       # softmax window is separable
    vid = rs.uniform(0, 1, (12, 12, 12, 1)).astype(np.float32)
    v3, p3 = ssim.smoe_ssim(vid, vid, use_yuv=False)
    np.testing.assert_allclose(p3, 1.0, atol=1e-5)
    assert abs(ssim.psnr(65536.0 * 0.01, 8) - 20.0) < 1e-9


def test_radial_quantizer_round_trip_bit_exact_vs_reference():
    """radial_as: A_diagonal is a (K,) vector, A_corr is left out (quantizer.py:11, 45, 61, 80, 132-136)."""
    z = np.load(os.path.join(GOLDEN, "quant_radial_cases.npz"))
    for ci in range(int(z["num_cases"])):
        pre = f"case{ci}_"
        qm, qp, d, C, K = [int(v) for v in z[pre + "meta"][:5]]
        s = _Shim()
        s.quantization_mode, s.quantize_pis, s.radial_as, s.dim_domain = qm, bool(qp), True, d
        s.image = np.zeros((4,) * d + (C,), np.float32)
        s.lower_bounds, s.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
        s.bit_depths = [int(v) for v in z[pre + "meta"][5:]]
        s.use_diff_center, s.musX_init = False, None
        p = {k: z[pre + "in_" + k] for k in ("pis", "musX", "A_diagonal", "A_corr", "gamma_e", "nu_e")}
        q = quant.quantize_params(s, copy.deepcopy(p))
        r = quant.rescaler(s, q)
        assert "A_corr" not in q
        for k in ("A_diagonal", "musX", "nu_e", "pis", "gamma_e"):
            np.testing.assert_array_equal(q[k], z[pre + "q_" + k], err_msg=f"case {ci} codes {k}")
            assert q[k].dtype == z[pre + "q_" + k].dtype
            np.testing.assert_array_equal(np.asarray(q["lower_bounds"][k]), z[pre + "lb_" + k])
        for k in ("A", "musX", "nu_e", "pis", "gamma_e"):
            np.testing.assert_array_equal(r[k], z[pre + "r_" + k], err_msg=f"case {ci} rescaled {k}")
