"""GPU parity tests: the CUDA path (through the C ABI, via the Smoe host mirror) against the oracle.

Tolerances (BASELINE.json north_star): reconstruction max-abs <= 1e-5 on [0,1] pixels compared
before output quantisation; parameter gradients <= 1e-4 relative (to the tensor's
max-norm); PSNR after a fixed iteration count within 0.05 dB; pruned index sets and quantiser codes
bit-exact.  The graph has two discontinuities -- the gate threshold w > tau (smoe.py:825-827) and
the output rounding (smoe.py:899).  A float32 evaluation can land on the other side of either when
the float64 value is within rounding noise of it, so pixels whose oracle margin to a discontinuity
is below 1e-3 (threshold, relative) / 2e-3 (rounding, in code units) are compared with the loose
bound that a flip implies (tau resp. one code) and must be rare.
"""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PARAM_KEYS = ("pis", "musX", "A_diagonal", "A_corr", "gamma_e", "nu_e")
TAU = 0.5 / 256


def _mk(img, k, **kw):
    from smoe_b200 import Smoe, AdamOptimizer
    m = Smoe(img, kernels_per_dim=k, **kw)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    return m


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def _train_pass_both(m, o, **kw):
    """One training pass on both implementations; returns (gpu result, oracle result).

    The loss is discontinuous at the output rounding boundaries (smoe.py:899) and a few pixels per
    ten thousand sit within float32 noise of one, so the oracle evaluates its gradient at the GPU's
    rounding decisions (resq_override) -- after checking that the decisions differ only rarely."""
    o_state = [k.copy() for k in o.kernel_list_per_batch]
    o.run_batched(train=False, update_reconstruction=True, **kw)
    o.kernel_list_per_batch = o_state
    rg = m.run_batched(train=True, update_reconstruction=True, **kw)
    rec_gpu = m.get_reconstruction()
    flips = np.round(rec_gpu * 255) != np.round(o.get_reconstruction() * 255)
    assert flips.mean() < 5e-3
    ro = o.run_batched(train=True, update_reconstruction=True, resq_override=rec_gpu, **kw)
    return rg, ro


@pytest.mark.parametrize("name", ["g21", "g23", "g33", "g21tic", "g31"])
def test_golden_graph_cases(name):
    z = np.load(os.path.join(GOLDEN, "graph_cases.npz"))
    n = name + "_"
    img = z[n + "image"]
    tic, det, yuv = [bool(v) for v in z[n + "flags"]]
    m = _mk(img, [int(v) for v in z[n + "k"]], use_determinant=det, train_inverse_cov=tic, use_yuv=yuv)
    m.set_params({k: z[n + "p_" + k] for k in PARAM_KEYS})
    m.kernel_list_per_batch = [z[n + "kernel_list"]]
    m._enable_res_pre()
    loss, mse, num_pi, _ = m.run_batched(pis_l1=0.3, u_l1=1e-6, train=True, update_reconstruction=True)
    C = img.shape[-1]
    pre = m._d_res_pre.cpu().numpy().reshape(-1, C)
    ok = z[n + "thr_margin"] > 1e-3
    assert ok.mean() > 0.97
    assert np.abs(pre[ok] - z[n + "r_pre"][ok]).max() <= 1e-5
    assert np.abs(pre - z[n + "r_pre"]).max() <= 2 * TAU
    rq = m.get_reconstruction().reshape(-1, C)
    code_ref = z[n + "res"] * 255
    qok = ok & (np.abs(code_ref - np.floor(code_ref) - 0.5).min(axis=1) > 2e-3)
    np.testing.assert_array_equal(np.round(rq[qok] * 255), np.round(z[n + "resq"][qok] * 255))
    assert np.abs(np.round(rq * 255) - np.round(z[n + "resq"] * 255)).max() <= 1
    assert abs(loss - float(z[n + "loss"])) <= 2e-5 * max(1.0, abs(float(z[n + "loss"])))
    assert abs(mse - float(z[n + "mse_op"])) <= 2e-3 * float(z[n + "mse_op"]) + 1e-3
    assert num_pi == int((z[n + "p_pis"] > 0).sum())
    # active / influential index sets: bit-exact given identical pis
    K = int(m._counts[0, 0].item())
    np.testing.assert_array_equal(m.get_active_indices(), z[n + "indices"])
    infl_gpu = np.nonzero(m.kernel_list_per_batch[0])[0]
    sym = set(infl_gpu.tolist()) ^ set(z[n + "indices_infl"].tolist())
    assert len(sym) <= 1            # a kernel whose only passing gate sits on the threshold may flip
    am = m.get_weight_matrix_argmax().reshape(-1)
    assert (am[ok] == z[n + "w_e_max"][ok]).mean() > 0.995
    g = m.get_gradients()
    for k in PARAM_KEYS:
        assert _rel(g[k], z[n + "g_" + k]) < 1e-4 if ok.all() else _rel(g[k], z[n + "g_" + k]) < 2e-2, k


def _c1():
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    return z["c1_image"]


def test_config1_forward_backward_and_100_iterations_psnr():
    """BASELINE config 1: 128x128 grayscale, 16x16 grid, forward + 100 Adam iterations."""
    from oracle.model import OracleAdam, OracleSmoe
    img = _c1()
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False, normalize_pis=True)
    m = _mk(img, [16, 16], **kw)
    o = OracleSmoe(img, kernels_per_dim=[16, 16], dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    l_g, mse_g, _, _ = m.run_batched(train=False, update_reconstruction=True)
    l_o, mse_o, _, _ = o.run_batched(train=False, update_reconstruction=True)
    assert abs(l_g - l_o) < 1e-6 and abs(mse_g - mse_o) < 1e-3 * mse_o
    assert (np.round(m.get_reconstruction() * 255) != np.round(o.get_reconstruction() * 255)).mean() < 2e-3
    # one training pass: gradients.  The loss is discontinuous at the output rounding boundaries, and
    # with 16k pixels a handful sit within float32 noise of one (P ~ 2 * 1e-6 * 255 per pixel); the
    # oracle therefore evaluates the gradient at the GPU's rounding decisions (resq_override), after
    # checking that every differing pixel is such a boundary case (pre-quantisation values agree).
    _train_pass_both(m, o)
    g = m.get_gradients()
    for k, ref in o.last_grads.items():
        assert _rel(g[k], ref.numpy()) < 1e-4, k
    for _ in range(99):
        m.run_batched(train=True)
        o.run_batched(train=True)
    _, mse_g, _, _ = m.run_batched(train=False)
    _, mse_o, _, _ = o.run_batched(train=False)
    psnr_g, psnr_o = 10 * np.log10(65536 / mse_g), 10 * np.log10(65536 / mse_o)
    assert abs(psnr_g - psnr_o) < 0.05, (psnr_g, psnr_o)
    pg, po = m.get_params(), o.get_params()
    # beyond the PSNR bar: the float32 and float64 Adam trajectories (lr 1.0 on A) stay close, not identical
    assert _rel(pg["musX"], po["musX"]) < 3e-3 and _rel(pg["nu_e"], po["nu_e"]) < 1e-2


def test_compaction_bit_exact_and_packed_records():
    """pi-mask stream compaction (smoe.py:738-753): index sets bit-exact for identical pis."""
    rs = np.random.RandomState(5)
    img = rs.uniform(0, 1, (40, 40, 3)).astype(np.float32)
    m = _mk(img, [37, 41], use_determinant=True, train_inverse_cov=False, use_yuv=True)
    K = m.start_pis
    p = m.get_params()
    p["pis"] = rs.uniform(-0.5, 1.0, K).astype(np.float32)
    p["pis"][rs.choice(K, 50)] = 0.0
    p["A_corr"][:, 1, 0] = rs.normal(0, 20, K)
    m.set_params(p)
    kl = rs.uniform(size=K) < 0.7
    m.kernel_list_per_batch = [kl]
    _, _, num_pi, _ = m.run_batched(train=False)
    Kact = int(m._counts[0, 0].item())
    want = np.nonzero(kl & (p["pis"] > 0))[0]
    assert Kact == want.size and num_pi == int((p["pis"] > 0).sum())
    np.testing.assert_array_equal(m.get_active_indices(), want)
    # the records are packed in Hilbert order of the centres: bring them back to ascending index for the comparison
    order = np.argsort(m._indices[:Kact].cpu().numpy())
    rec = m._packed[:Kact].cpu().numpy()[order]
    assert (m._pos.cpu().numpy()[want] >= 0).all()
    np.testing.assert_array_equal(rec[:, 0:2], p["musX"][want])
    A = p["A_diagonal"][want] + p["A_corr"][want]
    Q = 0.72134752044448170368 * np.einsum("klj,kmj->klm", A.astype(np.float64), A.astype(np.float64))
    np.testing.assert_allclose(rec[:, 2:5], np.stack([Q[:, 0, 0], Q[:, 0, 1], Q[:, 1, 1]], 1), rtol=1e-6)
    coef = p["pis"][want].astype(np.float64) * A[:, 0, 0] * A[:, 1, 1] / (2 * np.pi)
    np.testing.assert_allclose(rec[:, 5], np.log2(coef), atol=1e-5)
    np.testing.assert_array_equal(rec[:, 6:9], p["nu_e"][want])
    # with fake-quantised pis (CLI default, smoe_test.py:304): 10 bits on [0,2]
    m2 = _mk(img, [37, 41], use_determinant=True, train_inverse_cov=False, quantize_pis=True,
             lower_bounds=[-2500, -.3, -5, 0, -32], upper_bounds=[2500, 1.3, 5, 2, 32], bit_depths=[20, 18, 6, 10, 10])
    p2 = m2.get_params()
    raw = rs.uniform(-0.002, 0.004, K).astype(np.float32)
    m2.set_params({"pis": raw})
    _, _, num_pi2, _ = m2.run_batched(train=False)
    from oracle.graph import fake_quant_args
    q = fake_quant_args(torch.tensor(raw), 0, 2, 10).numpy()
    assert num_pi2 == int((q > 0).sum())
    np.testing.assert_array_equal(m2.get_active_indices(), np.nonzero(q > 0)[0])
    np.testing.assert_array_equal(m2.get_params()["pis"], q)


def test_batches_accumulate_like_reference():
    """4 spatial batches: loss is the pixel-weighted sum, gradient the SUM of per-batch means."""
    from oracle.model import OracleAdam, OracleSmoe
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["rgb_image"]
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, start_batches=4)
    m = _mk(img, [6, 8], **kw)
    o = OracleSmoe(img, kernels_per_dim=[6, 8], dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    assert m.start_batches == o.start_batches == 4 and m.batch_size_valued == o.batch_size_valued
    lg, lo = _train_pass_both(m, o, pis_l1=0.1)
    assert abs(lg[0] - lo[0]) < 1e-6 and lg[2] == lo[2]
    g = m.get_gradients()
    for k, ref in o.last_grads.items():
        assert _rel(g[k], ref.numpy()) < 1e-4, k
    for a, b in zip(m.kernel_list_per_batch, o.kernel_list_per_batch):
        assert (a != b).sum() <= 1
    # Adam step (TF1 form) after the pass
    pg, po = m.get_params(), o.get_params()
    for k in PARAM_KEYS:
        assert np.abs(pg[k] - po[k]).max() <= 2e-6 * max(1.0, np.abs(po[k]).max()) + 1e-3 * (k in ("A_diagonal", "A_corr")), k


def test_video_and_pruning_state_machine():
    from oracle.model import OracleAdam, OracleSmoe
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["vid_image"]
    kw = dict(use_determinant=False, train_inverse_cov=False, use_yuv=False)
    m = _mk(img, [3, 4, 2], **kw)
    o = OracleSmoe(img, kernels_per_dim=[3, 4, 2], dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    p = m.get_params()
    p["pis"][5] = -1.0
    m.set_params({"pis": p["pis"]})
    o.vars["pis"][5] = -1.0
    a, b = _train_pass_both(m, o)
    assert abs(a[0] - b[0]) < 1e-6 and a[2] == b[2] == 23
    assert not m.kernel_list_per_batch[0][5]
    g = m.get_gradients()
    for k, ref in o.last_grads.items():
        assert _rel(g[k], ref.numpy()) < 1e-4, k
    p = m.get_params()
    p["pis"][5] = 1.0
    m.set_params({"pis": p["pis"]})
    _, _, num_pi, _ = m.run_batched(train=False)
    assert num_pi == 24 and not m.kernel_list_per_batch[0][5]      # back above 0 but off the list


@pytest.mark.parametrize("case", ["img", "img_tic", "video"])
def test_culling_skipping_and_dense_execution_are_bit_identical_and_deterministic(case):
    """dense_exec 0 (exact culling + exact-zero skipping), 1 (every pair executed), 2 (skipping only): the
    skipped terms are exact zeros, so reconstruction, arg-max, influence lists and gradients must agree
    BITWISE -- on a grid fine enough that most (tile, kernel) pairs really are culled."""
    if case == "video":
        import bench
        img = bench.synth_image((40, 48, 16, 3), 5)
        k, kw = [10, 12, 4], dict(use_determinant=True, train_inverse_cov=False, use_yuv=False)
    else:
        import bench
        img = bench.synth_image((96, 160, 3), 6)
        k = [24, 40]
        kw = dict(use_determinant=True, train_inverse_cov=(case == "img_tic"), use_yuv=True)
    rs = np.random.RandomState(11)
    out = []
    for mode in (0, 1, 2, 0):
        m = _mk(img, k, dense_exec=mode, **kw)
        p = m.get_params()
        d = m.dim_domain
        steer = np.random.RandomState(3).normal(0, 0.2 * p["A_diagonal"].max(), p["A_corr"].shape).astype(np.float32)
        p["A_corr"] = np.where(np.tril(np.ones((d, d), bool), -1)[None], steer, 0).astype(np.float32)
        if case == "img_tic":
            p["A_corr"] *= 0.3
        p["gamma_e"] = np.random.RandomState(4).normal(0, 0.3, p["gamma_e"].shape).astype(np.float32)
        m.set_params(p)
        m._enable_res_pre()
        m.run_batched(pis_l1=0.2, train=True, update_reconstruction=True)
        out.append((m._d_res_pre.cpu().numpy().copy(), m._grads.cpu().numpy().copy(), m._d_argmax.cpu().numpy().copy(),
                    m._klist.cpu().numpy().copy(), m._theta.cpu().numpy().copy()))
    for b in (1, 2, 3):
        for q in range(5):
            np.testing.assert_array_equal(out[0][q], out[b][q])
    assert np.isfinite(out[0][1]).all() and np.abs(out[0][1]).max() > 0


def test_fed_quantized_params_reconstruction():
    """quantize_params -> rescaler -> run_batched(with_quantized_params=True) (smoe.py:1688-1689)."""
    from oracle.model import OracleSmoe
    from oracle import quant as oq
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["rgb_image"]
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, quantization_mode=1,
              bit_depths=[20, 18, 6, 10, 10])
    m = _mk(img, [6, 8], **kw)
    o = OracleSmoe(img, kernels_per_dim=[6, 8], dtype=torch.float64, **kw)
    from smoe_b200 import quantize_params, rescaler
    m.qparams = quantize_params(m, m.get_params())
    m.rparams = rescaler(m, m.qparams)
    o.qparams = oq.quantize_params(o, o.get_params())
    o.rparams = oq.rescaler(o, o.qparams)
    for k in ("A_diagonal", "A_corr", "musX", "nu_e", "pis", "gamma_e"):
        np.testing.assert_array_equal(m.qparams[k], o.qparams[k], err_msg=k)
    for k in m.rparams:
        np.testing.assert_array_equal(m.rparams[k], o.rparams[k], err_msg=k)
    lg = m.run_batched(train=False, update_reconstruction=True, with_quantized_params=True)
    lo = o.run_batched(train=False, update_reconstruction=True, with_quantized_params=True)
    assert abs(lg[0] - lo[0]) < 2e-6
    assert (np.round(m.get_qreconstruction() * 255) != np.round(o.qreconstruction_image * 255)).mean() < 5e-3


def test_quantizer_codes_bit_exact_vs_reference_vectors():
    from smoe_b200 import quantize_params, rescaler
    z = np.load(os.path.join(GOLDEN, "quant_cases.npz"))

    class Shim:
        pass
    for ci in range(int(z["num_cases"])):
        pre = f"case{ci}_"
        qm, qp, d, C, K = [int(v) for v in z[pre + "meta"][:5]]
        s = Shim()
        s.quantization_mode, s.quantize_pis, s.radial_as, s.dim_domain = qm, bool(qp), False, d
        s.image = np.zeros((4,) * d + (C,), np.float32)
        s.lower_bounds, s.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
        s.bit_depths = [int(v) for v in z[pre + "meta"][5:]]
        s.use_diff_center, s.musX_init = False, None
        p = {k: z[pre + "in_" + k].copy() for k in PARAM_KEYS}
        q = quantize_params(s, p)
        r = rescaler(s, q)
        for k in ("A_diagonal", "A_corr", "musX", "nu_e", "pis", "gamma_e"):
            np.testing.assert_array_equal(q[k], z[pre + "q_" + k], err_msg=f"case {ci} codes {k}")
            assert q[k].dtype == z[pre + "q_" + k].dtype, (ci, k)
            np.testing.assert_array_equal(np.asarray(q["lower_bounds"][k]), z[pre + "lb_" + k])
            np.testing.assert_array_equal(np.asarray(q["upper_bounds"][k]), z[pre + "ub_" + k])
        for k in ("A", "musX", "nu_e", "pis", "gamma_e"):
            np.testing.assert_array_equal(r[k], z[pre + "r_" + k], err_msg=f"case {ci} rescaled {k}")


def test_ssim_and_psnr_kernels():
    from oracle import ssim as ossim
    from smoe_b200.ops.image_ops_impl import mse_gpu, smoe_ssim
    rs = np.random.RandomState(3)
    a = rs.uniform(0, 1, (37, 53, 3)).astype(np.float32)
    b = np.clip(a + 0.08 * rs.standard_normal(a.shape), 0, 1).astype(np.float32)
    v, per = smoe_ssim(a, b, use_yuv=True)
    vo, pero = ossim.smoe_ssim(a, b, use_yuv=True, dtype=np.float64)
    np.testing.assert_allclose(per, pero, atol=2e-5)
    assert abs(v - vo) < 2e-5
    assert abs(smoe_ssim(a, a, use_yuv=False)[0] - 1.0) < 1e-6
    vid_a = rs.uniform(0, 1, (14, 13, 12, 1)).astype(np.float32)
    vid_b = np.clip(vid_a + 0.1 * rs.standard_normal(vid_a.shape), 0, 1).astype(np.float32)
    np.testing.assert_allclose(smoe_ssim(vid_a, vid_b, use_yuv=False)[1],
                               ossim.smoe_ssim(vid_a, vid_b, use_yuv=False, dtype=np.float64)[1], atol=2e-5)
    assert abs(mse_gpu(a, b) - float(((a.astype(np.float64) - b) ** 2).mean())) < 1e-9


def test_full_size_properties_config2():
    """BASELINE config 2 shape (512x512, 64x64 grid): size-independent properties."""
    z = _c1()
    img = np.tile(z, (4, 4, 1))
    m = _mk(img, [64, 64], use_determinant=True, train_inverse_cov=False, use_yuv=False)
    l0, mse0, num_pi, _ = m.run_batched(train=False, update_reconstruction=True)
    assert num_pi == 4096 and np.isfinite(l0)
    rec = m.get_reconstruction()
    assert rec.min() >= 0 and rec.max() <= 1
    assert np.abs(rec * 255 - np.round(rec * 255)).max() < 1e-3          # on the 8-bit lattice
    # mse_op reported by the fused epilogue == GPU metric kernel on the stored reconstruction
    from smoe_b200.ops.image_ops_impl import mse_gpu
    assert abs(mse_gpu(rec, img) * 65536 - mse0) < 1e-4 * mse0
    # idempotence / determinism: same launch twice -> bitwise equal
    l1, mse1, _, _ = m.run_batched(train=False, update_reconstruction=True)
    assert l0 == l1 and mse0 == mse1
    np.testing.assert_array_equal(rec, m.get_reconstruction())
    # the 4x4 tiling of a periodic pattern with a matching grid keeps block-mean experts: loss drops under training
    for _ in range(5):
        lt = m.run_batched(train=True)[0]
    assert lt < l0
    # K = 1: reconstruction is the clipped linear expert
    one = _mk(img[:64, :64], [1, 1], use_determinant=False, train_inverse_cov=False, use_yuv=False)
    one.set_params({"gamma_e": np.array([[[0.3], [-0.2]]], np.float32), "nu_e": np.array([[0.4]], np.float32)})
    one._enable_res_pre()
    one.run_batched(train=False, update_reconstruction=True)
    yy, xx = np.meshgrid(np.linspace(0, 1, 64), np.linspace(0, 1, 64), indexing="ij")
    np.testing.assert_allclose(one._d_res_pre.cpu().numpy().reshape(64, 64), 0.4 + 0.3 * yy - 0.2 * xx, atol=2e-6)


def _decoded_dict(H, W, C, seed, keep=0.125):
    """Synthetic decoded-bitstream dict per SURVEY.md 8d (config 5), any size."""
    rs = np.random.RandomState(seed)
    gh, gw = H // 4, W // 4
    used = rs.uniform(size=gh * gw) < keep
    K = int(used.sum())
    return {"shape_of_img": [np.array([H, W])], "dim_of_output": [np.array([C])], "used_determinants": 1,
            "used_kernels": [used.astype(np.float64)], "pis": [rs.uniform(.5, 1.5, K).astype(np.float32)],
            "musX": (rs.uniform(-.5, .5, (K, 2)) / np.array([gh, gw])).astype(np.float32),
            "A_diagonal": rs.uniform(0.3 * gw, 1.5 * gw, (K, 2)).astype(np.float32),
            "A_corr": rs.normal(0, 0.1 * gw, (K, 1)).astype(np.float32),
            "nu_e": rs.uniform(0, 1, (K, C)).astype(np.float32), "gamma_e": rs.normal(0, 1, (K, 2, C)).astype(np.float32)}


def test_decoded_entry_point_matches_oracle():
    """smoe_reconstruction_decoded.py:16-62 (BASELINE config 5 path) at a size the oracle can hold."""
    from smoe_b200 import smoe_reconstruction_decoded as dec
    from oracle.graph import GraphCfg, graph_forward
    from oracle import init_ref
    H, W, C = 72, 96, 3
    cp = _decoded_dict(H, W, C, seed=1005)
    smoe, rec, loss, mse = dec.main(cp=cp, write=False)
    assert rec.shape == (H, W, C)
    mus_grid, _ = init_ref.kernel_grid([H // 4, W // 4], 2, False)
    rp = dec.decode_params(cp, mus_grid)
    np.testing.assert_array_equal(rp["A"][:, 0, 1], 0)
    np.testing.assert_array_equal(rp["A"][:, 1, 0], cp["A_corr"][:, 0])
    K = rp["pis"].shape[0]
    cfg = GraphCfg(dim_domain=2, num_channels=C, use_determinant=True, train_inverse_cov=False, use_yuv=True, start_pis=K)
    jd = init_ref.gen_domain(np.zeros((H, W, C), np.float32), 2).reshape(-1, 2 + C).astype(np.float32).astype(np.float64)
    feed = {k: torch.tensor(np.asarray(v, np.float64)) for k, v in rp.items()}
    dummy = {"pis": torch.ones(K, dtype=torch.float64), "musX": feed["musX"], "A_diagonal": feed["A"], "A_corr": feed["A"] * 0,
             "gamma_e": feed["gamma_e"], "nu_e": feed["nu_e"]}
    out = graph_forward(dummy, np.ones(K, bool), torch.tensor(jd[:, :2]), torch.tensor(jd[:, 2:]), cfg, feed=feed)
    ref = out["resq"].numpy().reshape(H, W, C)
    assert (np.round(rec * 255) != np.round(ref * 255)).mean() < 5e-3
    assert np.abs(rec - ref).max() <= 1.0 / 255 + 1e-6
    assert abs(loss - float(out["loss"])) < 2e-5


def test_reconstruction_entry_point_round_trip(tmp_path):
    """train -> save_model -> smoe_reconstruction.main with quantised parameters (smoe_reconstruction.py:15-79)."""
    from smoe_b200 import save_model, smoe_reconstruction
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["rgb_image"]
    m = _mk(img, [6, 8], use_determinant=True, train_inverse_cov=False, use_yuv=False)
    for _ in range(10):
        m.run_batched(train=True)
    save_model(m, str(tmp_path / "00000010_params.pkl"), quantize=False)
    np.save(str(tmp_path / "img.npy"), np.round(img * 255).astype(np.uint8))
    os.mkdir(str(tmp_path / "out"))
    s2, loss, mse, path = smoe_reconstruction.main(str(tmp_path / "img.npy"), str(tmp_path / "out"),
                                                   str(tmp_path / "00000010_params.pkl"))
    assert os.path.exists(path + ".png") and path.endswith("_20_18_6_10_10")
    assert os.path.exists(str(tmp_path / "out" / "00000010_params_20_18_6_10_10.pkl"))
    # quantisation at these bit depths costs little: PSNR within 1 dB of the unquantised model
    _, mse_f, _, _ = m.run_batched(train=False)
    assert abs(10 * np.log10(mse / mse_f)) < 1.0
    # steering survives the save / load round trip
    p_saved = m.get_params()
    p_loaded = s2.get_params()
    np.testing.assert_array_equal(p_saved["A_corr"], p_loaded["A_corr"])
    np.testing.assert_array_equal(p_saved["A_diagonal"], p_loaded["A_diagonal"])


def test_negative_determinant_weight_is_reported():
    """pi * prod(diag A) < 0 under use_determinant: the reference would give the kernel a NEGATIVE weight (smoe.py:809-820);
    the log-domain kernels evaluate |weight| and the host says so (counts[2] -> one warning per model)."""
    rs = np.random.RandomState(2)
    img = rs.uniform(0, 1, (32, 32, 1)).astype(np.float32)
    m = _mk(img, [4, 4], use_determinant=True, train_inverse_cov=False, use_yuv=False)
    p = m.get_params()
    p["A_diagonal"][5, 0, 0] *= -1
    m.set_params(p)
    with pytest.warns(UserWarning, match="pi \\* prod"):
        m.run_batched(train=False)
    assert m._last_nonpos == 1
    m2 = _mk(img, [4, 4], use_determinant=True, train_inverse_cov=False, use_yuv=False)
    m2.run_batched(train=False)
    assert m2._last_nonpos == 0


def test_reference_written_checkpoint_loads_and_reconstructs(tmp_path):
    """a checkpoint pickle written by the REFERENCE's utils.save_model (tests/golden/, made by
    oracle/make_golden.py from the real code) goes through the product's smoe_reconstruction.main unchanged, and the
    reconstruction from its parameters matches the float64 oracle."""
    import shutil
    from smoe_b200 import smoe_reconstruction
    from smoe_b200.utils import load_params
    pkl = os.path.join(GOLDEN, "ref_saved_model_00000100_params.pkl")
    shutil.copy(pkl, str(tmp_path / "00000100_params.pkl"))
    shutil.copy(os.path.join(GOLDEN, "ref_saved_model_image.npy"), str(tmp_path / "img.npy"))
    s2, loss, mse, path = smoe_reconstruction.main(str(tmp_path / "img.npy"), str(tmp_path / "out"),
                                                   str(tmp_path / "00000100_params.pkl"))
    assert os.path.exists(path + ".png") and s2.start_pis == 39 and s2.use_yuv and s2.use_determinant
    params = load_params(pkl)
    img = np.load(os.path.join(GOLDEN, "ref_saved_model_image.npy")).astype(np.float32) / 255.
    pre = s2.get_pre_clip_reconstruction().reshape(-1, 3)
    idx = np.arange(img.shape[0] * img.shape[1])
    out = _oracle_at_pixels(img, params, idx, dict(use_determinant=True, train_inverse_cov=False, use_yuv=True))
    _check_sampled_forward(pre, out, tol_frac=0.9)
    assert abs(loss - float(out["loss"])) < 2e-5 and abs(mse / float(out["mse_op"]) - 1) < 2e-3


def _oracle_at_pixels(img, params, flat_idx, cfgkw, feed=None):
    """float64 oracle forward at a sample of pixels (each pixel's output depends on no other pixel)."""
    from oracle.graph import GraphCfg, graph_forward
    d, C = img.ndim - 1, img.shape[-1]
    K = params["pis"].shape[0]
    axes = [np.linspace(0, 1, n).astype(np.float32).astype(np.float64) for n in img.shape[:d]]
    sub = np.unravel_index(flat_idx, img.shape[:d])
    dom = np.stack([axes[a][sub[a]] for a in range(d)], axis=1)
    tgt = img.reshape(-1, C)[flat_idx].astype(np.float64)
    cfg = GraphCfg(dim_domain=d, num_channels=C, start_pis=K, **cfgkw)
    tp = {k: torch.tensor(np.asarray(v, np.float64)) for k, v in params.items()}
    fd = None if feed is None else {k: torch.tensor(np.asarray(v, np.float64)) for k, v in feed.items()}
    return graph_forward(tp, np.ones(K, bool), torch.tensor(dom), torch.tensor(tgt), cfg, feed=fd)


def _check_sampled_forward(pre_gpu, out, tol_frac=0.97, tol=1e-5):
    thr = (out["w_full"] / TAU - 1).abs().min(dim=0).values.numpy()
    ok = thr > 1e-3
    assert ok.mean() > tol_frac
    # the reconstruction is the clipped mixture (smoe.py:857); compared before the output rounding (D4)
    ref = np.clip(out["r_pre"].detach().numpy(), 0, 1)
    got = np.clip(pre_gpu, 0, 1)
    assert np.abs(got[ok] - ref[ok]).max() <= tol
    assert np.abs(got - ref).max() <= 2 * TAU * max(1.0, float(np.abs(out["r_pre"].detach().numpy()).max()))


@pytest.mark.parametrize("workload", ["c3", "c4s"])
def test_full_size_culling_is_exact_and_forward_matches_oracle(workload):
    """BASELINE config 3 (1080p RGB, 32,768 kernels) and config 4 at 1/8 scale (video, 3x3 A): one training
    step with exact culling vs exact-zero skipping only -- gradients, updated parameters, kernel lists and the
    reconstruction must agree BITWISE; the reconstruction is checked against the float64 oracle at a random
    sample of pixels (a pixel's output depends on no other pixel)."""
    import bench
    shape, kgrid, seed, _ = bench.WORKLOADS[workload]
    img = bench.synth_image(shape, seed)
    res = []
    for mode in (0, 2, 1):          # exact culling | exact-zero skipping only | EVERY pair executed
        m = _mk(img, kgrid, dense_exec=mode, **bench.SMOE_KW)
        m._enable_res_pre()
        p0 = m.get_params() if mode == 0 else None
        loss = m.run_batched(pis_l1=0.1, train=True, update_reconstruction=True)
        res.append((m._grads.cpu().numpy(), m._theta.cpu().numpy(), m._klist.cpu().numpy(), m._d_res_pre.cpu().numpy(),
                    m._d_argmax.cpu().numpy(), loss))
        if mode == 0:
            params0 = p0
        del m
        torch.cuda.empty_cache()
    for other in (1, 2):
        for q in range(5):
            np.testing.assert_array_equal(res[0][q], res[other][q])
        assert res[0][5] == res[other][5]
    assert np.isfinite(res[0][0]).all() and np.abs(res[0][0]).max() > 0
    rs = np.random.RandomState(1)
    idx = rs.choice(int(np.prod(shape[:-1])), 160, replace=False)
    out = _oracle_at_pixels(img, params0, idx, dict(use_determinant=True, train_inverse_cov=False, use_yuv=False))
    _check_sampled_forward(res[0][3][idx], out)


def _oracle_kernel_rows(img, params, kernel_ids, cfgkw, rec_gpu, nsig=5.5):
    """float64 closed-form gradient rows of a SAMPLE of kernels of a full-size model.

    A kernel's gradient is a sum over pixels of terms that vanish like its gate, so it is evaluated on the crop of
    +-nsig sigma around the kernel's centre; the crop is cut into small chunks and each chunk sees every kernel that
    can reach one of its pixels (the normaliser S needs them), so the cost per sampled kernel is a few million pair
    evaluations whatever the model size.  The loss mean runs over ALL pixels of the image (loss_count), and the
    rounding decisions are the GPU's (resq_override) as in _train_pass_both."""
    from oracle.graph import GraphCfg, closed_form_grads
    d, C = img.ndim - 1, img.shape[-1]
    shape = img.shape[:d]
    ntot = int(np.prod(shape))
    K = params["pis"].shape[0]
    axes = [np.linspace(0, 1, n).astype(np.float32).astype(np.float64) for n in shape]
    mu_px = params["musX"].astype(np.float64) * (np.array(shape) - 1)
    Ad = params["A_diagonal"].astype(np.float64)
    sig = np.array([(shape[a] - 1) / np.median(np.abs(Ad[:, a, a])) for a in range(d)])      # pixels
    R = np.maximum(np.ceil(nsig * sig), 2).astype(int)
    chunk = np.maximum(np.ceil(2 * sig), 4).astype(int)
    cfg = GraphCfg(dim_domain=d, num_channels=C, start_pis=K, **cfgkw)
    rows = {k: [] for k in PARAM_KEYS}
    for kid in kernel_ids:
        c = np.round(mu_px[kid]).astype(int)
        lo = np.maximum(c - R, 0)
        hi = np.minimum(c + R + 1, np.array(shape))
        acc = {k: 0.0 for k in PARAM_KEYS}
        starts = [range(lo[a], hi[a], chunk[a]) for a in range(d)]
        import itertools
        for org in itertools.product(*starts):
            end = [min(org[a] + chunk[a], hi[a]) for a in range(d)]
            sl = tuple(slice(org[a], end[a]) for a in range(d))
            mesh = np.stack(np.meshgrid(*[axes[a][sl[a]] for a in range(d)], indexing="ij"), axis=-1).reshape(-1, d)
            cc = np.array([(org[a] + end[a] - 1) / 2 for a in range(d)])
            half = np.array([(end[a] - org[a]) / 2 for a in range(d)])
            sel = np.all(np.abs(mu_px - cc) <= half + R + 1, axis=1)
            sub = np.nonzero(sel)[0]
            j = int(np.nonzero(sub == kid)[0][0])
            g, _ = closed_form_grads({k: params[k][sub] for k in PARAM_KEYS}, np.ones(sub.size, bool), mesh,
                                     img[sl].reshape(-1, C), cfg, resq_override=rec_gpu[sl].reshape(-1, C),
                                     loss_count=ntot)
            for k in PARAM_KEYS:
                acc[k] = acc[k] + g[k][j]
        for k in PARAM_KEYS:
            rows[k].append(acc[k])
    return {k: np.stack(v) for k, v in rows.items()}


@pytest.mark.parametrize("workload", ["c2", "c3", "c4s"])
def test_full_size_sampled_kernel_gradients_match_oracle(workload):
    """BASELINE configs 2, 3 and 4 (1/8 scale) at FULL size: one training pass on the GPU, then the gradient rows of a
    random sample of kernels against the float64 closed form evaluated on each kernel's reach (VERDICT r1, next #1a).
    Bar: 1e-4 of the tensor's max-norm over the sample (BASELINE.json: 'parameter gradients within 1e-4')."""
    import bench
    shape, kgrid, seed, _ = bench.WORKLOADS[workload]
    img = bench.synth_image(shape, seed)
    m = _mk(img, kgrid, **bench.SMOE_KW)
    p0 = m.get_params()
    m.run_batched(train=True, update_reconstruction=True)
    g = m.get_gradients()
    rec = m.get_reconstruction()
    rs = np.random.RandomState(11)
    n = {"c2": 40, "c3": 40, "c4s": 6}[workload]
    ids = rs.choice(m.start_pis, n, replace=False)
    ref = _oracle_kernel_rows(img, p0, ids, dict(use_determinant=True, train_inverse_cov=False, use_yuv=False), rec)
    for k in PARAM_KEYS:
        got = g[k][ids].astype(np.float64)
        scale = max(np.abs(ref[k]).max(), 1e-30)
        if k == "A_corr":            # zero at the initial (diagonal) state except through the data term
            scale = max(scale, np.abs(ref["A_diagonal"]).max() * 1e-3)
        err = np.abs(got - ref[k]).max() / scale
        assert err < 1e-4, (k, err)


def test_full_size_config4_culling_is_exact_and_forward_matches_oracle():
    """BASELINE config 4 at FULL size (1280x720x32 RGB video, 32x64x32 = 65,536 kernels, 3x3 steering): one training
    step with exact culling vs exact-zero skipping only must agree bitwise, and the reconstruction matches the float64
    oracle at a sample of pixels."""
    import bench
    shape, kgrid, seed, _ = bench.WORKLOADS["c4"]
    img = bench.synth_image(shape, seed)
    res = []
    params0 = None
    for mode in (0, 2):
        m = _mk(img, kgrid, dense_exec=mode, **bench.SMOE_KW)
        m._enable_res_pre()
        if mode == 0:
            params0 = m.get_params()
        loss = m.run_batched(pis_l1=0.1, train=True, update_reconstruction=False)
        res.append((m._grads.cpu().numpy(), m._theta.cpu().numpy(), m._klist.cpu().numpy(), m._d_res_pre.cpu().numpy(),
                    loss))
        del m
        torch.cuda.empty_cache()
    for q in range(4):
        np.testing.assert_array_equal(res[0][q], res[1][q])
    assert res[0][4] == res[1][4]
    assert np.isfinite(res[0][0]).all() and np.abs(res[0][0]).max() > 0
    rs = np.random.RandomState(4)
    idx = rs.choice(int(np.prod(shape[:-1])), 100, replace=False)
    out = _oracle_at_pixels(img, params0, idx, dict(use_determinant=True, train_inverse_cov=False, use_yuv=False))
    _check_sampled_forward(res[0][3][idx], out)


@pytest.mark.parametrize("workload", ["c2", "c4s"])
def test_epsilon_culling_is_opt_in_and_stays_inside_the_parity_bars(workload):
    """eps_bits = 48 (OPT-IN; the default is exact): terms below 2^-48 of a pixel's normaliser are dropped.  Against
    the exact mode on the same model: pre-quantisation reconstruction within 1e-6, gradients within 1e-5 of the tensor
    max-norm (the parity bars are 1e-5 / 1e-4), identical pruned index sets, over several passes (the forward's bound
    is the previous pass's tile minimum, re-checked after every sweep)."""
    import bench
    shape, kgrid, seed, _ = bench.WORKLOADS[workload]
    img = bench.synth_image(shape, seed)
    me = _mk(img, kgrid, eps_bits=48, **bench.SMOE_KW)
    mx = _mk(img, kgrid, **bench.SMOE_KW)
    me._enable_res_pre()
    mx._enable_res_pre()
    for it in range(3):
        le = me.run_batched(pis_l1=0.1, train=True, update_reconstruction=True)
        lx = mx.run_batched(pis_l1=0.1, train=True, update_reconstruction=True)
        assert abs(le[0] - lx[0]) <= 1e-6 * max(1.0, abs(lx[0])) and le[2] == lx[2]
        assert np.abs(me._d_res_pre.cpu().numpy() - mx._d_res_pre.cpu().numpy()).max() <= 1e-6
        ge, gx = me.get_gradients(), mx.get_gradients()
        for k in PARAM_KEYS:
            assert _rel(ge[k], gx[k]) < 1e-5, (it, k)
        np.testing.assert_array_equal(me.get_active_indices(), mx.get_active_indices())
        # keep both on the exact trajectory -- raw device copies, so that both models also keep the SAME packing order
        # (set_params would re-derive the Hilbert order of one of them and with it the summation order)
        me._theta.copy_(mx._theta)
        me._adam_m.copy_(mx._adam_m)
        me._adam_v.copy_(mx._adam_v)
        me._klist.copy_(mx._klist)
    with pytest.raises(ValueError):
        _mk(img[:32, :32], [4, 4] + ([2] if img.ndim == 4 else []), eps_bits=8, **bench.SMOE_KW)


def test_full_size_decoder_config5():
    """BASELINE config 5: 3840x2160 RGB decoded reconstruction (smoe_reconstruction_decoded.py), forward only,
    ~64.8k surviving kernels of a 540x960 grid, with GPU PSNR / SSIM against a synthetic 4K target."""
    from smoe_b200 import smoe_reconstruction_decoded as dec
    from smoe_b200.ops.image_ops_impl import mse_gpu, smoe_ssim
    import bench
    H, W, C = 2160, 3840, 3
    cp = _decoded_dict(H, W, C, seed=1005)
    smoe, rec, loss, mse = dec.main(cp=cp, write=False)
    assert rec.shape == (H, W, C) and rec.min() >= 0 and rec.max() <= 1
    assert np.abs(rec * 255 - np.round(rec * 255)).max() < 1e-3
    smoe._enable_res_pre()
    smoe.run_batched(train=False, update_reconstruction=True, with_quantized_params=True)
    pre = smoe._d_res_pre.cpu().numpy()
    rs = np.random.RandomState(2)
    idx = rs.choice(H * W, 120, replace=False)
    rp = smoe.rparams
    K = rp["pis"].shape[0]
    dummy = {"pis": np.ones(K), "musX": rp["musX"], "A_diagonal": rp["A"], "A_corr": rp["A"] * 0, "gamma_e": rp["gamma_e"],
             "nu_e": rp["nu_e"]}
    out = _oracle_at_pixels(np.zeros((H, W, C), np.float32), dummy, idx,
                            dict(use_determinant=True, train_inverse_cov=False, use_yuv=True), feed=rp)
    # The SURVEY 8d recipe for config 5 draws A up to 1.5 * 960 (sigma ~ 2.7 px of 3840) and slopes ~ N(0,1):
    # logits carry terms of magnitude ~50-100 whose float32 rounding alone is ~1e-5 in the gate (a float32
    # evaluation on absolute coordinates, as TensorFlow's, is ~10x worse), so the bar here is 5e-5.
    _check_sampled_forward(pre[idx], out, tol_frac=0.9, tol=5e-5)
    # GPU metrics of north_star item 4 against a synthetic 4K frame
    target = bench.synth_image((H, W, C), 1005)
    m = mse_gpu(rec, target)
    assert abs(m / float(((rec.astype(np.float64) - target) ** 2).mean()) - 1) < 1e-7
    s, per = smoe_ssim(rec, target, use_yuv=True)
    assert -1 <= s <= 1 and per.shape == (3,)


def test_train_loop_cadence_histories_and_best_params_match_oracle():
    """Smoe.train (smoe.py:1485-1603): initial evaluation, validation / kernel-list cadence, quantisation
    mode 1 side path, best-parameter shadow copy, history lists."""
    from oracle.model import OracleAdam, OracleSmoe
    from smoe_b200 import AdamOptimizer
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["rgb_image"][:40, :48]
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, quantization_mode=1,
              bit_depths=[20, 18, 6, 10, 10])
    m = _mk(img, [5, 6], **kw)
    o = OracleSmoe(img, kernels_per_dim=[5, 6], dtype=torch.float64, **kw)
    seen = []
    m.train(12, val_iter=4, ukl_iter=3, pis_l1=0.05, callbacks=[lambda s: seen.append(s.iter)])
    o.train(12, val_iter=4, ukl_iter=3, optimizer1=OracleAdam(1e-3), optimizer2=OracleAdam(1e-5), optimizer3=OracleAdam(1.0),
            pis_l1=0.05)
    assert seen == [0, 4, 8, 12] and m.iter == 12
    assert [i for i, _ in m.get_losses()] == [i for i, _ in o.losses] == [0, 4, 8, 12]
    for (_, a), (_, b) in zip(m.get_losses(), o.losses):
        assert abs(a - b) < 5e-5 * max(1.0, abs(b))
    for (_, a), (_, b) in zip(m.get_mses(), o.mses):
        assert abs(10 * np.log10(a / b)) < 0.05            # PSNR within 0.05 dB at every validation
    assert [n for _, n in m.get_num_pis()] == [n for _, n in o.num_pis]
    assert abs(m.get_best_loss() - o.best_loss) < 5e-5 and len(m.get_qlosses()) == 4
    bp, ob = m.get_best_params(), {k: v.numpy() for k, v in o.best.items()}
    for k in ("musX", "nu_e", "pis"):
        assert np.abs(bp[k] - ob[k]).max() < 2e-4 * max(1.0, np.abs(ob[k]).max()), k
    assert m.qparams is not None and m.rparams is not None and m.get_qreconstruction().shape == img.shape


@pytest.mark.parametrize("flags", [dict(train_pis=False), dict(train_musx=False), dict(train_gammas=False),
                                   dict(grad_clip=1e-4), dict(precision=10), dict(only_y_gamma=True)])
def test_trainable_flags_clip_and_precision(flags):
    """One Adam step under the constructor's trainability flags (smoe.py:389-396, 1112-1117), gradient clipping
    (smoe.py:1152-1153), a 10-bit output precision and only_y_gamma (smoe.py:725-729), against the oracle."""
    from oracle.model import OracleAdam, OracleSmoe
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["rgb_image"][:34, :38]      # ragged vs the 16x32 tiles
    clip = flags.pop("grad_clip", None)
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, **flags)
    m = _mk(img, [4, 5], **kw)
    o = OracleSmoe(img, kernels_per_dim=[4, 5], dtype=torch.float64, **kw)
    m.set_optimizer(m.optimizer1, m.optimizer2, m.optimizer3, grad_clip_value_abs=clip)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0), grad_clip_value_abs=clip)
    rs = np.random.RandomState(8)
    p = m.get_params()
    p["gamma_e"] = rs.normal(0, 0.2, p["gamma_e"].shape).astype(np.float32)
    m.set_params({"gamma_e": p["gamma_e"]})
    o.vars["gamma_e"] = torch.tensor(p["gamma_e"].astype(np.float64))
    p0 = m.get_params()
    a, b = _train_pass_both(m, o)
    assert abs(a[0] - b[0]) < 1e-6
    pg, po = m.get_params(), o.get_params()
    for k in PARAM_KEYS:
        assert np.abs(pg[k] - po[k]).max() <= 3e-6 * max(1.0, np.abs(po[k]).max()) + 2e-3 * (k in ("A_diagonal", "A_corr")), k
    if flags.get("train_pis") is False:
        np.testing.assert_array_equal(pg["pis"], p0["pis"])
    if flags.get("train_musx") is False:
        np.testing.assert_array_equal(pg["musX"], p0["musX"])
    if flags.get("train_gammas") is False:
        np.testing.assert_array_equal(pg["gamma_e"], p0["gamma_e"])


def test_degenerate_inputs():
    """No active kernel at all, a single-row image, kernels that influence nothing."""
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["c1_image"][:20, :24]
    m = _mk(img, [3, 3], use_determinant=False, train_inverse_cov=False, use_yuv=False)
    m.set_params({"pis": -np.ones(9, np.float32)})
    loss, mse, num_pi, _ = m.run_batched(train=True, update_reconstruction=True)
    assert num_pi == 0 and np.isfinite(loss) and np.all(m.get_reconstruction() == 0)
    assert not np.any(m.kernel_list_per_batch[0]) and np.all(m._grads.cpu().numpy() == 0)
    expect = float(np.mean((np.abs(img.astype(np.float64)) - 0.5 / 256) ** 2))
    assert abs(loss - expect) < 1e-7
    row = np.load(os.path.join(GOLDEN, "init_cases.npz"))["c1_image"][:1, :50]
    one = _mk(row, [1, 5], use_determinant=True, train_inverse_cov=False, use_yuv=False)
    l1, _, n1, _ = one.run_batched(train=True, update_reconstruction=True)
    assert n1 == 5 and np.isfinite(l1) and one.get_reconstruction().shape == (1, 50, 1)
    far = _mk(img, [3, 3], use_determinant=False, train_inverse_cov=False, use_yuv=False)
    p = far.get_params()
    p["musX"][4] = (7.0, -3.0)                       # a kernel far outside the image influences no pixel
    far.set_params({"musX": p["musX"]})
    far.run_batched(train=True)
    kl = far.kernel_list_per_batch[0]
    assert not kl[4] and kl.sum() == 8


def test_training_cli_driver(tmp_path):
    """The minimal `smoe_test.py` driver: grid model, three Adam optimizers, pi-sparsification, pickles."""
    from smoe_b200 import smoe_test, load_params
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["c1_image"][:64, :64]
    np.save(str(tmp_path / "img.npy"), np.round(img * 255).astype(np.uint8))
    m = smoe_test.main(str(tmp_path / "img.npy"), str(tmp_path / "out"), 30, 10, [8, 8], None, 1.0, 1e-3, 1, 100.0, 1000.0,
                       True, True, 0, [20, 18, 6, 10, 10], False, [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32], False,
                       False, None)
    assert [i for i, _ in m.get_losses()] == [0, 10, 20, 30]
    assert m.get_losses()[-1][1] < m.get_losses()[0][1]
    p = load_params(str(tmp_path / "out" / "params_last.pkl"))
    assert p["pis"].shape[0] == m.get_num_pis()[-1][1] and (p["pis"] > 0).all()
    assert os.path.exists(str(tmp_path / "out" / "reconstruction.png"))


def test_pi_sparsification_prunes_and_index_sets_follow_pis():
    """BASELINE config 2 path at reduced size: training with an L1 penalty on pi (smoe.py:1027) drives pis
    through 0 and the kernels drop out of the compaction (smoe.py:480, 738-753).  After every step the active
    index set must equal `kernel_list & (pis > 0)` evaluated on the parameters the step started from, and the
    pruning trajectory must track the float64 oracle."""
    from oracle.model import OracleAdam, OracleSmoe
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["c1_image"][:64, :64]
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False, normalize_pis=True)
    m = _mk(img, [16, 16], **kw)
    o = OracleSmoe(img, kernels_per_dim=[16, 16], dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    counts_g, counts_o = [], []
    for it in range(450):
        pis_before = m.get_params()["pis"] if it % 50 == 0 else None
        kl_before = m.kernel_list_per_batch[0] if it % 50 == 0 else None
        counts_g.append(m.run_batched(pis_l1=10.0, train=True)[2])
        if pis_before is not None:
            K = int(m._counts[0, 0].item())
            np.testing.assert_array_equal(m.get_active_indices(), np.nonzero(kl_before & (pis_before > 0))[0])
            assert counts_g[-1] == int((pis_before > 0).sum())
        counts_o.append(o.run_batched(pis_l1=10.0, train=True)[2])
        if it == 379:
            # just before the first pi reaches 0 (pi = 1/256 shrinks by ~lr2 = 1e-5 per step): the pis themselves
            # -- a continuous quantity -- must track the float64 oracle to a few Adam steps
            dp = np.abs(m.get_params()["pis"] - o.get_params()["pis"])
            assert counts_g[-1] == counts_o[-1] == 256 and np.quantile(dp, 0.99) < 3e-5 and dp.max() < 1e-4, dp.max()
    assert counts_g[0] == 256 and counts_g[-1] <= 0.7 * 256          # >= 30 % pruned
    # Same trajectory.  Most pis cross 0 within a few iterations of each other (iterations ~390-420), so the COUNT
    # is steep there and a crossing that happens two iterations earlier or later moves it by tens; the float32
    # and float64 runs of the oracle itself differ by up to 12 kernels on this case.
    assert np.abs(np.array(counts_g) - np.array(counts_o)).max() <= 40 and abs(counts_g[-1] - counts_o[-1]) <= 26
    assert np.isfinite(m.run_batched(train=False)[0])


@pytest.mark.parametrize("case", ["img_yuv", "video"])
def test_fake_quant_training_mode2_diff_center_and_kernel_count_norm(case):
    """quantization_mode 2 (smoe.py:482-496): every variable is fake-quantised with fixed bounds before use,
    gradients pass straight through inside the bounds; use_diff_center (smoe.py:390-394, 746-747): the musX
    variable holds offsets from the fixed grid; kernel_count_as_norm_l1 (smoe.py:1022-1025)."""
    from oracle.model import OracleAdam, OracleSmoe
    from oracle.graph import _nudge
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    img, k = (z["rgb_image"], [6, 8]) if case == "img_yuv" else (z["vid_image"], [3, 4, 2])
    lb, ub, bd = [-40.0, -0.3, 0.1, 0.0, -2.0], [40.0, 1.3, 0.9, 2.0, 2.0], [12, 12, 7, 10, 8]
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=(case == "img_yuv"), normalize_pis=False,
              quantization_mode=2, lower_bounds=lb, upper_bounds=ub, bit_depths=bd, use_diff_center=True,
              kernel_count_as_norm_l1=True)
    m = _mk(img, k, **kw)
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    K, d, C = m.start_pis, m.dim_domain, img.shape[-1]
    rs = np.random.RandomState(11)
    pert = {"musX": rs.uniform(-0.02, 0.02, (K, d)), "pis": rs.uniform(0.3, 1.7, K),
            "gamma_e": rs.normal(0, 0.3, (K, d, C)), "nu_e": o.vars["nu_e"].numpy() + rs.normal(0, 0.05, (K, C)),
            "A_corr": np.tril(rs.normal(0, 2.0, (K, d, d)), -1)}
    pert["pis"][[2, 9]] = -0.2                                  # pruned: fake-quant clamps to 0, mask is qpis > 0
    pert["nu_e"][0], pert["nu_e"][1] = 0.02, 0.97               # outside [lb, ub]: clamped, no gradient
    pert = {kk: v.astype(np.float32) for kk, v in pert.items()}
    m.set_params(pert)
    for kk, v in pert.items():
        o.vars[kk] = torch.tensor(v.astype(np.float64))
    # get_params returns the fake-quantised tensors (smoe.py:1796-1798), bit for bit
    pg, po = m.get_params(), o.get_params()
    for kk in PARAM_KEYS:
        np.testing.assert_array_equal(pg[kk], po[kk], err_msg=kk)
    nmin, nmax, _ = _nudge(lb[2], ub[2], bd[2])
    outside = (pert["nu_e"] < nmin) | (pert["nu_e"] > nmax)
    assert outside.any()
    (lg, mg, npg, _), (lo, mo, npo, _) = _train_pass_both(m, o, pis_l1=0.3, u_l1=1e-5)
    assert npg == npo == K - 2
    assert abs(lg - lo) < 2e-6 * max(1.0, abs(lo)) and abs(mg - mo) < 2e-3 * mo + 1e-3
    g = m.get_gradients()
    for kk, ref in o.last_grads.items():
        assert _rel(g[kk], ref.numpy()) < 1e-4, kk
    assert np.abs(g["nu_e"][outside]).max() == 0               # straight-through mask outside the bounds
    assert np.abs(g["pis"][[2, 9]]).max() == 0
    # 4 more iterations: the quantised parameters follow the oracle up to a few code flips.  (Not more: with
    # the reference's lr of 1.0 on A the quantised trajectory is chaotic -- one flipped code or one kernel
    # dropping off the influence list at the gate threshold and float32 / float64 runs part ways; on this
    # case they stay together for 13 iterations, on the video case for fewer than 8.)
    for _ in range(4):
        m.run_batched(train=True, pis_l1=0.3, u_l1=1e-5)
        o.run_batched(train=True, pis_l1=0.3, u_l1=1e-5)
    pg, po = m.get_params(), o.get_params()
    for kk, grp in (("pis", 3), ("musX", 1), ("A_diagonal", 0), ("A_corr", 0), ("gamma_e", 4), ("nu_e", 2)):
        step = (ub[grp] - lb[grp]) / (2 ** bd[grp] - 1)
        dcode = np.abs(pg[kk] - po[kk]) / step
        assert dcode.max() < 4 and (dcode > 0.5).mean() < 0.05, (kk, dcode.max(), (dcode > 0.5).mean())
    assert abs(m.psnr() - o_psnr(o)) < 0.05


def o_psnr(o):
    _, mse, _, _ = o.run_batched(train=False, update_reconstruction=True)
    return 10 * np.log10((2 ** o.precision) ** 2 / mse)


@pytest.mark.parametrize("case", ["img", "video"])
def test_loss_mask_weights_the_pixel_loss(case):
    """loss_mask / use_loss_mask (smoe.py:550, 932, 1674-1677): per-pixel weights on the loss term only --
    mse and the reconstruction do not change, gradients of zero-weight pixels vanish."""
    from oracle.model import OracleAdam, OracleSmoe
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    img, k = (z["rgb_image"], [6, 8]) if case == "img" else (z["vid_image"], [3, 4, 2])
    rs = np.random.RandomState(3)
    mask = rs.uniform(0, 2, img.shape[:-1]).astype(np.float32)
    mask[rs.uniform(size=mask.shape) < 0.3] = 0.0
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=(case == "img"), loss_mask=mask,
              start_batches=4 if case == "img" else 1)
    m = _mk(img, k, **kw)
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    m2 = _mk(img, k, **{kk: v for kk, v in kw.items() if kk != "loss_mask"})
    l0, mse0, _, _ = m2.run_batched(train=False, pis_l1=0.1)
    (lg, mg, _, _), (lo, mo, _, _) = _train_pass_both(m, o, pis_l1=0.1, use_loss_mask=True)
    assert abs(lg - lo) < 2e-6 * max(1.0, abs(lo)) and abs(mg - mo) < 2e-3 * mo + 1e-3
    assert abs(mg - mse0) < 1e-6 * mse0 and abs(lg - l0) > 1e-4 * l0          # weights touch the loss, not the mse
    g = m.get_gradients()
    for kk, ref in o.last_grads.items():
        assert _rel(g[kk], ref.numpy()) < 1e-4, kk
    # the train loop forwards use_loss_mask like the reference (smoe.py:1507-1558) and replays it as a CUDA graph
    for _ in range(3):
        a = m.run_batched(train=True, pis_l1=0.1, use_loss_mask=True)
        b = o.run_batched(train=True, pis_l1=0.1, use_loss_mask=True)
        assert abs(a[0] - b[0]) < 1e-5 * max(1.0, abs(b[0]))
    # without a mask object the flag is an error, as feeding a missing mask is in the reference
    with pytest.raises(ValueError):
        m2.run_batched(train=False, use_loss_mask=True)


def test_random_sampling_of_training_pixels():
    """sampling_percentage < 100 (smoe.py:1664-1667): a training pass feeds round(N_b * pct / 100) pixels per batch,
    drawn by np.random.choice with the error-proportional probabilities of the last reconstruction pass
    (smoe.py:906-907, 1768-1769).  With the same NumPy seed both implementations draw the same pixels."""
    from oracle.model import OracleAdam, OracleSmoe
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    img, k = z["rgb_image"], [6, 8]
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, start_batches=4)
    m = _mk(img, k, **kw)
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    # before any reconstruction pass the probabilities are uniform (smoe.py:271-273)
    np.random.seed(5)
    rg = m.run_batched(train=True, sampling_percentage=30, pis_l1=0.1)
    np.random.seed(5)
    ro = o.run_batched(train=True, sampling_percentage=30, pis_l1=0.1)
    nb = int(np.prod(m.batch_size_valued))
    assert all(len(s) == round(nb * 0.3) for s in m.last_samples)
    np.testing.assert_array_equal(m.last_samples[-1], o.last_samples)
    assert abs(rg[0] - ro[0]) < 2e-6 * max(1.0, abs(ro[0])) and abs(rg[1] - ro[1]) < 2e-3 * ro[1] + 1e-3
    g = m.get_gradients()
    for kk, ref in o.last_grads.items():
        assert _rel(g[kk], ref.numpy()) < 2e-3, kk          # no resq_override here: a few rounding flips allowed
    for a, b in zip(m.kernel_list_per_batch, o.kernel_list_per_batch):
        assert (a != b).sum() <= 1                            # influence lists come from the sampled pixels only
    # a reconstruction pass installs err_map / sum(err_map) per batch
    m.run_batched(train=False, update_reconstruction=True)
    o.run_batched(train=False, update_reconstruction=True)
    for pg, po in zip(m.random_sampling_per_batch, o.random_sampling_per_batch):
        pg = pg.cpu().numpy()
        assert abs(pg.sum() - 1) < 1e-4
        flips = np.abs(pg - po) > 1e-3 * np.maximum(po, po.mean())
        assert flips.mean() < 5e-3                            # pixels whose output code flipped
    # statistical property: pixels with larger error are drawn more often
    np.random.seed(6)
    m.run_batched(train=True, sampling_percentage=20)
    p0 = m.random_sampling_per_batch[0].cpu().numpy()
    drawn = np.zeros(nb, bool)
    drawn[m.last_samples[0]] = True
    assert p0[drawn].mean() > 1.1 * p0[~drawn].mean()
    # full training with sampling still converges
    l_start = m.run_batched(train=False)[0]
    m.train(30, val_iter=10, sampling_percentage=50, pis_l1=0.0)
    assert m.run_batched(train=False)[0] < l_start


@pytest.mark.parametrize("case", ["mse_overlap", "ssim", "ssim_overlap", "ssim_video", "ssim_gray_yuv"])
def test_ssim_loss_and_overlap_of_batches(case):
    """ssim_opt (smoe.py:981-1010): loss = 1 - weighted custom_ssim of the SYMMETRIC-padded batch, gradient through
    the SSIM map, the output fake-quant and the clip.  overlap_of_batches (smoe.py:18-35, 909-923, 985-991): every
    window is forwarded with a halo that is cropped before the loss; halo pixels (and the zero padding at
    coordinate 0 of border windows) only reach the influence lists.
    Gradient tolerance with SSIM is 1e-3: sigma^2 = E[x^2] - mu^2 cancels in float32 (as it does in TF)."""
    from oracle.model import OracleAdam, OracleSmoe
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    img, k = (z["vid_image"], [3, 4, 2]) if case == "ssim_video" else (z["rgb_image"], [6, 8])
    if case == "ssim_gray_yuv":
        img = np.ascontiguousarray(img[..., :1])
    ssim = case.startswith("ssim")
    ov = {"mse_overlap": 3, "ssim_overlap": 2}.get(case, 0)
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=case != "ssim_video", ssim_opt=ssim,
              overlap_of_batches=ov, start_batches=1 if case == "ssim_video" else 4)
    m = _mk(img, k, **kw)
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    (lg, mg, npg, _), (lo, mo, npo, _) = _train_pass_both(m, o, pis_l1=0.2, u_l1=1e-6)
    assert npg == npo
    assert abs(lg - lo) < (2e-5 if ssim else 2e-6) * max(1.0, abs(lo)) and abs(mg - mo) < 2e-3 * mo + 1e-3
    if ssim:
        assert 0 < lg < 1.2
    g = m.get_gradients()
    for kk, ref in o.last_grads.items():
        assert _rel(g[kk], ref.numpy()) < (1e-3 if ssim else 1e-4), (kk, _rel(g[kk], ref.numpy()))
    for a, b in zip(m.kernel_list_per_batch, o.kernel_list_per_batch):
        assert (a != b).sum() <= 1
    if ov:
        # the halo only widens the influence lists
        m0 = _mk(img, k, **dict(kw, overlap_of_batches=0))
        m0.run_batched(train=True, update_reconstruction=True, pis_l1=0.2, u_l1=1e-6)
        assert all(((a | b) == a).all() for a, b in zip(m.kernel_list_per_batch, m0.kernel_list_per_batch))
        assert sum(int(a.sum()) - int(b.sum()) for a, b in zip(m.kernel_list_per_batch, m0.kernel_list_per_batch)) > 0
        if not ssim:
            np.testing.assert_array_equal(m.get_reconstruction(), m0.get_reconstruction())
        m.update_kernel_list()
        o.update_kernel_list()
        for a, b in zip(m.kernel_list_per_batch, o.kernel_list_per_batch):
            assert (a != b).sum() <= 1
    # training (CUDA-graph replay from the second plain step on) follows the oracle and lowers the loss
    first = None
    for _ in range(6):
        a = m.run_batched(train=True, pis_l1=0.2, u_l1=1e-6)
        b = o.run_batched(train=True, pis_l1=0.2, u_l1=1e-6)
        first = first if first is not None else a[0]
        assert abs(a[0] - b[0]) < 5e-4 * max(1.0, abs(b[0]))
    assert a[0] < first
    if ssim:
        # the loss value is the SSIM metric kernel's value on the same reconstruction (1 batch only)
        if m.start_batches == 1:
            l, _, _, _ = m.run_batched(train=False, update_reconstruction=True)
            assert abs((1 - m.ssim()[0]) - l) < 1e-5


@pytest.mark.parametrize("case", ["img_diff_center", "video_abs_centres"])
def test_fake_quant_training_mode3_ranges_of_the_surviving_kernels(case):
    """quantization_mode 3 (smoe.py:497-531): fake_quant_with_min_max_vars with reduce_min / reduce_max over the
    kernels whose quantised pi is positive -- shifted form for A_diagonal / nu_e, plain form (with TF's nudging,
    clamping and gradient routing to the extremes) for A_corr / musX / gamma_e, fixed bounds for pis."""
    from oracle.model import OracleAdam, OracleSmoe
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    img, k = (z["rgb_image"], [6, 8]) if case == "img_diff_center" else (z["vid_image"], [3, 4, 2])
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False, normalize_pis=False,
              quantization_mode=3, lower_bounds=[0, 0, 0, 0.0, 0], upper_bounds=[0, 0, 0, 2.0, 0],
              bit_depths=[8, 9, 6, 10, 5], use_diff_center=(case == "img_diff_center"))
    m = _mk(img, k, **kw)
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    K, d, C = m.start_pis, m.dim_domain, img.shape[-1]
    rs = np.random.RandomState(12)
    pert = {"pis": rs.uniform(0.3, 1.7, K), "gamma_e": rs.normal(0, 0.3, (K, d, C)),
            "nu_e": o.vars["nu_e"].numpy() + rs.normal(0, 0.05, (K, C)),
            "A_corr": np.tril(rs.normal(0, 2.0, (K, d, d)), -1),
            "A_diagonal": o.vars["A_diagonal"].numpy() * (1 + 0.2 * rs.uniform(-1, 1, (K, 1, 1)))}
    if case == "img_diff_center":
        pert["musX"] = rs.uniform(-0.02, 0.02, (K, d))
    pert["pis"][[1, 7]] = -0.3
    pert["nu_e"][1] = 9.0                                   # a pruned kernel must not widen the range
    pert["gamma_e"][7] = -50.0
    pert = {kk: v.astype(np.float32) for kk, v in pert.items()}
    m.set_params(pert)
    for kk, v in pert.items():
        o.vars[kk] = torch.tensor(v.astype(np.float64))
    pg, po = m.get_params(), o.get_params()
    keep = po["pis"] > 0
    assert keep.sum() == K - 2
    for kk in PARAM_KEYS:                                   # bit for bit on the rows the graph reads
        np.testing.assert_array_equal(pg[kk][keep], po[kk][keep], err_msg=kk)
    assert pg["nu_e"][keep].max() < 2 and abs(pg["gamma_e"][keep]).max() < 5
    clamped = sum(int((np.abs(po[kk][keep] - pert[kk][keep]) > 0.5001 * (pert[kk][keep].max() - pert[kk][keep].min())
                       / (2 ** b - 1)).sum()) for kk, b in (("gamma_e", 5),))
    (lg, mg, npg, _), (lo, mo, npo, _) = _train_pass_both(m, o, pis_l1=0.3, u_l1=1e-5)
    assert npg == npo == K - 2
    assert abs(lg - lo) < 2e-6 * max(1.0, abs(lo)) and abs(mg - mo) < 2e-3 * mo + 1e-3
    g = m.get_gradients()
    for kk, ref in o.last_grads.items():
        assert _rel(g[kk], ref.numpy()) < 1e-4, (kk, _rel(g[kk], ref.numpy()), clamped)
    assert np.abs(g["pis"][[1, 7]]).max() == 0 and np.abs(g["nu_e"][[1, 7]]).max() == 0
    for _ in range(3):
        a = m.run_batched(train=True, pis_l1=0.3, u_l1=1e-5)
        b = o.run_batched(train=True, pis_l1=0.3, u_l1=1e-5)
        assert abs(a[0] - b[0]) < 1e-3 * max(1.0, abs(b[0]))
    pg, po = m.get_params(), o.get_params()
    for kk in ("nu_e", "musX", "pis"):
        span = po[kk][keep].max() - po[kk][keep].min() + 1e-9
        assert (np.abs(pg[kk][keep] - po[kk][keep]) > 0.02 * span).mean() < 0.1, kk


@pytest.mark.parametrize("case", ["img", "video_mode2"])
def test_radial_kernels(case):
    """radial_as (smoe.py:429-434, 714-721): one trainable scalar a per kernel, A = a * I, A_corr frozen at 0;
    get_params returns A_diagonal as a (K,) vector; the quantiser leaves A_corr out (quantizer.py:11, 45, 132-136)."""
    from oracle.model import OracleAdam, OracleSmoe
    from smoe_b200 import quantize_params, rescaler
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    img, k = (z["rgb_image"], [6, 8]) if case == "img" else (z["vid_image"], [3, 4, 2])
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False, radial_as=True, normalize_pis=False)
    if case == "video_mode2":
        kw.update(quantization_mode=2, lower_bounds=[-40.0, -0.3, -1.0, 0.0, -2.0], upper_bounds=[40.0, 1.3, 2.0, 2.0, 2.0],
                  bit_depths=[12, 12, 8, 10, 8])
    m = _mk(img, k, **kw)
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    K, d = m.start_pis, m.dim_domain
    pg, po = m.get_params(), o.get_params()
    assert pg["A_diagonal"].shape == (K,) and pg["A_corr"].shape == (K, d, d)
    for kk in PARAM_KEYS:
        np.testing.assert_array_equal(pg[kk], po[kk], err_msg=kk)
    a = (pg["A_diagonal"] * (1 + 0.2 * np.random.RandomState(4).uniform(-1, 1, K))).astype(np.float32)
    m.set_params({"A_diagonal": a})
    o.vars["A_diagonal"] = torch.tensor(a.astype(np.float64))
    (lg, mg, _, _), (lo, mo, _, _) = _train_pass_both(m, o, pis_l1=0.1, u_l1=1e-5)
    assert abs(lg - lo) < 2e-6 * max(1.0, abs(lo)) and abs(mg - mo) < 2e-3 * mo + 1e-3
    g = m.get_gradients()
    assert g["A_diagonal"].shape == (K,)
    for kk, ref in o.last_grads.items():
        assert _rel(g[kk], ref.numpy()) < 1e-4, kk
    assert np.abs(g["A_corr"]).max() == 0
    for _ in range(4):
        x = m.run_batched(train=True, pis_l1=0.1, u_l1=1e-5)
        y = o.run_batched(train=True, pis_l1=0.1, u_l1=1e-5)
        assert abs(x[0] - y[0]) < 1e-3 * max(1.0, abs(y[0]))
    p1 = m.get_params()
    assert np.abs(p1["A_corr"]).max() == 0 and np.abs(p1["A_diagonal"] - a).max() > 0
    if case == "img":
        # quantiser round trip of the radial dict: bit-exact with the reference's own vectors, and usable as rparams
        zq = np.load(os.path.join(GOLDEN, "quant_radial_cases.npz"))

        class Shim:
            pass
        for ci in range(int(zq["num_cases"])):
            pre = f"case{ci}_"
            qm, qp, dd, C, _ = [int(v) for v in zq[pre + "meta"][:5]]
            s = Shim()
            s.quantization_mode, s.quantize_pis, s.radial_as, s.dim_domain = qm, bool(qp), True, dd
            s.image = np.zeros((4,) * dd + (C,), np.float32)
            s.lower_bounds, s.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
            s.bit_depths = [int(v) for v in zq[pre + "meta"][5:]]
            s.use_diff_center, s.musX_init = False, None
            q = quantize_params(s, {kk: zq[pre + "in_" + kk].copy() for kk in PARAM_KEYS})
            r = rescaler(s, q)
            assert "A_corr" not in q
            for kk in ("A_diagonal", "musX", "nu_e", "pis", "gamma_e"):
                np.testing.assert_array_equal(q[kk], zq[pre + "q_" + kk], err_msg=f"case {ci} codes {kk}")
                assert q[kk].dtype == zq[pre + "q_" + kk].dtype
            for kk in ("A", "musX", "nu_e", "pis", "gamma_e"):
                np.testing.assert_array_equal(r[kk], zq[pre + "r_" + kk], err_msg=f"case {ci} rescaled {kk}")
        m.quantization_mode, m.bit_depths = 1, [12, 12, 8, 10, 8]
        m.qparams = quantize_params(m, m.get_params())
        m.rparams = rescaler(m, m.qparams)
        lq, _, _, _ = m.run_batched(train=False, update_reconstruction=True, with_quantized_params=True)
        l0, _, _, _ = m.run_batched(train=False, update_reconstruction=True)
        assert abs(lq - l0) < 0.2 * l0 + 1e-3


@pytest.mark.parametrize("case", ["rgb_whole", "gray_halo", "rgb_region", "rgb_small"])
def test_ssim_tile_kernels_match_the_separable_passes(case, monkeypatch):
    """smoe_ssim_loss on images runs as two shared-memory tile kernels (csrc/ssim_tile.cuh); the separable
    global-memory passes (video path, SMOE_SSIM_GENERIC=1) are the same arithmetic in the same order: identical
    gradient planes, SSIM sums equal up to the order of the block sums.  Sizes with interior and border tiles,
    an overlap halo, and a caller-given compute region (the sharded SSIM of SURVEY.md 8 f-4)."""
    import ctypes as C
    from smoe_b200 import _ffi
    L = _ffi.lib()
    dev = torch.device("cuda:0")
    Hh, Ww, Cc = {"rgb_whole": (150, 203, 3), "gray_halo": (131, 97, 1), "rgb_region": (120, 140, 3),
                  "rgb_small": (37, 21, 3)}[case]
    cfg = _ffi.Cfg()
    cfg.d, cfg.C, cfg.precision, cfg.use_yuv = 2, Cc, 8, 1
    b = _ffi.Batch()
    b.dims[:] = [Hh, Ww, 1]
    b.tile[:] = [16, 32, 1]
    region = None
    if case == "gray_halo":                       # a window in the middle of the image: its halo is cropped on all sides
        b.origin[:] = [16, 32, 0]
        b.extent[:] = [96, 64, 1]
        b.halo = 3
    elif case == "rgb_region":                    # block + ring inside a larger resident buffer
        b.origin[:] = [10, 0, 0]
        b.extent[:] = [80, 96, 1]
        region = _ffi.SsimRegion()
        region.lo[:] = [0, 0, 0]
        region.n[:] = [100, 106, 1]
        region.inv_count = 1.0 / (300 * 400)
    else:
        b.origin[:] = [0, 0, 0]
        b.extent[:] = [Hh, Ww, 1]
    b.inv_count = 1.0 / (b.extent[0] * b.extent[1])
    g = torch.Generator(device="cpu").manual_seed(5)
    img = torch.rand((Hh, Ww, Cc), generator=g)
    pre = img + 0.1 * torch.randn((Hh, Ww, Cc), generator=g)          # some values leave [0, 1]: clip mask
    res = (pre.clamp(0, 1) * 255).round() / 255
    img, pre, res = img.to(dev), pre.to(dev), res.to(dev)
    ntiles = L.smoe_num_tiles(C.byref(b))
    stride = L.smoe_pix_stride(2, Cc, C.byref(b))
    L.smoe_ssim_loss_workspace_bytes.restype = C.c_size_t
    ws = torch.zeros(L.smoe_ssim_loss_workspace_bytes(C.byref(cfg), C.byref(b)) // 4 + 64, dtype=torch.float32, device=dev)
    pix0 = torch.randn((ntiles * stride,), generator=g).to(dev) * 30       # qthr plane: some pixels below log2(1e-11)
    out = {}
    for mode in ("tile", "generic"):
        if mode == "generic":
            monkeypatch.setenv("SMOE_SSIM_GENERIC", "1")
        else:
            monkeypatch.delenv("SMOE_SSIM_GENERIC", raising=False)
        pix = pix0.clone()
        scal = torch.zeros(16, dtype=torch.float32, device=dev)
        _ffi.check(L.smoe_ssim_loss(C.byref(cfg), C.byref(b), C.byref(region) if region is not None else None,
                                    _ffi.ptr(res), _ffi.ptr(img), _ffi.ptr(pre), _ffi.ptr(pix), _ffi.ptr(scal),
                                    _ffi.ptr(ws), _ffi.stream_ptr()), "smoe_ssim_loss")
        torch.cuda.synchronize()
        out[mode] = (pix.cpu().numpy(), scal.cpu().numpy())
    assert np.abs(out["generic"][0] - pix0.cpu().numpy()).max() > 0          # the call wrote gradient planes
    np.testing.assert_array_equal(out["tile"][0], out["generic"][0])
    np.testing.assert_allclose(out["tile"][1][8:8 + Cc], out["generic"][1][8:8 + Cc], rtol=2e-6)
    assert np.all(out["generic"][1][8:8 + Cc] > 0)
