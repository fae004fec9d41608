"""Generates tests/golden/*.npz from the REAL reference code.  Run in the authoring container only:

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  `/root/reference` does not exist on the GPU box, so nothing at test
or bench time imports this module; the vectors it wrote are committed under tests/golden/.

What is executed from the reference (unmodified, imported from /root/reference under stub
modules for tensorflow / matplotlib / skimage / progressbar / hdf5storage, with the NumPy<1.24
aliases `np.int`, `np.bool` restored):
  * quantizer.quantize_params, quantizer.rescaler, utils.reduce_params
  * smoe.Smoe.gen_domain, generate_kernel_grid, generate_experts, generate_pis,
    get_batch_shape and smoe.sliding_window
  * utils.save_model (a checkpoint pickle written by the reference itself: ref_saved_model_*.pkl)
The TensorFlow graph itself cannot be executed here (no TensorFlow, no network); the graph
fixtures (graph_*.npz) are therefore produced by the float64 restatement in oracle/graph.py and
are marked `pinned=False` inside the file.
"""
from __future__ import annotations

import copy
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    if not hasattr(np, "int"):
        np.int = int          # noqa: reference uses the pre-1.24 aliases
    if not hasattr(np, "bool"):
        np.bool = bool
    for name in ["tensorflow", "tensorflow.python", "tensorflow.python.ops", "tensorflow.python.framework",
                 "tensorflow.python.ops.array_ops", "tensorflow.python.ops.math_ops",
                 "tensorflow.python.ops.nn", "tensorflow.python.ops.nn_ops",
                 "tensorflow.python.ops.control_flow_ops", "tensorflow.python.framework.ops",
                 "tensorflow.python.framework.dtypes", "tensorflow.python.framework.constant_op",
                 "tensorflow.python.ops.image_ops_impl",
                 "skimage", "skimage.feature", "skimage.measure", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.gridspec", "progressbar", "hdf5storage", "skvideo", "skvideo.io"]:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    sys.path.insert(0, REF)
    import quantizer as ref_quantizer    # noqa: E402
    import utils as ref_utils            # noqa: E402
    import smoe as ref_smoe              # noqa: E402
    return ref_smoe, ref_quantizer, ref_utils


def synth_image(shape, seed):
    """SURVEY.md 8d synthetic pattern, quantised to uint8 then /255 as utils.py:126-128."""
    rs = np.random.RandomState(seed)
    H, W = shape[0], shape[1]
    v, u = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    C = shape[-1]
    frames = shape[2] if len(shape) == 4 else None

    def base(uu, vv):
        return (0.5 + 0.25 * np.sin(2 * np.pi * (3 * uu + 2 * vv)) + 0.2 * (uu > vv)
                + 0.15 * np.exp(-((uu - .3) ** 2 + (vv - .6) ** 2) / 0.02))
    shifts = [0.0, 0.11, 0.23]
    if frames is None:
        img = np.stack([base(u + shifts[c], v) for c in range(C)], axis=-1)
    else:
        ts = np.linspace(0, 1, frames)
        img = np.stack([np.stack([base(u + shifts[c] + 0.1 * t, v) for c in range(C)], axis=-1) for t in ts], axis=2)
    img = np.clip(img + 0.03 * rs.standard_normal(img.shape), 0, 1)
    return (np.round(img * 255).astype(np.uint8).astype(np.float32) / 255.).astype(np.float32)


class _Shim:
    """Attribute bag standing in for a constructed Smoe (the ctor itself needs TensorFlow)."""


def golden_init(ref_smoe):
    Smoe = ref_smoe.Smoe
    cases = {
        "c1": dict(shape=(128, 128, 1), k=[16, 16], seed=1001, tic=False, norm=True),
        "rgb": dict(shape=(48, 64, 3), k=[6, 8], seed=7, tic=True, norm=False),
        "vid": dict(shape=(24, 32, 8, 3), k=[3, 4, 2], seed=9, tic=False, norm=True),
        "one": dict(shape=(40, 40, 1), k=[5], seed=11, tic=False, norm=True),
    }
    out = {}
    for name, c in cases.items():
        img = synth_image(c["shape"], c["seed"])
        s = object.__new__(Smoe)
        s.image = img
        s.dim_domain = img.ndim - 1
        s.train_inverse_cov = c["tic"]
        s.musX_init = s.A_init = None
        s.generate_kernel_grid(c["k"])
        s.generate_experts()
        s.generate_pis(c["norm"])
        jd = Smoe.gen_domain(img, s.dim_domain)
        out[f"{name}_image"] = img
        out[f"{name}_k"] = np.array(c["k"])
        out[f"{name}_tic"] = np.array(c["tic"])
        out[f"{name}_norm"] = np.array(c["norm"])
        out[f"{name}_joint_domain"] = jd
        out[f"{name}_musX"] = s.musX_init
        out[f"{name}_A"] = s.A_init
        out[f"{name}_nu_e"] = s.nu_e_init
        out[f"{name}_gamma_e"] = s.gamma_e_init
        out[f"{name}_pis"] = s.pis_init
    np.savez_compressed(os.path.join(OUT, "init_cases.npz"), **out)

    bs = {}
    for nb, shape in [(1, (128, 128, 3)), (4, (512, 512, 3)), (16, (1080, 1920, 5)), (8, (720, 1280, 32, 6)),
                      (6, (48, 64, 5)), (3, (30, 42, 3)), (5, (24, 32, 8, 6)), (2, (17, 19, 3))]:
        bs[f"{nb}_" + "x".join(map(str, shape))] = np.array(Smoe.get_batch_shape(nb, shape))
    np.savez_compressed(os.path.join(OUT, "batch_shapes.npz"), **bs)

    sw = {}
    img = np.arange(6 * 8 * 3, dtype=np.float64).reshape(6, 8, 3)
    coords = [c for c, _ in ref_smoe.sliding_window(img, 0, (3, 4))]
    wins = [w for _, w in ref_smoe.sliding_window(img, 0, (3, 4))]
    sw["img2"] = img
    sw["coords2"] = np.array(coords)
    sw["wins2"] = np.array(wins)
    vid = np.arange(4 * 6 * 4 * 2, dtype=np.float64).reshape(4, 6, 4, 2)
    sw["img3"] = vid
    sw["coords3"] = np.array([c for c, _ in ref_smoe.sliding_window(vid, 0, (2, 3, 2))])
    sw["wins3"] = np.array([w for _, w in ref_smoe.sliding_window(vid, 0, (2, 3, 2))])
    sw["coords2_ov"] = np.array([c for c, _ in ref_smoe.sliding_window(img, 1, (3, 4))])
    sw["wins2_ov"] = np.array([w for _, w in ref_smoe.sliding_window(img, 1, (3, 4))])
    np.savez_compressed(os.path.join(OUT, "sliding_window.npz"), **sw)


def random_params(rs, K, d, C, dtype=np.float32):
    A_diag = np.zeros((K, d, d), dtype)
    A_corr = np.zeros((K, d, d), dtype)
    for i in range(d):
        A_diag[:, i, i] = rs.uniform(20, 300, K)
        for j in range(i):
            A_corr[:, i, j] = rs.normal(0, 30, K)
    pis = rs.uniform(-0.2, 1.0, K).astype(dtype)      # some <= 0 -> dropped by reduce_params
    return {"pis": pis, "musX": rs.uniform(0, 1, (K, d)).astype(dtype), "A_diagonal": A_diag, "A_corr": A_corr,
            "gamma_e": rs.normal(0, 1, (K, d, C)).astype(dtype), "nu_e": rs.uniform(0, 1, (K, C)).astype(dtype)}


def golden_quant(ref_quantizer):
    out = {}
    cases = []
    for qm in (0, 1, 2, 3):
        for qp in (False, True):
            if qm == 3 and not qp:
                continue          # reference raises UnboundLocalError there (quantizer.py:36-41)
            for (d, C) in ((2, 1), (2, 3), (3, 3)):
                cases.append((qm, qp, d, C))
    for ci, (qm, qp, d, C) in enumerate(cases):
        rs = np.random.RandomState(100 + ci)
        K = 37 + ci
        p = random_params(rs, K, d, C)
        s = _Shim()
        s.quantization_mode, s.quantize_pis, s.radial_as, s.dim_domain = qm, qp, False, d
        s.image = np.zeros((4,) * d + (C,), np.float32)
        s.lower_bounds, s.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
        s.bit_depths = [20, 18, 6, 10, 10] if ci % 2 == 0 else [12, 11, 8, 7, 9]
        s.use_diff_center = False
        s.musX_init = None
        q = ref_quantizer.quantize_params(s, copy.deepcopy(p))
        r = ref_quantizer.rescaler(s, q)
        pre = f"case{ci}_"
        out[pre + "meta"] = np.array([qm, int(qp), d, C, K] + list(s.bit_depths))
        for k, v in p.items():
            out[pre + "in_" + k] = v
        for k in ("A_diagonal", "A_corr", "musX", "nu_e", "pis", "gamma_e"):
            out[pre + "q_" + k] = q[k]
            out[pre + "lb_" + k] = np.asarray(q["lower_bounds"][k])
            out[pre + "ub_" + k] = np.asarray(q["upper_bounds"][k])
        for k, v in r.items():
            out[pre + "r_" + k] = v
    out["num_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "quant_cases.npz"), **out)


def golden_quant_radial(ref_quantizer):
    """radial_as (smoe.py:429-434, 714-721): A_diagonal is one scalar per kernel, A_corr is not quantised
    (quantizer.py:11, 45, 61, 80, 97-136)."""
    out = {}
    cases = [(qm, qp, d, C) for qm in (1, 2, 3) for qp in (False, True) if not (qm == 3 and not qp)
             for (d, C) in ((2, 3), (3, 1))]
    for ci, (qm, qp, d, C) in enumerate(cases):
        rs = np.random.RandomState(300 + ci)
        K = 29 + ci
        p = random_params(rs, K, d, C)
        p["A_diagonal"] = rs.uniform(5, 60, K).astype(np.float32)
        p["A_corr"] = np.zeros((K, d, d), np.float32)
        s = _Shim()
        s.quantization_mode, s.quantize_pis, s.radial_as, s.dim_domain = qm, qp, True, d
        s.image = np.zeros((4,) * d + (C,), np.float32)
        s.lower_bounds, s.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
        s.bit_depths = [12, 11, 8, 7, 9] if ci % 2 == 0 else [20, 18, 6, 10, 10]
        s.use_diff_center = False
        s.musX_init = None
        q = ref_quantizer.quantize_params(s, copy.deepcopy(p))
        r = ref_quantizer.rescaler(s, q)
        pre = f"case{ci}_"
        out[pre + "meta"] = np.array([qm, int(qp), d, C, K] + list(s.bit_depths))
        for k, v in p.items():
            out[pre + "in_" + k] = v
        assert "A_corr" not in q
        for k in ("A_diagonal", "musX", "nu_e", "pis", "gamma_e"):
            out[pre + "q_" + k] = q[k]
            out[pre + "lb_" + k] = np.asarray(q["lower_bounds"][k])
            out[pre + "ub_" + k] = np.asarray(q["upper_bounds"][k])
        for k, v in r.items():
            out[pre + "r_" + k] = v
    out["num_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "quant_radial_cases.npz"), **out)


def golden_graph():
    """Graph fixtures from the float64 restatement (NOT TensorFlow output: pinned=False)."""
    import torch
    from .graph import GraphCfg, graph_grads, PARAM_KEYS
    from . import init_ref
    out = {"pinned": np.array(False)}
    cases = [
        dict(name="g21", shape=(24, 28, 1), k=[4, 5], tic=False, det=True, yuv=False, seed=21, steer=True),
        dict(name="g23", shape=(20, 24, 3), k=[4, 4], tic=False, det=True, yuv=True, seed=22, steer=True),
        dict(name="g33", shape=(10, 12, 6, 3), k=[3, 3, 2], tic=False, det=False, yuv=False, seed=23, steer=True),
        dict(name="g21tic", shape=(16, 16, 1), k=[4, 4], tic=True, det=False, yuv=True, seed=24, steer=True),
        dict(name="g31", shape=(8, 10, 6, 1), k=[2, 3, 2], tic=False, det=True, yuv=False, seed=25, steer=False),
    ]
    for c in cases:
        img = synth_image(c["shape"], c["seed"])
        d = img.ndim - 1
        C = img.shape[-1]
        rs = np.random.RandomState(c["seed"])
        mus, A = init_ref.kernel_grid(c["k"], d, c["tic"])
        nu, ga = init_ref.experts(img, mus)
        K = mus.shape[0]
        p = {"pis": init_ref.pis(K, True).astype(np.float64) * rs.uniform(0.5, 1.5, K),
             "musX": mus + rs.normal(0, 0.01, mus.shape),
             "A_diagonal": np.where(np.eye(d, dtype=bool)[None], A * rs.uniform(0.8, 1.2, A.shape), 0.0),
             "A_corr": np.where(np.tril(np.ones((d, d), bool), -1)[None],
                                rs.normal(0, 0.15 * A.max(), A.shape) if c["steer"] else np.zeros(A.shape), 0.0),
             "gamma_e": rs.normal(0, 0.3, ga.shape), "nu_e": nu.astype(np.float64)}
        p["pis"][rs.choice(K, 2, replace=False)] = [0.0, -0.01]          # pruned kernels
        # round to float32 so float32 implementations start from identical values
        p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
        jd = init_ref.gen_domain(img, d).reshape(-1, d + C)
        jd32 = jd.astype(np.float32).astype(np.float64)
        cfg = GraphCfg(dim_domain=d, num_channels=C, use_determinant=c["det"], train_inverse_cov=c["tic"],
                       use_yuv=c["yuv"], start_pis=K)
        klist = np.ones(K, bool)
        klist[rs.choice(K, 1)] = False
        tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
        dom = torch.tensor(jd32[:, :d])
        tgt = torch.tensor(jd32[:, d:])
        o, g = graph_grads(tp, klist, dom, tgt, cfg, pis_l1=0.3, u_l1=1e-6)
        n = c["name"] + "_"
        out[n + "image"] = img
        out[n + "k"] = np.array(c["k"])
        out[n + "flags"] = np.array([c["tic"], c["det"], c["yuv"]])
        out[n + "kernel_list"] = klist
        for k in PARAM_KEYS:
            out[n + "p_" + k] = p[k].astype(np.float32)
            out[n + "g_" + k] = g[k].numpy()
        for k in ("S", "r_pre", "res", "resq", "w_e_max", "indices", "indices_infl"):
            out[n + k] = o[k].detach().numpy()
        out[n + "loss"] = np.array(float(o["loss"].detach()))
        out[n + "mse_op"] = np.array(float(o["mse_op"].detach()))
        wf = o["w_full"].detach().numpy()
        tau = 0.5 / 256
        out[n + "thr_margin"] = np.abs(wf / tau - 1).min(axis=0)
    np.savez_compressed(os.path.join(OUT, "graph_cases.npz"), **out)


def golden_saved_model(ref_smoe, ref_quantizer, ref_utils):
    """A checkpoint pickle written by the REFERENCE's own `utils.save_model` (utils.py:18-59) -- with its
    `reduce_params` and `quantizer.quantize_params` -- from a shim holding reference-initialised parameters, plus the
    image it belongs to.  tests load it through the product's load_params / smoe_reconstruction.main (SURVEY.md 8 f-2:
    existing parameter files must work unchanged)."""
    Smoe = ref_smoe.Smoe
    img = synth_image((48, 64, 3), 1201)
    s = object.__new__(Smoe)
    s.image, s.dim_domain, s.train_inverse_cov = img, 2, False
    s.musX_init = s.A_init = None
    s.generate_kernel_grid([6, 8])
    s.generate_experts()
    s.generate_pis(False)
    rs = np.random.RandomState(7)
    K = s.musX_init.shape[0]
    A = np.asarray(s.A_init, np.float32)
    A_diag = np.zeros_like(A)
    A_corr = np.zeros_like(A)
    for i in range(2):
        A_diag[:, i, i] = A[:, i, i] * rs.uniform(0.8, 1.2, K)
    A_corr[:, 1, 0] = rs.normal(0, 3, K)
    pis = np.ones(K, np.float32)
    pis[rs.choice(K, 9, replace=False)] = 0.0          # pruned kernels: reduce_params drops them
    params = {"pis": pis, "musX": np.asarray(s.musX_init, np.float32) + rs.normal(0, 0.004, (K, 2)).astype(np.float32),
              "A_diagonal": A_diag, "A_corr": A_corr,
              "gamma_e": rs.normal(0, 0.2, (K, 2, 3)).astype(np.float32), "nu_e": np.asarray(s.nu_e_init, np.float32)}
    shim = _Shim()
    shim.quantization_mode, shim.quantize_pis = 1, False
    shim.bit_depths = [20, 18, 6, 10, 10]
    shim.lower_bounds, shim.upper_bounds = [-2500, -.3, -5, 0, -32], [2500, 1.3, 5, 2, 32]
    shim.use_yuv, shim.only_y_gamma, shim.ssim_opt = True, False, False
    shim.use_determinant, shim.use_diff_center, shim.radial_as = True, False, False
    shim.train_gammas, shim.train_musx, shim.train_pis = True, True, True
    shim.dim_domain, shim.image, shim.train_trafo, shim.affines = 2, img, False, None
    shim.musX_init = np.asarray(s.musX_init)
    shim.qparams = ref_quantizer.quantize_params(shim, copy.deepcopy(params))
    shim.get_params = lambda: copy.deepcopy(params)
    shim.get_best_params = shim.get_params
    shim.get_mses = lambda: [(0, 812.5), (100, 301.25)]
    shim.get_losses = lambda: [(0, 0.0123), (100, 0.0045)]
    shim.get_num_pis = lambda: [(0, K), (100, K - 9)]
    path = os.path.join(OUT, "ref_saved_model_00000100_params.pkl")
    ref_utils.save_model(shim, path, best=False, reduce=True, quantize=True)
    np.save(os.path.join(OUT, "ref_saved_model_image.npy"), np.round(img * 255).astype(np.uint8))
    print("wrote", path)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_smoe, ref_quantizer, ref_utils = import_reference()
    if "--saved-model-only" in sys.argv:
        golden_saved_model(ref_smoe, ref_quantizer, ref_utils)
        return
    golden_saved_model(ref_smoe, ref_quantizer, ref_utils)
    golden_init(ref_smoe)
    golden_quant(ref_quantizer)
    golden_quant_radial(ref_quantizer)
    golden_graph()
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
