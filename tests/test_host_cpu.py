"""CPU-side checks of the product: the C-ABI library builds, loads and exports every symbol that
include/smoe_b200.h declares; the host mirror of the reference's NumPy helpers is bit-exact with
vectors produced by the reference itself.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__ as ge
    ge.build()
    from smoe_b200 import _ffi
    return _ffi.LIB_PATH


def test_library_exports_every_declared_symbol(libpath):
    hdr = open(os.path.join(ROOT, "include", "smoe_b200.h")).read()
    declared = set(re.findall(r"\b(smoe_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"smoe_cfg", "smoe_batch", "smoe_adam", "smoe_peers", "smoe_halo_map", "smoe_ssim_region"}
    from smoe_b200 import _ffi
    assert declared == set(_ffi.EXPORTS), declared ^ set(_ffi.EXPORTS)
    h = ctypes.CDLL(libpath)
    for s in declared:
        assert hasattr(h, s), s
    assert h.smoe_abi_version() == _ffi.ABI_VERSION == int(re.search(r"#define SMOE_ABI_VERSION (\d+)", hdr).group(1))
    assert h.smoe_param_count(2, 1) == 9 and h.smoe_param_count(2, 3) == 15 and h.smoe_param_count(3, 3) == 22
    assert h.smoe_packed_stride(2, 3) == 20 and h.smoe_packed_stride(3, 3) == 28 and h.smoe_packed_stride(2, 1) == 12


def test_struct_sizes_match_header(libpath):
    from smoe_b200 import _ffi
    assert ctypes.sizeof(_ffi.Cfg) == (14 + 1 + 15 + 3 + 1) * 4
    assert ctypes.sizeof(_ffi.Peers) == 8 + 8 * _ffi.MAX_PEERS
    assert ctypes.sizeof(_ffi.Batch) == 14 * 4
    assert ctypes.sizeof(_ffi.Adam) == 15 * 4


def test_no_cpu_fallback_and_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "steered-mixture-of-experts_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, f
    import torch
    if not torch.cuda.is_available():
        from smoe_b200 import Smoe
        with pytest.raises(RuntimeError):
            Smoe(np.zeros((8, 8, 1), np.float32), kernels_per_dim=[2, 2])


def _shell(img, tic):
    from smoe_b200 import Smoe
    s = object.__new__(Smoe)
    s.image = img
    s.dim_domain = img.ndim - 1
    s.train_inverse_cov = tic
    s.musX_init = s.A_init = None
    return s


def test_init_helpers_bit_exact_vs_reference_vectors():
    from smoe_b200 import Smoe
    z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
    for name in ("c1", "rgb", "vid", "one"):
        img = z[f"{name}_image"]
        d = img.ndim - 1
        s = _shell(img, bool(z[f"{name}_tic"]))
        s.generate_kernel_grid([int(v) for v in z[f"{name}_k"]])
        s.generate_experts()
        s.generate_pis(bool(z[f"{name}_norm"]))
        np.testing.assert_array_equal(Smoe.gen_domain(img, d), z[f"{name}_joint_domain"])
        np.testing.assert_array_equal(s.musX_init, z[f"{name}_musX"])
        np.testing.assert_array_equal(s.A_init, z[f"{name}_A"])
        np.testing.assert_array_equal(s.nu_e_init, z[f"{name}_nu_e"])
        assert s.nu_e_init.dtype == np.float32
        np.testing.assert_array_equal(s.gamma_e_init, z[f"{name}_gamma_e"])
        np.testing.assert_array_equal(s.pis_init, z[f"{name}_pis"])
        assert s.pis_init.dtype == np.float32


def test_batch_shapes_and_sliding_window_vs_reference_vectors():
    from smoe_b200 import Smoe, sliding_window
    z = np.load(os.path.join(GOLDEN, "batch_shapes.npz"))
    for key in z.files:
        nb, shp = key.split("_")
        assert Smoe.get_batch_shape(int(nb), tuple(int(v) for v in shp.split("x"))) == tuple(z[key]), key
    s = np.load(os.path.join(GOLDEN, "sliding_window.npz"))
    for img, bs, ov, ck, wk in ((s["img2"], (3, 4), 0, "coords2", "wins2"), (s["img3"], (2, 3, 2), 0, "coords3", "wins3"),
                                (s["img2"], (3, 4), 1, "coords2_ov", "wins2_ov")):
        got = list(sliding_window(img, ov, bs))
        np.testing.assert_array_equal(np.array([c for c, _ in got]), s[ck])
        np.testing.assert_array_equal(np.array([w for _, w in got]), s[wk])


def test_reduce_params_mutates_like_reference():
    from smoe_b200 import reduce_params
    p = {"pis": np.array([0.5, 0.0, -1.0, 2.0], np.float32), "A_diagonal": np.zeros((4, 2, 2)), "A_corr": np.zeros((4, 2, 2)),
         "nu_e": np.arange(4.)[:, None], "gamma_e": np.zeros((4, 2, 1)), "musX": np.zeros((4, 2))}
    q, idx = reduce_params(p)
    assert q is p and idx.tolist() == [True, False, False, True] and p["nu_e"][:, 0].tolist() == [0.0, 3.0]


def test_adam_shim_alpha_matches_tf_formula():
    from smoe_b200 import AdamOptimizer
    o = AdamOptimizer(1e-3)
    assert o._lr == 1e-3
    a1 = o._step_alpha()
    assert abs(a1 / (1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)) - 1) < 1e-4   # float32 beta powers, as TF
    a2 = o._step_alpha()
    assert abs(a2 / (1e-3 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)) - 1) < 1e-4


def test_c_abi_argument_errors_are_reported_without_a_device(libpath):
    """Every entry point validates its arguments before touching CUDA: a bad call returns SMOE_E_BADARG /
    SMOE_E_UNSUPPORTED and leaves a message in smoe_last_error() (the reference raises Python exceptions,
    smoe.py:237-241; the host mirror turns these codes into RuntimeError)."""
    from smoe_b200 import _ffi
    h = ctypes.CDLL(libpath)
    h.smoe_last_error.restype = ctypes.c_char_p
    null = ctypes.c_void_p(0)
    cfg = _ffi.Cfg()
    cfg.d, cfg.C, cfg.precision = 2, 3, 8
    b = _ffi.Batch()
    assert h.smoe_pack(ctypes.byref(cfg), null, null, null, null, null, 16, null, null, null, null, null, null, null,
                       null, null, null, null) == -1
    assert b"null argument" in h.smoe_last_error()
    assert h.smoe_forward(null, null, null, null, null, null, 16, null, null, null, null, null, null, null, null, null,
                          null, null) == -1
    assert h.smoe_loss(ctypes.byref(cfg), ctypes.byref(b), one16 := (ctypes.c_float * 16)(), one16, one16, null, one16,
                       null, one16, one16, (ctypes.c_int32 * 1)(), null) == -1
    assert b"exactly one of image" in h.smoe_last_error()
    assert h.smoe_backward(null, null, null, null, 0, null, null, null, null, null, 1, null, null, null, null) == -1
    # the peer exchange validates its peer set before touching CUDA
    pr = _ffi.Peers()
    pr.world, pr.rank = 9, 0
    assert h.smoe_xchg_publish(ctypes.byref(cfg), ctypes.byref(pr), one8 := (ctypes.c_int32 * 8)(), 4, 1, null, null,
                               (ctypes.c_float * 16)(), (ctypes.c_ubyte * 4)(), null) == -1
    assert b"bad peer set" in h.smoe_last_error()
    assert h.smoe_xchg_window_bytes(32768, 15) == 256 + 2 * 4 * (32768 * 15 + 16 + 32768 + 1024)
    assert h.smoe_adam_step(ctypes.byref(cfg), null, null, null, null, null, null, 0, null, null, null) == -1
    assert h.smoe_ssim_loss(ctypes.byref(cfg), ctypes.byref(b), null, null, null, null, null, null, null) == -1
    cfg3 = _ffi.Cfg()
    cfg3.d, cfg3.C, cfg3.quantization_mode = 2, 3, 3
    one = (ctypes.c_float * 64)()
    assert h.smoe_pack(ctypes.byref(cfg3), one, null, null, one, null, 4, one, one, one, one, one, one, one, null, null,
                       null, null) == -1
    assert b"quant_ranges" in h.smoe_last_error()
    assert h.smoe_quant_ranges(ctypes.byref(cfg), one, 4, 1, one, null) == -1      # only meaningful for mode 3
    assert h.smoe_quant_ranges_bytes() > 0
    # the host mirror raises with the library's message
    with pytest.raises(RuntimeError, match="null argument"):
        _ffi.check(h.smoe_pack(ctypes.byref(cfg), null, null, null, null, null, 16, null, null, null, null, null, null,
                               null, null, null, null, null), "smoe_pack")


def test_absorbed_terms_leave_a_float32_sum_unchanged():
    """The exactness argument of the forward's sweep A, part 2 (csrc/forward.cu): S only grows, and a term 2^q with
    q < log2(S) - 25.25 is below half an ulp of S, so adding it -- as the dense mode does -- returns S bit for bit.
    Checked in numpy float32 for sums of any magnitude, with the term at the very edge of the cut (and 2 ulp of
    ex2.approx error on top), and for a whole tail of such terms added one by one."""
    rs = np.random.RandomState(0)
    S = np.exp2(rs.uniform(-100, 100, 200000)).astype(np.float32)
    S = np.concatenate([S, np.exp2(np.arange(-100, 100)).astype(np.float32),                  # exact powers of two
                        np.nextafter(np.exp2(np.arange(-100, 100)).astype(np.float32), np.float32(0))])   # just below
    L = np.log2(S.astype(np.float64))
    t = (np.exp2(L - 25.25) * (1 + 2.0 ** -21)).astype(np.float32)
    assert np.array_equal(S + t, S)
    acc = S.copy()
    for _ in range(64):                        # many absorbed terms in a row never add up
        acc = acc + t
    assert np.array_equal(acc, S)
    # the bound is about as tight as it can be: two binades up, the term does change some of the sums
    t_big = np.exp2(L - 23.0).astype(np.float32)
    assert not np.array_equal(S + t_big, S)
