"""On-disk formats, CPU: the checkpoint schema against a pickle WRITTEN BY THE REFERENCE's own
utils.save_model (tests/golden/ref_saved_model_*.pkl, made by oracle/make_golden.py), and the readers / writers of
utils.py:68-162."""
import copy
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN


def _ref_cp():
    with open(os.path.join(GOLDEN, "ref_saved_model_00000100_params.pkl"), "rb") as fd:
        return pickle.load(fd)


def test_load_params_reads_a_reference_written_checkpoint():
    from smoe_b200.utils import load_params
    p = load_params(os.path.join(GOLDEN, "ref_saved_model_00000100_params.pkl"))
    assert set(p) == {"pis", "musX", "A_diagonal", "A_corr", "gamma_e", "nu_e"}
    K = p["pis"].shape[0]
    assert K == 39 and (p["pis"] > 0).all()                # reduce_params dropped the 9 pruned kernels
    assert p["A_diagonal"].shape == (K, 2, 2) and p["gamma_e"].shape == (K, 2, 3) and p["nu_e"].dtype == np.float32


def test_save_model_writes_the_reference_schema(tmp_path):
    """Our save_model on a model-like object with the same state writes a pickle equal to the reference-written one,
    key by key (values, dtypes, nesting)."""
    from smoe_b200.utils import save_model
    ref = _ref_cp()
    used = np.asarray(ref["qparams"]["used_kernels"])
    K_all = used.size
    full = {}
    for k, v in ref["params"].items():                      # un-reduce: pruned rows get pi = 0
        arr = np.zeros((K_all,) + v.shape[1:], v.dtype)
        arr[used] = v
        full[k] = arr

    class M:
        pass
    m = M()
    m.quantization_mode, m.quantize_pis = ref["quantization_mode"], ref["quantized_pis"]
    m.lower_bounds, m.upper_bounds = ref["lower_bounds"], ref["upper_bounds"]
    m.use_yuv, m.only_y_gamma, m.ssim_opt = ref["use_yuv"], ref["only_y_gamma"], ref["ssim_opt"]
    m.use_determinant, m.use_diff_center = ref["use_determinant"], ref["use_diff_center"]
    m.radial_as, m.train_gammas, m.train_musx, m.train_pis = False, True, True, True
    m.dim_domain = 2
    m.image = np.zeros(tuple(ref["qparams"]["shape_of_img"]) + (ref["qparams"]["dim_of_output"],), np.float32)
    m.qparams = {k: copy.deepcopy(ref["qparams"][k]) for k in
                 ("lower_bounds", "upper_bounds", "steps", "A_diagonal", "musX", "nu_e", "pis", "gamma_e", "A_corr")}
    m.get_params = lambda: copy.deepcopy(full)
    m.get_best_params = m.get_params
    m.get_mses, m.get_losses, m.get_num_pis = (lambda: ref["mses"]), (lambda: ref["losses"]), (lambda: ref["num_pis"])
    out = str(tmp_path / "00000100_params.pkl")
    save_model(m, out)
    with open(out, "rb") as fd:
        got = pickle.load(fd)

    def same(a, b, path=""):
        assert type(a) is type(b) or (np.isscalar(a) and np.isscalar(b)), (path, type(a), type(b))
        if isinstance(a, dict):
            assert list(a.keys()) == list(b.keys()), path
            for k in a:
                same(a[k], b[k], path + "/" + str(k))
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                same(x, y, f"{path}[{i}]")
        elif isinstance(a, np.ndarray):
            assert a.dtype == b.dtype and a.shape == b.shape, (path, a.dtype, b.dtype, a.shape, b.shape)
            np.testing.assert_array_equal(a, b, err_msg=path)
        else:
            assert a == b, path
    same(ref, got)


def test_read_write_image_formats(tmp_path):
    import cv2
    from smoe_b200.utils import read_image, write_image
    rs = np.random.RandomState(0)
    # still image round trip through .png (YUV conversion both ways: +-2 codes)
    img = rs.uniform(0, 1, (24, 32, 3)).astype(np.float32)
    write_image(img, str(tmp_path / "a"), 2, False, 8)
    back, prec, aff = read_image(str(tmp_path / "a.png"), use_yuv=False)
    assert prec == 8 and aff is None and back.dtype == np.float32
    np.testing.assert_array_equal(np.round(back * 255), np.round(img * 255))
    gray = np.repeat(rs.uniform(0, 1, (24, 32, 1)).astype(np.float32), 3, axis=-1)
    write_image(gray, str(tmp_path / "g"), 2, False, 8)
    assert read_image(str(tmp_path / "g.png"))[0].shape == (24, 32, 1)           # utils.py:72-77: grey -> 1 channel
    # .npz frame stack with affines (utils.py:114-122)
    imgs = rs.randint(0, 256, (5, 24, 32, 3)).astype(np.uint8)
    np.savez(str(tmp_path / "v.npz"), imgs=imgs, affines=np.eye(3)[None].repeat(5, 0))
    vid, prec, aff = read_image(str(tmp_path / "v.npz"), use_yuv=False)
    assert vid.shape == (24, 32, 5, 3) and aff.shape == (5, 3, 3) and prec == 8
    np.testing.assert_array_equal(np.round(vid[:, :, 2] * 255), imgs[2])
    vy, _, _ = read_image(str(tmp_path / "v.npz"), use_yuv=True)
    np.testing.assert_array_equal(np.round(vy[:, :, 1] * 255), cv2.cvtColor(imgs[1], cv2.COLOR_RGB2YUV))
    # raw .yuv input is refused as in the reference; unknown extensions too
    open(str(tmp_path / "x.yuv"), "wb").write(b"0")
    with pytest.raises(ValueError, match="Raw Video"):
        read_image(str(tmp_path / "x.yuv"))
    with pytest.raises(ValueError, match="Unknown data format"):
        read_image(str(tmp_path / "x.foo"))
    # video output: an I420 stream (or the .npy fallback when this OpenCV build cannot write one)
    video = rs.uniform(0, 1, (24, 32, 4, 3)).astype(np.float32)
    write_image(video, str(tmp_path / "out"), 3, False, 8)
    assert os.path.exists(str(tmp_path / "out.yuv")) or os.path.exists(str(tmp_path / "out.npy"))
    # a video container written by OpenCV reads back as (rows, cols, frames, C)
    path = str(tmp_path / "c.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (32, 24))
    if wr.isOpened():
        for f in range(3):
            wr.write(imgs[f])
        wr.release()
        got, prec, _ = read_image(path, use_yuv=False)
        assert got.shape[:2] == (24, 32) and got.shape[2] == 3 and got.shape[3] in (1, 3) and prec == 8
