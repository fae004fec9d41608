"""Host glue with the reference's names (utils.py:7-162, plotter.py:14-15): parameter reduction,
the pickle schema of save_model / load_params, image read / write.  No arithmetic lives here."""
from __future__ import annotations

import pickle

import numpy as np


def reduce_params(params):
    """Drop kernels with pi <= 0 from every tensor of the dict -- in place, like the reference
    (utils.py:7-15) -- and return (params, boolean index)."""
    idx = params["pis"] > 0
    for key in ("pis", "A_diagonal", "A_corr", "nu_e", "gamma_e", "musX"):
        params[key] = params[key][idx]
    return params, idx


def psnr(mse, precision):
    return 10 * np.log10((2 ** precision) ** 2 / mse)


def save_model(smoe, path, best=False, reduce=True, quantize=True):
    """Checkpoint pickle with the reference's schema (utils.py:18-59)."""
    params = smoe.get_best_params() if best else smoe.get_params()
    bool_idx = None
    if reduce:
        params, bool_idx = reduce_params(params)
    cp = {"params": params, "mses": smoe.get_mses(), "losses": smoe.get_losses(), "num_pis": smoe.get_num_pis(),
          "quantization_mode": smoe.quantization_mode, "quantized_pis": smoe.quantize_pis,
          "lower_bounds": smoe.lower_bounds, "upper_bounds": smoe.upper_bounds, "use_yuv": smoe.use_yuv,
          "only_y_gamma": smoe.only_y_gamma, "ssim_opt": smoe.ssim_opt, "use_determinant": smoe.use_determinant,
          "use_diff_center": smoe.use_diff_center}
    if quantize:
        q = smoe.qparams
        q.update({"dim_of_domain": smoe.dim_domain, "dim_of_output": smoe.image.shape[-1],
                  "shape_of_img": smoe.image.shape[:-1], "used_ranges": False, "quantized_tria_params": True,
                  "trained_gamma": smoe.train_gammas, "trained_musx": smoe.train_musx, "radial_as": smoe.radial_as,
                  "trained_pis": smoe.train_pis, "use_yuv": smoe.use_yuv, "only_y_gamma": smoe.only_y_gamma,
                  "use_determinant": smoe.use_determinant, "use_diff_center": smoe.use_diff_center})
        if reduce:
            q.update({"used_kernels": bool_idx})
        cp["qparams"] = q
    with open(path, "wb") as fd:
        pickle.dump(cp, fd)


def load_params(path):
    with open(path, "rb") as fd:
        return pickle.load(fd)["params"]


def read_image(path, use_yuv=True):
    """(image float32 in [0,1], precision, affines) as utils.py:68-134: still images, video containers (through
    OpenCV, frames stacked along the third axis), `.npz` frame stacks with their `affines`, light-field `.mat`
    (needs hdf5storage, as the reference), and -- an addition -- `.npy` arrays.  `.yuv` raw video is refused exactly
    as the reference refuses it (utils.py:111-113)."""
    affines = None
    low = path.lower()
    if low.endswith(".npy"):
        orig = np.load(path)
    elif low.endswith((".png", ".tif", ".tiff", ".pgm", ".ppm", ".jpg", ".jpeg")):
        import cv2
        orig = cv2.imread(path)
        same = np.logical_and(orig[:, :, 0] == orig[:, :, 1], orig[:, :, 0] == orig[:, :, 2])
        if int(same.sum()) == orig.shape[0] * orig.shape[1]:
            orig = orig[:, :, :1]
        if orig.shape[2] == 3 and use_yuv:
            orig = cv2.cvtColor(orig, cv2.COLOR_BGR2YUV)
    elif low.endswith((".mp4", ".avi", ".mov", ".mkv", ".flv")):
        import cv2
        cap = cv2.VideoCapture(path)
        frames = []
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2YUV) if use_yuv else frame)
        cap.release()
        if not frames:
            raise ValueError("no frames could be read from " + path)
        orig = np.uint8(np.stack(frames, axis=2))                       # (rows, cols, frames, 3)
        # grayscale video: U and V planes (nearly) equal -- the reference's experimental 90 % test (utils.py:100-104)
        if int((orig[:, :, :, 1] == orig[:, :, :, 2]).sum()) > np.prod(orig.shape[0:3]) * 0.9:
            orig = orig[:, :, :, :1]
    elif low.endswith(".mat"):
        try:
            import hdf5storage
        except ImportError as exc:
            raise ImportError("reading light-field .mat files needs the hdf5storage package, as in the reference") from exc
        import cv2
        orig = hdf5storage.loadmat(path)["LF"][:, :, :, :, 0:3]
        if use_yuv:
            for ii in range(orig.shape[0]):
                for jj in range(orig.shape[1]):
                    orig[ii, jj] = cv2.cvtColor(orig[ii, jj], cv2.COLOR_RGB2YUV)
    elif low.endswith(".yuv"):
        raise ValueError("Raw Video Data is not supported yet!")
    elif low.endswith(".npz"):
        npz = np.load(path)
        orig = np.ascontiguousarray(np.moveaxis(npz["imgs"], 0, -2))    # (frames, rows, cols, C) -> (rows, cols, frames, C)
        if use_yuv and orig.shape[-1] == 3:
            import cv2
            for ii in range(orig.shape[2]):
                orig[:, :, ii, :] = cv2.cvtColor(np.ascontiguousarray(orig[:, :, ii, :]), cv2.COLOR_RGB2YUV)
        if "affines" in npz.files:
            affines = npz["affines"]
    else:
        raise ValueError("Unknown data format")
    precision = 8
    if orig.dtype == np.uint8:
        orig = orig.astype(np.float32) / 255.
    elif orig.dtype == np.uint16:
        orig = orig.astype(np.float32) / 2 ** 16.
        precision = 16
    return orig, precision, affines


def write_image(img, path, type, yuv, precision):
    """utils.py:136-162: `.png` for images (type 2), an I420 `.yuv` stream written by OpenCV for video (type 3; falls
    back to a `.npy` stack when this OpenCV build has no I420 writer), `.mat` for light fields (type 4)."""
    if precision == 8:
        img = np.uint8(np.round(img * 255))
    elif precision == 16:
        img = np.uint16(np.round(img * 2 ** precision))
    if type == 2:
        import cv2
        if yuv and img.shape[-1] == 3:
            img = cv2.cvtColor(img, cv2.COLOR_YUV2BGR)
        cv2.imwrite(path + ".png", img)
    elif type == 3:
        import cv2
        # (width, height): the reference passes img.shape[0:2], which only works for square frames
        out = cv2.VideoWriter(path + ".yuv", cv2.VideoWriter_fourcc(*"I420"), 25, (img.shape[1], img.shape[0]))
        ok = out.isOpened()
        for ii in range(img.shape[2]):
            frame = np.ascontiguousarray(img[:, :, ii, :])
            if frame.shape[-1] == 1:                       # the reference's TODO: grayscale frames as 3 equal planes
                frame = np.repeat(frame, 3, axis=-1)
            elif yuv:
                frame = cv2.cvtColor(frame, cv2.COLOR_YUV2BGR)
            if ok:
                out.write(frame)
        out.release()
        if not ok:
            np.save(path + ".npy", img)
    elif type == 4:
        try:
            import hdf5storage
        except ImportError as exc:
            raise ImportError("writing light-field .mat files needs the hdf5storage package, as in the reference") from exc
        import cv2
        if yuv:
            for ii in range(img.shape[0]):
                for jj in range(img.shape[1]):
                    img[ii, jj] = cv2.cvtColor(img[ii, jj], cv2.COLOR_YUV2RGB)
        hdf5storage.write({"LF": img}, "/", path + ".mat", matlab_compatible=True)
    else:
        np.save(path + ".npy", img)
