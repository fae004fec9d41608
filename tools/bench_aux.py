"""HBM-bound pieces and the config-5 decoder, timed with CUDA events (python tools/bench_aux.py)."""
import ctypes as C
import json
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
from smoe_b200 import Smoe, smoe_reconstruction_decoded as dec
from smoe_b200._ffi import check, lib, ptr, stream_ptr
from smoe_b200.ops.image_ops_impl import smoe_ssim, mse_gpu

sys.path.insert(0, 'tests')
out = {}
peak = json.load(open('MEASURED_PEAKS.json'))["hbm_gbs"] if __import__('os').path.exists('MEASURED_PEAKS.json') else 6650.0


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- config 5 decoder: 3840x2160 RGB, forward only ---------------------------------------------------
from test_gpu_parity import _decoded_dict
H, W, Cc = 2160, 3840, 3
cp = _decoded_dict(H, W, Cc, seed=1005)
smoe, rec, loss, mse = dec.main(cp=cp, write=False)
K = int(smoe.rparams["pis"].shape[0])
ms = timeit(lambda: smoe.run_batched(train=False, update_reconstruction=False, with_quantized_params=False) and None, 5)
# forward with the FED parameters (the decoder path proper), device-resident
def fed():
    smoe._enqueue(0, 0, False, True, True)
ms_fed = timeit(fed, 5)
out["c5_decoder"] = {"pixels": H * W, "kernels_fed": K, "grid_kernels": smoe.start_pis, "forward_ms": ms_fed,
                     "mpixel_per_s": H * W / ms_fed / 1e3, "evals_per_s": H * W * K / (ms_fed / 1e3)}
target = torch.from_numpy(bench.synth_image((H, W, Cc), 1005)).cuda()
recd = torch.from_numpy(rec).cuda()
# ---- SSIM / PSNR kernels -----------------------------------------------------------------------------
dims = (C.c_int32 * 3)(H, W, 1)
ws = torch.empty((lib().smoe_ssim_workspace_bytes(2, dims, Cc) + 7) // 8, dtype=torch.float64, device="cuda")
o = torch.zeros(4, dtype=torch.float64, device="cuda")
ms_ssim = timeit(lambda: check(lib().smoe_ssim(2, dims, Cc, ptr(recd), ptr(target), ptr(o), ptr(ws), stream_ptr()), "ssim"))
alg = 2 * H * W * Cc * 4
out["ssim_4k"] = {"ms": ms_ssim, "algorithmic_GBps": alg / ms_ssim / 1e6, "frac_of_measured_hbm": alg / ms_ssim / 1e6 / peak,
                  "ssim": [float(v) for v in o[:Cc].cpu()]}
ws2 = torch.empty(1024, dtype=torch.float64, device="cuda")
o2 = torch.zeros(1, dtype=torch.float64, device="cuda")
ms_sq = timeit(lambda: check(lib().smoe_sqerr(ptr(recd), ptr(target), C.c_size_t(recd.numel()), ptr(o2), ptr(ws2), stream_ptr()), "sq"))
out["psnr_4k"] = {"ms": ms_sq, "algorithmic_GBps": alg / ms_sq / 1e6, "frac_of_measured_hbm": alg / ms_sq / 1e6 / peak,
                  "psnr_db": float(10 * np.log10(1.0 / (o2.item() / recd.numel())))}
# ---- pi-mask compaction on the 518,400-kernel grid of config 5 ----------------------------------------
Kall, P, PK = smoe.start_pis, smoe._P, smoe._PK
smoe._theta[:, smoe._off["pi"]] = torch.where(torch.rand(Kall, device="cuda") < 0.7, 1.0, -1.0)
ms_pack = timeit(lambda: check(lib().smoe_pack(C.byref(smoe._cfg), ptr(smoe._theta), ptr(smoe._mus_grid), ptr(None), ptr(smoe._klist[0]), ptr(smoe._perm), Kall, ptr(smoe._packed),
                                               ptr(smoe._indices), ptr(smoe._pos), ptr(smoe._counts[0]), ptr(smoe._regsums[0]),
                                               ptr(smoe._chunk_bounds), ptr(smoe._pack_ws), ptr(None), ptr(None), ptr(None), stream_ptr()), "pack"))
Ka = int(smoe._counts[0, 0])
alg_pack = Kall * (P * 4 + 1) + Ka * (PK * 4 + 4) + Kall * 4
out["compaction_518k"] = {"ms": ms_pack, "K_all": Kall, "K_active": Ka, "algorithmic_GBps": alg_pack / ms_pack / 1e6,
                          "frac_of_measured_hbm": alg_pack / ms_pack / 1e6 / peak}
# ---- widened rows on config 3: SSIM loss, fake-quant-aware training (one GPU, graph replay) --------------
del smoe, recd, target
torch.cuda.empty_cache()
from smoe_b200 import AdamOptimizer
shape, kgrid, seed, desc = bench.WORKLOADS["c3"]
img3 = bench.synth_image(shape, seed)


def step_ms(**kw):
    kws = dict(bench.SMOE_KW)
    kws.update(kw)
    mm = Smoe(img3, kernels_per_dim=kgrid, **kws)
    mm.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    for _ in range(4):
        mm.run_batched(train=True)
    ms_ = timeit(lambda: mm.run_batched(train=True), 30)
    return mm, ms_


qkw = dict(normalize_pis=False, lower_bounds=[-2500, -.3, -5, 0, -32], upper_bounds=[2500, 1.3, 5, 2, 32],
           bit_depths=[16, 16, 8, 10, 10])
m0, ms0 = step_ms()
out["c3_step_ms"] = {"squared_error_loss": ms0}
del m0
ms_s = None
mS, ms_s = step_ms(ssim_opt=True)
out["c3_step_ms"]["ssim_opt"] = ms_s
b = mS._batches[0]
ms_sl = timeit(lambda: check(lib().smoe_ssim_loss(C.byref(mS._cfg), C.byref(b), ptr(None), ptr(mS._d_res), ptr(mS._d_image),
                                                  ptr(mS._d_res_pre), ptr(mS._pix), ptr(mS._scalars[0]), ptr(mS._ssim_ws),
                                                  stream_ptr()), "ssim_loss"))
npx = mS.num_pixel
alg_sl = npx * 3 * 4 * 3 + npx * (3 + 1) * 4 * 2          # read res, image, res_pre; read-modify-write g_c, gr planes
out["ssim_loss_c3"] = {"ms": ms_sl, "algorithmic_GBps": alg_sl / ms_sl / 1e6, "frac_of_measured_hbm": alg_sl / ms_sl / 1e6 / peak,
                       "launches": 2 * 2 + 2}
del mS
for qm in (2, 3):
    mq, msq = step_ms(quantization_mode=qm, use_diff_center=True, **qkw)
    out["c3_step_ms"][f"quantization_mode_{qm}"] = msq
    del mq
print(json.dumps(out, indent=1))
