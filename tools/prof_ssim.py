"""One launch each of the SSIM metric (4K RGB) and the SSIM loss (1080p RGB) for an ncu capture / a quick timing:
    python tools/prof_ssim.py [time]"""
import ctypes as C
import sys
sys.path.insert(0, '.')
import torch
from smoe_b200 import _ffi
from smoe_b200._ffi import check, lib, ptr, stream_ptr

L = lib()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)


def run(H, W, Cc, loss):
    img = torch.rand((H, W, Cc), generator=g).to(dev)
    pre = (img.cpu() + 0.05 * torch.randn((H, W, Cc), generator=g)).to(dev)
    res = (pre.clamp(0, 1) * 255).round() / 255
    if not loss:
        dims = (C.c_int32 * 3)(H, W, 1)
        L.smoe_ssim_workspace_bytes.restype = C.c_size_t
        ws = torch.empty((L.smoe_ssim_workspace_bytes(2, dims, Cc) + 7) // 8, dtype=torch.float64, device=dev)
        o = torch.zeros(4, dtype=torch.float64, device=dev)
        return lambda: check(L.smoe_ssim(2, dims, Cc, ptr(res), ptr(img), ptr(o), ptr(ws), stream_ptr()), "ssim")
    cfg = _ffi.Cfg()
    cfg.d, cfg.C, cfg.precision, cfg.use_yuv = 2, Cc, 8, 1
    b = _ffi.Batch()
    b.dims[:] = [H, W, 1]
    b.origin[:] = [0, 0, 0]
    b.extent[:] = [H, W, 1]
    b.tile[:] = [16, 32, 1]
    b.inv_count = 1.0 / (H * W)
    L.smoe_ssim_loss_workspace_bytes.restype = C.c_size_t
    ws = torch.zeros(L.smoe_ssim_loss_workspace_bytes(C.byref(cfg), C.byref(b)) // 4 + 64, dtype=torch.float32, device=dev)
    pix = torch.zeros((L.smoe_num_tiles(C.byref(b)) * L.smoe_pix_stride(2, Cc, C.byref(b)),), device=dev)
    scal = torch.zeros(16, device=dev)
    return lambda: check(L.smoe_ssim_loss(C.byref(cfg), C.byref(b), None, ptr(res), ptr(img), ptr(pre), ptr(pix),
                                          ptr(scal), ptr(ws), stream_ptr()), "ssim_loss")


for name, fn in (("ssim_4k_rgb", run(2160, 3840, 3, False)), ("ssim_1080p_rgb", run(1080, 1920, 3, False)),
                 ("ssim_loss_1080p_rgb", run(1080, 1920, 3, True))):
    fn()
    torch.cuda.synchronize()
    if len(sys.argv) > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            fn()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(name, "%.4f ms" % (e0.elapsed_time(e1) / 20))
