"""Diagnostics: why does the culled backward take longer inside the step than alone?  (python tools/ab_probe.py)"""
import ctypes as C
import json
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
from smoe_b200 import Smoe, AdamOptimizer
from smoe_b200._ffi import check, lib, ptr, stream_ptr


def mhz():
    n = 20_000_000
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.cuda._sleep(n); e1.record(); torch.cuda.synchronize()
    return n / e0.elapsed_time(e1) / 1e3


shape, kgrid, seed, desc = bench.WORKLOADS["c3"]
m = Smoe(bench.synth_image(shape, seed), kernels_per_dim=kgrid, **bench.SMOE_KW)
m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
out = {"mhz_idle": mhz()}
for _ in range(5):
    m.run_batched(train=True)
out["kernel_times"] = {k: v for k, v in bench.kernel_times(m, steps=10, with_step=True).items() if k != "pairs"}
L, st = lib(), stream_ptr()
b = m._batches[0]
counts = m._counts[0]


def bwd():
    check(L.smoe_backward(C.byref(m._cfg), C.byref(b), ptr(m._packed), ptr(counts), m.start_pis, ptr(m._pix),
                          ptr(m._tile_qmin[0]), ptr(m._d_axes[0]), ptr(m._d_axes[1]), ptr(None), m._splits,
                          ptr(m._raw_part), ptr(None), st), "backward")


out["bwd_alone_ms"] = bench.event_time(bwd, 10, warm=2)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def bwd_flushed():
    flush.fill_(1.0)


ts = []
for _ in range(6):
    flush.fill_(1.0)
    torch.cuda.synchronize()
    ts.append(bench.event_time(bwd, 1, warm=0))
out["bwd_after_l2_flush_ms"] = float(np.mean(ts))
for _ in range(300):
    m.run_batched(train=True)
out["mhz_after_300_steps"] = mhz()
out["splits"] = m._splits
for ns in (17, 23, 31, 47):
    m._splits = ns
    m._raw_part = torch.zeros((ns * m.start_pis * m._P,), dtype=torch.float32, device=m.device)
    m._graphs = {}
    out[f"ktimes_splits_{ns}"] = {k: round(v, 4) for k, v in bench.kernel_times(m, steps=5, with_step=False).items() if k.endswith("_ms")}
print(json.dumps(out))
