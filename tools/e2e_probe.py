import json, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from smoe_b200 import Smoe, AdamOptimizer
shape, kgrid, seed, desc = bench.WORKLOADS["c3"]
img = bench.synth_image(shape, seed)
m = Smoe(img, kernels_per_dim=kgrid, **bench.SMOE_KW)
m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
u8 = torch.from_numpy(np.round(img * 255).astype(np.uint8)).pin_memory()
dst = torch.empty_like(u8, device="cuda")
out = {}
out["h2d_6MB_ms"] = bench.event_time(lambda: dst.copy_(u8, non_blocking=True), 20, warm=3)
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device="cuda")
out["h2d_256MB_GBps"] = 0.268 / (bench.event_time(lambda: dbig.copy_(big, non_blocking=True), 5, warm=1) / 1e3)
for _ in range(5): m.run_batched(train=True)
out["step_ms"] = bench.event_time(lambda: m.run_batched(train=True), 20, warm=3)
def e2e():
    m.set_image(u8); m.run_batched(train=True)
for _ in range(5): e2e()
out["e2e_ms"] = bench.event_time(e2e, 20, warm=3)
# host-side cost of the calls
t0 = time.perf_counter()
for _ in range(50): m.set_image(u8)
torch.cuda.synchronize(); out["set_image_wall_ms"] = (time.perf_counter() - t0) / 50 * 1e3
t0 = time.perf_counter()
for _ in range(50): e2e()
out["e2e_wall_ms"] = (time.perf_counter() - t0) / 50 * 1e3
t0 = time.perf_counter()
for _ in range(50): m.run_batched(train=True)
out["step_wall_ms"] = (time.perf_counter() - t0) / 50 * 1e3
# no-overlap variant: copy on the main stream
def e2e_serial():
    m._d_image_u8.copy_(u8, non_blocking=True); m._use_u8 = True; m.run_batched(train=True)
for _ in range(3): e2e_serial()
out["e2e_serial_ms"] = bench.event_time(e2e_serial, 20, warm=3)
print(json.dumps(out))
