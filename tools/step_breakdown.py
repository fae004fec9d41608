"""Per-launch device times of one eager training step (events between the C-ABI calls), optionally under torchrun.
    python tools/step_breakdown.py c3            |  torchrun ... tools/step_breakdown.py c3"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
from smoe_b200 import Smoe, AdamOptimizer, _ffi
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
shape, kgrid, seed, desc = bench.WORKLOADS[wl]
m = Smoe(bench.synth_image(shape, seed), kernels_per_dim=kgrid, **bench.SMOE_KW)
m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
for _ in range(10):
    m.run_batched(train=True)
# wrap every library entry point with an event pair
L = _ffi.lib()
names = [n for n in _ffi.EXPORTS if n not in ("smoe_last_error", "smoe_abi_version")]
log = []
class Wrap:
    def __init__(self, fn, name): self.fn, self.name = fn, name
    def __call__(self, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = self.fn(*a); e1.record(); log.append((self.name, e0, e1)); return r
class Proxy:
    def __getattr__(self, n):
        f = getattr(L, n)
        return Wrap(f, n) if n in names and not n.endswith("_bytes") else f
_ffi._lib = Proxy()
m.use_cuda_graphs = False
acc = {}
host = []
for it in range(12):
    log.clear()
    if world > 1: torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); E0 = torch.cuda.Event(enable_timing=True); E1 = torch.cuda.Event(enable_timing=True)
    E0.record(); m.run_batched(train=True); E1.record(); torch.cuda.synchronize()
    if it < 2: continue
    host.append((time.perf_counter() - t0) * 1e3)
    for n, a, b in log:
        acc.setdefault(n, []).append(a.elapsed_time(b))
    acc.setdefault("TOTAL_eager_step", []).append(E0.elapsed_time(E1))
_ffi._lib = L
m.use_cuda_graphs = True
m._graphs = {}
for _ in range(3): m.run_batched(train=True)
g = bench.event_time(lambda: m.run_batched(train=True), 20, warm=2)
# host-side time of run_batched before/after the GPU work (graph path)
t0 = time.perf_counter()
for _ in range(50): m.run_batched(train=True)
wall = (time.perf_counter() - t0) / 50 * 1e3
if rank == 0:
    print(json.dumps({"world": world, "graph_step_ms": g, "graph_wall_ms": wall, "eager_wall_ms": float(np.mean(host)),
                      **{k: round(float(np.mean(v)), 4) for k, v in acc.items()}}, indent=0))
if world > 1:
    m.close(); torch.distributed.destroy_process_group()
