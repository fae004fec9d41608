"""A few training steps of a bench workload, for ncu (python scratch/prof_step.py c3 3 [dense])."""
import sys
sys.path.insert(0, '.')
import torch
import bench
from smoe_b200 import Smoe, AdamOptimizer
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"
shape, kgrid, seed, desc = bench.WORKLOADS[wl]
m = Smoe(bench.synth_image(shape, seed), kernels_per_dim=kgrid, dense_exec=dense, **bench.SMOE_KW)
m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
for _ in range(steps):
    out = m.run_batched(train=True)
torch.cuda.synchronize()
print(desc, out)
