import json, sys
sys.path.insert(0, '.')
import torch, bench
from smoe_b200 import Smoe, AdamOptimizer
shape, kgrid, seed, desc = bench.WORKLOADS["c3"]
img = bench.synth_image(shape, seed)
for mode in (1, 0):
    m = Smoe(img, kernels_per_dim=kgrid, dense_exec=mode, **bench.SMOE_KW)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    m.run_batched(train=True)
    k = bench.kernel_times(m, steps=4 if mode else 10, with_step=False)
    print(json.dumps({"mode": mode, **{a: round(b, 4) for a, b in k.items() if a.endswith("_ms")}}))
    del m
