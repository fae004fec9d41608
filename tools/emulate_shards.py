"""Per-rank work of a pixel-sharded step, measured on ONE GPU: rank r of R is emulated by a model that owns only that
rank's pixel block (no exchange), so ncu / event timing of the sharded kernels does not need R GPUs.

    python tools/emulate_shards.py c3 8          # every rank of an 8-rank run of config 3
"""
import json
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
from smoe_b200 import Smoe, AdamOptimizer

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ranks = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else range(R)
shape, kgrid, seed, desc = bench.WORKLOADS[wl]
img = bench.synth_image(shape, seed)
rows = []
for r in ranks:
    m = Smoe(img, kernels_per_dim=kgrid, _emulate_shard=(r, R), **bench.SMOE_KW)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    for _ in range(4):
        m.run_batched(train=True)
    step = bench.event_time(lambda: m.run_batched(train=True), 20, warm=2)
    k = bench.kernel_times(m, steps=5, with_step=False)
    rows.append({"rank": r, "block": [list(b) for b in m._block], "grid": list(m._block_grid), "splits": m._splits,
                 "tiles": m._max_tiles, "step_ms": step, "forward_ms": k["forward_ms"], "backward_ms": k["backward_ms"],
                 "other_ms": step - k["forward_ms"] - k["backward_ms"], "pairs": k["pairs"]})
    print(json.dumps(rows[-1]))
    del m
    torch.cuda.empty_cache()
print(json.dumps({"workload": wl, "ranks": R, "max_step_ms": max(x["step_ms"] for x in rows),
                  "mean_step_ms": float(np.mean([x["step_ms"] for x in rows]))}))
