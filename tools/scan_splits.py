"""Backward segment-count scan (python tools/scan_splits.py c3 [rank world])."""
import json, os, sys
sys.path.insert(0, '.')
import torch, bench
from smoe_b200 import Smoe, AdamOptimizer
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
emu = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
shape, kgrid, seed, desc = bench.WORKLOADS[wl]
img = bench.synth_image(shape, seed)
for ns in (0, 13, 17, 23, 31, 47, 61):
    os.environ["SMOE_SPLITS"] = str(ns)
    m = Smoe(img, kernels_per_dim=kgrid, _emulate_shard=emu, **bench.SMOE_KW)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    for _ in range(30):
        m.run_batched(train=True)
    k = bench.kernel_times(m, steps=5, with_step=True)
    print(json.dumps({"splits_req": ns, "splits": m._splits, **{a: round(b, 4) for a, b in k.items() if a.endswith("_ms")}}))
    del m
    torch.cuda.empty_cache()
