"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank builds the row-sharded model and, next to it, an unsharded copy of the same model
(distributed=False) on its own GPU, runs three training steps on both and compares loss, accumulated
gradients (<= 1e-4 relative to the tensor max-norm, SURVEY.md 8e: summation order differs with R),
kernel lists and the gathered reconstruction."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    import bench
    from smoe_b200 import Smoe, AdamOptimizer
    ok = True
    for shape, k in (((135, 96, 3), [12, 10]), ((40, 48, 12, 3), [4, 4, 3])):
        img = bench.synth_image(shape, 77)
        kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False)
        ms = Smoe(img, kernels_per_dim=k, **kw)
        m1 = Smoe(img, kernels_per_dim=k, distributed=False, **kw)
        assert ms._world == world and m1._world == 1
        for m in (ms, m1):
            m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
        for step in range(3):
            a = ms.run_batched(pis_l1=0.1, train=True)
            b = m1.run_batched(pis_l1=0.1, train=True)
            ga, gb = ms.get_gradients(), m1.get_gradients()
            for key in ga:
                rel = np.abs(ga[key] - gb[key]).max() / max(np.abs(gb[key]).max(), 1e-30)
                if rel > 1e-4:
                    ok = False
                    print(f"rank {rank} step {step} {key} rel {rel:.3e}")
            if abs(a[0] - b[0]) > 1e-6 or a[2] != b[2]:
                ok = False
                print(f"rank {rank} step {step} loss {a[0]} vs {b[0]}")
        ra, rb = ms.get_reconstruction(), m1.get_reconstruction()
        if ra.shape != rb.shape or (np.round(ra * 255) != np.round(rb * 255)).mean() > 2e-3:
            ok = False
            print(f"rank {rank} reconstruction mismatch")
        kla, klb = ms.kernel_list_per_batch[0], m1.kernel_list_per_batch[0]
        if (kla != klb).sum() > 1:
            ok = False
            print(f"rank {rank} kernel list mismatch {(kla != klb).sum()}")
        pa, pb = ms.get_params(), m1.get_params()
        for key in pa:
            if np.abs(pa[key] - pb[key]).max() > 1e-3 * max(1.0, np.abs(pb[key]).max()):
                ok = False
                print(f"rank {rank} params {key} diverged")
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_CHECK", "OK" if t.item() == 1.0 else "FAILED", f"world={world}")
    torch.distributed.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
