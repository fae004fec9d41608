"""world_size-2 gloo test of the sharded path's host logic, on CPU.

The GPU engine shards contiguous row bands over ranks, every rank computes UN-normalised sums over
its own pixels with the global 1/N, and one all-reduce(sum) of [per-kernel statistics | loss
scalars | influence flags] makes every rank hold the 1-rank result.  Here the per-shard arithmetic is
done by the oracle (there is no GPU in this container); what is tested is the decomposition: band
bounds, the global normalisation, the packed exchange buffer, the OR of the influence flags."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import init_ref
    from oracle.graph import GraphCfg, PARAM_KEYS, graph_forward
    img = np.load(os.path.join(GOLDEN, "init_cases.npz"))["rgb_image"]          # 48 x 64 x 3
    H = img.shape[0]
    b0, b1 = H * rank // world, H * (rank + 1) // world                          # the engine's band rule
    mus, A = init_ref.kernel_grid([6, 8], 2, False)
    nu, ga = init_ref.experts(img, mus)
    K = mus.shape[0]
    p = {"pis": torch.tensor(init_ref.pis(K, True), dtype=torch.float64), "musX": torch.tensor(mus),
         "A_diagonal": torch.tensor(A), "A_corr": torch.zeros(K, 2, 2, dtype=torch.float64),
         "gamma_e": torch.tensor(ga), "nu_e": torch.tensor(nu, dtype=torch.float64)}
    jd = init_ref.gen_domain(img, 2)
    cfg = GraphCfg(dim_domain=2, num_channels=3, use_determinant=True, train_inverse_cov=False, use_yuv=False, start_pis=K)

    def grads(rows, scale):
        leaf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        d = torch.tensor(jd[rows].reshape(-1, 5).astype(np.float32).astype(np.float64))
        out = graph_forward(leaf, np.ones(K, bool), d[:, :2], d[:, 2:], cfg)
        # the graph's loss is a MEAN over the fed pixels; a shard contributes sum/N_total = mean * n_shard/N
        g = torch.autograd.grad(out["loss"] * scale, [leaf[k] for k in PARAM_KEYS], allow_unused=True)
        flat = torch.cat([(torch.zeros_like(leaf[k]) if gi is None else gi).reshape(-1) for k, gi in zip(PARAM_KEYS, g)])
        return flat, float(out["loss"].detach()) * scale, out["kernel_list_batch"].to(torch.float64)

    n_band = b1 - b0
    flat, loss_part, infl = grads(slice(b0, b1), n_band / H)
    xbuf = torch.cat([flat, torch.tensor([loss_part], dtype=torch.float64), infl])   # one packed buffer
    dist.all_reduce(xbuf)
    if rank == 0:
        ref_flat, ref_loss, ref_infl = grads(slice(0, H), 1.0)
        np.save(os.path.join(out_dir, "sharded.npy"), xbuf.numpy())
        np.save(os.path.join(out_dir, "single.npy"), torch.cat([ref_flat, torch.tensor([ref_loss], dtype=torch.float64), ref_infl]).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_band_sharding_equals_single_rank(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = np.load(str(tmp_path / "sharded.npy"))
    b = np.load(str(tmp_path / "single.npy"))
    K = 48
    ng = a.size - 1 - K
    np.testing.assert_allclose(a[:ng], b[:ng], rtol=1e-9, atol=1e-16)
    assert abs(a[ng] - b[ng]) < 1e-12
    np.testing.assert_array_equal(a[ng + 1:] > 0, b[ng + 1:] > 0)          # influence flags: OR over ranks


def test_band_bounds_cover_every_row_once():
    for H in (135, 1080, 720, 7):
        for world in (1, 2, 3, 4, 8):
            bands = [(H * r // world, H * (r + 1) // world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
