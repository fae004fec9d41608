"""Host-logic checks of the oracle's Smoe emulation (batches, kernel lists, Adam, training)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle.model import OracleAdam, OracleSmoe


def _img(name="rgb"):
    return np.load(os.path.join(GOLDEN, "init_cases.npz"))[f"{name}_image"]


def test_tf_adam_first_steps():
    var = torch.tensor([1.0, -2.0], dtype=torch.float64)
    opt = OracleAdam(0.1)
    g = torch.tensor([0.5, -0.25], dtype=torch.float64)
    opt.apply([("v", g, var)])
    # first step of Adam moves by ~lr*sign(g) (epsilon outside the bias correction)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = np.array([1.0, -2.0]) - lr_t * (0.1 * g.numpy()) / (np.sqrt(0.001 * g.numpy() ** 2) + 1e-8)
    np.testing.assert_allclose(var.numpy(), exp, rtol=1e-12)


def test_training_reduces_loss_and_batches_scale_gradients():
    img = _img("rgb")[:32, :32]
    a = OracleSmoe(img, kernels_per_dim=[4, 4], use_yuv=False, train_inverse_cov=False, use_determinant=True,
                   dtype=torch.float64)
    a.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    l0 = a.run_batched(train=False)[0]
    a.run_batched(train=True)
    g1 = {k: v.clone() for k, v in a.last_grads.items()}
    for _ in range(15):
        l1 = a.run_batched(train=True)[0]
    assert l1 < l0
    # 4 equal batches: accumulated gradient = sum of per-batch means = 4x the 1-batch gradient
    b = OracleSmoe(img, kernels_per_dim=[4, 4], use_yuv=False, train_inverse_cov=False, use_determinant=True,
                   start_batches=4, dtype=torch.float64)
    assert b.start_batches == 4 and b.batch_size_valued == (16, 16)
    b.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    lb = b.run_batched(train=True)[0]
    assert abs(lb - l0) < 1e-12
    for k in g1:
        np.testing.assert_allclose(b.last_grads[k].numpy(), 4 * g1[k].numpy(), rtol=1e-9, atol=1e-15)


def test_pruning_state_machine():
    img = _img("c1")[:32, :32]
    m = OracleSmoe(img, kernels_per_dim=[4, 4], use_yuv=False, train_inverse_cov=False, dtype=torch.float64)
    m.vars["pis"][5] = -1.0
    m.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    _, _, num_pi, _ = m.run_batched(train=True)
    assert num_pi == 15 and not m.kernel_list_per_batch[0][5]
    m.vars["pis"][5] = 1.0          # comes back > 0 but stays off the list (smoe.py:1763-1766)
    _, _, num_pi, _ = m.run_batched(train=False)
    assert num_pi == 16 and not m.kernel_list_per_batch[0][5]
