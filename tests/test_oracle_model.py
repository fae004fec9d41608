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


def test_loss_mask_and_pixel_sampling_in_the_oracle_model():
    """loss_mask (smoe.py:932, 1674-1677): zero-weight pixels contribute no gradient, mse is unweighted.
    sampling_percentage (smoe.py:1664-1667): round(N_b * pct / 100) pixels per batch from np.random.choice with the
    probabilities of the last reconstruction pass; the global NumPy seed makes the draw reproducible."""
    rs = np.random.RandomState(2)
    v, u = np.meshgrid(np.linspace(0, 1, 16), np.linspace(0, 1, 20), indexing="ij")
    img = np.clip(0.5 + 0.3 * np.sin(5 * u + 3 * v)[..., None] + 0.05 * rs.standard_normal((16, 20, 1)), 0, 1)
    img = (np.round(img * 255) / 255).astype(np.float32)
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False, dtype=torch.float64)
    mask = np.ones((16, 20), np.float32)
    mask[:, 10:] = 0.0
    o = OracleSmoe(img, kernels_per_dim=[3, 4], loss_mask=mask, **kw)
    o.set_optimizer(OracleAdam(0.0), OracleAdam(0.0), OracleAdam(0.0))      # lr 0: groups are skipped (smoe.py:1120)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    l0, m0, _, _ = o.run_batched(train=False)
    l1, m1, _, _ = o.run_batched(train=True, use_loss_mask=True)
    assert abs(m0 - m1) < 1e-12 and l1 < l0                     # half of the pixels carry no loss
    # kernels whose influential pixels all lie in the masked half get no expert gradient
    g_nu = o.last_grads["nu_e"].numpy()[:, 0].reshape(3, 4)
    assert np.abs(g_nu[:, 3]).max() < 1e-3 * np.abs(g_nu[:, 0]).max()
    o2 = OracleSmoe(img, kernels_per_dim=[3, 4], start_batches=2, **kw)
    o2.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    np.random.seed(3)
    a = o2.run_batched(train=True, sampling_percentage=25)
    s_a = o2.last_samples.copy()
    o3 = OracleSmoe(img, kernels_per_dim=[3, 4], start_batches=2, **kw)
    o3.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    np.random.seed(3)
    b = o3.run_batched(train=True, sampling_percentage=25)
    assert len(s_a) == 40 and np.array_equal(s_a, o3.last_samples) and a == b
    # after a reconstruction pass the probabilities are the normalised error map
    o2.run_batched(train=False, update_reconstruction=True)
    for p in o2.random_sampling_per_batch:
        assert p.shape == (160,) and abs(float(p.sum()) - 1) < 1e-5 and (p >= 0).all()
