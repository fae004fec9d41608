"""Oracle self-checks: KAT-1, closed-form backward vs autograd, analytic properties."""
import numpy as np
import pytest
import torch

from oracle import init_ref
from oracle.graph import GraphCfg, PARAM_KEYS, closed_form_grads, graph_forward, graph_grads


def kat1(use_det):
    I = np.array([[round(255 * (4 * y + x) / 15) / 255 for x in range(4)] for y in range(4)], np.float64)[..., None]
    mus, A = init_ref.kernel_grid([2, 2], 2, False)
    nu, ga = init_ref.experts(I, mus)
    p = dict(pis=init_ref.pis(4, True).astype(np.float64), musX=mus, A_diagonal=A, A_corr=np.zeros_like(A),
             gamma_e=ga, nu_e=nu.astype(np.float64))
    jd = init_ref.gen_domain(I, 2).reshape(-1, 3)
    cfg = GraphCfg(dim_domain=2, num_channels=1, use_determinant=use_det, train_inverse_cov=False,
                   use_yuv=False, start_pis=4)
    return p, jd, cfg


def test_kat1_forward_and_grads():
    p, jd, cfg = kat1(False)
    tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
    out, g = graph_grads(tp, np.ones(4, bool), torch.tensor(jd[:, :2]), torch.tensor(jd[:, 2:]), cfg)
    np.testing.assert_allclose(p["nu_e"][:, 0], [0.16666667, 0.3, 0.7, 0.83333333], rtol=5e-6)
    np.testing.assert_allclose(out["S"][:4].detach(), [0.02635631, 0.07520154, 0.07520154, 0.02635631], rtol=2e-7)
    np.testing.assert_allclose(out["w_full"][0, :4].detach(), [0.999753226, 0.952456584, 0.0474200211, 1.23379350e-4], rtol=1e-7)
    assert int((~out["infl"]).sum()) == 28
    np.testing.assert_allclose(out["r_pre"][:4, 0].detach(), [0.16662554, 0.17296877, 0.29364031, 0.29992597], atol=1e-8)
    np.testing.assert_allclose(out["r_pre"][[5, 10, 15], 0].detach(), [0.198283915452, 0.801716084548, 0.833127688395], atol=1e-10)
    np.testing.assert_array_equal(np.round(out["resq"][:4, 0].detach().numpy() * 255), [42, 44, 75, 76])
    assert abs(float(out["loss"].detach()) - 0.015144070248) < 2e-9   # KAT inputs were float32-rounded
    assert abs(float(out["mse_op"].detach()) - 1023.10237601) < 1e-4
    assert abs(10 * np.log10(65536 / float(out["mse_op"].detach())) - 18.06561) < 1e-5
    np.testing.assert_allclose(g["pis"], [-4.4520e-4, 4.7775e-4, 4.5346e-4, -4.8601e-4], rtol=2e-4)
    np.testing.assert_allclose(g["nu_e"][:, 0], [0.0109761, 0.00629165, -0.00678173, -0.01146617], rtol=1e-5)
    np.testing.assert_allclose(g["musX"], [(-0.00601605, 0.0011727), (-0.00913828, 0.00215535),
                                           (0.00922202, -0.0020942), (0.00607877, -0.00119669)], rtol=1e-5)
    np.testing.assert_allclose(torch.diagonal(g["A_diagonal"], dim1=1, dim2=2),
                               [(7.37267938e-4, 1.44157150e-4), (5.16439388e-4, -6.60195321e-5),
                                (5.31619440e-4, -6.96875999e-5), (7.52599965e-4, 1.48230052e-4)], rtol=5e-6)
    np.testing.assert_allclose(g["A_corr"][:, 1, 0], [-2.80623275e-4, -7.57960392e-5, -7.78221117e-5, -2.77211053e-4], rtol=5e-6)
    np.testing.assert_allclose(g["gamma_e"].reshape(-1), [-0.00712058, -0.00034182, -0.00801585, 0.00284964,
                                                          -0.01479758, -0.00344201, -0.01858676, -0.01180799], rtol=1e-5)
    assert float(g["A_diagonal"][:, 0, 1].abs().max()) == 0 and float(g["A_corr"][:, 0, 0].abs().max()) == 0


def test_kat1_determinant():
    p, jd, cfg = kat1(True)
    tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
    out, g = graph_grads(tp, np.ones(4, bool), torch.tensor(jd[:, :2]), torch.tensor(jd[:, 2:]), cfg)
    np.testing.assert_allclose(out["S"][:2].detach(), [0.15101053, 0.43087307], rtol=2e-7)
    np.testing.assert_allclose(torch.diagonal(g["A_diagonal"], dim1=1, dim2=2),
                               [(7.18717790e-4, 1.25607002e-4), (5.36345629e-4, -4.61132916e-5),
                                (5.50513632e-4, -5.07934087e-5), (7.32349682e-4, 1.27979768e-4)], rtol=5e-6)
    assert abs(float(out["loss"].detach()) - 0.015144070248) < 2e-9   # KAT inputs were float32-rounded


def _rand_case(d, C, tic, det, yuv, seed, K=9, N=150, tg=True):
    rs = np.random.RandomState(seed)
    p = dict(pis=rs.uniform(0.05, 1, K), musX=rs.uniform(0, 1, (K, d)), gamma_e=rs.normal(0, .3, (K, d, C)),
             nu_e=rs.uniform(0, 1, (K, C)))
    Ad = np.zeros((K, d, d)); Ac = np.zeros((K, d, d))
    for i in range(d):
        Ad[:, i, i] = rs.uniform(2, 6, K)
        for j in range(i):
            Ac[:, i, j] = rs.normal(0, 1.0, K)
        for j in range(i + 1, d):
            Ac[:, i, j] = rs.normal(0, 1.0, K)     # must be ignored (upper part of A_corr)
            Ad[:, i, j] = rs.normal(0, 1.0, K)     # must be ignored (off-diagonal of A_diagonal)
    p["A_diagonal"], p["A_corr"] = Ad, Ac
    p["pis"][1] = -0.1
    x = rs.uniform(0, 1, (N, d)); t = rs.uniform(0, 1, (N, C))
    kl = np.ones(K, bool); kl[3] = False
    cfg = GraphCfg(dim_domain=d, num_channels=C, use_determinant=det, train_inverse_cov=tic, use_yuv=yuv,
                   train_gammas=tg, start_pis=K)
    return p, x, t, kl, cfg


@pytest.mark.parametrize("d,C,tic,det,yuv,tg", [(2, 1, False, False, False, True), (2, 3, False, True, True, True),
                                               (3, 3, False, True, False, True), (2, 1, True, False, True, True),
                                               (3, 1, True, True, False, True), (2, 3, False, True, True, False)])
def test_closed_form_matches_autograd(d, C, tic, det, yuv, tg):
    p, x, t, kl, cfg = _rand_case(d, C, tic, det, yuv, seed=d * 10 + C, tg=tg)
    tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
    out, g = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg, pis_l1=0.2, u_l1=1e-3)
    cf, extra = closed_form_grads(p, kl, x, t, cfg, pis_l1=0.2, u_l1=1e-3)
    for k in PARAM_KEYS:
        np.testing.assert_allclose(cf[k], g[k].numpy(), rtol=1e-9, atol=1e-14, err_msg=k)
    np.testing.assert_allclose(extra["r_pre"], out["r_pre"].detach().numpy(), rtol=1e-12, atol=1e-14)
    assert float(np.abs(cf["pis"][[1, 3]]).max()) == 0          # pruned / unlisted kernels get no gradient


def test_einsum_broadcast_mode_equals_einsum():
    for tic in (False, True):
        p, x, t, kl, cfg = _rand_case(3, 3, tic, True, False, seed=5)
        tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
        a = graph_forward(tp, kl, torch.tensor(x), torch.tensor(t), cfg)
        cfg.einsum_mode = "broadcast"
        b = graph_forward(tp, kl, torch.tensor(x), torch.tensor(t), cfg)
        np.testing.assert_allclose(a["maha"].numpy(), b["maha"].numpy(), rtol=1e-12)


def test_single_kernel_is_clipped_linear_expert():
    rs = np.random.RandomState(0)
    x = rs.uniform(0, 1, (50, 2)); t = rs.uniform(0, 1, (50, 1))
    p = dict(pis=np.array([0.7]), musX=np.array([[.4, .6]]), A_diagonal=np.array([[[3., 0], [0, 4.]]]),
             A_corr=np.array([[[0., 0], [1., 0]]]), gamma_e=np.array([[[0.9], [-1.4]]]), nu_e=np.array([[0.5]]))
    cfg = GraphCfg(dim_domain=2, num_channels=1, train_inverse_cov=False, use_yuv=False, start_pis=1)
    out = graph_forward({k: torch.tensor(v) for k, v in p.items()}, np.ones(1, bool), torch.tensor(x), torch.tensor(t), cfg)
    np.testing.assert_allclose(out["res"][:, 0].numpy(), np.clip(0.5 + 0.9 * x[:, 0] - 1.4 * x[:, 1], 0, 1), atol=1e-12)


def test_mirrored_kernels_gate_half_on_bisector_and_pruned_absent():
    p = dict(pis=np.array([0.5, 0.5, 0.0]), musX=np.array([[.3, .5], [.7, .5], [.5, .5]]),
             A_diagonal=np.tile(np.diag([5., 5.]), (3, 1, 1)), A_corr=np.zeros((3, 2, 2)),
             gamma_e=np.zeros((3, 2, 1)), nu_e=np.array([[.2], [.8], [.9]]))
    x = np.array([[.5, .1], [.5, .9]]); t = np.zeros((2, 1))
    cfg = GraphCfg(dim_domain=2, num_channels=1, train_inverse_cov=False, use_yuv=False, start_pis=3)
    out = graph_forward({k: torch.tensor(v) for k, v in p.items()}, np.ones(3, bool), torch.tensor(x), torch.tensor(t), cfg)
    np.testing.assert_allclose(out["w"].numpy(), 0.5, atol=1e-12)
    assert out["indices"].tolist() == [0, 1] and out["num_pi"] == 2
    assert out["w_e_max"].tolist() == [0, 0]          # argmax tie -> lowest index


# ----------------------------------------------------------------------------------------------------
# quantization_mode 2 (fake-quant-aware training with fixed bounds, smoe.py:482-496)
# ----------------------------------------------------------------------------------------------------
LB, UB = [-40.0, -0.3, 0.1, 0.0, -2.0], [40.0, 1.3, 0.9, 2.0, 2.0]       # order: A, musX, nu_e, pis, gamma_e


def test_mode2_values_lie_on_the_code_grid_and_gradients_pass_inside_the_bounds_only():
    p, x, t, kl, cfg = _rand_case(2, 3, False, True, True, 5)
    cfg.quantization_mode, cfg.lower_bounds, cfg.upper_bounds, cfg.bit_depths = 2, LB, UB, [10, 12, 6, 10, 8]
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}      # TF variables are float32
    tp = {k: torch.tensor(v) for k, v in p.items()}
    out, g = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg, pis_l1=0.2, u_l1=1e-3)
    # the graph sees values on the nudged code grid
    from oracle.graph import _nudge
    for key, grp in (("nu_e", 2), ("musX", 1), ("gamma_e", 4)):
        nmin, nmax, scale = _nudge(LB[grp], UB[grp], cfg.bit_depths[grp])
        code = (out[key].detach().numpy() - nmin) / scale
        assert np.abs(code - np.round(code)).max() < 1e-3
    nmin, nmax, _ = _nudge(LB[2], UB[2], 6)
    outside = (p["nu_e"] < nmin) | (p["nu_e"] > nmax)
    assert outside.any() and (~outside).any()
    assert float(g["nu_e"][torch.tensor(outside)].abs().max()) == 0          # straight-through mask
    live = out["indices"].numpy()
    assert float(g["nu_e"][live][torch.tensor(~outside[live])].abs().max()) > 0
    # pis are fake-quantised in mode 2 even without quantize_pis (smoe.py:474)
    code = out["pis"].detach().numpy() / (2.0 / 1023)
    assert np.abs(code - np.round(code)).max() < 1e-3


def test_mode2_with_wide_bounds_and_16_bits_approaches_mode0():
    p, x, t, kl, cfg = _rand_case(2, 1, False, False, False, 6)
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
    tp = {k: torch.tensor(v) for k, v in p.items()}
    out0, g0 = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg)
    cfg.quantization_mode, cfg.bit_depths = 2, [16] * 5
    cfg.lower_bounds, cfg.upper_bounds = [-8.0, -0.5, -0.5, -1.0, -2.0], [8.0, 1.5, 1.5, 1.0, 2.0]
    out2, g2 = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg)
    assert np.abs(out0["r_pre"].detach().numpy() - out2["r_pre"].detach().numpy()).max() < 2e-3
    assert out0["num_pi"] == out2["num_pi"]


def test_diff_center_and_kernel_count_norm():
    p, x, t, kl, cfg = _rand_case(2, 1, False, True, False, 7)
    tp = {k: torch.tensor(v) for k, v in p.items()}
    out, g = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg, pis_l1=0.5)
    grid = torch.tensor(np.random.RandomState(1).uniform(0, 1, p["musX"].shape))
    tp2 = dict(tp, musX=tp["musX"] - grid)
    cfg.use_diff_center = True
    out2, g2 = graph_grads(tp2, kl, torch.tensor(x), torch.tensor(t), cfg, pis_l1=0.5, musX_grid=grid)
    assert abs(float(out["loss"].detach()) - float(out2["loss"].detach())) < 1e-12
    np.testing.assert_allclose(g["musX"], g2["musX"], atol=1e-12)
    cfg.kernel_count_as_norm_l1 = True                                        # smoe.py:1022-1025
    out3, g3 = graph_grads(tp2, kl, torch.tensor(x), torch.tensor(t), cfg, pis_l1=0.5, musX_grid=grid)
    live = out3["indices"].numpy()
    num_pi = int((p["pis"] > 0).sum())
    np.testing.assert_allclose((g3["pis"] - g2["pis"])[live], 0.5 / num_pi - 0.5 / cfg.start_pis, atol=1e-12)


# ----------------------------------------------------------------------------------------------------
# SSIM loss (ssim_opt, smoe.py:981-1010) and overlap_of_batches (smoe.py:18-35, 909-923)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nd", [2, 3])
def test_torch_ssim_matches_the_numpy_restatement(nd):
    from oracle.graph import _symmetric_pad, custom_ssim_torch
    from oracle.ssim import smoe_ssim
    rs = np.random.RandomState(nd)
    shp = (14, 17, 3) if nd == 2 else (7, 12, 6, 3)
    a, b = rs.uniform(0, 1, shp), rs.uniform(0, 1, shp)
    per = custom_ssim_torch(_symmetric_pad(torch.tensor(a), nd), _symmetric_pad(torch.tensor(b), nd), nd).numpy()
    _, ref = smoe_ssim(a, b, use_yuv=False, dtype=np.float64)
    np.testing.assert_allclose(per, ref, rtol=1e-10)
    assert abs(float(custom_ssim_torch(_symmetric_pad(torch.tensor(a), nd), _symmetric_pad(torch.tensor(a), nd), nd)
                     .mean()) - 1) < 1e-12


def test_ssim_loss_and_overlap_in_the_oracle_model():
    from oracle.model import OracleAdam, OracleSmoe
    rs = np.random.RandomState(0)
    v, u = np.meshgrid(np.linspace(0, 1, 24), np.linspace(0, 1, 32), indexing="ij")
    img = np.clip(0.5 + 0.3 * np.sin(5 * u + 3 * v)[..., None] + 0.05 * rs.standard_normal((24, 32, 3)), 0, 1)
    img = (np.round(img * 255) / 255).astype(np.float32)
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, start_batches=4, dtype=torch.float64)
    o0 = OracleSmoe(img, kernels_per_dim=[4, 4], **kw)
    o1 = OracleSmoe(img, kernels_per_dim=[4, 4], overlap_of_batches=3, **kw)
    for o in (o0, o1):
        o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    r0 = o0.run_batched(train=True, update_reconstruction=True)
    r1 = o1.run_batched(train=True, update_reconstruction=True)
    # the halo is cropped before the loss: same loss, same gradients, same reconstruction ...
    assert abs(r0[0] - r1[0]) < 1e-12 and abs(r0[1] - r1[1]) < 1e-9
    for k in o0.last_grads:
        np.testing.assert_allclose(o0.last_grads[k], o1.last_grads[k], atol=1e-12)
    np.testing.assert_array_equal(o0.reconstruction_image, o1.reconstruction_image)
    # ... but the influence lists also see the halo pixels (and the zero padding at coordinate 0)
    assert all((b1 | b0 == b1).all() for b0, b1 in zip(o0.kernel_list_per_batch, o1.kernel_list_per_batch))
    assert sum(int(b1.sum() - b0.sum()) for b0, b1 in zip(o0.kernel_list_per_batch, o1.kernel_list_per_batch)) > 0
    # SSIM loss: 1 - weighted SSIM per batch; decreases under training
    o2 = OracleSmoe(img, kernels_per_dim=[4, 4], ssim_opt=True, overlap_of_batches=2, **kw)
    o2.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(0.1))
    l_first = o2.run_batched(train=True)[0]
    assert 0 < l_first < 1
    for _ in range(15):
        l_last = o2.run_batched(train=True)[0]
    assert l_last < l_first


# ----------------------------------------------------------------------------------------------------
# quantization_mode 3 (fake-quant with the min / max of the surviving kernels, smoe.py:497-531)
# ----------------------------------------------------------------------------------------------------
def test_mode3_ranges_follow_the_surviving_kernels_and_route_clipped_gradients():
    from oracle.graph import effective_params
    p, x, t, kl, cfg = _rand_case(2, 3, False, True, True, 8, K=12)
    cfg.quantization_mode, cfg.bit_depths = 3, [6, 7, 5, 10, 4]
    cfg.lower_bounds, cfg.upper_bounds = [0, 0, 0, 0.0, 0], [0, 0, 0, 2.0, 0]      # only the pis bounds are used
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
    p["nu_e"][1] = 7.0                                  # kernel 1 has pi < 0: it must not widen any range
    tp = {k: torch.tensor(v) for k, v in p.items()}
    eff = effective_params(tp, cfg)
    keep = p["pis"] > 1.0 / 1023                        # fake-quantised pi > 0
    # shifted form: the extremes of the kept rows are reproduced exactly, values lie on min + k * scale
    nu_k = p["nu_e"][keep]
    q = eff["nu_e"].numpy()[keep]
    assert q.min() == nu_k.min() and abs(q.max() - nu_k.max()) < 1e-6
    scale = (np.float32(nu_k.max()) - np.float32(nu_k.min())) / np.float32(31)
    code = (q - nu_k.min()) / scale
    assert np.abs(code - np.round(code)).max() < 1e-3 and np.abs(q - nu_k).max() <= scale / 2 * 1.001
    # plain form (gamma_e, 4 bits): zero is a code; elements beyond the nudged range are clamped
    ga = eff["gamma_e"].numpy()[keep]
    gscale = (p["gamma_e"][keep].max() - p["gamma_e"][keep].min()) / 15
    assert np.abs(ga / gscale - np.round(ga / gscale)).max() < 1e-3
    # gradients: shifted groups pass straight through; in a plain group the gradient of a clamped element goes
    # to the extreme element of the kept rows, and the total is conserved
    out, g = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg, pis_l1=0.1)
    cfg0 = GraphCfg(**{**cfg.__dict__, "quantization_mode": 0})
    eff_leaf = {k: v.detach().clone().requires_grad_(True) for k, v in eff.items()}
    eff_leaf["pis"] = out["pis"].new_tensor(p["pis"]).requires_grad_(True)
    cfg0.quantize_pis = True
    out0, g0 = graph_grads({k: v.detach() for k, v in eff_leaf.items()}, kl, torch.tensor(x), torch.tensor(t), cfg0,
                           pis_l1=0.1)
    np.testing.assert_allclose(g["nu_e"], g0["nu_e"], atol=1e-14)
    np.testing.assert_allclose(torch.diagonal(g["A_diagonal"], dim1=1, dim2=2),
                               torch.diagonal(g0["A_diagonal"], dim1=1, dim2=2), atol=1e-14)
    assert abs(float(g["gamma_e"].sum()) - float(g0["gamma_e"].sum())) < 1e-12 * max(1, float(g0["gamma_e"].abs().sum()))
    assert abs(float(out["loss"].detach()) - float(out0["loss"].detach())) < 1e-12


def test_mode3_all_zero_group_is_passed_through():
    p, x, t, kl, cfg = _rand_case(2, 1, False, True, False, 9)
    p["A_corr"][:] = 0.0
    p["gamma_e"][:] = 0.0
    cfg.quantization_mode, cfg.bit_depths = 3, [8, 8, 8, 10, 8]
    cfg.lower_bounds, cfg.upper_bounds = [0, 0, 0, 0.0, 0], [0, 0, 0, 2.0, 0]
    tp = {k: torch.tensor(v.astype(np.float32).astype(np.float64)) for k, v in p.items()}
    out, g = graph_grads(tp, kl, torch.tensor(x), torch.tensor(t), cfg)
    assert float(out["gamma_e"].detach().abs().max()) == 0
    assert float(g["gamma_e"].abs().max()) > 0 and float(g["A_corr"][:, 1, 0].abs().max()) > 0
