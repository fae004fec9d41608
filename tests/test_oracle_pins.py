"""Independent pins for the oracle pieces that restate TensorFlow ops (TensorFlow itself is not installable here, so
`oracle/graph.py` and `oracle/ssim.py` are "parity unpinned" against TF).  Each test checks the
restatement against an implementation that shares no code with it:

  * fake quantisation (`_nudge`, `fq_values`; smoe.py:475, 899 -> TF FakeQuantWithMinMaxArgs) against
    `torch.fake_quantize_per_tensor_affine`, PyTorch's own implementation of the same affine nudged quantiser;
  * the SSIM window, the VALID moments and the SSIM map (`oracle/ssim.py`; ops/image_ops_impl.py:131-233) against
    `scipy.ndimage.correlate1d` Gaussian moments and the textbook SSIM formula;
  * the gradient rules written into the graph's ops (`graph_grads`; tf.gradients at smoe.py:1148) against central finite
    differences of the float64 forward, on inputs kept away from the two discontinuities (gate threshold, output
    rounding), where the loss is a smooth function of the parameters.
"""
import numpy as np
import pytest
import torch

from oracle.graph import GraphCfg, PARAM_KEYS, _nudge, fq_values, graph_forward, graph_grads
from oracle import ssim as ossim


@pytest.mark.parametrize("mn,mx,bits", [(0.0, 1.0, 8), (0.0, 2.0, 10), (-2500.0, 2500.0, 20), (-0.3, 1.3, 18),
                                        (-5.0, 5.0, 6), (-32.0, 32.0, 10), (0.1, 0.9, 4), (-1.0, -0.25, 5)])
def test_fake_quant_matches_torch_affine_quantiser(mn, mx, bits):
    """TF's Nudge(): scale = (max-min)/(2^b-1), zero point rounded and clamped, nudged_min = -zp*scale; the values
    are floor((clamp(x) - nudged_min) / scale + 0.5) * scale + nudged_min.  torch's op computes
    (clamp(nearbyint(x / scale) + zp, qmin, qmax) - zp) * scale: the same quantiser except on exact half-code ties
    (floor(.+0.5) vs round-half-even), which the comparison skips."""
    nmin, nmax, scale = _nudge(mn, mx, bits)
    qmax = 2 ** bits - 1
    zp = int(round(-nmin / scale))
    assert 0 <= zp <= qmax
    assert abs(-zp * scale - nmin) <= 1e-6 * max(1.0, abs(nmin)) and abs((qmax - zp) * scale - nmax) <= 1e-6 * max(1.0, abs(nmax))
    # the nudged range contains a representable zero and has the width of the requested one
    assert abs((nmax - nmin) - (mx - mn)) <= 1e-6 * (mx - mn)
    assert nmin <= 0.0 <= nmax or mn > 0 or mx < 0
    rs = np.random.RandomState(bits)
    x = np.concatenate([rs.uniform(mn - 0.2 * (mx - mn), mx + 0.2 * (mx - mn), 20000),
                        nmin + scale * rs.randint(0, qmax + 1, 2000)]).astype(np.float64)
    ours = fq_values(torch.tensor(x), nmin, nmax, scale).numpy()
    ref = torch.fake_quantize_per_tensor_affine(torch.tensor(x), float(scale), zp, 0, qmax).numpy()
    code = (np.clip(x, nmin, nmax) - nmin) / scale
    # TF multiplies by the float32 reciprocal inv_scale = 1.0f / scale (relative error 6e-8), torch divides: at code
    # ~2^bits the two can disagree within 2^bits * 1.2e-7 of a half-code boundary
    tie = np.abs(code - np.floor(code) - 0.5) < max(1e-6, 2.0 ** bits * 2.4e-7)
    assert tie.mean() < 0.6
    # same CODE everywhere away from ties; the VALUE is TF's float32(k * scale) + nudged_min, i.e. torch's float64
    # value rounded to float32
    np.testing.assert_array_equal(np.round((ours[~tie] - nmin) / scale), np.round((ref[~tie] - nmin) / scale))
    np.testing.assert_allclose(ours[~tie], ref[~tie], rtol=0, atol=2.4e-7 * max(1.0, abs(mx), abs(mn)))
    # and the quantised values are codes of the nudged grid
    k = (ours - nmin) / scale
    tolk = 1e-4 + 2.0 ** bits * 1.2e-7       # values are float32(k * scale) + float32(nudged_min)
    assert np.abs(k - np.round(k)).max() < tolk and k.min() >= -tolk and k.max() <= qmax + tolk


def test_fake_quant_float32_matches_torch_float32():
    """The float32 form the CUDA epilogue reproduces (k * f32(1/255)): same codes as torch's float32 op away from ties."""
    nmin, nmax, scale = _nudge(0.0, 1.0, 8)
    x = torch.tensor(np.random.RandomState(3).uniform(-0.1, 1.1, 50000).astype(np.float32))
    ours = fq_values(x, nmin, nmax, scale)
    ref = torch.fake_quantize_per_tensor_affine(x, float(np.float32(scale)), 0, 0, 255)
    code = np.clip(x.numpy().astype(np.float64), 0, 1) * 255
    ok = np.abs(code - np.floor(code) - 0.5) > 1e-4
    assert np.array_equal(np.round(ours.numpy()[ok] * 255), np.round(ref.numpy()[ok] * 255))


def _scipy_ssim(a, b, ndim, max_val=1.0):
    """SSIM per channel with scipy.ndimage Gaussian moments (size 11, sigma 1.5), VALID region, float64."""
    from scipy import ndimage
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
    g /= g.sum()
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2

    def blur(x):
        for ax in range(ndim):
            x = ndimage.correlate1d(x, g, axis=ax, mode="constant")
        return x[tuple(slice(5, n - 5) for n in x.shape[:ndim])]

    out = []
    for c in range(a.shape[-1]):
        x, y = a[..., c].astype(np.float64), b[..., c].astype(np.float64)
        mx, my = blur(x), blur(y)
        sxx, syy, sxy = blur(x * x) - mx * mx, blur(y * y) - my * my, blur(x * y) - mx * my
        l = (2 * mx * my + c1) / (mx * mx + my * my + c1)
        cs = (2 * sxy + c2) / (sxx + syy + c2)
        out.append((l * cs).mean())
    return np.array(out)


@pytest.mark.parametrize("shape", [(40, 52, 1), (33, 47, 3), (20, 24, 18, 3)])
def test_ssim_matches_scipy_gaussian_moments(shape):
    rs = np.random.RandomState(len(shape))
    nd = len(shape) - 1
    a = rs.uniform(0, 1, shape)
    b = np.clip(a + 0.1 * rs.standard_normal(shape), 0, 1)
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
    g /= g.sum()
    outer = g[:, None] * g[None, :] if nd == 2 else g[:, None, None] * g[None, :, None] * g[None, None, :]
    # the reference's softmax over the 11^d logits == outer product of the normalised 1-D Gaussians
    np.testing.assert_allclose(ossim.gauss_window(nd, dtype=np.float64), outer, rtol=1e-12)
    got = ossim.custom_ssim(a, b, max_val=1.0, ndim=nd, dtype=np.float64)
    np.testing.assert_allclose(got, _scipy_ssim(a, b, nd), rtol=1e-9, atol=1e-12)
    # the caller's SYMMETRIC pad by 5 (smoe.py:994-1004) == numpy 'symmetric' padding, then the same VALID SSIM
    pad = [(5, 5)] * nd + [(0, 0)]
    ap, bp = np.pad(a, pad, mode="symmetric"), np.pad(b, pad, mode="symmetric")
    want = _scipy_ssim(ap, bp, nd)
    per = want
    s_ref = float((per * np.array([6, 1, 1]) / 8).sum()) if shape[-1] == 3 else float(per[0])
    s_got, per_got = ossim.smoe_ssim(a, b, use_yuv=True, dtype=np.float64)
    np.testing.assert_allclose(per_got, want, rtol=1e-9)
    assert abs(s_got - s_ref) < 1e-10


def _smooth_case(d, C, tic, det, seed):
    """A case whose gates are all far from the threshold and whose parameters sit in the smooth region."""
    rs = np.random.RandomState(seed)
    K, N = 7, 60
    p = dict(pis=rs.uniform(0.2, 1, K), musX=rs.uniform(0.1, 0.9, (K, d)), gamma_e=rs.normal(0, .3, (K, d, C)),
             nu_e=rs.uniform(0.2, 0.8, (K, C)))
    Ad = np.zeros((K, d, d)); Ac = np.zeros((K, d, d))
    for i in range(d):
        Ad[:, i, i] = rs.uniform(1.5, 3.0, K)
        for j in range(i):
            Ac[:, i, j] = rs.normal(0, 0.5, K)
    p["A_diagonal"], p["A_corr"] = Ad, Ac
    x = rs.uniform(0, 1, (N, d))
    t = rs.uniform(0, 1, (N, C))
    cfg = GraphCfg(dim_domain=d, num_channels=C, use_determinant=det, train_inverse_cov=tic, use_yuv=(C == 3),
                   start_pis=K)
    return p, x, t, cfg


@pytest.mark.parametrize("d,C,tic,det", [(2, 1, False, True), (2, 3, False, False), (3, 3, False, True), (2, 1, True, True)])
def test_graph_gradient_rules_match_finite_differences(d, C, tic, det):
    """f(theta) = sum(G * r_pre(theta)) for a fixed random cotangent G: autograd through the graph's ops (incl. the
    hand-written rules for max(1e-11, S), the threshold mask and the A assembly) vs central differences of the
    float64 forward.  The threshold mask is piecewise constant, so the check runs where no gate is within 1e-2
    (relative) of tau and verifies that the perturbation flips none."""
    p, x, t, cfg = _smooth_case(d, C, tic, det, seed=10 * d + C)
    kl = np.ones(p["pis"].shape[0], bool)
    tx, tt = torch.tensor(x), torch.tensor(t)
    G = torch.tensor(np.random.RandomState(1).standard_normal(t.shape))

    def f(params):
        out = graph_forward({k: torch.as_tensor(v, dtype=torch.float64) for k, v in params.items()}, kl, tx, tt, cfg)
        return (G * out["r_pre"]).sum(), out

    leaf = {k: torch.tensor(p[k], dtype=torch.float64, requires_grad=True) for k in PARAM_KEYS}
    val, out0 = f(leaf)
    tau = 0.5 / 256
    margin = (out0["w_full"].detach() / tau - 1).abs().min().item()
    assert margin > 1e-2
    grads = torch.autograd.grad(val, [leaf[k] for k in PARAM_KEYS], allow_unused=True)
    mask0 = (out0["w_full"].detach() > tau)
    rs = np.random.RandomState(2)
    for k, g in zip(PARAM_KEYS, grads):
        g = torch.zeros_like(leaf[k]) if g is None else g
        flat = p[k].reshape(-1)
        for j in rs.choice(flat.size, min(flat.size, 12), replace=False):
            if k == "A_diagonal" or k == "A_corr":
                idx = np.unravel_index(j, p[k].shape)
                structural = (idx[1] != idx[2]) if k == "A_diagonal" else (idx[1] <= idx[2])
            else:
                structural = False
            h = 1e-6 * max(1.0, abs(flat[j]))
            vals = []
            for sgn in (+1, -1):
                q = {kk: vv.copy() for kk, vv in p.items()}
                q[k].reshape(-1)[j] += sgn * h
                v, o = f(q)
                assert bool(((o["w_full"] > tau) == mask0).all())
                vals.append(float(v))
            fd = (vals[0] - vals[1]) / (2 * h)
            an = float(g.reshape(-1)[j])
            if structural:               # entries the A assembly never reads (band_part, smoe.py:732-733)
                assert an == 0.0 and abs(fd) < 1e-9
            else:
                assert abs(fd - an) <= 2e-6 * max(abs(an), abs(fd)) + 1e-9, (k, j, fd, an)


def test_loss_cotangent_formula_is_the_derivative_of_the_loss_in_resq():
    """dL/d(resq) = 2 (|diff| - eps) sign(diff) * channel weight / N (smoe.py:931-937): the loss is a smooth function
    of the quantiser OUTPUT; checked by finite differences on resq_override (the straight-through estimator then copies
    this cotangent onto r where 0 <= r <= 1, which is a definition, not something differentiable)."""
    p, x, t, cfg = _smooth_case(2, 3, False, True, seed=5)
    kl = np.ones(p["pis"].shape[0], bool)
    tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
    tx, tt = torch.tensor(x), torch.tensor(t)
    out = graph_forward(tp, kl, tx, tt, cfg)
    resq = out["resq"].detach().clone()
    diff = (resq - tt).numpy()
    eps = cfg.margin / 2 ** cfg.precision
    N = x.shape[0]
    cw = np.array([6 / 8, 1 / 8, 1 / 8]) / N
    want = 2 * (np.abs(diff) - eps) * np.sign(diff) * cw[None]
    rs = np.random.RandomState(0)
    for _ in range(10):
        n, c = rs.randint(N), rs.randint(3)
        h = 1e-7
        vals = []
        for sgn in (+1, -1):
            rq = resq.clone()
            rq[n, c] += sgn * h
            vals.append(float(graph_forward(tp, kl, tx, tt, cfg, resq_override=rq)["loss"]))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - want[n, c]) < 1e-7 * max(1.0, abs(want[n, c])) + 1e-10
