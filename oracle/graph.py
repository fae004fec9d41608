"""Restatement of the reference's TensorFlow graph (single-model path) on torch-CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED for this file
(TensorFlow is not available; see the package docstring).

Follows, op for op:
  * variable assembly + pi-mask compaction ...... smoe.py:466-480, 732-753
  * Mahalanobis / kernel value .................. smoe.py:777-817
      einsum 'abli,alm,anm,abnj->ab' (ops/special_math_ops.py:36-149)
  * gating, threshold, influence list, argmax ... smoe.py:819-838
  * experts + mixture + clip .................... smoe.py:840-858
  * output fake-quant, loss, mse ................ smoe.py:899-937, 1012-1056
Gradients come from torch autograd through ops whose backward rules are written to
match the TF rules at those call sites (SURVEY.md section 8c "TF quirks"), and from an
independent closed-form backward (`closed_form_grads`) used as the kernel spec.

The single-model math is the reference's model-0 branch with every kernel assigned
to model 0 (smoe.py:764-765, 796): SURVEY decision D1.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

PARAM_KEYS = ("pis", "musX", "A_diagonal", "A_corr", "gamma_e", "nu_e")


@dataclass
class GraphCfg:
    """Static configuration of the graph; names follow Smoe.__init__ (smoe.py:38-41)."""
    dim_domain: int = 2
    num_channels: int = 1
    precision: int = 8
    margin: float = 0.5
    use_determinant: bool = False
    train_inverse_cov: bool = True
    use_yuv: bool = True
    train_gammas: bool = True
    only_y_gamma: bool = False
    use_diff_center: bool = False
    quantize_pis: bool = False
    quantization_mode: int = 0
    lower_bounds: Optional[list] = None
    upper_bounds: Optional[list] = None
    bit_depths: Optional[list] = None
    kernel_count_as_norm_l1: bool = False
    ssim_opt: bool = False      # SSIM instead of the squared-error loss (smoe.py:981-1010)
    start_pis: int = 0          # smoe.py:264 (K at construction)
    einsum_mode: str = "einsum"  # "einsum" | "broadcast" (TF materialisation strategy)


# ----------------------------------------------------------------------------------
# TF op restatements with TF gradient rules
# ----------------------------------------------------------------------------------

def _nudge(mn: float, mx: float, bits: int):
    """TF `Nudge()` of fake_quant_with_min_max_args (narrow_range=False), float32 math."""
    f = np.float32
    quant_min, quant_max = f(0.0), f(2 ** bits - 1)
    scale = (f(mx) - f(mn)) / (quant_max - quant_min)
    zp_from_min = quant_min - f(mn) / scale
    if zp_from_min < quant_min:
        nudged_zp = quant_min
    elif zp_from_min > quant_max:
        nudged_zp = quant_max
    else:
        nudged_zp = f(np.floor(zp_from_min + f(0.5)))   # std::round for positives
    nudged_min = (quant_min - nudged_zp) * scale
    nudged_max = (quant_max - nudged_zp) * scale
    return float(nudged_min), float(nudged_max), float(scale)


def fq_values(x, nmin, nmax, scale):
    """Output values of TF's FakeQuantWithMinMaxArgs functor (TF 1.13-1.15 form with
    inv_scale = 1.0f/scale).  The code k is decided in x's dtype; the returned VALUE is always
    what float32 TF would store, float32(k * scale) + nmin, so that `resq - target` has TF's
    sign even in the float64 oracle (126 of the 256 8-bit codes differ from k/255 by one ulp,
    which makes sign(diff) non-zero on exactly-matched pixels)."""
    f = np.float32
    inv_scale = float(f(1.0) / f(scale))
    c = torch.clamp(x, nmin, nmax)
    if x.dtype == torch.float32:
        k = torch.floor((c - nmin) * inv_scale + 0.5)
        return k * float(f(scale)) + nmin
    k = torch.floor((c - nmin) * inv_scale + 0.5)
    return (k.to(torch.float32) * float(f(scale)) + float(f(nmin))).to(x.dtype)


class _FakeQuantArgs(torch.autograd.Function):
    """tf.quantization.fake_quant_with_min_max_args (smoe.py:475, 899).

    forward : floor((clamp(x, nmin, nmax) - nmin) * inv_scale + 0.5) * scale + nmin
    backward: gradient passes iff nmin <= x <= nmax (straight-through)."""

    @staticmethod
    def forward(ctx, x, nmin, nmax, scale):
        ctx.save_for_backward((x >= nmin) & (x <= nmax))
        return fq_values(x, nmin, nmax, scale)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask.to(g.dtype), None, None, None


def fake_quant_args(x, mn, mx, bits):
    nmin, nmax, scale = _nudge(mn, mx, bits)
    return _FakeQuantArgs.apply(x, nmin, nmax, scale)


def fake_quant_var(x, mn, mx, bits):
    """fake_quant_with_min_max_args on a VARIABLE (smoe.py:482-496).  TF variables are float32, so the
    code is decided on the float32 value whatever dtype the oracle computes in."""
    nmin, nmax, scale = _nudge(mn, mx, bits)
    x32 = x.detach().to(torch.float32)
    q = fq_values(x32, nmin, nmax, scale).to(x.dtype)
    mask = ((x32 >= nmin) & (x32 <= nmax)).to(x.dtype)
    return x * mask + (q - x * mask).detach()                   # value q, gradient = straight-through mask


class _FakeQuantVarsMasked(torch.autograd.Function):
    """tf.quantization.fake_quant_with_min_max_vars whose min / max are reduce_min / reduce_max over the rows
    with pis > 0 (smoe.py:497-531), in the two forms the reference uses:
      plain   q = fq(x; min, max)                      (A_corr, musX, gamma_e)
      shifted q = fq(x - min; 0, max - min) + min      (A_diagonal over its diagonal entries, nu_e)
    TF kernel semantics (FakeQuantWithMinMaxVarsFunctor / ...GradientFunctor, float32): min == max == 0 returns
    zeros and passes the gradient; otherwise Nudge(), clamp, round; the gradient passes where
    nudged_min <= x <= nudged_max, the gradients of elements below / above go to `min` / `max`, and from there
    through reduce_min / reduce_max (equal shares among ties) back to the extreme kept elements.  For the shifted
    form all of that cancels to a plain pass-through for the kept elements (d q_i / d min = -1 + 1)."""

    @staticmethod
    def forward(ctx, x, keep_rows, bits, shifted, sel):
        f = np.float32
        x32 = x.detach().to(torch.float32)
        keep = keep_rows.reshape((-1,) + (1,) * (x.dim() - 1)).expand_as(x32)
        pool = keep if sel is None else (keep & sel)          # elements the reduce_min / reduce_max run over
        if not bool(pool.any()):
            ctx.mode = "pass"
            return x.clone()
        mn, mx = f(x32[pool].min().item()), f(x32[pool].max().item())
        lo, hi = (f(0.0), f(mx - mn)) if shifted else (mn, mx)
        shift = mn if shifted else f(0.0)
        if lo == 0 and hi == 0:
            ctx.mode = "pass"
            return (torch.zeros_like(x32) + float(shift)).to(x.dtype)
        nmin, nmax, scale = _nudge(float(lo), float(hi), bits)
        xs = x32 - float(shift)
        q = fq_values(xs, nmin, nmax, scale) + float(shift)
        if shifted:
            ctx.mode = "pass"
        else:
            ctx.mode = "route"
            below, above = xs < nmin, xs > nmax
            ctx.save_for_backward(below, above, pool & (x32 == float(mn)), pool & (x32 == float(mx)))
        return q.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        if ctx.mode == "pass":
            return g, None, None, None, None
        below, above, is_min, is_max = ctx.saved_tensors
        gin = g * (~(below | above)).to(g.dtype)
        gin = gin + is_min.to(g.dtype) * ((g * below.to(g.dtype)).sum() / is_min.sum())
        gin = gin + is_max.to(g.dtype) * ((g * above.to(g.dtype)).sum() / is_max.sum())
        return gin, None, None, None, None


def effective_params(params, cfg, train_musx=True):
    """The q* tensors of smoe.py:482-538: mode 2 fake-quantises every variable with the fixed bounds (order
    A, musX, nu_e, pis, gamma_e: smoe_test.py:302-309), mode 3 with the min / max over the kernels whose
    quantised pi is positive (pis keep their fixed bounds); modes 0/1 use the variables as they are.
    `pis` itself is handled by the caller (smoe.py:474-478)."""
    if cfg.quantization_mode == 3:
        bd = cfg.bit_depths
        qpis = fake_quant_args(params["pis"].detach(), cfg.lower_bounds[3], cfg.upper_bounds[3], bd[3])
        keep = qpis > 0                                                          # pis_mask, smoe.py:480
        d = params["A_diagonal"].shape[-1]
        eye = torch.eye(d, dtype=torch.bool).expand_as(params["A_diagonal"])
        out = dict(params)
        fqv = _FakeQuantVarsMasked.apply
        out["A_diagonal"] = fqv(params["A_diagonal"], keep, bd[0], True, eye)    # smoe.py:506-511 (diag_part)
        out["A_corr"] = fqv(params["A_corr"], keep, bd[0], False, None)          # smoe.py:512-515
        if train_musx:
            out["musX"] = fqv(params["musX"], keep, bd[1], False, None)          # smoe.py:516-522
        out["nu_e"] = fqv(params["nu_e"], keep, bd[2], True, None)               # smoe.py:524-527
        out["gamma_e"] = fqv(params["gamma_e"], keep, bd[4], False, None)        # smoe.py:529-532
        return out
    if cfg.quantization_mode != 2:
        return params
    lb, ub, bd = cfg.lower_bounds, cfg.upper_bounds, cfg.bit_depths
    out = dict(params)
    out["A_diagonal"] = fake_quant_var(params["A_diagonal"], lb[0], ub[0], bd[0])
    out["A_corr"] = fake_quant_var(params["A_corr"], lb[0], ub[0], bd[0])
    out["musX"] = fake_quant_var(params["musX"], lb[1], ub[1], bd[1])
    out["nu_e"] = fake_quant_var(params["nu_e"], lb[2], ub[2], bd[2])
    out["gamma_e"] = fake_quant_var(params["gamma_e"], lb[4], ub[4], bd[4])
    return out


class _ClipByValue01(torch.autograd.Function):
    """tf.clip_by_value(x, 0, 1) (smoe.py:857): gradient passes iff 0 <= x <= 1."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward((x >= 0) & (x <= 1))
        return torch.clamp(x, 0.0, 1.0)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask.to(g.dtype)


def _gauss_window_torch(ndim, dtype, size=11, sigma=1.5):
    """ops/image_ops_impl.py:131-151: softmax of -x^2/(2 sigma^2) over the full size^ndim window."""
    c = torch.arange(size, dtype=dtype) - (size - 1) / 2.0
    g = -(c * c) / (2.0 * sigma * sigma)
    full = g[:, None] + g[None, :] if ndim == 2 else g[:, None, None] + g[None, :, None] + g[None, None, :]
    return torch.softmax(full.reshape(-1), dim=0).reshape(full.shape)


def _symmetric_pad(x, ndim, pad=5):
    """tf.pad(..., "SYMMETRIC") on the first ndim axes (smoe.py:994-1003); differentiable gather."""
    for ax in range(ndim):
        idx = np.pad(np.arange(x.shape[ax]), pad, mode="symmetric")
        x = x.index_select(ax, torch.as_tensor(idx, dtype=torch.long))
    return x


def custom_ssim_torch(img1, img2, ndim, max_val=1.0):
    """ops/image_ops_impl.py:235-293 (helpers :77-233) on already padded (spatial..., C) tensors: depthwise
    VALID correlation with the Gaussian window, luminance * contrast-structure, mean over positions -> (C,)."""
    import torch.nn.functional as F
    win = _gauss_window_torch(ndim, img1.dtype)[None, None]
    conv = F.conv2d if ndim == 2 else F.conv3d

    def reducer(x):                                            # channels -> batch (image_ops_impl.py:206-222)
        return conv(x.movedim(-1, 0)[:, None], win)[:, 0]
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mean0, mean1 = reducer(img1), reducer(img2)
    num0 = mean0 * mean1 * 2.0
    den0 = mean0 * mean0 + mean1 * mean1
    lum = (num0 + c1) / (den0 + c1)
    num1 = reducer(img1 * img2) * 2.0
    den1 = reducer(img1 * img1 + img2 * img2)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).reshape(lum.shape[0], -1).mean(dim=1)


def assemble_A(A_diagonal, A_corr, train_inverse_cov):
    """smoe.py:732-735: band_part(A_diag,0,0) + strict-lower(A_corr) (+ its transpose)."""
    if A_diagonal.dim() == 1:        # radial_as: tile the scalar, band_part keeps the diagonal (smoe.py:714-721)
        d = A_corr.shape[-1]
        diag = A_diagonal[:, None, None] * torch.eye(d, dtype=A_diagonal.dtype)
    else:
        diag = torch.diag_embed(torch.diagonal(A_diagonal, dim1=-2, dim2=-1))
    low = torch.tril(A_corr, diagonal=-1)
    A = diag + low
    if train_inverse_cov:
        A = A + low.transpose(-1, -2)
    return A


def _maha(x_sub_mu, A, train_inverse_cov, mode):
    """smoe.py:791-797.  x_sub_mu: (K,N,d), A: (K,d,d) -> (K,N)."""
    if mode == "broadcast":
        # the reference's exponential_space_einsum: broadcast every index into one tensor,
        # multiply, reduce_sum (special_math_ops.py:144-149) -- (K,N,d,d[,d]) temporaries.
        if train_inverse_cov:     # 'abli,alm,abmj->ab'
            prod = x_sub_mu[:, :, :, None] * A[:, None, :, :] * x_sub_mu[:, :, None, :]
            return prod.sum(dim=(2, 3))
        # 'abli,alm,anm,abnj->ab' : axes (a,b,l,m,n)
        prod = (x_sub_mu[:, :, :, None, None] * A[:, None, :, :, None]
                * A.transpose(1, 2)[:, None, None, :, :] * x_sub_mu[:, :, None, None, :])
        return prod.sum(dim=(2, 3, 4))
    if train_inverse_cov:
        return torch.einsum("knl,klm,knm->kn", x_sub_mu, A, x_sub_mu)
    y = torch.einsum("kl m,knl->knm".replace(" ", ""), A, x_sub_mu)
    return (y * y).sum(-1)


def graph_forward(params: Dict[str, torch.Tensor], kernel_list, domain, target, cfg: GraphCfg,
                  pis_l1=0.0, u_l1=0.0, loss_weights=None, musX_grid=None,
                  feed: Optional[Dict[str, torch.Tensor]] = None,
                  resq_override: Optional[torch.Tensor] = None, crop=None,
                  train_musx: bool = True) -> Dict[str, torch.Tensor]:
    """One `session.run` of the reference graph on one batch of pixels.

    params : K_all-sized variables (pis, musX, A_diagonal, A_corr, gamma_e, nu_e)
    kernel_list : (K_all,) bool feed (smoe.py:552)
    domain : (N,d) pixel coordinates, target : (N,C) colours
    feed : optional {A, musX, nu_e, gamma_e, pis} fed *over* the compacted tensors
           (with_quantized_params, smoe.py:1688-1689)
    crop : optional (window shape incl. halo, overlap): the fed pixels are a window padded by `overlap` on every
           side (smoe.py:18-35); the halo is cropped before the loss (smoe.py:909-923, 985-991)
    resq_override : optional (N,C) values to use as the fake-quant OUTPUT (the straight-through
           gradient path is unchanged).  Lets a test evaluate the gradient conditional on another
           implementation's rounding decisions, which differ legitimately for pixels that sit
           within float32 noise of a rounding boundary.
    """
    dt = domain.dtype
    d, C = cfg.dim_domain, cfg.num_channels
    if cfg.quantization_mode > 3:
        raise ValueError("quantization_mode must be 0..3")

    pis_var = params["pis"]
    # smoe.py:474-480
    if cfg.quantize_pis or cfg.quantization_mode >= 2:
        qpis = fake_quant_args(pis_var, cfg.lower_bounds[3], cfg.upper_bounds[3], cfg.bit_depths[3])
    else:
        qpis = pis_var
    pis_mask = qpis > 0
    bool_mask = torch.as_tensor(kernel_list, dtype=torch.bool) & pis_mask.detach()
    indices = torch.nonzero(bool_mask).flatten()

    params = effective_params(params, cfg, train_musx)              # smoe.py:482-538 (modes 2, 3), else identity
    gamma_all = params["gamma_e"]
    if cfg.use_yuv and cfg.train_gammas and cfg.only_y_gamma:       # smoe.py:725-729
        gmask = torch.zeros(d, C, dtype=dt)
        gmask[:, 0] = 1
        gamma_all = gamma_all * gmask
    A_all = assemble_A(params["A_diagonal"], params["A_corr"], cfg.train_inverse_cov)

    musX_all = params["musX"] + musX_grid if cfg.use_diff_center else params["musX"]
    musX = musX_all[bool_mask]
    nu_e = params["nu_e"][bool_mask]
    gamma_e = gamma_all[bool_mask]
    A = A_all[bool_mask]
    pis = qpis[bool_mask]
    if feed is not None:
        musX, nu_e, gamma_e, A, pis = (feed["musX"], feed["nu_e"], feed["gamma_e"], feed["A"], feed["pis"])

    x_sub_mu = domain[None, :, :] - musX[:, None, :]                       # (K,N,d) smoe.py:777-782
    maha = _maha(x_sub_mu, A, cfg.train_inverse_cov, cfg.einsum_mode)       # smoe.py:791-805
    n_exp = torch.exp(-0.5 * maha)                                           # smoe.py:807
    if cfg.use_determinant:                                                  # smoe.py:809-815
        n_div = torch.prod(torch.diagonal(A, dim1=-2, dim2=-1), dim=-1)
        n_dis = math.sqrt((2 * math.pi) ** d)
        Nk = (n_div / n_dis)[:, None] * n_exp
    else:
        Nk = n_exp
    n_w = Nk * pis[:, None]                                                  # smoe.py:819
    n_w_sum = n_w.sum(dim=0)
    floor = torch.tensor(10e-12, dtype=dt)
    S = torch.where(n_w_sum > floor, n_w_sum, floor)                         # smoe.py:821 (grad iff S>1e-11)
    w_full = n_w / S                                                         # smoe.py:823
    tau = 0.5 * 1 / (2 ** cfg.precision)                                     # smoe.py:825
    infl = (w_full > tau)
    w = w_full * infl.to(dt)                                                 # smoe.py:827 (no renormalisation)

    kernel_list_batch = infl.sum(dim=1) > 0                                  # smoe.py:829
    indices_infl = indices[kernel_list_batch]                                # smoe.py:836
    if kernel_list_batch.any():
        w_sel = w[kernel_list_batch]
        # tf.argmax: first maximal index
        arg_local = torch.argmax(w_sel.detach(), dim=0)
        w_e_max = indices_infl[arg_local]
        w_e_max_local = arg_local
    else:
        w_e_max = torch.zeros(domain.shape[0], dtype=torch.long)
        w_e_max_local = w_e_max

    if cfg.train_gammas:                                                     # smoe.py:841-846
        sloped = torch.einsum("kdc,nd->ckn", gamma_e, domain)
        experts = sloped + nu_e.t()[:, :, None]
    else:                                                                    # smoe.py:848
        experts = nu_e.t()[:, :, None].expand(C, nu_e.shape[0], domain.shape[0])
    r_pre = (w[None, :, :] * experts).sum(dim=1)                             # (C,N)
    res = _ClipByValue01.apply(r_pre).t()                                    # smoe.py:857-858 -> (N,C)

    resq = fake_quant_args(res, 0.0, 1.0, cfg.precision)                     # smoe.py:899
    if resq_override is not None:
        resq = resq + (resq_override.to(dt) - resq).detach()
    diff = resq - target                                                     # smoe.py:905
    sq = diff * diff
    err_map = sq.mean(dim=1)                                                 # smoe.py:906
    sampl_prob = err_map / err_map.sum()
    inner = None
    if crop is not None and crop[1] > 0:                                     # smoe.py:909-923
        shp, ov = tuple(crop[0]), int(crop[1])
        inner = (slice(ov, -ov),) * d
        diff = diff.reshape(shp + (C,))[inner].reshape(-1, C)
        sq = diff * diff
    mse = sq.mean()                                                          # smoe.py:927
    if not cfg.ssim_opt:
        eps = cfg.margin * 1 / (2 ** cfg.precision)                          # smoe.py:931
        # (with a halo HEAD multiplies the cropped loss by un-cropped weights and fails; intent: weights of ones)
        lw = torch.ones(diff.shape[0], 1, dtype=dt) if loss_weights is None else loss_weights
        loss_px = torch.clamp_min((diff.abs() - eps) ** 2, 0.0) * lw         # smoe.py:932
        if cfg.use_yuv:                                                      # smoe.py:933-935
            loss_pixel = 6 / 8 * loss_px[:, 0].mean() + 1 / 8 * loss_px[:, 1:].mean(dim=0).sum()
        else:
            loss_pixel = loss_px.mean()                                      # smoe.py:937
    else:                                                                    # smoe.py:981-1010
        assert crop is not None, "ssim_opt needs the window shape"
        shp = tuple(crop[0])
        res_img, tgt_img = resq.reshape(shp + (C,)), target.reshape(shp + (C,))
        if inner is not None:
            res_img, tgt_img = res_img[inner], tgt_img[inner]
        ssim_c = custom_ssim_torch(_symmetric_pad(res_img, d), _symmetric_pad(tgt_img, d), d)
        if cfg.use_yuv:
            ssim = (ssim_c * torch.tensor([6.0, 1.0, 1.0], dtype=dt)[:max(C, 1)]).sum() / 8 if C == 3 \
                else (ssim_c * torch.tensor([6.0, 1.0, 1.0], dtype=dt)).sum() / 8
        else:
            ssim = ssim_c.mean()
        loss_pixel = 1 - ssim

    num_pi = int(pis_mask.sum())                                             # smoe.py:1012
    norm = float(num_pi) if cfg.kernel_count_as_norm_l1 else float(cfg.start_pis)   # smoe.py:1022-1025
    l1 = pis_l1 * pis.sum() / norm if norm > 0 else pis.sum() * 0.0          # smoe.py:1027
    ul1 = u_l1 * torch.diagonal(A, dim1=-2, dim2=-1).sum()                   # smoe.py:1044
    loss = loss_pixel + l1 + ul1                                             # smoe.py:1051
    mse_op = mse * ((2 ** cfg.precision) ** 2)                               # smoe.py:1053

    return dict(loss=loss, mse_op=mse_op, num_pi=num_pi, indices=indices, indices_infl=indices_infl,
                kernel_list_batch=kernel_list_batch, S=S, S_raw=n_w_sum, w_full=w_full, w=w, infl=infl,
                r_pre=r_pre.t(), res=res, resq=resq, diff=diff, w_e_max=w_e_max,
                w_e_max_local=w_e_max_local, sampl_prob=sampl_prob, loss_pixel=loss_pixel,
                maha=maha, A=A, musX=musX, nu_e=nu_e, gamma_e=gamma_e, pis=pis)


def graph_grads(params, kernel_list, domain, target, cfg, pis_l1=0.0, u_l1=0.0, **kw):
    """tf.gradients(loss_op, variables) (smoe.py:1148) w.r.t. the K_all-sized variables."""
    leaf = {k: params[k].detach().clone().requires_grad_(True) for k in PARAM_KEYS}
    out = graph_forward(leaf, kernel_list, domain, target, cfg, pis_l1, u_l1, **kw)
    grads = torch.autograd.grad(out["loss"], [leaf[k] for k in PARAM_KEYS], allow_unused=True)
    g = {k: (torch.zeros_like(leaf[k]) if gi is None else gi) for k, gi in zip(PARAM_KEYS, grads)}
    return out, g


# ----------------------------------------------------------------------------------
# Independent closed-form backward (the spec the CUDA backward is written from)
# ----------------------------------------------------------------------------------

def closed_form_grads(params_np: Dict[str, np.ndarray], kernel_list, domain, target, cfg: GraphCfg,
                      pis_l1=0.0, u_l1=0.0, resq_override=None, loss_count=None):
    """SURVEY.md section 8a-8, float64 NumPy, train_inverse_cov False or True.

    t_nk = w_nk (m_nk gE_nk - gr_n);  dpi = sum_n t/pi (+l1);  dmu = A sum_n t y;
    dA[l,m] (l>=m) = -sum_n t delta_l y_m (+ sum_n t / A_ii, + u_l1 on the diagonal);
    dnu = sum_n m w g;  dgamma = sum_n m w g x.

    resq_override : optional (N,C) fake-quant OUTPUT to use (see graph_forward).
    loss_count : optional pixel count the loss mean runs over (default: the N fed pixels) -- lets a test feed a
        crop of a larger batch and obtain that crop's share of the full-batch gradient.
    """
    f8 = np.float64
    d, C = cfg.dim_domain, cfg.num_channels
    if cfg.quantization_mode >= 2 or cfg.use_diff_center:
        raise NotImplementedError("closed form covers quantization_mode 0/1 without use_diff_center")
    pis_all = params_np["pis"].astype(f8)
    if cfg.quantize_pis:
        nmin, nmax, scale = _nudge(cfg.lower_bounds[3], cfg.upper_bounds[3], cfg.bit_depths[3])
        qp = fq_values(torch.tensor(pis_all), nmin, nmax, scale).numpy()
        ste = (pis_all >= nmin) & (pis_all <= nmax)
    else:
        qp, ste = pis_all, np.ones_like(pis_all, dtype=bool)
    mask = np.asarray(kernel_list, bool) & (qp > 0)
    idx = np.nonzero(mask)[0]
    Ad = params_np["A_diagonal"].astype(f8)[idx]
    Ac = params_np["A_corr"].astype(f8)[idx]
    A = np.zeros_like(Ad)
    for i in range(d):
        A[:, i, i] = Ad[:, i, i]
        for j in range(i):
            A[:, i, j] = Ac[:, i, j]
            if cfg.train_inverse_cov:
                A[:, j, i] = Ac[:, i, j]
    mu = params_np["musX"].astype(f8)[idx]
    nu = params_np["nu_e"].astype(f8)[idx]
    ga = params_np["gamma_e"].astype(f8)[idx].copy()
    if cfg.use_yuv and cfg.train_gammas and cfg.only_y_gamma:
        ga[:, :, 1:] = 0
    pi = qp[idx]
    x = np.asarray(domain, f8)
    tgt = np.asarray(target, f8)
    N = x.shape[0]
    delta = x[None] - mu[:, None]                           # (K,N,d)
    if cfg.train_inverse_cov:
        Ad_ = np.einsum("klm,knm->knl", A, delta)           # A delta
        maha = (delta * Ad_).sum(-1)
    else:
        y = np.einsum("klm,knl->knm", A, delta)             # A^T delta
        maha = (y * y).sum(-1)
    coef = pi.copy()
    if cfg.use_determinant:
        coef = coef * np.prod(np.diagonal(A, axis1=-2, axis2=-1), axis=-1) / math.sqrt((2 * math.pi) ** d)
    nw = coef[:, None] * np.exp(-0.5 * maha)
    Sraw = nw.sum(0)
    live = Sraw > 10e-12
    S = np.where(live, Sraw, 10e-12)
    wf = nw / S
    tau = 0.5 / 2 ** cfg.precision
    m = wf > tau
    w = wf * m
    if cfg.train_gammas:
        E = nu[:, None, :] + np.einsum("kdc,nd->knc", ga, x)         # (K,N,C)
    else:
        E = np.broadcast_to(nu[:, None, :], (nu.shape[0], N, C))
    r = np.einsum("kn,knc->nc", w, E)
    res = np.clip(r, 0, 1)
    nmin, nmax, scale = _nudge(0.0, 1.0, cfg.precision)
    resq = fq_values(torch.tensor(res), nmin, nmax, scale).numpy()
    if resq_override is not None:
        resq = np.asarray(resq_override, f8).reshape(resq.shape)
    diff = resq - tgt
    eps = cfg.margin / 2 ** cfg.precision
    Nl = float(N if loss_count is None else loss_count)
    if cfg.use_yuv:
        cw = np.array([6 / 8] + [1 / 8] * (C - 1)) / Nl
    else:
        cw = np.ones(C) / (Nl * C)
    g = 2 * (np.abs(diff) - eps) * np.sign(diff) * cw[None, :]
    g = g * ((r >= 0) & (r <= 1))                                     # clip + fake-quant STE
    gE = np.einsum("nc,knc->kn", g, E)
    gr = (g * r).sum(-1)
    t = wf * (m * gE - np.where(live, gr, 0.0)[None, :])              # dL/dlog(nw)
    K = idx.size
    out = {k: np.zeros(params_np[k].shape, f8) for k in PARAM_KEYS}
    norm = float((qp > 0).sum()) if cfg.kernel_count_as_norm_l1 else float(cfg.start_pis)
    dpi = t.sum(1) / pi + (pis_l1 / norm if norm > 0 else 0.0)
    out["pis"][idx] = dpi * ste[idx]
    if cfg.train_inverse_cov:
        # maha = delta^T A delta, A symmetric built from diag + strict lower (+ transpose)
        dmu = np.einsum("kn,knl->kl", t, 0.5 * (Ad_ + np.einsum("kml,knm->knl", A, delta)))
        out["musX"][idx] = dmu
        dA = -0.5 * np.einsum("kn,knl,knm->klm", t, delta, delta)
        for i in range(d):
            out["A_diagonal"][idx, i, i] = dA[:, i, i] + u_l1
            if cfg.use_determinant:
                out["A_diagonal"][idx, i, i] += t.sum(1) / A[:, i, i]
            for j in range(i):
                out["A_corr"][idx, i, j] = dA[:, i, j] + dA[:, j, i]
    else:
        ty = np.einsum("kn,knm->km", t, y)
        out["musX"][idx] = np.einsum("klm,km->kl", A, ty)
        dA = -np.einsum("kn,knl,knm->klm", t, delta, y)
        for i in range(d):
            out["A_diagonal"][idx, i, i] = dA[:, i, i] + u_l1
            if cfg.use_determinant:
                out["A_diagonal"][idx, i, i] += t.sum(1) / A[:, i, i]
            for j in range(i):
                out["A_corr"][idx, i, j] = dA[:, i, j]
    out["nu_e"][idx] = np.einsum("kn,nc->kc", w, g)
    if cfg.train_gammas:
        dga = np.einsum("kn,nc,nd->kdc", w, g, x)
        if cfg.use_yuv and cfg.only_y_gamma:
            dga[:, :, 1:] = 0
        out["gamma_e"][idx] = dga
    extra = dict(S=S, r_pre=r, res=res, resq=resq, w=w, wf=wf, m=m, idx=idx, g=g, t=t)
    return out, extra


def ambiguity_margins(out, cfg: GraphCfg):
    """Per-pixel distances from the graph's discontinuities, for tolerance-aware comparison.

    thr : min_k |w_nk/tau - 1|   (gate threshold, smoe.py:825-827)
    qnt : distance of res*(2^p-1) from the nearest half-integer (fake-quant, smoe.py:899)
    """
    tau = 0.5 / 2 ** cfg.precision
    wf = out["w_full"].detach() if isinstance(out["w_full"], torch.Tensor) else torch.as_tensor(out["w_full"])
    thr = (wf / tau - 1).abs().min(dim=0).values
    q = 2 ** cfg.precision - 1
    res = out["res"].detach()
    frac = res * q - torch.floor(res * q)
    qnt = (frac - 0.5).abs().min(dim=1).values
    return thr.numpy(), qnt.numpy()
