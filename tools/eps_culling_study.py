"""How much does dropping (pixel, kernel) terms below 2^-x of the pixel's normaliser change the graph?
Float64 study on config 1 (128x128, 16x16 kernels) with the oracle's formulas -- evidence for the
"epsilon-culling" item under "next" in DESIGN.md.  CPU only:  python scratch/eps_culling_study.py"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from oracle import init_ref
from oracle.graph import GraphCfg, assemble_A, fake_quant_args, _ClipByValue01

img = np.load("tests/golden/init_cases.npz")["c1_image"]
d, C, K = 2, 1, 256
mus, A0 = init_ref.kernel_grid([16, 16], d, False)
nu, ga = init_ref.experts(img, mus)
rs = np.random.RandomState(0)
p = {"pis": torch.tensor(init_ref.pis(K, True).astype(np.float64)), "musX": torch.tensor(mus + rs.normal(0, 0.005, mus.shape)),
     "A_diag": torch.tensor(np.stack([np.diag(A0[k]) for k in range(K)]) * (1 + 0.2 * rs.uniform(-1, 1, (K, 1)))),
     "A_low": torch.tensor(rs.normal(0, 3.0, K)), "nu": torch.tensor(nu.astype(np.float64)),
     "ga": torch.tensor(rs.normal(0, 0.3, (K, d, C)))}
jd = init_ref.gen_domain(img, d).reshape(-1, d + C)
x, tgt = torch.tensor(jd[:, :d]), torch.tensor(jd[:, d:])


def run(xcull):
    leaf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    A = torch.diag_embed(leaf["A_diag"])
    A = A + torch.zeros_like(A).index_put((torch.arange(K), torch.ones(K, dtype=torch.long), torch.zeros(K, dtype=torch.long)), leaf["A_low"])
    delta = x[None] - leaf["musX"][:, None]
    y = torch.einsum("klm,knl->knm", A, delta)
    q = -0.5 * (y * y).sum(-1) + torch.log(leaf["pis"] * torch.diagonal(A, dim1=1, dim2=2).prod(-1) / (2 * np.pi))[:, None]
    nw = torch.exp(q)
    if xcull is not None:                       # drop terms below 2^-x of the (exact) normaliser, everywhere they appear
        keep = (nw.detach() > nw.detach().sum(0, keepdim=True) * 2.0 ** (-xcull)).double()
        nw = nw * keep
    S = nw.sum(0).clamp_min(1e-11)
    w = nw / S
    w = w * (w > 0.5 / 256).double()
    E = leaf["nu"].t()[:, :, None] + torch.einsum("kdc,nd->ckn", leaf["ga"], x)
    r = (w[None] * E).sum(1)
    res = _ClipByValue01.apply(r).t()
    resq = fake_quant_args(res, 0.0, 1.0, 8)
    diff = resq - tgt
    loss = ((diff.abs() - 0.5 / 256) ** 2).mean()
    g = torch.autograd.grad(loss, list(leaf.values()))
    return r.detach().t(), [gi.detach() for gi in g], float((nw.detach() > 0).double().mean())


r0, g0, f0 = run(None)
print(f"exact: fraction of non-zero float64 terms {f0:.3f}")
for xc in (126, 64, 48, 40, 32, 24):
    r1, g1, f1 = run(xc)
    gerr = max(float((a - b).abs().max() / a.abs().max()) for a, b in zip(g0, g1))
    print(f"x = {xc:3d}: kept pairs {f1:.4f}  max |dr| = {float((r0 - r1).abs().max()):.2e}  "
          f"max gradient change / tensor max-norm = {gerr:.2e}")
