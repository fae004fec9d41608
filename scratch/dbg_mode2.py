import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smoe_b200 import Smoe, AdamOptimizer
from oracle.model import OracleAdam, OracleSmoe
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z = np.load(os.path.join(GOLDEN, "init_cases.npz"))
img, k = z["rgb_image"], [6, 8]
lb, ub, bd = [-40.0, -0.3, 0.1, 0.0, -2.0], [40.0, 1.3, 0.9, 2.0, 2.0], [12, 12, 7, 10, 8]
for mode in (2,):
    kw = dict(use_determinant=True, train_inverse_cov=False, use_yuv=True, normalize_pis=False,
              quantization_mode=mode, lower_bounds=lb, upper_bounds=ub, bit_depths=bd, use_diff_center=True,
              kernel_count_as_norm_l1=True)
    m = Smoe(img, kernels_per_dim=k, **kw)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    o = OracleSmoe(img, kernels_per_dim=k, dtype=torch.float64, **kw)
    o.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    K, d, C = m.start_pis, m.dim_domain, img.shape[-1]
    rs = np.random.RandomState(11)
    pert = {"musX": rs.uniform(-0.02, 0.02, (K, d)), "pis": rs.uniform(0.3, 1.7, K),
            "gamma_e": rs.normal(0, 0.3, (K, d, C)), "nu_e": o.vars["nu_e"].numpy() + rs.normal(0, 0.05, (K, C)),
            "A_corr": np.tril(rs.normal(0, 2.0, (K, d, d)), -1)}
    pert["pis"][[2, 9]] = -0.2
    pert = {kk: v.astype(np.float32) for kk, v in pert.items()}
    m.set_params(pert)
    for kk, v in pert.items():
        o.vars[kk] = torch.tensor(v.astype(np.float64))
    for it in range(22):
        m.run_batched(train=True, pis_l1=0.3, u_l1=1e-5)
        o.run_batched(train=True, pis_l1=0.3, u_l1=1e-5)
        g = m.get_gradients()
        th = m._theta.cpu().numpy()
        o_mu = o.vars["musX"].numpy(); o_nu = o.vars["nu_e"].numpy(); o_pi = o.vars["pis"].numpy()
        off = m._off
        ge = {kk: float(np.abs(g[kk] - o.last_grads[kk].numpy()).max() / max(np.abs(o.last_grads[kk].numpy()).max(), 1e-30)) for kk in o.last_grads}
        pg, po = m.get_params(), o.get_params()
        print({kk: int((np.abs(pg[kk] - po[kk]) > 1e-6).sum()) for kk in pg}, "klist diff", int((m.kernel_list_per_batch[0] != o.kernel_list_per_batch[0]).sum()),
              "n raw mu > 1e-5:", int((np.abs(th[:, :d] - o_mu) > 1e-5).sum()))
        print(mode, it, "raw var diff mu %.3e nu %.3e pi %.3e" % (np.abs(th[:, :d] - o_mu).max(), np.abs(th[:, off["nu"]:off["nu"] + C] - o_nu).max(),
              np.abs(th[:, off["pi"]] - o_pi).max()), "grad rel", {a: "%.1e" % b for a, b in ge.items()})
