import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from smoe_b200 import Smoe, AdamOptimizer
PARAM_KEYS = ("pis", "musX", "A_diagonal", "A_corr", "gamma_e", "nu_e")
z = np.load('tests/golden/graph_cases.npz')
def rel(a,b): return float(np.abs(np.asarray(a,np.float64)-b).max()/max(np.abs(b).max(),1e-30))
for name in ["g21","g23","g33","g21tic","g31"]:
    n=name+'_'
    img=z[n+'image']; tic,det,yuv=[bool(v) for v in z[n+'flags']]
    m=Smoe(img, kernels_per_dim=[int(v) for v in z[n+'k']], use_determinant=det, train_inverse_cov=tic, use_yuv=yuv)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    m.set_params({k: z[n+'p_'+k] for k in PARAM_KEYS})
    m.kernel_list_per_batch=[z[n+'kernel_list']]
    m._enable_res_pre()
    loss,mse,num_pi,_=m.run_batched(pis_l1=0.3,u_l1=1e-6,train=True,update_reconstruction=True)
    C=img.shape[-1]
    pre=m._d_res_pre.cpu().numpy().reshape(-1,C)
    rq=m.get_reconstruction().reshape(-1,C)
    flips=(np.round(rq*255)!=np.round(z[n+'resq']*255)).sum()
    g=m.get_gradients()
    print(name, 'pre err', np.abs(pre-z[n+'r_pre']).max(), 'flips', flips, 'loss', loss-float(z[n+'loss']), {k: rel(g[k], z[n+'g_'+k]) for k in PARAM_KEYS})
    if name=='g31':
        print(g['A_corr'][:4], z[n+'g_A_corr'][:4])
