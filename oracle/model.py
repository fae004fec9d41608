"""CPU emulation of the reference's `Smoe` host logic around the restated graph.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for the graph arithmetic (oracle/__init__.py).

  ctor / init ............. smoe.py:223-274 (+ init helpers, oracle/init_ref.py)
  set_optimizer ........... smoe.py:1079-1204, optimizers as smoe_test.py:84-88
  TF1 AdamOptimizer ....... lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
                            v += (g^2-v)(1-b2); var -= lr_t*m/(sqrt(v)+eps)
  run_batched ............. smoe.py:1606-1793 (zero accumulators, per-batch session.run,
                            loss weighting, kernel-list upkeep 1763-1766, train_op)
  train ................... smoe.py:1485-1603 (cadence, best checkpoint, divergence stop)

HEAD defects are not reproduced: the ctor works without `affines`,
`batch_size=None` is accepted, and `init_params` A is split into its diagonal and
strictly-lower parts (SURVEY decisions D1, D2).
"""
from __future__ import annotations

import numpy as np
import torch

from . import init_ref
from .graph import GraphCfg, PARAM_KEYS, graph_forward
from .quant import quantize_params, rescaler


class OracleAdam:
    """tf.train.AdamOptimizer(lr) restated; state is attached to the optimizer object."""

    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self._lr = learning_rate
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.t = 0
        self.slots = {}
        # TF1 keeps beta1_power / beta2_power as float32 variables multiplied by float32(beta) in `_finish`
        self.b1p32, self.b2p32 = np.float32(1), np.float32(1)

    def apply(self, named_grads_and_vars):
        self.t += 1
        b1p, b2p = self.beta1 ** self.t, self.beta2 ** self.t
        self.b1p32 = np.float32(self.b1p32 * np.float32(self.beta1))
        self.b2p32 = np.float32(self.b2p32 * np.float32(self.beta2))
        for name, g, var in named_grads_and_vars:
            if name not in self.slots:
                self.slots[name] = (torch.zeros_like(var), torch.zeros_like(var))
            m, v = self.slots[name]
            dt = var.dtype
            if dt == torch.float32:
                f = np.float32
                alpha = float(f(self._lr) * np.sqrt(f(1) - self.b2p32) / (f(1) - self.b1p32))
            else:
                alpha = self._lr * np.sqrt(1 - b2p) / (1 - b1p)
            m += (g - m) * (1 - self.beta1)
            v += (g * g - v) * (1 - self.beta2)
            var -= (m * alpha) / (v.sqrt() + self.epsilon)


class OracleSmoe:
    def __init__(self, image, kernels_per_dim=None, train_pis=True, init_params=None, start_batches=1,
                 batch_size=None, train_gammas=True, train_musx=True, use_diff_center=False,
                 use_determinant=False, normalize_pis=True, quantization_mode=0, bit_depths=None,
                 quantize_pis=False, lower_bounds=None, upper_bounds=None, use_yuv=True,
                 only_y_gamma=False, precision=8, iter_offset=0, margin=0.5,
                 kernel_count_as_norm_l1=False, train_inverse_cov=True, dtype=torch.float32,
                 einsum_mode="einsum", loss_mask=None, ssim_opt=False, overlap_of_batches=0, radial_as=False):
        self.image = np.asarray(image)
        self.dtype = dtype
        self.dim_domain = self.image.ndim - 1
        self.num_pixel = int(np.prod(self.image.shape[:self.dim_domain]))
        self.precision = precision
        self.use_yuv = use_yuv
        self.radial_as = radial_as
        self.use_diff_center = use_diff_center
        self.quantization_mode = quantization_mode
        self.quantize_pis = quantize_pis
        self.bit_depths, self.lower_bounds, self.upper_bounds = bit_depths, lower_bounds, upper_bounds
        self.train_pis, self.train_gammas, self.train_musx = train_pis, train_gammas, train_musx
        self.iter = iter_offset
        self.loss_mask = loss_mask
        self.ssim_opt, self.overlap = ssim_opt, int(overlap_of_batches)
        self.joint_domain = init_ref.gen_domain(self.image, self.dim_domain)      # float64
        self.batch_shape = init_ref.get_batch_shape(start_batches, self.joint_domain.shape)
        if batch_size is not None and batch_size[0] is not None:
            bs = list(batch_size) if len(batch_size) == self.dim_domain else [batch_size[0]] * self.dim_domain
            for i in range(self.dim_domain):
                if self.joint_domain.shape[i] % bs[i] > 0:
                    raise ValueError("Required BatchSize is not compatible to input dimensions")
            self.batch_size_valued = tuple(bs)
        else:
            self.batch_size_valued = tuple(self.batch_shape[:-1])
        self.start_batches = int(np.prod(np.ceil(np.array(self.image.shape[:-1]) / np.array(self.batch_size_valued))))
        assert kernels_per_dim is not None or init_params is not None
        if init_params:
            pis0, mus0 = init_params["pis"], init_params["musX"]
            A0 = init_params["A_diagonal"] + init_params["A_corr"]
            ga0, nu0 = init_params["gamma_e"], init_params["nu_e"]
        else:
            mus0, A0 = init_ref.kernel_grid(kernels_per_dim, self.dim_domain, train_inverse_cov)
            nu0, ga0 = init_ref.experts(self.image, mus0)
            pis0 = init_ref.pis(mus0.shape[0], normalize_pis)
        self.musX_init = np.asarray(mus0)
        self.start_pis = int(np.asarray(pis0).size)
        K, d = self.start_pis, self.dim_domain
        A0 = np.asarray(A0, np.float64)
        eye = np.eye(d, dtype=bool)[None]
        tt = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=dtype)
        self.vars = {
            "pis": tt(pis0),
            "musX": tt(np.zeros_like(mus0) if use_diff_center else mus0),
            "A_diagonal": tt(A0[:, 0, 0]) if radial_as else tt(np.where(eye, A0, 0.0)),     # smoe.py:429-434
            "A_corr": tt(np.zeros_like(A0)) if radial_as
            else tt(np.where(np.tril(np.ones((d, d), bool), -1)[None], A0, 0.0)),
            "gamma_e": tt(ga0),
            "nu_e": tt(nu0),
        }
        self.musX_grid = tt(mus0) if use_diff_center else None
        self.best = {k: v.clone() for k, v in self.vars.items()}
        self.cfg = GraphCfg(dim_domain=d, num_channels=self.image.shape[-1], precision=precision,
                            margin=margin, use_determinant=use_determinant,
                            train_inverse_cov=train_inverse_cov, use_yuv=use_yuv,
                            train_gammas=train_gammas, only_y_gamma=only_y_gamma,
                            use_diff_center=use_diff_center, quantize_pis=quantize_pis,
                            quantization_mode=quantization_mode, lower_bounds=lower_bounds,
                            upper_bounds=upper_bounds, bit_depths=bit_depths,
                            kernel_count_as_norm_l1=kernel_count_as_norm_l1, start_pis=K, ssim_opt=ssim_opt,
                            einsum_mode=einsum_mode)
        self.kernel_list_per_batch = [np.ones((K,), dtype=bool) for _ in range(self.start_batches)]
        nb = int(np.prod(self.batch_size_valued))                                   # smoe.py:271-273
        self.random_sampling_per_batch = [np.ones((nb,), dtype=np.float32) / nb] * self.start_batches
        self.optimizers = None
        self.grad_clip = None
        self.losses, self.mses, self.num_pis = [], [], []
        self.best_loss = self.best_mse = None
        self.qparams = self.rparams = None
        self.reconstruction_image = self.qreconstruction_image = None
        self.weight_matrix_argmax = None
        self.valid = self.qvalid = False
        self.last_grads = None

    # -- optimizers -----------------------------------------------------------------
    def set_optimizer(self, optimizer1, optimizer2=None, optimizer3=None, grad_clip_value_abs=None):
        self.optimizers = (optimizer1, optimizer2 or optimizer1, optimizer3 or optimizer1)
        self.grad_clip = grad_clip_value_abs

    def _groups(self):
        o1, o2, o3 = self.optimizers
        g1 = ["nu_e"] + (["gamma_e"] if self.train_gammas else []) + (["musX"] if self.train_musx else [])
        g2 = ["pis"] if self.train_pis else []
        g3 = ["A_diagonal"] if self.radial_as else ["A_diagonal", "A_corr"]
        return [(o, names) for o, names in ((o1, g1), (o2, g2), (o3, g3)) if not o._lr == 0]

    # -- the batched executor -----------------------------------------------------------
    def run_batched(self, pis_l1=0, u_l1=0, train=True, update_reconstruction=False,
                    with_quantized_params=False, resq_override=None, sampling_percentage=100,
                    use_loss_mask=False):
        self.valid = False
        if with_quantized_params:
            self.qvalid = False
        trainable = [n for _, names in (self._groups() if train else []) for n in names]
        accum = {n: torch.zeros_like(self.vars[n]) for n in trainable}
        loss_val = mse_val = 0.0
        num_pi = -1
        rec = np.zeros_like(self.image)
        amax = np.zeros(self.image.shape[:-1])
        d = self.dim_domain
        ov = self.overlap
        for ii, (coord, batch) in enumerate(init_ref.sliding_window(self.joint_domain, ov, self.batch_size_valued)):
            coord = coord + ov                                          # smoe.py:1719-1720
            img_patch = batch.reshape(-1, batch.shape[-1])
            samples = None
            if train and sampling_percentage < 100:                     # smoe.py:1664-1667
                num_samples = np.uint32(np.round(img_patch.shape[0] * sampling_percentage / 100))
                samples = np.random.choice(img_patch.shape[0], (num_samples,), replace=False,
                                           p=self.random_sampling_per_batch[ii])
                img_patch = img_patch[samples, :]
            self.last_samples = samples
            patch = torch.tensor(img_patch, dtype=self.dtype)   # f64 -> f32 feed
            domain, target = patch[:, :d], patch[:, d:]
            lw = None
            if use_loss_mask:                                           # smoe.py:1674-1677 (2-D: evident intent)
                sl = tuple(slice(int(c), int(c) + b) for c, b in zip(coord, self.batch_size_valued))
                lw = torch.tensor(np.asarray(self.loss_mask)[sl].reshape(-1, 1), dtype=self.dtype)
            leaf = {k: v.detach().clone().requires_grad_(k in trainable) for k, v in self.vars.items()}
            feed = None
            if with_quantized_params and update_reconstruction:
                feed = {k: torch.tensor(np.asarray(v), dtype=self.dtype) for k, v in self.rparams.items()}
            ovr = None
            if resq_override is not None:
                # the window with its halo, out of the zero-padded override image (halo values are cropped
                # before the loss, so only the interior ones matter)
                arr = np.pad(np.asarray(resq_override), ((ov, ov),) * d + ((0, 0),))
                sl = tuple(slice(int(c), int(c) + b + 2 * ov) for c, b in zip(coord, self.batch_size_valued))
                ovr = torch.tensor(arr[sl].reshape(-1, self.image.shape[-1]), dtype=self.dtype)
                if samples is not None:
                    ovr = ovr[torch.as_tensor(samples)]
            out = graph_forward(leaf, self.kernel_list_per_batch[ii], domain, target, self.cfg,
                                pis_l1, u_l1, musX_grid=self.musX_grid, feed=feed, resq_override=ovr,
                                loss_weights=lw, crop=(batch.shape[:-1], ov), train_musx=self.train_musx)
            if train and trainable:
                gs = torch.autograd.grad(out["loss"], [leaf[n] for n in trainable], allow_unused=True)
                for n, g in zip(trainable, gs):
                    if g is not None:
                        accum[n] += g                                   # assign_add, smoe.py:1150
            if update_reconstruction:
                sl = tuple(slice(int(c), int(c) + b) for c, b in zip(coord, self.batch_size_valued))
                inner = tuple(slice(ov, ov + b) for b in self.batch_size_valued)      # smoe.py:1725-1744
                rec[sl] = out["resq"].detach().numpy().reshape(tuple(batch.shape[:-1]) + (-1,))[inner]
                amax[sl] = out["w_e_max"].numpy().reshape(tuple(batch.shape[:-1]))[inner]
                self.random_sampling_per_batch[ii] = out["sampl_prob"].detach().numpy().astype(np.float32)  # smoe.py:1768-1769
            frac = np.prod(self.batch_size_valued) / self.num_pixel
            loss_val += float(out["loss"].detach()) * frac
            mse_val += float(out["mse_op"].detach()) * frac
            num_pi = out["num_pi"]
            if not with_quantized_params:                               # smoe.py:1763-1766
                bm = np.zeros_like(self.kernel_list_per_batch[ii])
                bm[out["indices_infl"].numpy()] = True
                self.kernel_list_per_batch[ii] = bm
        if update_reconstruction:
            if with_quantized_params:
                self.qreconstruction_image, self.qvalid = rec, True
            else:
                self.reconstruction_image, self.weight_matrix_argmax, self.valid = rec, amax, True
        if train:
            self.last_grads = {n: g.clone() for n, g in accum.items()}
            for opt, names in self._unique_optimizers():
                gv = []
                for n in names:
                    g = accum[n]
                    if self.grad_clip is not None:
                        g = torch.clamp(g, -self.grad_clip, self.grad_clip)
                    gv.append((n, g, self.vars[n]))
                opt.apply(gv)
        return loss_val, mse_val, num_pi, 0

    def update_kernel_list(self):
        """smoe.py:2287-2365 (single-model part): OR into every batch's list the pi>0 kernels whose Mahalanobis
        distance is < 800 at one of the 3^d corner / mid points of the batch."""
        from itertools import product as _product
        from .graph import assemble_A, effective_params
        d = self.dim_domain
        eff = effective_params({k: v.detach() for k, v in self.vars.items()}, self.cfg, self.train_musx)
        A = assemble_A(eff["A_diagonal"], eff["A_corr"], self.cfg.train_inverse_cov).double()
        mu = eff["musX"].double()
        if self.use_diff_center:
            mu = mu + self.musX_grid.double()
        pis = torch.tensor(self.get_params()["pis"], dtype=torch.float64)
        for k, (coord, batch) in enumerate(init_ref.sliding_window(self.joint_domain, self.overlap,
                                                                    self.batch_size_valued)):
            flat = batch.reshape(-1, batch.shape[-1])[:, :d]
            mn, mx = flat.min(axis=0), flat.max(axis=0)
            pts = np.array(list(_product(*[[mn[a], mx[a], (mn[a] + mx[a]) / 2] for a in range(d)])))
            probe = torch.tensor(pts.astype(np.float32).astype(np.float64))
            delta = probe[None] - mu[:, None]
            if self.cfg.train_inverse_cov:
                maha = torch.einsum("knl,klm,knm->kn", delta, A, delta)
            else:
                y = torch.einsum("klm,knl->knm", A, delta)
                maha = (y * y).sum(-1)
            near = ((maha < 800).any(dim=1) & (pis > 0)).numpy()
            self.kernel_list_per_batch[k] = np.logical_or(self.kernel_list_per_batch[k], near)

    def _unique_optimizers(self):
        # one apply_gradients per group (smoe.py:1173-1184); a shared optimizer object advances
        # its beta powers once per apply_gradients call, as in TF
        return self._groups()

    # -- train loop -----------------------------------------------------------------
    def train(self, num_iter, val_iter=100, optimizer1=None, optimizer2=None, optimizer3=None,
              grad_clip_value_abs=None, pis_l1=0, u_l1=0, callbacks=(), ukl_iter=None):
        if ukl_iter is None:
            ukl_iter = val_iter
        if optimizer1:
            self.set_optimizer(optimizer1, optimizer2, optimizer3, grad_clip_value_abs)
        assert self.optimizers is not None, "no optimizer found, you have to specify one!"
        if self.quantization_mode >= 1:
            self.qparams = quantize_params(self, self.get_params())
        self.best_loss, self.best_mse, num_pi, _ = self.run_batched(pis_l1, u_l1, train=False,
                                                                    update_reconstruction=True)
        self.losses.append((self.iter, self.best_loss))
        self.mses.append((self.iter, self.best_mse))
        self.num_pis.append((self.iter, num_pi))
        for cb in callbacks:
            cb(self)
        for i in range(1, num_iter + 1):
            self.iter += 1
            validate = i % val_iter == 0
            loss_val, mse_val, num_pi, _ = self.run_batched(pis_l1, u_l1, train=True)
            if i % ukl_iter == 0:                                        # smoe.py:1531-1536
                self.update_kernel_list()
                if not validate:
                    loss_val, mse_val, num_pi, _ = self.run_batched(pis_l1, u_l1, train=False)
            if validate:
                if self.quantization_mode >= 1:
                    self.qparams = quantize_params(self, self.get_params())
                if self.quantization_mode == 1:
                    self.rparams = rescaler(self, self.qparams)
                    self.run_batched(pis_l1, u_l1, train=False, update_reconstruction=True,
                                     with_quantized_params=True)
                loss_val, mse_val, num_pi, _ = self.run_batched(pis_l1, u_l1, train=False,
                                                                update_reconstruction=True)
            if np.isnan(loss_val) or (len(self.losses) > 0 and loss_val + 1 > (self.losses[0][1] + 100) * 10):
                break
            if validate:
                if not self.best_loss or loss_val < self.best_loss:
                    self.best_loss = loss_val
                    self.best = {k: v.clone() for k, v in self.vars.items()}
                self.losses.append((self.iter, loss_val))
                if not self.best_mse or mse_val < self.best_mse:
                    self.best_mse = mse_val
                self.mses.append((self.iter, mse_val))
                self.num_pis.append((self.iter, num_pi))
                for cb in callbacks:
                    cb(self)
        return loss_val, mse_val

    # -- getters --------------------------------------------------------------------
    def get_params(self):
        from .graph import effective_params
        eff = effective_params({k: v.detach() for k, v in self.vars.items()}, self.cfg, self.train_musx)   # smoe.py:1796-1798
        out = {k: v.numpy().astype(np.float32).copy() for k, v in eff.items()}
        if self.quantize_pis or self.quantization_mode >= 2:
            from .graph import fake_quant_args
            out["pis"] = fake_quant_args(self.vars["pis"].detach(), self.lower_bounds[3], self.upper_bounds[3],
                                         self.bit_depths[3]).numpy().astype(np.float32)
        return out

    def get_reconstruction(self):
        if not self.valid:
            self.run_batched(train=False, update_reconstruction=True)
        return self.reconstruction_image
