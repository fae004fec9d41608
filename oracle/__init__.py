"""CPU oracle for the SMoE hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, op for op, the arithmetic of the reference
(roljon/Steered-Mixture-of-Experts, `/root/reference`) for the path that
`BASELINE.json:north_star` names.  It exists so that the CUDA path can be
checked; it is never part of the product:

* only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
  / `--impl reference` legs may import it;
* nothing under `steered-mixture-of-experts_b200/` imports it, and the product
  raises when the CUDA library is missing instead of falling back to this.

PARITY STATUS
-------------
* `quant.py`, `init_ref.py` (quantizer round trip, domain / kernel-grid / expert
  / pi initialisers, batch-shape and sliding-window helpers): PINNED -- checked
  bit-for-bit against the reference's own NumPy code executed in the authoring
  container (`oracle/make_golden.py` imported `/root/reference/quantizer.py`,
  `utils.reduce_params` and the static/NumPy methods of `smoe.Smoe` under stub
  modules); the vectors live in `tests/golden/`.
* `graph.py`, `adam.py`, `ssim.py`, `model.py` (the TensorFlow-1.x graph of
  `smoe.py:714-1056`, its gradients, TF Adam, `custom_ssim`): **PARITY
  UNPINNED**.  The arithmetic lives in TensorFlow 1.x (unpinned third-party
  dependency, not under `/root/reference`, not installable offline) and the
  reference ships no tests, fixtures or golden vectors for it.  The restatement
  follows the published semantics of the TF ops at the reference's call sites
  and is cross-checked against (a) the known-answer vector KAT-1 of SURVEY.md
  section 8c, (b) an independent closed-form backward vs autograd in float64,
  (c) analytic properties (K=1, mirrored kernels, SSIM(x,x)=1 ...).
"""
