"""Restatement of the reference's NumPy initialisers and batching helpers.

TEST INFRASTRUCTURE ONLY.  PINNED: every function here is checked bit-for-bit against
the reference's own code (imported under stubs by oracle/make_golden.py) through the
vectors in tests/golden/init_*.npz.

  gen_domain ............ smoe.py:2395-2426
  kernel grid / A ....... smoe.py:2146-2163
  experts (nu, gamma) ... smoe.py:2165-2235
  pis ................... smoe.py:2237-2242
  get_batch_shape ....... smoe.py:2459-2543
  sliding_window ........ smoe.py:18-35
"""
from __future__ import annotations

from itertools import product

import numpy as np


def gen_domain(in_, dim):
    """Pixel grid (array input: linspace(0,1,n) per axis + colours appended) or kernel-centre
    grid (list input: linspace(1/(2k), 1-1/(2k), k), flattened row-major)."""
    is_arr = isinstance(in_, np.ndarray)
    if is_arr:
        n = [int(in_.shape[i]) for i in range(dim)]
        axes = [np.linspace(0, 1, k) for k in n]
    else:
        n = [int(in_[i] if len(in_) > 1 else in_[0]) for i in range(dim)]
        # np.int32 scalars in the reference: 1/np.int32 is float64, same values
        axes = [np.linspace((1 / k) / 2, 1 - (1 / k) / 2, k) for k in n]
    grids = np.meshgrid(*axes, indexing="ij")
    if is_arr:
        return np.append(np.stack(grids, axis=-1), in_, axis=-1)
    return np.reshape(np.stack(grids, axis=-1), (int(np.prod(n)), dim))


def kernel_grid(kernels_per_dim, dim, train_inverse_cov):
    musX = gen_domain(list(kernels_per_dim), dim)
    if len(kernels_per_dim) > 1:
        vals = np.array([2 * (kernels_per_dim[i] + 1) for i in range(dim)], dtype=np.float64)
        proto = np.diag(vals)
        K = int(np.prod(kernels_per_dim))
    else:
        proto = np.zeros((dim, dim))
        np.fill_diagonal(proto, 2 * (kernels_per_dim[0] + 1))
        K = kernels_per_dim[0] ** dim
    A = np.tile(proto, (K, 1, 1))
    if train_inverse_cov:
        A = A ** 2
    return musX, A


def experts(image, musX):
    """nu = float32 block means around each centre, gamma = 0 (d = 2 or 3)."""
    dim = image.ndim - 1
    C = image.shape[-1]
    gamma = np.zeros((musX.shape[0], dim, C))
    stride = musX[0]
    ext = image.shape[:dim]
    mean = np.empty((musX.shape[0], C), dtype=np.float32)
    for k in range(musX.shape[0]):
        sl = []
        for a in range(dim):
            lo = int(round((musX[k, a] - stride[a]) * ext[a]))
            hi = int(round((musX[k, a] + stride[a]) * ext[a]))
            sl.append(slice(lo, hi))
        mean[k] = np.mean(image[tuple(sl)], axis=tuple(range(dim)))
    return mean, gamma


def pis(K, normalize):
    p = np.ones((K,), dtype=np.float32)
    return p / K if normalize else p


def _divisors(n):
    out = [1]
    nn, i, fac = n, 2, {}
    while i * i <= nn:
        while nn % i == 0:
            fac[i] = fac.get(i, 0) + 1
            nn //= i
        i += 1
    if nn > 1:
        fac[nn] = 1
    primes = list(fac.keys())

    def gen(k):
        if k == len(primes):
            yield 1
        else:
            for f in gen(k + 1):
                p = 1
                for _ in range(fac[primes[k]] + 1):
                    yield f * p
                    p *= primes[k]
    return list(gen(0))


def get_batch_shape(desired_batches, shape):
    """Closest divisor tiling with >= desired batches, most cube-like (smallest divisor sum);
    enumeration order (hence tie-breaking) as in the reference."""
    factors = [_divisors(shape[i]) for i in range(len(shape) - 1)] + [[1]]
    if len(shape) > 4:
        factors[0] = [1]
        factors[1] = [1]
    shapes = list(product(*factors))
    nb = np.array([[np.prod(s[:-1])] for s in shapes], dtype=np.float64)
    diff = nb - desired_batches
    diff[diff < 0] = np.inf
    aimed = nb[np.argmin(diff)]
    cand = [shapes[i] for i in np.where(nb == aimed)[0]]
    sums = np.array([[np.sum(c[2:3]) if len(c) > 4 else np.sum(c)] for c in cand], dtype=np.float64)
    div = cand[int(np.argmin(sums))]
    return tuple(int(shape[i] / div[i]) for i in range(len(shape)))


def sliding_window(image, overlap, batch):
    """Yields (coord, window) y-outer, x, then z innermost; zero halo of `overlap`."""
    if image.ndim == 3:
        pad = np.pad(image, ((overlap, overlap), (overlap, overlap), (0, 0)), "constant")
        for y in range(0, pad.shape[0] - 2 * overlap, batch[0]):
            for x in range(0, pad.shape[1] - 2 * overlap, batch[1]):
                yield (np.array([y - overlap, x - overlap]),
                       pad[y:y + batch[0] + 2 * overlap, x:x + batch[1] + 2 * overlap, :])
    elif image.ndim == 4:
        pad = np.pad(image, ((overlap, overlap),) * 3 + ((0, 0),), "constant")
        for y in range(0, pad.shape[0] - 2 * overlap, batch[0]):
            for x in range(0, pad.shape[1] - 2 * overlap, batch[1]):
                for z in range(0, pad.shape[2] - 2 * overlap, batch[2]):
                    yield (np.array([y - overlap, x - overlap, z - overlap]),
                           pad[y:y + batch[0] + 2 * overlap, x:x + batch[1] + 2 * overlap,
                               z:z + batch[2] + 2 * overlap, :])
