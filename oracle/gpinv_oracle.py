"""CPU fp64 restatement of the GPinv StVGP / GPMC hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker* for the CUDA path in ``gpinv_b200``: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing in ``gpinv_b200`` imports it, and the product
path has no CPU fallback.

What it restates (citations are into /root/reference, which is NOT read at run time):

* ``GPinv/stvgp.py:96-107``   ELBO = (sum logp - KL_MC)/S  or  sum logp/S - KL_analytic
* ``GPinv/stvgp.py:144-185``  u = q_mu + tril(q_sqrt) eps, KL_MC, f = L u + mean
* ``GPinv/stvgp.py:187-195``  analytic whitened KL (GPflow ``gauss_kl_white[_diag]``)
* ``GPinv/kernels.py:42-63``  K, Kdiag, Cholesky (jitter on the core *before* variance scaling)
* ``GPinv/kernels.py:105-157`` RBF / RBF_csym / RBF_casym cores
* ``GPinv/kernels.py:182-189`` Stack
* ``GPinv/kronecker.py:10-22,49-66`` kronecker_product, RBF2D.Cholesky
* ``GPinv/likelihoods.py:70-82`` Gaussian likelihood
* ``GPinv/mean_functions.py:31-72`` Zero / Constant / Stack
* ``GPinv/gpmc.py:39-46,71-82`` GPMC log-density, f = L V + mean
* ``GPinv/conditionals.py:38-78`` whitened conditional
* ``notebooks/Abel_inversion.ipynb:268-287`` AbelLikelihood
* ``notebooks/Spectroscopic_Abel_inversion.ipynb:306-356`` SpecAbelLikelihood
* ``testing/make_LosMatrix.py:6-63`` line-of-sight geometry

Third-party arithmetic that is NOT in /root/reference (GPflow ~0.3, Oct 2016, unpinned:
``setup.py:33`` has ``GPflow>=0.3.0`` commented out) is restated from its published
formulas: ``densities.gaussian``, ``kernels.Stationary.square_dist`` (expansion form),
``kullback_leiblers.gauss_kl_white[_diag]``, ``transforms.positive`` (softplus + 1e-6),
``settings.numerics.jitter_level`` = 1e-6, ``priors.Gaussian``.

PINNING STATUS: **pinned against outputs of the reference's own code.**  TensorFlow<=0.11 and
GPflow 0.3 are not installable, but ``oracle/run_reference.py`` imports the unmodified
``/root/reference/GPinv`` (and ``exec``s the notebook likelihood cells) over the torch-backed
``tensorflow`` / ``GPflow`` stand-ins in ``oracle/refshim`` and commits what it returns as
``tests/golden/ref_*.npz`` (ELBO / log-density, packed gradient, ``predict_f`` mean / variance /
full covariance, LOS matrices; 10 model cases incl. BASELINE configs[0], and compact fixtures
``refc_*.npz`` for BASELINE configs[1] at full size n=1024, S=256 and an Abel case at n=512).  This restatement
reproduces every one of them (``tests/test_oracle.py``: |df| <= 1e-9|f|, ||dg|| <= 1e-8||g||; measured
<= 2e-11 and <= 1e-10).  The stand-ins themselves are validated by the reference's OWN unit tests:
all 28 in-scope tests of ``testing/test_kernels.py``, ``test_kronecker.py``, ``test_mean.py``,
``test_stvgp.py`` (incl. the optimisation against exact GP regression) and ``test_gpmc.py`` pass
when run over them (``python -m oracle.run_reference_tests``).  The GPflow formulas themselves (not vendored in the reference) are pinned by
(G1) the golden value published in ``notebooks/Stochastic_approximation_demo.ipynb:142-150``
(19.49947035345194), which fixes the RBF formula, density constants, softplus+1e-6 transform
and parameter order; further checks: (G2) the zero-variance identity ELBO_MC == log p(Y) for the
exact whitened posterior; (G4) explicit double-loop kernels in the style of
``testing/test_kernels.py:9-78``.

Gradients come from torch-CPU autograd of this restatement (the reference's own backward
is ``tf.gradients`` of the same graph, ``GPinv/model.py:35-37``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

DT = torch.float64
JITTER = 1e-6          # GPflow settings.numerics.jitter_level
POS_EPS = 1e-6         # GPflow transforms.positive lower bound
LOG2PI = math.log(2.0 * math.pi)


def _t(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(DT)
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=DT)


# ----------------------------------------------------------------------------
# transforms (GPflow transforms.positive = Log1pe)
# ----------------------------------------------------------------------------
def positive_forward(x: torch.Tensor) -> torch.Tensor:
    """y = log(1 + exp(x)) + 1e-6."""
    return torch.nn.functional.softplus(x, beta=1.0, threshold=1e9) + POS_EPS


def positive_backward_np(y: np.ndarray) -> np.ndarray:
    """Inverse of positive_forward, numpy."""
    y = np.asarray(y, dtype=np.float64) - POS_EPS
    return np.where(y > 35.0, y, np.log(np.expm1(np.minimum(y, 35.0))))


# ----------------------------------------------------------------------------
# densities (GPflow densities.gaussian)
# ----------------------------------------------------------------------------
def gaussian_density(x, mu, var):
    """-0.5 log(2 pi) - 0.5 log(var) - 0.5 (mu - x)^2 / var"""
    return -0.5 * LOG2PI - 0.5 * torch.log(var) - 0.5 * torch.square(mu - x) / var


# ----------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------
def square_dist(X: torch.Tensor, X2: Optional[torch.Tensor], ls: torch.Tensor) -> torch.Tensor:
    """GPflow Stationary.square_dist: expansion form after dividing by the lengthscales.
    D_ij = -2 <x_i, x'_j> + |x_i|^2 + |x'_j|^2  (this order of additions)."""
    X = X / ls
    Xs = torch.sum(torch.square(X), 1)
    if X2 is None:
        return -2.0 * (X @ X.T) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
    X2 = X2 / ls
    X2s = torch.sum(torch.square(X2), 1)
    return -2.0 * (X @ X2.T) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)


def _slice(X, active_dims):
    if active_dims is None:
        return X
    return X[:, list(active_dims)]


def kcore(kind: str, X: torch.Tensor, X2: Optional[torch.Tensor], ls: torch.Tensor,
          active_dims=None) -> torch.Tensor:
    """Unit kernel shared by all outputs.  GPinv/kernels.py:105-107 (rbf), 115-120 (csym),
    140-145 (casym).  csym/casym always take the two-argument square_dist path."""
    X = _slice(X, active_dims)
    X2 = None if X2 is None else _slice(X2, active_dims)
    if kind == "rbf":
        return torch.exp(-square_dist(X, X2, ls) / 2)
    if X2 is None:
        X2 = X
    Xa, X2a = torch.abs(X), torch.abs(X2)
    e1 = torch.exp(-square_dist(Xa, X2a, ls) / 2)
    e2 = torch.exp(-square_dist(Xa, -X2a, ls) / 2)
    if kind == "rbf_csym":
        return e1 + e2
    if kind == "rbf_casym":
        return e1 - e2
    raise ValueError(kind)


def kern_K(kind, X, X2, ls, variance, active_dims=None):
    """[n, n2, R] = variance_r * core.  GPinv/kernels.py:42-48."""
    core = kcore(kind, X, X2, ls, active_dims)
    return core[:, :, None] * variance[None, None, :]


def kern_Kdiag(kind, X, ls, variance, active_dims=None):
    """[n, R].  GPinv/kernels.py:50-54 (rbf), 122-132 (csym), 147-157 (casym)."""
    n = X.shape[0]
    var = variance[None, :].expand(n, -1)
    if kind == "rbf":
        return var.clone()
    Xa = torch.abs(_slice(X, active_dims))
    sq = torch.sum(torch.square((Xa + Xa) / ls), 1)
    d = torch.exp(-0.5 * sq)
    if kind == "rbf_csym":
        return var * (1.0 + d)[:, None]
    return var * (1.0 - d)[:, None]


def chol_core(kind, X, ls, active_dims=None, jitter=JITTER):
    core = kcore(kind, X, None, ls, active_dims) + torch.eye(X.shape[0], dtype=DT) * jitter
    return torch.linalg.cholesky(core)


def kern_cholesky(kind, X, ls, variance, active_dims=None):
    """[n, n, R] = sqrt(variance_r) * chol(core + jitter I).  GPinv/kernels.py:56-63."""
    chol = chol_core(kind, X, ls, active_dims)
    return chol[:, :, None] * torch.sqrt(variance)[None, None, :]


def kronecker_product(A, B):
    """GPinv/kronecker.py:10-22: [m*p, n*q], row index i*p + k."""
    m, n = A.shape
    p, q = B.shape
    return (A[:, None, :, None] * B[None, :, None, :]).reshape(m * p, n * q)


def rbf2d_cholesky(dim1, dim2, ls, variance):
    """GPinv/kronecker.py:49-66: chol(K(dim1)+jI) (x) chol(K(dim2)+jI), dense, one shared ls."""
    c1 = chol_core("rbf", dim1, ls)
    c2 = chol_core("rbf", dim2, ls)
    chol = kronecker_product(c1, c2)
    return chol[:, :, None] * torch.sqrt(variance)[None, None, :]


# ----------------------------------------------------------------------------
# explicit loop references in the style of testing/test_kernels.py:9-78 (G4)
# ----------------------------------------------------------------------------
def loop_rbf_core(X, X2, ls):
    X, X2, ls = np.asarray(X), np.asarray(X2), np.asarray(ls)
    out = np.zeros((X.shape[0], X2.shape[0]))
    for i in range(X.shape[0]):
        for j in range(X2.shape[0]):
            d = (X[i] - X2[j]) / ls
            out[i, j] = np.exp(-0.5 * np.sum(d * d))
    return out


def loop_core(kind, X, X2, ls):
    if X2 is None:
        X2 = X
    if kind == "rbf":
        return loop_rbf_core(X, X2, ls)
    Xa, X2a = np.abs(X), np.abs(X2)
    a = loop_rbf_core(Xa, X2a, ls)
    b = loop_rbf_core(Xa, -X2a, ls)
    return a + b if kind == "rbf_csym" else a - b


def loop_kronecker(A, B):
    m, n = A.shape
    p, q = B.shape
    out = np.zeros((m * p, n * q))
    for i in range(m):
        for j in range(n):
            out[i * p:(i + 1) * p, j * q:(j + 1) * q] = A[i, j] * B
    return out


# ----------------------------------------------------------------------------
# line-of-sight geometry (testing/make_LosMatrix.py)
# ----------------------------------------------------------------------------
def make_los_matrix_loops(r, z):
    """Literal loop form of testing/make_LosMatrix.py:6-46 (small sizes only)."""
    n, N = r.shape[0], z.shape[0]
    A = np.zeros((N, n))
    for i in range(N):
        az = np.abs(z[i])
        for j in range(1, n):
            r_in = 0.5 * (r[j - 1] + r[j])
            r_out = 0.5 * (r[j + 1] + r[j]) if j < n - 1 else r[n - 1]
            if az < r_out:
                if az < r_in:
                    A[i, j] = 2. * np.sqrt(r_out ** 2. - z[i] ** 2) - 2. * np.sqrt(r_in ** 2. - z[i] ** 2)
                elif az > r_in:
                    A[i, j] = 2. * np.sqrt(r_out ** 2. - z[i] ** 2)
    return A


def make_los_matrix(r, z):
    """Vectorised, formula-identical version (column 0 zero; exact tie |z|==r_in gives 0)."""
    r = np.asarray(r, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    n = r.shape[0]
    r_in = np.zeros(n)
    r_out = np.zeros(n)
    r_in[1:] = 0.5 * (r[:-1] + r[1:])
    r_out[1:n - 1] = 0.5 * (r[2:] + r[1:n - 1])
    r_out[n - 1] = r[n - 1]
    az = np.abs(z)[:, None]
    z2 = (z ** 2)[:, None]
    with np.errstate(invalid="ignore"):
        so = 2. * np.sqrt(r_out[None, :] ** 2. - z2)
        si = 2. * np.sqrt(r_in[None, :] ** 2. - z2)
    both = (az < r_out[None, :]) & (az < r_in[None, :])
    outer = (az < r_out[None, :]) & (az > r_in[None, :])
    A = np.where(both, so - si, np.where(outer, so, 0.0))
    A[:, 0] = 0.0
    return np.ascontiguousarray(A)


def make_cos_theta(r, z):
    """testing/make_LosMatrix.py:48-63."""
    r = np.asarray(r, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = z[:, None] / r[None, :]
    return np.ascontiguousarray(np.where(np.abs(z)[:, None] < r[None, :], c, 0.0))


# ----------------------------------------------------------------------------
# StVGP sampling + KL  (GPinv/stvgp.py:144-195)
# ----------------------------------------------------------------------------
def stvgp_sample(q_mu, q_sqrt, L, mean, eps, q_diag=False):
    """GEMM-form restatement.  q_mu [n,R]; q_sqrt [n,n,R] (full) or [n,R] (diag, already
    positive); L [n,n,R]; mean [n,R]; eps [R,S,n].  Returns f [S,n,R], KL_MC (scalar)."""
    R, S, n = eps.shape
    if q_diag:
        sqrt = torch.diag_embed(q_sqrt.T)                       # [R,n,n]
    else:
        sqrt = torch.tril(q_sqrt.permute(2, 0, 1))              # [R,n,n]
    logdet_S = float(S) * torch.sum(torch.log(torch.square(torch.diagonal(sqrt, dim1=1, dim2=2))))
    u = q_mu.T[:, None, :] + torch.matmul(eps, sqrt.transpose(1, 2))   # [R,S,n]
    KL = -0.5 * logdet_S - 0.5 * torch.sum(torch.square(eps)) + 0.5 * torch.sum(torch.square(u))
    Lr = L.permute(2, 0, 1)                                     # [R,n,n]
    f = torch.matmul(u, Lr.transpose(1, 2))                     # [R,S,n]
    return f.permute(1, 2, 0) + mean[None, :, :], KL


def stvgp_sample_literal(q_mu, q_sqrt, L, mean, eps, q_diag=False):
    """Literal tile-and-batch-matmul form of GPinv/stvgp.py:150-185 (small sizes only)."""
    R, S, n = eps.shape
    if q_diag:
        sqrt = torch.diag_embed(q_sqrt.T)
    else:
        sqrt = torch.tril(q_sqrt.permute(2, 0, 1))
    logdet_S = float(S) * torch.sum(torch.log(torch.square(torch.diagonal(sqrt, dim1=1, dim2=2))))
    sqrt_t = sqrt[:, None].expand(R, S, n, n)
    v = eps[..., None]                                           # [R,S,n,1]
    mu = q_mu.T[:, None, :, None].expand(R, S, n, 1)
    u = mu + torch.matmul(sqrt_t, v)
    KL = -0.5 * logdet_S - 0.5 * torch.sum(torch.square(v)) + 0.5 * torch.sum(torch.square(u))
    Lt = L.permute(2, 0, 1)[:, None].expand(R, S, n, n)
    f = torch.matmul(Lt, u).squeeze(-1).permute(1, 2, 0) + mean[None].expand(S, n, R)
    return f, KL


def gauss_kl_white(q_mu, q_sqrt):
    """GPflow kullback_leiblers.gauss_kl_white: q_sqrt [n,n,R]."""
    Lq = torch.tril(q_sqrt.permute(2, 0, 1))
    KL = 0.5 * torch.sum(torch.square(q_mu))
    KL = KL - 0.5 * float(q_mu.numel())
    KL = KL - 0.5 * torch.sum(torch.log(torch.square(torch.diagonal(Lq, dim1=1, dim2=2))))
    KL = KL + 0.5 * torch.sum(torch.square(Lq))
    return KL


def gauss_kl_white_diag(q_mu, q_sqrt):
    """GPflow kullback_leiblers.gauss_kl_white_diag: q_sqrt [n,R]."""
    KL = 0.5 * torch.sum(torch.square(q_mu))
    KL = KL - 0.5 * float(q_mu.numel())
    KL = KL - 0.5 * torch.sum(torch.log(torch.square(q_sqrt)))
    KL = KL + 0.5 * torch.sum(torch.square(q_sqrt))
    return KL


# ----------------------------------------------------------------------------
# likelihoods
# ----------------------------------------------------------------------------
def lik_gaussian_logp(F, Y, var):
    """GPinv/likelihoods.py:70-76.  F [S,n,R], Y [n,R]."""
    return gaussian_density(F, Y[None], var)


def abel_transform(F, A):
    """notebooks/Abel_inversion.ipynb:279-283: A . exp(F) per sample; F [S,n,1] -> [S,M,1].
    GEMM form exp(F)[S,n] @ A^T: the same sums as the notebook's tile-and-batch-matmul
    (``abel_transform_literal``; tests assert they agree to 1e-12), but A is streamed once per
    evaluation instead of once per sample -- the form the CPU baseline is timed in."""
    return torch.matmul(torch.exp(F[:, :, 0]), A.T)[:, :, None]


def abel_transform_literal(F, A):
    """Literal form of the notebook cell: tf.tile(A) to [S,M,n], one mat-vec per sample."""
    return torch.matmul(A[None].expand(F.shape[0], -1, -1), torch.exp(F))


def lik_abel_logp(F, Y, A, var):
    """notebooks/Abel_inversion.ipynb:274-277.  Y [M,1]."""
    return gaussian_density(abel_transform(F, A), Y[None], var)


def spec_abel_transform(F, A, cosT, lam):
    """notebooks/Spectroscopic_Abel_inversion.ipynb:316-346.
    F [S,n,3] = (log a, log sigma, v); A, cosT [m,n]; lam [k]; returns [S,k,m]."""
    a = torch.exp(F[:, :, 0])[:, None, None, :]       # [S,1,1,n]
    s = torch.exp(F[:, :, 1])[:, None, None, :]
    v = F[:, :, 2][:, None, None, :]
    lam_ = lam.reshape(1, -1, 1, 1)
    f = a / (math.sqrt(2 * math.pi) * s) * torch.exp(-0.5 * torch.square((lam_ - v * cosT[None, None]) / s))
    return torch.sum(A[None, None] * f, 3)


def lik_spec_abel_logp(F, Y, A, cosT, lam, var):
    """ibid :348-356.  Y [k,m]."""
    return gaussian_density(spec_abel_transform(F, A, cosT, lam), Y[None], var)


# ----------------------------------------------------------------------------
# whitened conditional (GPinv/conditionals.py:38-78, whiten=True)
# ----------------------------------------------------------------------------
def conditional_whitened(Kmn, Lm, Knn_or_diag, f, q_sqrt=None, full_cov=False):
    """Kmn [R,n,n2]; Lm [R,n,n]; Knn [R,n2,n2] or Kdiag [R,n2]; f [n,R];
    q_sqrt [n,R] / [n,n,R] / None.  Returns fmean [n2,R], fvar [n2,R] or [n2,n2,R]."""
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    if full_cov:
        fvar = Knn_or_diag - torch.matmul(A.transpose(1, 2), A)
    else:
        fvar = Knn_or_diag - torch.sum(torch.square(A), 1)
    fmean = torch.matmul(A.transpose(1, 2), f.T[:, :, None]).squeeze(-1).T
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A * q_sqrt.T[:, :, None]
        else:
            Lq = torch.tril(q_sqrt.permute(2, 0, 1))
            LTA = torch.matmul(Lq.transpose(1, 2), A)
        if full_cov:
            fvar = fvar + torch.matmul(LTA.transpose(1, 2), LTA)
        else:
            fvar = fvar + torch.sum(torch.square(LTA), 1)
    if full_cov:
        fvar = fvar.permute(1, 2, 0)
    else:
        fvar = fvar.T
    return fmean, fvar


# ----------------------------------------------------------------------------
# model-level objective driven by a plain-dict description
# ----------------------------------------------------------------------------
# spec = {
#   'model': 'stvgp' | 'gpmc',
#   'X': [n,D], 'Y': array,
#   'kern':  {'type': 'rbf'|'rbf_csym'|'rbf_casym', 'input_dim', 'output_dim', 'ARD', 'active_dims'}
#          | {'type': 'stack', 'kern_list': [...]}
#          | {'type': 'rbf2d', 'output_dim', 'dim1', 'dim2', 'ARD': False}
#   'mean':  {'type': 'zero'|'constant', 'output_dim'} | {'type': 'stack', 'mean_list': [...]}
#   'lik':   {'type': 'gaussian'} | {'type': 'abel', 'A'} | {'type': 'spec_abel', 'A', 'cosT', 'lam'}
#   'q_diag', 'KL_analytic', 'num_latent'
# }
# Parameters are packed in GPflow's ``sorted_params`` order: attributes sorted by name at each
# level of the tree (ASCII order, so GPMC's ``V`` comes first); list containers keep list order.

def _kern_layout(k, prefix):
    t = k["type"]
    if t == "stack":
        out = []
        for i, sub in enumerate(k["kern_list"]):
            out += _kern_layout(sub, "%s.kern_list.%d" % (prefix, i))
        return out
    nls = k["input_dim"] if k.get("ARD", False) else 1
    return [(prefix + ".lengthscales", (nls,), "positive"),
            (prefix + ".variance", (k["output_dim"],), "positive")]


def _mean_layout(m, prefix):
    t = m["type"]
    if t == "zero":
        return []
    if t == "constant":
        return [(prefix + ".c", (m["output_dim"],), "identity")]
    out = []
    for i, sub in enumerate(m["mean_list"]):
        out += _mean_layout(sub, "%s.mean_list.%d" % (prefix, i))
    return out


def param_layout(spec) -> List[Tuple[str, Tuple[int, ...], str]]:
    n = spec["X"].shape[0]
    R = spec["num_latent"]
    lay = []
    if spec["model"] == "gpmc":
        lay.append(("V", (n, R), "identity"))
    lay += _kern_layout(spec["kern"], "kern")
    lay.append(("likelihood.variance", (1,), "positive"))
    lay += _mean_layout(spec["mean"], "mean_function")
    if spec["model"] == "stvgp":
        lay.append(("q_mu", (n, R), "identity"))
        if spec.get("q_diag", False):
            lay.append(("q_sqrt", (n, R), "positive"))
        else:
            lay.append(("q_sqrt", (n, n, R), "identity"))
    return lay


def unpack(spec, x: torch.Tensor) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for name, shape, tr in param_layout(spec):
        sz = int(np.prod(shape))
        v = x[off:off + sz].reshape(shape)
        out[name] = positive_forward(v) if tr == "positive" else v
        off += sz
    assert off == x.numel(), (off, x.numel())
    return out


def state_size(spec) -> int:
    return int(sum(np.prod(s) for _, s, _ in param_layout(spec)))


def pack_constrained(spec, values: Dict[str, np.ndarray]) -> np.ndarray:
    """Constrained-space values -> free state vector."""
    parts = []
    for name, shape, tr in param_layout(spec):
        v = np.asarray(values[name], dtype=np.float64).reshape(shape)
        parts.append((positive_backward_np(v) if tr == "positive" else v).ravel())
    return np.concatenate(parts)


def _kern_chol(k, prefix, P, X):
    t = k["type"]
    if t == "stack":
        return torch.cat([_kern_chol(s, "%s.kern_list.%d" % (prefix, i), P, X)
                          for i, s in enumerate(k["kern_list"])], 2)
    ls, var = P[prefix + ".lengthscales"], P[prefix + ".variance"]
    if t == "rbf2d":
        return rbf2d_cholesky(_t(k["dim1"]), _t(k["dim2"]), ls, var)
    return kern_cholesky(t, X, ls, var, k.get("active_dims"))


def _kern_K(k, prefix, P, X, X2):
    t = k["type"]
    if t == "stack":
        return torch.cat([_kern_K(s, "%s.kern_list.%d" % (prefix, i), P, X, X2)
                          for i, s in enumerate(k["kern_list"])], 2)
    ls, var = P[prefix + ".lengthscales"], P[prefix + ".variance"]
    kind = "rbf" if t == "rbf2d" else t
    return kern_K(kind, X, X2, ls, var, k.get("active_dims"))


def _kern_Kdiag(k, prefix, P, X):
    t = k["type"]
    if t == "stack":
        return torch.cat([_kern_Kdiag(s, "%s.kern_list.%d" % (prefix, i), P, X)
                          for i, s in enumerate(k["kern_list"])], 1)
    ls, var = P[prefix + ".lengthscales"], P[prefix + ".variance"]
    kind = "rbf" if t == "rbf2d" else t
    return kern_Kdiag(kind, X, ls, var, k.get("active_dims"))


def _mean(m, prefix, P, n):
    t = m["type"]
    if t == "zero":
        return torch.zeros(n, m["output_dim"], dtype=DT)
    if t == "constant":
        return P[prefix + ".c"][None, :].expand(n, -1)
    return torch.cat([_mean(s, "%s.mean_list.%d" % (prefix, i), P, n)
                      for i, s in enumerate(m["mean_list"])], 1)


def _logp(spec, P, F):
    lk = spec["lik"]
    var = P["likelihood.variance"]
    Y = _t(spec["Y"])
    if lk["type"] == "gaussian":
        return lik_gaussian_logp(F, Y, var)
    if lk["type"] == "abel":
        return lik_abel_logp(F, Y, _t(lk["A"]), var)
    if lk["type"] == "spec_abel":
        return lik_spec_abel_logp(F, Y, _t(lk["A"]), _t(lk["cosT"]), _t(lk["lam"]).reshape(-1), var)
    raise ValueError(lk["type"])


def sample_F_transform(spec, P, F):
    lk = spec["lik"]
    if lk["type"] == "gaussian":
        return F
    if lk["type"] == "abel":
        return abel_transform(F, _t(lk["A"]))
    return spec_abel_transform(F, _t(lk["A"]), _t(lk["cosT"]), _t(lk["lam"]).reshape(-1))


def latent_samples(spec, x, eps):
    """f samples [S,n,R] and KL_MC for the state x."""
    P = unpack(spec, _t(x))
    X = _t(spec["X"])
    L = _kern_chol(spec["kern"], "kern", P, X)
    mean = _mean(spec["mean"], "mean_function", P, X.shape[0])
    f, KL = stvgp_sample(P["q_mu"], P["q_sqrt"], L, mean, _t(eps), spec.get("q_diag", False))
    return P, f, KL


def objective_value(spec, x: torch.Tensor, eps=None) -> torch.Tensor:
    """Scalar that GPflow maximises: StVGP ELBO (GPinv/stvgp.py:96-107) or GPMC log joint
    (GPinv/gpmc.py:39-46 + prior N(V|0,1))."""
    P = unpack(spec, x)
    X = _t(spec["X"])
    n = X.shape[0]
    L = _kern_chol(spec["kern"], "kern", P, X)
    mean = _mean(spec["mean"], "mean_function", P, n)
    if spec["model"] == "stvgp":
        eps = _t(eps)
        S = eps.shape[1]
        f, KL = stvgp_sample(P["q_mu"], P["q_sqrt"], L, mean, eps, spec.get("q_diag", False))
        lik = torch.sum(_logp(spec, P, f))
        if not spec.get("KL_analytic", False):
            return (lik - KL) / S
        if spec.get("q_diag", False):
            return lik / S - gauss_kl_white_diag(P["q_mu"], P["q_sqrt"])
        return lik / S - gauss_kl_white(P["q_mu"], P["q_sqrt"])
    V = P["V"]
    F = torch.matmul(L.permute(2, 0, 1), V.T[:, :, None]).squeeze(-1).T + mean
    lik = torch.sum(_logp(spec, P, F[None]))
    prior = torch.sum(gaussian_density(torch.zeros((), dtype=DT), V, torch.ones((), dtype=DT)))
    return lik + prior


def objective(spec, x, eps=None) -> Tuple[float, np.ndarray]:
    """The reference's ``_objective(x)``: returns (-f, -df/dx) (GPinv/model.py:64-75 pattern)."""
    xt = _t(x).clone().requires_grad_(True)
    f = objective_value(spec, xt, eps)
    (g,) = torch.autograd.grad(f, xt)
    return -float(f.detach()), -g.numpy()


def predict_f(spec, x, Xnew, full_cov=False):
    """GPinv/stvgp.py:109-122, GPinv/gpmc.py:48-51."""
    P = unpack(spec, _t(x))
    X, Xn = _t(spec["X"]), _t(Xnew)
    Kmn = _kern_K(spec["kern"], "kern", P, X, Xn).permute(2, 0, 1)
    Lm = _kern_chol(spec["kern"], "kern", P, X).permute(2, 0, 1)
    if full_cov:
        Knn = _kern_K(spec["kern"], "kern", P, Xn, None).permute(2, 0, 1)
    else:
        Knn = _kern_Kdiag(spec["kern"], "kern", P, Xn).T
    if spec["model"] == "stvgp":
        mu, var = conditional_whitened(Kmn, Lm, Knn, P["q_mu"], P["q_sqrt"], full_cov)
    else:
        mu, var = conditional_whitened(Kmn, Lm, Knn, P["V"], None, full_cov)
    mean = _mean(spec["mean"], "mean_function", P, Xn.shape[0])
    return (mu + mean).numpy(), var.numpy()


# ----------------------------------------------------------------------------
# exact GPR (used by the golden value G1 and the zero-variance identity G2)
# ----------------------------------------------------------------------------
def gpr_nlml(x_free, X, Y):
    """Negative log marginal likelihood of GPflow GPR(RBF) at free state
    [lengthscale, kernel variance, noise variance] (softplus + 1e-6)."""
    ls, v, s2 = [positive_forward(_t(x_free)[i]) for i in range(3)]
    X, Y = _t(X), _t(Y)
    n = X.shape[0]
    K = v * torch.exp(-square_dist(X, None, ls.reshape(1)) / 2) + s2 * torch.eye(n, dtype=DT)
    L = torch.linalg.cholesky(K)
    a = torch.linalg.solve_triangular(L, Y, upper=False)
    return float(0.5 * torch.sum(a * a) + torch.sum(torch.log(torch.diagonal(L)))
                 + 0.5 * n * Y.shape[1] * LOG2PI)
