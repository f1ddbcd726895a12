"""Kronecker-structured 2-D RBF kernel (mirror of GPinv/kronecker.py).

The reference multiplies the two small Cholesky factors out into a dense [n1 n2, n1 n2] matrix
(GPinv/kronecker.py:61).  Here the product stays factored: (L1 (x) L2) u = vec(L1 U L2^T) with
U = mat(u) [n1,n2] (row index i*n2 + j, the order pinned by testing/test_kronecker.py:37-41),
i.e. two DMMA GEMMs of cost 2 n (n1 + n2) per sample instead of 2 n^2.
"""
from __future__ import annotations

import torch

from . import kernels, ops
from .param import DataHolder
from .settings import settings


def kronecker_product(A, B):
    """Dense Kronecker product [m*p, n*q] (GPinv/kronecker.py:10-22); API parity / tests only."""
    m, n = A.shape
    p, q = B.shape
    return (A[:, None, :, None] * B[None, :, None, :]).reshape(m * p, n * q)


def kron_apply(L1, L2, U, trans=False):
    """(L1 (x) L2) u for every row u of U [B, n1*n2] without forming the product
    (``trans``: the transposed operator).  sm_100a op ``gpinv::kron_apply``."""
    return ops.kron_apply(L1.contiguous(), L2.contiguous(), U.contiguous(), bool(trans))


def kron_sample(core, sqrt_v, Q, q_mu, mean, eps, q_diag, want_exp):
    """Factored counterpart of ``ops.stvgp_sample`` for a Kronecker core (L1, L2):
    returns (F [R,S,n], expF or None, KL [1]).  Full and diagonal q_sqrt (the reference handles
    both through its dense factor, GPinv/kronecker.py:49-66 + GPinv/stvgp.py:154-183)."""
    L1, L2 = core
    F, expF, KL, _, _ = ops.kron_sample(L1.contiguous(), L2.contiguous(), sqrt_v.contiguous(), Q.contiguous(),
                                        q_mu.contiguous(), mean.contiguous(), eps.contiguous(), bool(q_diag),
                                        bool(want_exp))
    return F, (expF if want_exp else None), KL


class RBF2D(kernels.RBF):
    """Separable 2-D RBF on the grid dim1 x dim2 with ONE shared lengthscale
    (GPinv/kronecker.py:24-66).  X must be the grid in row order i*n2 + j."""

    def __init__(self, input_dim, output_dim, dim1, dim2, variance=None, lengthscales=None,
                 active_dims=None, ARD=False):
        kernels.RBF.__init__(self, input_dim, output_dim, variance, lengthscales, active_dims, ARD)
        self.dim1 = DataHolder(dim1, "recompile")
        self.dim2 = DataHolder(dim2, "recompile")

    def Cholesky_factor(self, X):
        j = settings.numerics.jitter_level
        c1, i1, _ = ops.chol_core(self.dim1.contiguous(), self.lengthscales, self._mode, float(j), False)
        c2, i2, _ = ops.chol_core(self.dim2.contiguous(), self.lengthscales, self._mode, float(j), False)
        return kernels.CholeskyFactor([(c1, c2)], torch.sqrt(self.variance), [0] * self.output_dim,
                                      [i1, i2], kron=True)

    def _Kcore(self, X, X2=None, jitter=0.0):
        # the dense 2-D kernel itself (input_dim=1 for the factors, 2 columns in X)
        return ops.kcore(X.contiguous(), None if X2 is None else X2.contiguous(), self.lengthscales,
                         self._mode, float(jitter) if X2 is None else 0.0)
