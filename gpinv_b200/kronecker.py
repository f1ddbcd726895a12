"""Kronecker-structured 2-D RBF kernel (mirror of GPinv/kronecker.py).

The reference multiplies the two small Cholesky factors out into a dense [n1 n2, n1 n2] matrix
(GPinv/kronecker.py:61).  Here the product stays factored: (L1 (x) L2) u = vec(L1 U L2^T) with
U = mat(u) [n1,n2] (row index i*n2 + j, the order pinned by testing/test_kronecker.py:37-41),
i.e. two DMMA GEMMs of cost 2 n (n1 + n2) per sample instead of 2 n^2.
"""
from __future__ import annotations

import math

import torch

from . import kernels, ops
from .param import DataHolder
from .settings import settings


def kronecker_product(A, B):
    """Dense Kronecker product [m*p, n*q] (GPinv/kronecker.py:10-22); API parity / tests only."""
    m, n = A.shape
    p, q = B.shape
    return (A[:, None, :, None] * B[None, :, None, :]).reshape(m * p, n * q)


class _KronApply(torch.autograd.Function):
    """F[b] = L1 U[b] L2^T for a batch of [n1,n2] matrices stored as rows of U [B, n1*n2]."""

    @staticmethod
    def forward(ctx, L1, L2, U):
        n1, n2 = L1.shape[0], L2.shape[0]
        B = U.shape[0]
        T = ops.gemm(U.reshape(B * n1, n2), L2, False, True, False)            # U L2^T
        Tp = T.reshape(B, n1, n2).permute(1, 0, 2).reshape(n1, B * n2).contiguous()
        Fp = ops.gemm(L1, Tp, False, False, False)                               # L1 (.)
        F = Fp.reshape(n1, B, n2).permute(1, 0, 2).reshape(B, n1 * n2).contiguous()
        ctx.save_for_backward(L1, L2, U, Tp)
        return F

    @staticmethod
    def backward(ctx, Fbar):
        L1, L2, U, Tp = ctx.saved_tensors
        n1, n2 = L1.shape[0], L2.shape[0]
        B = U.shape[0]
        Fbar = Fbar.contiguous()
        Fbp = Fbar.reshape(B, n1, n2).permute(1, 0, 2).reshape(n1, B * n2).contiguous()
        L1bar = torch.tril(ops.gemm(Fbp, Tp, False, True, False))               # sum_b Fbar_b T_b^T
        Gp = ops.gemm(L1, Fbp, True, False, False)                               # L1^T Fbar
        G = Gp.reshape(n1, B, n2).permute(1, 0, 2).reshape(B * n1, n2).contiguous()
        Ubar = ops.gemm(G, L2, False, False, False).reshape(B, n1 * n2)          # (L1^T Fbar) L2
        # F_b = V_b L2^T with V_b = L1 U_b  =>  L2bar = sum_b Fbar_b^T V_b
        Up = U.reshape(B, n1, n2).permute(1, 0, 2).reshape(n1, B * n2).contiguous()
        Vp = ops.gemm(L1, Up, False, False, False)
        V = Vp.reshape(n1, B, n2).permute(1, 0, 2).reshape(B * n1, n2).contiguous()
        L2bar = torch.tril(ops.gemm(Fbar.reshape(B * n1, n2), V, True, False, False))
        return L1bar, L2bar, Ubar


def kron_sample(core, sqrt_v, Q, q_mu, mean, eps, q_diag, want_exp):
    """Factored counterpart of ``ops.stvgp_sample`` for a Kronecker core (L1, L2):
    returns (F [R,S,n], expF or None, KL [1])."""
    L1, L2 = core
    R, S, n = eps.shape
    if q_diag:
        U = q_mu[:, None, :] + Q[:, None, :] * eps
        logdet = torch.sum(torch.log(torch.square(Q)))
    else:
        U = q_mu[:, None, :] + torch.stack(
            [ops.gemm(eps[r].contiguous(), Q[r].contiguous(), False, True, False) for r in range(R)], 0)
        logdet = torch.sum(torch.log(torch.square(torch.diagonal(Q, dim1=1, dim2=2))))
        if Q.requires_grad:
            raise NotImplementedError("full q_sqrt with a Kronecker kernel: use q_diag=True")
    KL = (-0.5 * float(S) * logdet - 0.5 * torch.sum(torch.square(eps))
          + 0.5 * torch.sum(torch.square(U))).reshape(1)
    F = _KronApply.apply(L1, L2, U.reshape(R * S, n)).reshape(R, S, n)
    F = F * sqrt_v[:, None, None] + mean[:, None, :]
    return F, (torch.exp(F) if want_exp else None), KL


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
        c1, i1 = ops.chol_core(self.dim1.contiguous(), self.lengthscales, self._mode, float(j))
        c2, i2 = ops.chol_core(self.dim2.contiguous(), self.lengthscales, self._mode, float(j))
        return kernels.CholeskyFactor([(c1, c2)], torch.sqrt(self.variance), [0] * self.output_dim,
                                      [i1, i2], kron=True)

    def _Kcore(self, X, X2=None, jitter=0.0):
        # the dense 2-D kernel itself (input_dim=1 for the factors, 2 columns in X)
        return ops.kcore(X.contiguous(), None if X2 is None else X2.contiguous(), self.lengthscales,
                         self._mode, float(jitter) if X2 is None else 0.0)
