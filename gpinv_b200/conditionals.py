"""Whitened GP conditional (mirror of GPinv/conditionals.py:6-78, ``whiten=True`` as used by
StVGP / GPMC ``build_predict``).  The projection A = Lm^{-1} Kmn is computed once per kernel
core on the transposed problem (A^T = Knm Lm^{-T}) with the substitution TRSM and the DMMA
GEMM; the per-output sqrt(variance) factors are applied as scalars."""
from __future__ import annotations

import torch

from . import ops


def conditional(Xnew, X, kern, f, full_cov=False, q_sqrt=None, whiten=False):
    """Xnew [n2,D]; X [n,D]; f [n,R]; q_sqrt None | [n,R] | [n,n,R].
    Returns fmean [n2,R] and fvar [n2,R] (or [n2,n2,R] when full_cov)."""
    if not whiten:
        raise NotImplementedError("only the whitened conditional is on the StVGP/GPMC path")
    chol = kern.Cholesky_factor(X)
    if chol.kron:
        chol = type(chol)([torch.kron(c[0], c[1]) for c in chol.cores], chol.sqrt_var,
                          chol.core_index, chol.infos, False)
    R = chol.num_latent
    # A_r = Lm_r^{-1} Kmn_r with Lm_r = sqrt_var_r * core factor and Kmn_r = variance_r * core(X, Xnew):
    # the scalar is variance_r / sqrt_var_r (= sqrt(variance_r) for the built-in factored kernels;
    # a kernel that overrides Cholesky() hands over unit-variance cores that already carry it)
    kvar = torch.cat([k.variance.reshape(-1) for k in _core_kernels(kern)])
    Kdiag = kern.Kdiag(Xnew)                                     # [n2,R]
    Knn = kern.K(Xnew) if full_cov else None                     # [n2,n2,R]
    means, vars_ = [None] * R, [None] * R
    out_kern = [k for k in _core_kernels(kern) for _ in range(k.output_dim)]   # output r -> its kernel
    for core, r0, r1 in chol.groups():
        k = out_kern[r0]
        Knm = k._Kcore(Xnew, X)                                  # [n2,n] core(Xnew, X)
        AT = ops.trsm_rlt(core, Knm)                             # [n2,n] = (Lc^{-1} core_mn)^T
        for r in range(r0, r1):
            sv = kvar[r] / chol.sqrt_var[r]
            ATr = AT * sv                                        # A_r^T = (v_r / sqrt_var_r) Lc^{-1} core_mn
            fmean = ops.gemm(ATr.contiguous(), f[:, r:r + 1].contiguous(), False, False, False)[:, 0]
            if full_cov:
                fvar = Knn[:, :, r] - ops.gemm(ATr.contiguous(), ATr.contiguous(), False, True, False)
            else:
                fvar = Kdiag[:, r] - torch.sum(torch.square(ATr), 1)
            if q_sqrt is not None:
                if q_sqrt.dim() == 2:
                    LTA_T = ATr * q_sqrt[:, r][None, :]           # (diag(q) A)^T
                else:
                    Lq = torch.tril(q_sqrt[:, :, r]).contiguous()
                    LTA_T = ops.gemm(ATr.contiguous(), Lq, False, False, False)   # (Lq^T A)^T = A^T Lq
                if full_cov:
                    fvar = fvar + ops.gemm(LTA_T.contiguous(), LTA_T.contiguous(), False, True, False)
                else:
                    fvar = fvar + torch.sum(torch.square(LTA_T), 1)
            means[r], vars_[r] = fmean, fvar
    fmean = torch.stack(means, 1)
    fvar = torch.stack(vars_, -1)
    return fmean, fvar


def _core_kernels(kern):
    """Member kernels in the order of CholeskyFactor.groups() (one per Cholesky)."""
    from .kernels import Stack
    if isinstance(kern, Stack):
        out = []
        for k in kern.kern_list:
            out += _core_kernels(k)
        return out
    return [kern]
