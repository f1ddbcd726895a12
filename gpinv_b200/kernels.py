"""Multi-output kernels (mirror of GPinv/kernels.py).

One shared "core" matrix times a per-output variance; ONE Cholesky factorisation is shared by
all R outputs of a kernel (GPinv/kernels.py:56-63).  The core build, its Cholesky and their
reverse passes are the sm_100a ops ``gpinv::kcore`` / ``gpinv::potrf``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .param import Param, ParamList, Parameterized
from .settings import settings
from . import transforms


class CholeskyFactor(object):
    """Factored form of ``Kern.Cholesky(X)``: output r uses ``sqrt_var[r] * cores[core_index[r]]``.

    The reference materialises the dense [n,n,R] tile (GPinv/kernels.py:60-63); the hot path
    keeps it factored and folds sqrt(variance) into the GEMM epilogue.  ``dense()`` gives the
    reference's tensor.  ``kron`` is set for RBF2D: cores are then (L1, L2) pairs."""

    def __init__(self, cores, sqrt_var, core_index, infos=None, kron=False):
        self.cores = cores
        self.sqrt_var = sqrt_var
        self.core_index = list(core_index)
        self.infos = infos or []
        self.kron = kron

    @property
    def num_latent(self):
        return len(self.core_index)

    def groups(self):
        """[(core, r0, r1)] runs of consecutive outputs sharing one core."""
        out = []
        r0 = 0
        for r in range(1, self.num_latent + 1):
            if r == self.num_latent or self.core_index[r] != self.core_index[r0]:
                out.append((self.cores[self.core_index[r0]], r0, r))
                r0 = r
        return out

    def dense(self):
        mats = []
        for r in range(self.num_latent):
            c = self.cores[self.core_index[r]]
            if self.kron:
                c = torch.kron(c[0], c[1])
            mats.append(c * self.sqrt_var[r])
        return torch.stack(mats, dim=2)

    @staticmethod
    def concat(factors):
        cores, idx, infos = [], [], []
        for f in factors:
            base = len(cores)
            cores += list(f.cores)
            idx += [base + i for i in f.core_index]
            infos += list(f.infos)
        kron = any(f.kron for f in factors)
        if kron and not all(f.kron for f in factors):
            raise NotImplementedError("cannot stack Kronecker and dense kernels")
        return CholeskyFactor(cores, torch.cat([f.sqrt_var for f in factors]), idx, infos, kron)


class Kern(Parameterized):
    """GPinv/kernels.py:11-69."""

    def __init__(self, output_dim):
        Parameterized.__init__(self)
        self.output_dim = output_dim

    def K(self, X, X2=None):
        """[n, n2, R] = variance_r * core  (GPinv/kernels.py:42-48)."""
        core = self._Kcore(X, X2)
        return core[:, :, None] * self.variance[None, None, :]

    def Kdiag(self, X):
        """[n, R]  (GPinv/kernels.py:50-54)."""
        return self.variance[None, :].expand(X.shape[0], -1)

    def Cholesky(self, X):
        """Dense [n, n, R] like the reference (GPinv/kernels.py:56-63).  Models use
        ``Cholesky_factor`` instead and never build this tile."""
        return self.Cholesky_factor(X).dense()

    def Cholesky_factor(self, X):
        """Factored ``Cholesky(X)`` for any kernel written against the reference's plugin
        interface: ``_Kcore(self, X, X2=None)`` (GPinv/kernels.py:31-40) gives the core, the
        jitter goes on the core before the variance scaling (GPinv/kernels.py:57-58), and the
        factorisation / its reverse are the sm_100a ``potrf`` ops.  A subclass that overrides
        ``Cholesky`` itself (as the reference's RBF2D does, GPinv/kronecker.py:49-66) is
        honoured: its dense [n,n,R] tile becomes R unit-variance cores."""
        if type(self).Cholesky is not Kern.Cholesky:
            dense = type(self).Cholesky(self, X)                   # [n,n,R], user-defined
            R = dense.shape[2]
            cores = [dense[:, :, r].contiguous() for r in range(R)]
            ones = torch.ones(R, dtype=dense.dtype, device=dense.device)
            return CholeskyFactor(cores, ones, list(range(R)), [])
        core = self._Kcore(X, None)
        n = core.shape[0]
        core = core + settings.numerics.jitter_level * torch.eye(n, dtype=core.dtype, device=core.device)
        chol, info = ops.potrf(core.contiguous())
        return CholeskyFactor([chol], torch.sqrt(self.variance), [0] * self.output_dim, [info])

    def _Kcore(self, X, X2=None):
        raise NotImplementedError


class Stationary(Kern):
    """GPinv/kernels.py:72-99 + the GPflow ``kernels.Stationary`` it mixes in (variance and
    lengthscales are positive Params; ``active_dims`` selects input columns)."""
    _mode = None

    def __init__(self, input_dim, output_dim, variance=None, lengthscales=None,
                 active_dims=None, ARD=False):
        Kern.__init__(self, output_dim)
        if variance is None:
            variance = np.ones(output_dim)
        variance = np.asarray(variance, dtype=np.float64)
        assert variance.shape[0] == self.output_dim
        self.input_dim = int(input_dim)
        if active_dims is None:
            self.active_dims = slice(input_dim)
        else:
            self.active_dims = np.array(active_dims, dtype=np.int64)
        self.variance = Param(variance, transforms.positive)
        self.ARD = bool(ARD)
        if ARD:
            if lengthscales is None:
                lengthscales = np.ones(input_dim)
            lengthscales = lengthscales * np.ones(input_dim)
        elif lengthscales is None:
            lengthscales = 1.0
        self.lengthscales = Param(lengthscales, transforms.positive)

    def _slice(self, X, X2):
        ad = object.__getattribute__(self, "active_dims")
        if isinstance(ad, slice):
            if X.shape[1] != self.input_dim:
                X = X[:, ad]
                X2 = None if X2 is None else X2[:, ad]
        else:
            idx = torch.as_tensor(ad, device=X.device)
            X = X.index_select(1, idx)
            X2 = None if X2 is None else X2.index_select(1, idx)
        return X.contiguous(), (None if X2 is None else X2.contiguous())

    def square_dist(self, X, X2):
        """GPflow Stationary.square_dist (expansion form) in torch ops, for user code."""
        X = X / self.lengthscales
        Xs = torch.sum(torch.square(X), 1)
        if X2 is None:
            return -2 * (X @ X.T) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
        X2 = X2 / self.lengthscales
        X2s = torch.sum(torch.square(X2), 1)
        return -2 * (X @ X2.T) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)

    def _Kcore(self, X, X2=None, jitter=0.0):
        X, X2 = self._slice(X, X2)
        return ops.kcore(X, X2, self.lengthscales, self._mode, float(jitter) if X2 is None else 0.0)

    def Cholesky_factor(self, X):
        """chol(core + jitter I) through the fused build-and-factor op (lower triangle only, in
        place); its reverse reduces the blocked sweep's lower triangle straight into d/d(lengthscales)."""
        if type(self)._Kcore is not Stationary._Kcore:
            return Kern.Cholesky_factor(self, X)      # a subclass with its own core: generic path
        Xs, _ = self._slice(X, None)
        # the reverse pass needs L^-1: have it built behind the factorisation, on the side stream
        need_inv = torch.is_grad_enabled() and self.lengthscales.requires_grad
        chol, info, _ = ops.chol_core(Xs, self.lengthscales, self._mode, float(settings.numerics.jitter_level),
                                      bool(need_inv))
        return CholeskyFactor([chol], torch.sqrt(self.variance), [0] * self.output_dim, [info])

    def __str__(self):
        return Parameterized.__str__(self)


class RBF(Stationary):
    """exp(-D/2)  (GPinv/kernels.py:101-107)."""
    _mode = ops.MODE["rbf"]


class _SymBase(RBF):
    _sign = +1.0

    def Kdiag(self, X):
        X, _ = self._slice(X, None)
        Xa = torch.abs(X)
        sq = torch.sum(torch.square((Xa + Xa) / self.lengthscales), 1)
        d = torch.exp(-0.5 * sq)
        return self.variance[None, :] * (1.0 + self._sign * d)[:, None]


class RBF_csym(_SymBase):
    """k(|x|,|x'|) + k(|x|,-|x'|)  (GPinv/kernels.py:109-132)."""
    _mode = ops.MODE["rbf_csym"]
    _sign = +1.0


class RBF_casym(_SymBase):
    """k(|x|,|x'|) - k(|x|,-|x'|)  (GPinv/kernels.py:134-157)."""
    _mode = ops.MODE["rbf_casym"]
    _sign = -1.0


class Stack(Kern):
    """Concatenate kernels along the output axis (GPinv/kernels.py:159-189); one Cholesky per
    member kernel."""

    def __init__(self, list_of_kerns):
        output_dim = 0
        for k in list_of_kerns:
            assert isinstance(k, Kern)
            output_dim += k.output_dim
        Kern.__init__(self, output_dim)
        self.kern_list = ParamList(list_of_kerns)

    def K(self, X, X2=None):
        return torch.cat([k.K(X, X2) for k in self.kern_list], 2)

    def Kdiag(self, X):
        return torch.cat([k.Kdiag(X) for k in self.kern_list], 1)

    def Cholesky_factor(self, X):
        return CholeskyFactor.concat([k.Cholesky_factor(X) for k in self.kern_list])
