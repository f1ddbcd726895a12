"""Likelihoods (mirror of GPinv/likelihoods.py) plus the two notebook likelihoods that define
the problem-specific transform F.

Interface kept from the reference (GPinv/likelihoods.py:13-62): ``logp(F[S,n,R], Y)`` returns
the elementwise log density (the caller sums it), ``sample_F(F)``, ``sample_Y(F)``.  Models call
``logp_sum`` — by default ``logp(...).sum()`` so user subclasses written against the reference
interface work unchanged (their torch ops run on the GPU and are differentiated by autograd);
the built-in likelihoods override it with fused sm_100a ops and a hand-written backward.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import densities, ops, transforms
from .param import DataHolder, Param, Parameterized

LOG2PI = math.log(2.0 * math.pi)


class Likelihood(Parameterized):
    """GPinv/likelihoods.py:13-62."""
    wants_expF = False   # ask the sampling op to also emit exp(F) from its GEMM epilogue

    def __init__(self):
        Parameterized.__init__(self)

    def logp(self, F, Y):
        raise NotImplementedError("implement the logp function for this likelihood")

    def logp_sum(self, F, Y, expF=None):
        """Scalar sum of ``logp``.  F is the [S,n,R] view of the internal [R,S,n] buffer."""
        return torch.sum(self.logp(F, Y))

    def sample_F(self, F):
        raise NotImplementedError("implement the sample_F function for this likelihood")

    def sample_Y(self, F, noise=None):
        raise NotImplementedError("implement the sample_Y function for this likelihood")


def _internal(F):
    """[S,n,R] view -> contiguous [R,S,n] (free when F came from the sampling op)."""
    return F.permute(2, 0, 1).contiguous()


class Gaussian(Likelihood):
    """GPinv/likelihoods.py:65-82."""

    def __init__(self):
        Likelihood.__init__(self)
        self.variance = Param(1.0, transforms.positive)

    def logp(self, F, Y):
        return densities.gaussian(F, Y[None].expand(F.shape[0], -1, -1), self.variance)

    def logp_sum(self, F, Y, expF=None):
        var = self.variance
        ss = ops.gauss_sumsq(_internal(F), Y.contiguous())
        return ops.gauss_logp(ss, var, float(F.numel())).reshape(())

    def sample_F(self, F):
        return F

    def sample_Y(self, F, noise=None):
        """GPinv/likelihoods.py:78-82: F + sqrt(variance) * N(0,1) (``noise`` fixes the draws)."""
        return F + torch.sqrt(self.variance) * (torch.randn_like(F) if noise is None else noise)


class AbelLikelihood(Likelihood):
    """y = A exp(f) + e  (notebooks/Abel_inversion.ipynb:268-287).  ``Amat`` [M,n] is the
    line-of-sight matrix of testing/make_LosMatrix.py; Y is [M,1]; one latent function."""
    wants_expF = True

    def __init__(self, Amat):
        Likelihood.__init__(self)
        self.Amat = DataHolder(Amat)
        self.variance = Param(np.ones(1), transforms.positive)

    def logp(self, F, Y):
        Af = self.sample_F(F)
        return densities.gaussian(Af, Y[None].expand(F.shape[0], -1, -1), self.variance)

    def logp_sum(self, F, Y, expF=None):
        assert F.shape[2] == 1, "AbelLikelihood is defined for one latent function"
        Fi = _internal(F)[0]
        Ei = torch.exp(Fi).detach() if expF is None else _internal(expF)[0]
        ss, _ = ops.abel_resid(Fi, Ei, self.Amat, Y.reshape(-1).contiguous())
        var = self.variance
        return ops.gauss_logp(ss, var, float(F.shape[0] * self.Amat.shape[0])).reshape(())

    def sample_F(self, F):
        X = torch.exp(F[:, :, 0]).contiguous()               # [S,n]
        return ops.gemm(X, self.Amat, False, True, False)[:, :, None]   # [S,M,1]

    def sample_Y(self, F, noise=None):
        f = self.sample_F(F)
        return f + (torch.randn_like(f) if noise is None else noise) * torch.sqrt(self.variance)


class SpecAbelLikelihood(Likelihood):
    """Spectroscopic Abel inversion, three latent functions (log a, log sigma, v)
    (notebooks/Spectroscopic_Abel_inversion.ipynb:306-356).  Amat, cosTheta [m,n]; lam [k];
    Y [k,m]."""

    def __init__(self, Amat, cosTheta, lam):
        Likelihood.__init__(self)
        self.Amat = DataHolder(Amat)
        self.cosT = DataHolder(cosTheta)
        self.lam = DataHolder(np.asarray(lam, dtype=np.float64).reshape(-1, 1))
        self.variance = Param(np.ones(1), transforms.positive)
        # kernel layout of the geometry: transposed to [n,m] (chord index contiguous).  Kept as data
        # holders that follow Amat / cosT (refreshed in place when those get new data of the same shape,
        # so a captured CUDA graph keeps reading valid memory)
        self._AmatT = DataHolder(np.ascontiguousarray(np.asarray(Amat, dtype=np.float64).T))
        self._cosTT = DataHolder(np.ascontiguousarray(np.asarray(cosTheta, dtype=np.float64).T))
        self._lam_step = self._uniform_step(lam)
        self.lam._on_change = lambda arr: setattr(self, "_lam_step", self._uniform_step(arr))
        self.Amat._on_change = lambda arr: self._AmatT.set_data(np.ascontiguousarray(arr.T))
        self.cosT._on_change = lambda arr: self._cosTT.set_data(np.ascontiguousarray(arr.T))

    def _geometry_T(self):
        return self._AmatT, self._cosTT

    @staticmethod
    def _uniform_step(lam):
        """Spacing of the wavelength grid when it is uniform (np.linspace, as in the notebook) -- the kernels then
        advance the Gaussian line shape along the wavelength index by a recurrence -- else 0.0."""
        lam = np.asarray(lam, dtype=np.float64).reshape(-1)
        if lam.size < 3:
            return 0.0
        step = (lam[-1] - lam[0]) / (lam.size - 1)
        if step == 0.0 or np.max(np.abs(np.diff(lam) - step)) > 1e-12 * max(abs(step), 1e-3 * np.max(np.abs(lam))):
            return 0.0
        return float(step)

    def _dlam(self):
        return self._lam_step

    def sample_F(self, F):
        """[S,k,m] transformed samples (fused sm_100a kernel)."""
        At, cTt = self._geometry_T()
        return ops.spec_transform(_internal(F), At, cTt, self.lam.reshape(-1).contiguous(), self._dlam())

    def sample_F_torch(self, F):
        """The notebook's tile-and-reduce formulation in torch ops (kept for user subclasses)."""
        a = torch.exp(F[:, :, 0])[:, None, None, :]
        s = torch.exp(F[:, :, 1])[:, None, None, :]
        v = F[:, :, 2][:, None, None, :]
        lam = self.lam.reshape(1, -1, 1, 1)
        f = a / (math.sqrt(2 * math.pi) * s) * torch.exp(-0.5 * torch.square((lam - v * self.cosT[None, None]) / s))
        return torch.sum(self.Amat[None, None] * f, 3)

    def logp(self, F, Y):
        f_samples = self.sample_F_torch(F)
        return densities.gaussian(f_samples, Y[None].expand(f_samples.shape[0], -1, -1), self.variance)

    def logp_sum(self, F, Y, expF=None):
        At, cTt = self._geometry_T()
        ss, P = ops.spec_resid(_internal(F), At, cTt, self.lam.reshape(-1).contiguous(), Y.contiguous(), self._dlam())
        var = self.variance
        return ops.gauss_logp(ss, var, float(P.numel())).reshape(())

    def sample_Y(self, F, noise=None):
        f = self.sample_F(F)
        return f + (torch.randn_like(f) if noise is None else noise) * torch.sqrt(self.variance)
