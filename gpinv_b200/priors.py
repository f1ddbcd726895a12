"""Priors (GPflow ``priors`` as re-exported at GPinv/__init__.py:1; GPflow's GPMC puts
``Gaussian(0, 1)`` on the whitened latent ``V``, inherited at GPinv/gpmc.py:37)."""
import torch

from . import densities


class Prior(object):
    def logp(self, x):
        raise NotImplementedError


class Gaussian(Prior):
    def __init__(self, mu, var):
        self.mu, self.var = float(mu), float(var)

    def logp(self, x):
        mu = torch.full((), self.mu, dtype=x.dtype, device=x.device)
        var = torch.full((), self.var, dtype=x.dtype, device=x.device)
        return torch.sum(densities.gaussian(mu, x, var))

    def __str__(self):
        return "N(%g,%g)" % (self.mu, self.var)


class LogNormal(Prior):
    def __init__(self, mu, var):
        self.mu, self.var = float(mu), float(var)

    def logp(self, x):
        mu = torch.full((), self.mu, dtype=x.dtype, device=x.device)
        var = torch.full((), self.var, dtype=x.dtype, device=x.device)
        return torch.sum(densities.lognormal(x, mu, var))

    def __str__(self):
        return "logN(%g,%g)" % (self.mu, self.var)
