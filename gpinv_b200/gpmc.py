"""GPMC: whitened MCMC model (mirror of GPinv/gpmc.py; prior and sampler from GPflow's
``gpmc.GPMC`` / ``hmc.sample_HMC``).  f = L V + mean(X), V ~ N(0, I)."""
from __future__ import annotations

import numpy as np
import torch

from . import conditionals, ops, priors
from .mean_functions import Zero
from .model import GPModel
from .param import DataHolder, Param


class GPMC(GPModel):
    def __init__(self, X, Y, kern, likelihood, mean_function=None, num_latent=None):
        """GPinv/gpmc.py:31-37."""
        num_latent = num_latent or Y.shape[1]
        if mean_function is None:
            mean_function = Zero(num_latent)
        X = DataHolder(X, on_shape_change="recompile")
        Y = DataHolder(Y, on_shape_change="recompile")
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function)
        self.num_data = X.shape[0]
        self.num_latent = num_latent
        self.V = Param(np.zeros((self.num_data, self.num_latent)))
        self.V.prior = priors.Gaussian(0., 1.)

    def build_likelihood(self):
        """log p(Y | V, theta)  (GPinv/gpmc.py:39-46); the prior on V is added by build_prior."""
        f, expF = self._get_f()
        return self.likelihood.logp_sum(f, self.Y, expF=expF)

    def build_predict(self, Xnew, full_cov=False):
        """GPinv/gpmc.py:48-51."""
        mu, var = conditionals.conditional(Xnew, self.X, self.kern, self.V,
                                           q_sqrt=None, full_cov=full_cov, whiten=True)
        return mu + self.mean_function(Xnew), var

    def sample_F(self, n_sample=1):
        """GPinv/gpmc.py:53-60 (the current state is one posterior sample)."""
        return self._run(lambda: self.likelihood.sample_F(self._get_f()[0]))

    def sample_Y(self, n_sample=1, noise=None):
        """GPinv/gpmc.py:62-69 (``noise`` fixes the observation-noise draws)."""
        def fn():
            f = self._get_f()[0]
            if noise is None:
                return self.likelihood.sample_Y(f)
            nz = torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float64)).to(f.device)
            return self.likelihood.sample_Y(f, noise=nz)
        return self._run(fn)

    def _get_f(self):
        """f = L V + mean, as a [1,n,R] sample set (GPinv/gpmc.py:71-82).  Reuses the sampling op
        with eps = 0 and a unit diagonal q_sqrt, so u = V and the same fused backward applies."""
        R, n = self.num_latent, self.num_data
        chol = self.kern.Cholesky_factor(self.X)
        self._pending_infos += chol.infos
        mean = self.mean_function(self.X).T.contiguous()
        V = self.V.T.contiguous()
        dev = V.device
        want_exp = bool(getattr(self.likelihood, "wants_expF", False))
        ones = torch.ones((R, n), dtype=torch.float64, device=dev)
        eps = torch.zeros((R, 1, n), dtype=torch.float64, device=dev)
        Fs, Es = [], []
        for core, r0, r1 in chol.groups():
            def part(t, r0=r0, r1=r1):       # (no slices when one group covers every latent function: see StVGP._sample)
                return t if (r0 == 0 and r1 == t.shape[0]) else t[r0:r1]
            if chol.kron:
                from . import kronecker
                F, expF, _ = kronecker.kron_sample(core, part(chol.sqrt_var), part(ones), part(V),
                                                   part(mean), part(eps), True, want_exp)
            else:
                F, expF, _, _ = ops.stvgp_sample(core, part(chol.sqrt_var).contiguous(), part(ones),
                                                 part(V), part(mean), part(eps), True, want_exp)
            Fs.append(F)
            Es.append(expF)
        F = Fs[0] if len(Fs) == 1 else torch.cat(Fs, 0)
        expF = None
        if want_exp:
            expF = (Es[0] if len(Es) == 1 else torch.cat(Es, 0)).permute(1, 2, 0)
        return F.permute(1, 2, 0), expF
