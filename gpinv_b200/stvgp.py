"""StVGP: stochastic approximation of the variational Gaussian process (mirror of
GPinv/stvgp.py).  ``build_likelihood`` / ``_sample`` read like the reference's graph code; the
arithmetic runs in the sm_100a ops of ``gpinv_b200.ops``.

One deliberate extension of the reference interface: ``_objective(x, eps=...)`` accepts the
standard-normal draws ``eps`` [R,S,n] (the reference draws them inside the TF graph,
GPinv/stvgp.py:164, which makes its ELBO irreproducible across calls); ``seed=`` fixes the
device Philox stream instead.
"""
from __future__ import annotations

import numpy as np
import torch

from . import conditionals, ops, transforms
from .mean_functions import Zero
from .model import GPModel
from .param import DataHolder, Param, default_device


class StVGP(GPModel):
    def __init__(self, X, Y, kern, likelihood, mean_function=None, num_latent=None,
                 q_diag=False, KL_analytic=False, num_samples=20, seed=None):
        """Arguments as GPinv/stvgp.py:38-50.  X [n,D]; Y any shape the likelihood accepts
        (``num_latent`` must be given when Y is not [n,R])."""
        self.num_data = X.shape[0]
        self.num_latent = num_latent or Y.shape[1]
        self.num_samples = num_samples
        if mean_function is None:
            mean_function = Zero(self.num_latent)
        Y = DataHolder(Y, on_shape_change="recompile")
        X = DataHolder(X, on_shape_change="recompile")
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function)
        self.q_diag = q_diag
        self.KL_analytic = KL_analytic
        self._init_variational()
        self._generator = None
        self._seed = seed
        self._KL = None

    def _init_variational(self):
        """GPinv/stvgp.py:62-74: q_mu = 0; q_sqrt = I per latent (no transform) or ones (positive)."""
        self.q_mu = Param(np.zeros((self.num_data, self.num_latent)))
        if self.q_diag:
            self.q_sqrt = Param(np.ones((self.num_data, self.num_latent)), transforms.positive)
        else:
            q_sqrt = np.array([np.eye(self.num_data) for _ in range(self.num_latent)]).swapaxes(0, 2)
            self.q_sqrt = Param(q_sqrt)
            if self.num_latent == 1:
                # only the lower triangle of each [n,n] slice is ever read (band_part,
                # GPinv/stvgp.py:157): lets the host API upload the lower trapezoids only
                self.q_sqrt._tril_rows = self.num_data

    def _prepare(self):
        """GPinv/stvgp.py:76-94: re-create the variational parameters when X changed size."""
        n = object.__getattribute__(self, "X").shape[0]
        if self.num_data != n:
            self.num_data = n
            self._init_variational()

    def _draw_eps(self, S):
        dev = default_device()
        if self._generator is None:
            self._generator = torch.Generator(device=dev)
            if self._seed is not None:
                self._generator.manual_seed(int(self._seed))
            else:
                self._generator.seed()
        return torch.randn((self.num_latent, S, self.num_data), dtype=torch.float64, device=dev,
                           generator=self._generator)

    def _eps_full(self, eps, S):
        """All S draws [R,S,n] on the device (drawn here when not given)."""
        if eps is None:
            eps = self._draw_eps(S)
        elif not isinstance(eps, torch.Tensor):
            eps = torch.as_tensor(np.ascontiguousarray(eps, dtype=np.float64)).to(default_device())
        return eps

    def _eps_tensor(self, eps, S):
        """This rank's share of the draws (all of them without sample sharding)."""
        eps = self._eps_full(eps, S)
        if self.dist is not None:
            eps = self.dist.shard_samples(eps)
        return eps.contiguous()

    def _graph_inputs(self, kw):
        # the graph's static input holds ALL draws: build_likelihood takes its shard inside the
        # capture and divides by the global S
        eps = kw.get("eps")
        S = self.num_samples if eps is None else eps.shape[1]
        return {"eps": self._eps_full(eps, S).contiguous()}

    def _refill_graph_inputs(self, dyn):
        if "eps" in dyn:
            if self._generator is None:
                self._draw_eps(1)                     # creates / seeds the device generator
            dyn["eps"].normal_(generator=self._generator)

    def build_likelihood(self, eps=None):
        """Variational lower bound with stochastic approximation (GPinv/stvgp.py:96-107)."""
        S_global = self.num_samples if eps is None else eps.shape[1]
        eps = self._eps_tensor(eps, S_global)
        f_samples, expF = self._sample(eps)
        lik = self.likelihood.logp_sum(f_samples, self.Y, expF=expF)
        share = 1.0 if self.dist is None else 1.0 / self.dist.world_size
        if not self.KL_analytic:
            return (lik - self._KL) / S_global
        return lik / S_global - share * self._analytical_KL()

    def build_predict(self, Xnew, full_cov=False):
        """GPinv/stvgp.py:109-122."""
        mu, var = conditionals.conditional(Xnew, self.X, self.kern, self.q_mu,
                                           q_sqrt=self.q_sqrt, full_cov=full_cov, whiten=True)
        return mu + self.mean_function(Xnew), var

    def sample_F(self, n_sample, eps=None):
        """Samples of the transformed latent function, [n_sample, ...] (GPinv/stvgp.py:124-132).
        ``eps`` [R,n_sample,n] fixes the draws (extension, like ``_objective(x, eps=)``)."""
        def fn():
            f, _ = self._sample(self._eps_tensor(eps, int(n_sample)))
            return self.likelihood.sample_F(f)
        return self._run(fn)

    def sample_Y(self, n_sample, eps=None, noise=None):
        """GPinv/stvgp.py:134-142.  ``noise`` (shape of the result) fixes the observation-noise
        draws of ``Likelihood.sample_Y`` (GPinv/likelihoods.py:78-82)."""
        def fn():
            f, _ = self._sample(self._eps_tensor(eps, int(n_sample)))
            if noise is None:
                return self.likelihood.sample_Y(f)
            nz = torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float64)).to(default_device())
            return self.likelihood.sample_Y(f, noise=nz)
        return self._run(fn)

    def _sample(self, eps):
        """f = mean + L (q_mu + tril(q_sqrt) eps) for every draw; stores the stochastic KL in
        ``self._KL`` (GPinv/stvgp.py:144-185).  Returns ([S,n,R] views of) F and exp(F)."""
        R = self.num_latent
        U_pre = kl_pre = None
        if self._early_u_applies(eps):
            # u = q_mu + tril(q_sqrt) eps does not depend on the Cholesky factor: prepare its operands first, and
            # run it on a low-priority stream from the middle of the factorisation on (ops.sample_u_early)
            q_mu = self.q_mu.T.contiguous()
            Q = ops.tril_unpack(self.q_sqrt)
            U_pre = torch.empty_like(eps)
            kl_pre = torch.empty(4, dtype=torch.float64, device=eps.device)
            ops.mark_inputs_ready()
            chol = self.kern.Cholesky_factor(self.X)
            ops.sample_u_early(Q.detach(), q_mu.detach(), eps, False, U_pre, kl_pre)
        else:
            chol = self.kern.Cholesky_factor(self.X)
            q_mu = self.q_mu.T.contiguous()                        # [R,n]
            if self.q_diag:
                Q = self.q_sqrt.T.contiguous()                     # [R,n]
            else:
                Q = ops.tril_unpack(self.q_sqrt)                   # [R,n,n], zeros above diag
        self._pending_infos += chol.infos
        mean = self.mean_function(self.X).T.contiguous()          # [R,n]
        want_exp = bool(getattr(self.likelihood, "wants_expF", False))
        Fs, Es, KLs = [], [], []
        for core, r0, r1 in chol.groups():
            # (one group over all latent functions: no slices -- the reverse of a slice is a zero fill plus a copy,
            #  eight extra launches per evaluation for the common single-kernel model)
            def part(t, r0=r0, r1=r1):
                return t if (r0 == 0 and r1 == t.shape[0]) else t[r0:r1]
            if chol.kron:
                from . import kronecker
                F, expF, KL = kronecker.kron_sample(core, part(chol.sqrt_var), part(Q), part(q_mu),
                                                    part(mean), part(eps), self.q_diag, want_exp)
            else:
                F, expF, KL, _ = ops.stvgp_sample(core, part(chol.sqrt_var).contiguous(), part(Q),
                                                  part(q_mu), part(mean), part(eps),
                                                  self.q_diag, want_exp, U_pre, kl_pre)
            Fs.append(F)
            Es.append(expF)
            KLs.append(KL)
        F = Fs[0] if len(Fs) == 1 else torch.cat(Fs, 0)
        expF = None
        if want_exp:
            expF = (Es[0] if len(Es) == 1 else torch.cat(Es, 0)).permute(1, 2, 0)
        KL = KLs[0]
        for k in KLs[1:]:
            KL = KL + k
        self._KL = KL
        return F.permute(1, 2, 0), expF

    EARLY_U_MIN_N = 2048          # factor size from which the first sampling product is hidden in the factorisation
    EARLY_U_MIN_WORK = 1 << 32    # ... and the minimum S n^2 of that product

    def _early_u_applies(self, eps):
        """One latent function, one plain stationary kernel (one factorisation, no Kronecker structure), a full
        q_sqrt that is already on the device (not the pending upload of the pipelined host path)."""
        import os
        from . import kernels
        if os.environ.get("GPINV_EARLY_U", "1") == "0" or self.q_diag or self.num_latent != 1:
            return False
        if self.dist is not None and self.dist.world_size > 1:
            return False          # (per-rank product is S/G rows: little to hide; not validated under sharding)
        n = self.num_data
        if n < self.EARLY_U_MIN_N or eps.shape[1] * n * n < self.EARLY_U_MIN_WORK:
            return False
        if type(self.kern) not in (kernels.RBF, kernels.RBF_csym, kernels.RBF_casym):
            return False
        return object.__getattribute__(self, "q_sqrt")._before_use is None

    def _analytical_KL(self):
        """GPinv/stvgp.py:187-195 -> GPflow gauss_kl_white[_diag]; one reduction kernel over q_mu and the
        lower triangle of q_sqrt in the reference's own [n,n,R] layout (and one for its gradient)."""
        return ops.kl_white(self.q_mu.contiguous(), self.q_sqrt.contiguous(), bool(self.q_diag)).reshape(())
