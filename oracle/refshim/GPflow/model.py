"""GPflow.model.Model / GPModel: objective = -(build_likelihood() + build_prior()), gradient by
reverse-mode differentiation of that graph (``tf.gradients``; here torch autograd, fp64)."""
import numpy as np
import tensorflow as tf
import torch

from .param import Parameterized, DataHolder, AutoFlow


class ObjectiveWrapper(object):
    def __init__(self, objective):
        self._objective = objective

    def __call__(self, x):
        return self._objective(x)


class Model(Parameterized):
    def __init__(self, name="model"):
        Parameterized.__init__(self)
        self.scoped_keys.extend(["build_likelihood", "build_prior"])
        self._name = name
        self._needs_recompile = True

    @property
    def name(self):
        return self._name

    def build_likelihood(self):
        raise NotImplementedError

    def _compile(self, optimizer=None, **kw):
        self._needs_recompile = False

    def _objective(self, x, random_normals=()):
        """(-f, -df/dx) at the free state x, like GPflow's compiled ``obj`` closure.  The draws
        that ``tf.random_normal`` hands out inside the graph are given explicitly."""
        self._compile()
        xt = torch.tensor(np.asarray(x, dtype=np.float64), dtype=torch.float64, requires_grad=True)
        if len(random_normals):
            tf.set_random_normal_queue(list(random_normals), strict=True)
        n = self.make_tf_array(xt)
        assert n == xt.numel(), (n, xt.numel())
        with self.tf_mode():
            f = self.build_likelihood() + self.build_prior()
        assert tf.random_normal_queue_len() == 0, "unused random draws"
        if len(random_normals):
            tf.set_random_normal_queue([], strict=False)
        f = tf.reduce_sum(f)
        (g,) = torch.autograd.grad(f, xt, allow_unused=True)
        g = torch.zeros_like(xt) if g is None else g
        return -float(f), -g.numpy()

    # ---- GPflow Model.optimize / sample, as far as the reference's tests use them -------------
    def optimize(self, method="L-BFGS-B", tol=None, callback=None, maxiter=1000, **kw):
        """scipy method name, or a ``tf.train`` optimizer object (Adam) for ``maxiter`` stochastic
        steps; random draws inside the graph are fresh per evaluation (non-strict queue)."""
        from scipy.optimize import minimize, OptimizeResult
        self._compile(optimizer=None if isinstance(method, str) else method)

        def obj(x):
            tf.set_random_normal_queue([], strict=False)
            return self._objective(x)

        if isinstance(method, str):
            r = minimize(fun=obj, x0=self.get_free_state(), method=method, jac=True, tol=tol, callback=callback,
                         options=dict(maxiter=maxiter))
            self.set_state(r.x)
            return r
        lr, b1, b2, eps = method.learning_rate, method.beta1, method.beta2, method.epsilon
        x = self.get_free_state().copy()
        m1, m2 = np.zeros_like(x), np.zeros_like(x)
        for t in range(1, maxiter + 1):
            _, g = obj(x)
            m1 = b1 * m1 + (1 - b1) * g
            m2 = b2 * m2 + (1 - b2) * g * g
            x = x - lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t) * m1 / (np.sqrt(m2) + eps)
            if callback is not None:
                callback(x)
        self.set_state(x)
        fun, jac = obj(x)
        return OptimizeResult(fun=fun, jac=jac, x=x, nit=maxiter, success=True, status="Finished iterations.")

    def sample(self, num_samples, Lmax=20, epsilon=0.01, thin=1, burn=0, verbose=False, return_logprobs=False,
               RNG=np.random.RandomState(0)):
        from . import hmc

        def obj(x):
            tf.set_random_normal_queue([], strict=False)
            return self._objective(x)
        self._compile()
        return hmc.sample_HMC(obj, num_samples, Lmax=Lmax, epsilon=epsilon, thin=thin, burn=burn,
                              x0=self.get_free_state(), verbose=verbose, return_logprobs=return_logprobs, RNG=RNG)

    def compute_log_likelihood(self, random_normals=()):
        xt = torch.tensor(self.get_free_state(), dtype=torch.float64)
        tf.set_random_normal_queue(list(random_normals))
        self.make_tf_array(xt)
        with self.tf_mode(), torch.no_grad():
            return float(self.build_likelihood())


class GPModel(Model):
    def __init__(self, X, Y, kern, likelihood, mean_function, name="model"):
        self.kern, self.likelihood, self.mean_function = kern, likelihood, mean_function
        Model.__init__(self, name)
        if isinstance(X, np.ndarray):
            X = DataHolder(X)
        if isinstance(Y, np.ndarray):
            Y = DataHolder(Y)
        self.X, self.Y = X, Y

    def build_predict(self, *args, **kwargs):
        raise NotImplementedError

    def predict_f(self, Xnew, full_cov=False):
        xt = torch.tensor(self.get_free_state(), dtype=torch.float64)
        self.make_tf_array(xt)
        with self.tf_mode(), torch.no_grad():
            mu, var = self.build_predict(tf.constant(Xnew), full_cov=full_cov)
        return mu.numpy(), var.numpy()
