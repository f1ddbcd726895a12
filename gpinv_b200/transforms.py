"""Parameter transforms (GPflow ``transforms`` as re-exported at GPinv/__init__.py:1).
``positive`` is GPflow's Log1pe: y = log(1 + exp(x)) + 1e-6 (used at GPinv/stvgp.py:69,
GPinv/likelihoods.py:68 and by GPflow Stationary for variance / lengthscales)."""
import numpy as np
import torch


class Transform(object):
    def forward(self, x):
        raise NotImplementedError

    def backward(self, y):
        raise NotImplementedError

    def torch_forward(self, x):
        raise NotImplementedError

    def torch_log_jacobian(self, x):
        raise NotImplementedError

    def __str__(self):
        raise NotImplementedError


class Identity(Transform):
    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def torch_forward(self, x):
        return x

    def torch_log_jacobian(self, x):
        return torch.zeros((), dtype=x.dtype, device=x.device)

    def __str__(self):
        return "(none)"


class Exp(Transform):
    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return np.exp(x) + self._lower

    def backward(self, y):
        return np.log(y - self._lower)

    def torch_forward(self, x):
        return torch.exp(x) + self._lower

    def torch_log_jacobian(self, x):
        return torch.sum(x)

    def __str__(self):
        return "+ve"


class Log1pe(Transform):
    """y = log(1 + exp(x)) + lower  (softplus)."""

    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return np.logaddexp(0.0, x) + self._lower

    def backward(self, y):
        y = np.asarray(y, dtype=np.float64) - self._lower
        return np.where(y > 35.0, y, np.log(np.expm1(np.minimum(y, 35.0))))

    def torch_forward(self, x):
        return torch.nn.functional.softplus(x, beta=1.0, threshold=1e9) + self._lower

    def torch_log_jacobian(self, x):
        return -torch.sum(torch.nn.functional.softplus(-x, beta=1.0, threshold=1e9))

    def __str__(self):
        return "+ve"


class Logistic(Transform):
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = float(a), float(b)

    def forward(self, x):
        return self.a + (self.b - self.a) / (1.0 + np.exp(-x))

    def backward(self, y):
        p = (np.asarray(y) - self.a) / (self.b - self.a)
        return -np.log(1.0 / p - 1.0)

    def torch_forward(self, x):
        return self.a + (self.b - self.a) * torch.sigmoid(x)

    def torch_log_jacobian(self, x):
        return torch.sum(x - 2.0 * torch.nn.functional.softplus(x) + np.log(self.b - self.a))

    def __str__(self):
        return "[%g, %g]" % (self.a, self.b)


positive = Log1pe()
