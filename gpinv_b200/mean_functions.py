"""Prior mean of the latent functions, evaluated on the device.

Drop-in for the three mean functions of the reference (GPinv/mean_functions.py): every mean maps
the latent inputs ``X [n, D]`` to an ``[n, R]`` tensor, one column per latent function (GPflow's
own mean functions return ``[n, 1]``; the multi-output shape is what ``StVGP._sample`` and
``GPMC._get_f`` add to ``L u``).  The values are tiny (``n x R`` doubles), so they are plain torch
expressions on CUDA tensors; their gradients (only ``Constant.c`` has one) come from autograd and
end up in the fused sampling kernel's ``mean_bar`` output.
"""
import numpy as np
import torch

from .param import Param, ParamList, Parameterized


class MeanFunction(Parameterized):
    """Base class: ``output_dim`` latent functions (GPinv/mean_functions.py:9-29)."""

    def __init__(self, output_dim):
        Parameterized.__init__(self)
        if int(output_dim) < 1:
            raise ValueError("output_dim must be a positive integer, got %r" % (output_dim,))
        self.output_dim = int(output_dim)

    def __call__(self, X):
        raise NotImplementedError("%s does not implement __call__(X) -> [n, output_dim]" % type(self).__name__)

    def _rows(self, X):
        if X.dim() != 2:
            raise ValueError("mean functions take X as an [n, D] matrix")
        return X.shape[0]


class Zero(MeanFunction):
    """m(x) = 0 for every latent function (GPinv/mean_functions.py:31-34)."""

    def __call__(self, X):
        return X.new_zeros((self._rows(X), self.output_dim))


class Constant(MeanFunction):
    """m_r(x) = c_r, a trainable offset per latent function, initialised to one
    (GPinv/mean_functions.py:36-45)."""

    def __init__(self, output_dim, c=None):
        MeanFunction.__init__(self, output_dim)
        value = np.ones(self.output_dim) if c is None else np.asarray(c, dtype=np.float64).reshape(-1)
        if value.size != self.output_dim:
            raise ValueError("Constant mean: c has %d entries for %d latent functions" % (value.size, self.output_dim))
        self.c = Param(value)

    def __call__(self, X):
        # a broadcast view, not a copy: [1, R] -> [n, R]
        return self.c.reshape(1, self.output_dim).expand(self._rows(X), self.output_dim)


class Stack(MeanFunction):
    """Column-wise concatenation of several mean functions; its ``output_dim`` is the sum of
    theirs (GPinv/mean_functions.py:47-72).  Used with ``kernels.Stack`` when groups of latent
    functions have different priors (spectroscopic notebook: two log-quantities around -2, one
    velocity around 0)."""

    def __init__(self, list_of_means):
        parts = list(list_of_means)
        if not parts or not all(isinstance(m, MeanFunction) for m in parts):
            raise TypeError("Stack expects a non-empty list of MeanFunction objects")
        MeanFunction.__init__(self, sum(m.output_dim for m in parts))
        self.mean_list = ParamList(parts)

    def __call__(self, X):
        return torch.cat([mean(X) for mean in self.mean_list], dim=1)
