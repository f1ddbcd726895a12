"""Mean functions returning [n, R] (mirror of GPinv/mean_functions.py)."""
import numpy as np
import torch

from .param import Param, ParamList, Parameterized


class MeanFunction(Parameterized):
    """GPinv/mean_functions.py:9-29."""

    def __init__(self, output_dim):
        Parameterized.__init__(self)
        self.output_dim = output_dim

    def __call__(self, X):
        raise NotImplementedError("Implement the __call__ method for this mean function")


class Zero(MeanFunction):
    """GPinv/mean_functions.py:31-34."""

    def __call__(self, X):
        return torch.zeros((X.shape[0], self.output_dim), dtype=torch.float64, device=X.device)


class Constant(MeanFunction):
    """GPinv/mean_functions.py:36-45."""

    def __init__(self, output_dim, c=None):
        MeanFunction.__init__(self, output_dim)
        if c is None:
            c = np.ones(output_dim, np.float64)
        self.c = Param(c)

    def __call__(self, X):
        return self.c[None, :].expand(X.shape[0], -1)


class Stack(MeanFunction):
    """GPinv/mean_functions.py:47-72."""

    def __init__(self, list_of_means):
        output_dim = 0
        for m in list_of_means:
            output_dim += m.output_dim
        MeanFunction.__init__(self, output_dim)
        self.mean_list = ParamList(list_of_means)

    def __call__(self, X):
        return torch.cat([l(X) for l in self.mean_list], 1)
