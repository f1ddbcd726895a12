"""GPflow.kernels.Kern / Stationary: parameters (variance, lengthscales: positive), ``_slice``
and the expansion-form ``square_dist`` (divide by lengthscales; -2 X X'^T + |X|^2_i + |X'|^2_j)."""
import numpy as np
import tensorflow as tf

from . import transforms
from .param import Param, Parameterized


class Kern(Parameterized):
    def __init__(self, input_dim, active_dims=None):
        Parameterized.__init__(self)
        self.scoped_keys.extend(["K", "Kdiag"])
        self.input_dim = input_dim
        if active_dims is None:
            self.active_dims = slice(input_dim)
        else:
            self.active_dims = np.array(active_dims, dtype=np.int32)
            assert len(active_dims) == input_dim

    def _slice(self, X, X2):
        if isinstance(self.active_dims, slice):
            X = X[:, self.active_dims]
            if X2 is not None:
                X2 = X2[:, self.active_dims]
        else:
            idx = [int(i) for i in self.active_dims]
            X = X[:, idx]
            if X2 is not None:
                X2 = X2[:, idx]
        return X, X2


class Stationary(Kern):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        Kern.__init__(self, input_dim, active_dims)
        self.scoped_keys.extend(["square_dist", "euclid_dist"])
        self.variance = Param(variance, transforms.positive)
        if ARD:
            if lengthscales is None:
                lengthscales = np.ones(input_dim, np.float64)
            else:
                lengthscales = lengthscales * np.ones(input_dim, np.float64)
            self.lengthscales = Param(lengthscales, transforms.positive)
            self.ARD = True
        else:
            if lengthscales is None:
                lengthscales = 1.0
            self.lengthscales = Param(lengthscales, transforms.positive)
            self.ARD = False

    def square_dist(self, X, X2):
        X = X / self.lengthscales
        Xs = tf.reduce_sum(tf.square(X), 1)
        if X2 is None:
            return -2 * tf.matmul(X, tf.transpose(X)) + \
                tf.reshape(Xs, (-1, 1)) + tf.reshape(Xs, (1, -1))
        X2 = X2 / self.lengthscales
        X2s = tf.reduce_sum(tf.square(X2), 1)
        return -2 * tf.matmul(X, tf.transpose(X2)) + \
            tf.reshape(Xs, (-1, 1)) + tf.reshape(X2s, (1, -1))

    def Kdiag(self, X):
        return self.variance * tf.ones((tf.shape(X)[0],), tf.float64)


class RBF(Stationary):
    """GPflow's own (single-output) RBF kernel: variance * exp(-d^2 / 2)."""

    def K(self, X, X2=None):
        X, X2 = self._slice(X, X2)
        return self.variance * tf.exp(-self.square_dist(X, X2) / 2)
