"""GPflow.transforms: Identity and Log1pe (``positive``: y = log(1 + exp(x)) + 1e-6)."""
import numpy as np
import tensorflow as tf


class Transform(object):
    def forward(self, x):
        raise NotImplementedError

    def backward(self, y):
        raise NotImplementedError

    def tf_forward(self, x):
        raise NotImplementedError

    def tf_log_jacobian(self, x):
        raise NotImplementedError


class Identity(Transform):
    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def tf_forward(self, x):
        return x

    def tf_log_jacobian(self, x):
        return tf.zeros((1,), tf.float64)

    def __str__(self):
        return "(none)"


class Log1pe(Transform):
    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x))) + self._lower

    def backward(self, y):
        ys = np.asarray(y, dtype=np.float64) - self._lower
        return np.where(ys > 35.0, ys, np.log(np.expm1(np.minimum(ys, 35.0))))

    def tf_forward(self, x):
        return tf.nn.softplus(x) + self._lower

    def tf_log_jacobian(self, x):
        return -tf.reduce_sum(tf.log(1. + tf.exp(-x)))

    def __str__(self):
        return "+ve"


positive = Log1pe()
