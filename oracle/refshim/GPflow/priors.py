"""GPflow.priors.Gaussian."""
import tensorflow as tf
from . import densities


class Gaussian(object):
    def __init__(self, mu, var):
        self.mu, self.var = float(mu), float(var)

    def logp(self, x):
        return tf.reduce_sum(densities.gaussian(x, self.mu, tf.constant(self.var)))
