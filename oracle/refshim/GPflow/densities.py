"""GPflow.densities (gaussian as published: -0.5 log 2pi - 0.5 log var - 0.5 (mu - x)^2 / var)."""
import numpy as np
import tensorflow as tf


def gaussian(x, mu, var):
    return -0.5 * np.log(2 * np.pi) - 0.5 * tf.log(var) - 0.5 * tf.square(mu - x) / var
