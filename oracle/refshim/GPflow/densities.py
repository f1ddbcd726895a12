"""GPflow.densities (gaussian as published: -0.5 log 2pi - 0.5 log var - 0.5 (mu - x)^2 / var)."""
import numpy as np
import tensorflow as tf


def gaussian(x, mu, var):
    return -0.5 * np.log(2 * np.pi) - 0.5 * tf.log(var) - 0.5 * tf.square(mu - x) / var


def multivariate_normal(x, mu, L):
    """log N(x | mu, L L^T) summed over the columns of x (GPflow.densities.multivariate_normal)."""
    d = x - mu
    alpha = tf.matrix_triangular_solve(L, d, lower=True)
    num_col = tf.shape(x)[1]
    num_dims = tf.shape(x)[0]
    ret = -0.5 * num_dims * num_col * np.log(2 * np.pi)
    ret += -num_col * tf.reduce_sum(tf.log(tf.batch_matrix_diag_part(L)))
    ret += -0.5 * tf.reduce_sum(tf.square(alpha))
    return ret
