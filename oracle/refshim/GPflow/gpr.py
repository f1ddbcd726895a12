"""GPflow.gpr.GPR (restated): exact GP regression, the model the reference's tests compare with."""
import tensorflow as tf

from . import densities, likelihoods
from .mean_functions import Zero
from .model import GPModel
from .param import DataHolder
from .tf_wraps import eye


class GPR(GPModel):
    def __init__(self, X, Y, kern, mean_function=None):
        likelihood = likelihoods.Gaussian()
        X = DataHolder(X, on_shape_change="pass")
        Y = DataHolder(Y, on_shape_change="pass")
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function or Zero())
        self.num_latent = Y.shape[1]

    def build_likelihood(self):
        K = self.kern.K(self.X) + eye(tf.shape(self.X)[0]) * self.likelihood.variance
        L = tf.cholesky(K)
        m = self.mean_function(self.X)
        return densities.multivariate_normal(self.Y, m, L)

    def build_predict(self, Xnew, full_cov=False):
        Kx = self.kern.K(self.X, Xnew)
        K = self.kern.K(self.X) + eye(tf.shape(self.X)[0]) * self.likelihood.variance
        L = tf.cholesky(K)
        A = tf.matrix_triangular_solve(L, Kx, lower=True)
        V = tf.matrix_triangular_solve(L, self.Y - self.mean_function(self.X))
        fmean = tf.matmul(tf.transpose(A), V) + self.mean_function(Xnew)
        if full_cov:
            fvar = self.kern.K(Xnew) - tf.matmul(tf.transpose(A), A)
            fvar = tf.tile(tf.expand_dims(fvar, 2), [1, 1, tf.shape(self.Y)[1]])
        else:
            fvar = self.kern.Kdiag(Xnew) - tf.reduce_sum(tf.square(A), 0)
            fvar = tf.tile(tf.reshape(fvar, (-1, 1)), [1, tf.shape(self.Y)[1]])
        return fmean, fvar
