"""GPflow.gpmc.GPMC constructor: whitened latent V [n, num_latent] with a N(0,1) prior."""
import numpy as np

from .model import GPModel
from .param import Param, DataHolder
from .priors import Gaussian


class GPMC(GPModel):
    def __init__(self, X, Y, kern, likelihood, mean_function=None, num_latent=None):
        X = DataHolder(X, on_shape_change="recompile")
        Y = DataHolder(Y, on_shape_change="recompile")
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function)
        self.num_data = X.shape[0]
        self.num_latent = num_latent or Y.shape[1]
        self.V = Param(np.zeros((self.num_data, self.num_latent)))
        self.V.prior = Gaussian(0., 1.)
