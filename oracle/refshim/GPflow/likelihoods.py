"""GPflow.likelihoods.Likelihood base (only the class identity is used by the reference)."""
from .param import Parameterized


class Likelihood(Parameterized):
    def __init__(self):
        Parameterized.__init__(self)
        self.scoped_keys.extend(["logp", "variational_expectations", "predict_mean_and_var",
                                 "predict_density"])


class Gaussian(Likelihood):
    """GPflow's Gaussian likelihood (used by GPR): noise variance, positive, initialised to 1."""

    def __init__(self):
        from .param import Param
        from . import transforms
        Likelihood.__init__(self)
        self.variance = Param(1.0, transforms.positive)
