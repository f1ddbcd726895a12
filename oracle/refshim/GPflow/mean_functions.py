"""GPflow.mean_functions.MeanFunction base."""
from .param import Parameterized


class MeanFunction(Parameterized):
    def __call__(self, X):
        raise NotImplementedError("Implement the __call__ method for this mean function")
