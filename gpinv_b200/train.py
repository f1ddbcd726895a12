"""Optimizer descriptors accepted by ``Model.optimize`` where the reference passes a
``tf.train`` optimizer object (testing/test_stvgp.py:31-35, notebooks/Abel_inversion.ipynb:460)."""
import torch


class _TorchOptimizer(object):
    _cls = None

    def __init__(self, learning_rate=0.001, **kw):
        self.learning_rate = learning_rate
        self.kw = kw

    def make(self, params):
        return self._cls(params, lr=self.learning_rate, **self.kw)


class AdamOptimizer(_TorchOptimizer):
    """tf.train.AdamOptimizer(learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8).  ``Model.optimize``
    runs it as TensorFlow's update fused on the device (``gpinv::adam_step``,
    ``Model._optimize_adam_device``); ``make`` (a torch.optim.Adam) is kept for user loops."""
    _cls = torch.optim.Adam

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        _TorchOptimizer.__init__(self, learning_rate, betas=(beta1, beta2), eps=epsilon)


class GradientDescentOptimizer(_TorchOptimizer):
    _cls = torch.optim.SGD


class AdagradOptimizer(_TorchOptimizer):
    _cls = torch.optim.Adagrad


class RMSPropOptimizer(_TorchOptimizer):
    _cls = torch.optim.RMSprop
