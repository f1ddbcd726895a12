"""Eager torch-CPU stand-in for the TensorFlow <= 0.11 ops used by /root/reference/GPinv and its
notebooks.  TEST INFRASTRUCTURE ONLY -- see oracle/refshim/README.md.

Every function takes / returns ``torch.float64`` CPU tensors (numpy arrays and Python scalars
are converted), so the reference's graph-building code runs eagerly and ``tf.gradients`` is
torch autograd.  Argument conventions are TF 0.10's (``tf.concat(dim, values)``,
``reduction_indices`` positional, ``tf.unpack``/``tf.pack``)."""
import numpy as np
import torch

float64 = torch.float64
float32 = torch.float32
int32 = torch.int32
__version__ = "0.10.0-refshim"

_RANDOM_NORMAL_QUEUE = []
_STRICT = [False]
_AUTO_GEN = [torch.Generator().manual_seed(0)]


def set_random_normal_queue(draws, strict=True):
    """Pre-set the values the next ``tf.random_normal`` calls return (in call order).  With
    ``strict`` (fixture generation) a call that finds the queue empty is an error; otherwise it
    draws from the module generator like the real op would (optimisation, sampling)."""
    del _RANDOM_NORMAL_QUEUE[:]
    _RANDOM_NORMAL_QUEUE.extend(draws)
    _STRICT[0] = bool(strict)


def random_normal_queue_len():
    return len(_RANDOM_NORMAL_QUEUE)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=torch.float64)


class _Shape(object):
    def __init__(self, shape):
        self.dims = tuple(int(s) for s in shape)
        self.ndims = len(self.dims)

    def as_list(self):
        return list(self.dims)


torch.Tensor.get_shape = lambda self: _Shape(self.shape)      # Tensor.get_shape().ndims


def shape(x):
    return tuple(int(s) for s in _t(x).shape)


def size(x):
    return int(_t(x).numel())


def constant(x, dtype=None):
    return _t(x)


def cast(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(x, dtype=dtype)


def expand_dims(x, dim):
    return torch.unsqueeze(_t(x), dim)


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*[int(p) for p in perm])


def reshape(x, shp):
    return _t(x).reshape(*[int(s) for s in shp])


def squeeze(x, squeeze_dims=None):
    x = _t(x)
    if squeeze_dims is None:
        return x.squeeze()
    for d in sorted([d % x.dim() for d in squeeze_dims], reverse=True):
        x = x.squeeze(d)
    return x


def concat(concat_dim, values):
    return torch.cat([_t(v) for v in values], int(concat_dim))


def pack(values, axis=0):
    return torch.stack([_t(v) for v in values], axis)


def unpack(value, num=None, axis=0):
    out = torch.unbind(_t(value), axis)
    assert num is None or len(out) == num
    return list(out)


def zeros(shp, dtype=float64):
    return torch.zeros(*[int(s) for s in shp], dtype=dtype)


def ones(shp, dtype=float64):
    return torch.ones(*[int(s) for s in shp], dtype=dtype)


def ones_like(x):
    return torch.ones_like(_t(x))


def zeros_like(x):
    return torch.zeros_like(_t(x))


def reduce_sum(x, reduction_indices=None, keep_dims=False):
    x = _t(x)
    if reduction_indices is None:
        return x.sum()
    return x.sum(dim=reduction_indices, keepdim=keep_dims)


def square(x):
    x = _t(x)
    return x * x


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def sqrt(x):
    return torch.sqrt(_t(x))


def abs(x):
    return torch.abs(_t(x))


def neg(x):
    return -_t(x)


def add_n(xs):
    out = _t(xs[0])
    for v in xs[1:]:
        out = out + _t(v)
    return out


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def batch_matmul(x, y, adj_x=False, adj_y=False):
    return matmul(x, y, transpose_a=adj_x, transpose_b=adj_y)


def batch_matrix_band_part(x, num_lower, num_upper):
    x = _t(x)
    if num_lower == -1 and num_upper == 0:
        return torch.tril(x)
    if num_lower == 0 and num_upper == -1:
        return torch.triu(x)
    raise NotImplementedError((num_lower, num_upper))


matrix_band_part = batch_matrix_band_part


def batch_matrix_diag(x):
    return torch.diag_embed(_t(x))


def batch_matrix_diag_part(x):
    return torch.diagonal(_t(x), dim1=-2, dim2=-1)


def batch_matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False):
    m = _t(matrix)
    if adjoint:
        m = m.transpose(-1, -2)
        lower = not lower
    return torch.linalg.solve_triangular(m, _t(rhs), upper=not lower)


matrix_triangular_solve = batch_matrix_triangular_solve


def cholesky(x):
    return torch.linalg.cholesky(_t(x))


batch_cholesky = cholesky


def random_normal(shp, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    shp = tuple(int(s) for s in shp)
    if not _RANDOM_NORMAL_QUEUE:
        if _STRICT[0]:
            raise RuntimeError("tf.random_normal%r called with an empty draw queue" % (shp,))
        return mean + stddev * torch.randn(shp, dtype=torch.float64, generator=_AUTO_GEN[0])
    draw = _t(_RANDOM_NORMAL_QUEUE.pop(0))
    if tuple(draw.shape) != shp:
        raise ValueError("queued draw has shape %r, the graph asks for %r" % (tuple(draw.shape), shp))
    return mean + stddev * draw


class _NN(object):
    @staticmethod
    def softplus(x):
        # log(1 + exp(x)) evaluated stably, no linear-branch threshold
        x = _t(x)
        return torch.clamp(x, min=0.0) + torch.log1p(torch.exp(-torch.abs(x)))


nn = _NN()


def gradients(ys, xs):
    return list(torch.autograd.grad(ys, xs, allow_unused=True))


# ---- session plumbing of the reference's own tests (testing/test_kernels.py etc.): eager here ----
class Session(object):
    def run(self, fetches, feed_dict=None):
        def one(v):
            return v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        if isinstance(fetches, (list, tuple)):
            return [one(v) for v in fetches]
        return one(fetches)

    def close(self):
        pass


def initialize_all_variables():
    return None


def set_random_seed(seed):
    torch.manual_seed(int(seed))
    _AUTO_GEN[0] = torch.Generator().manual_seed(int(seed))


class _AdamOptimizer(object):
    """tf.train.AdamOptimizer: the hyper-parameters only; GPflow.model.Model.optimize applies them."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon


class _Train(object):
    AdamOptimizer = _AdamOptimizer


train = _Train()


def ones_initializer(*a, **k):          # pragma: no cover  (name only)
    raise NotImplementedError
