"""GPflow.param: Parentable / Param / DataHolder / Parameterized / ParamList / AutoFlow.

Behaviour kept: ``value``; ``fixed``; ``prior``; free-state packing in ``sorted_params`` order
(alphabetical by attribute name); ``make_tf_array(x)`` slices the free vector and applies the
transform; in ``tf_mode`` attribute access yields tensors; assigning an array to an existing
Param / DataHolder attribute updates it in place."""
from contextlib import contextmanager

import numpy as np
import tensorflow as tf

from . import transforms


class Parentable(object):
    def __init__(self):
        self._parent = None

    @property
    def highest_parent(self):
        return self if self._parent is None else self._parent.highest_parent

    @property
    def name(self):
        if self._parent is None:
            return "unnamed"
        if isinstance(self._parent, ParamList):
            return "item%i" % self._parent._list.index(self)
        for key, val in self._parent.__dict__.items():
            if val is self and key != "_parent":
                return key
        return "unnamed"

    @property
    def long_name(self):
        if self._parent is None:
            return self.name
        return self._parent.long_name + "." + self.name


class Param(Parentable):
    def __init__(self, array, transform=transforms.Identity()):
        Parentable.__init__(self)
        self._array = np.asarray(np.atleast_1d(array), dtype=np.float64).copy()
        self._tf_array = None
        self._log_jacobian = None
        self.prior = None
        self.fixed = False
        self.transform = transform

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape

    @property
    def size(self):
        return self._array.size

    def get_free_state(self):
        if self.fixed:
            return np.empty((0,), np.float64)
        return np.asarray(self.transform.backward(self.value.flatten()), dtype=np.float64)

    def set_state(self, x):
        if self.fixed:
            return 0
        free_size = self.size
        self._array[...] = np.asarray(self.transform.forward(x[:free_size])).reshape(self.shape)
        return free_size

    def make_tf_array(self, free_array):
        if self.fixed:
            self._tf_array = tf.constant(self._array)
            self._log_jacobian = 0.0
            return 0
        x_free = free_array[:self.size]
        mapped = self.transform.tf_forward(x_free)
        self._tf_array = tf.reshape(mapped, self.shape)
        self._log_jacobian = self.transform.tf_log_jacobian(x_free)
        return self.size

    def build_prior(self):
        if self.prior is None:
            return tf.constant(0.0)
        return self.prior.logp(self._tf_array) + tf.reduce_sum(self._log_jacobian)


class DataHolder(Parentable):
    def __init__(self, array, on_shape_change="raise"):
        Parentable.__init__(self)
        self._array = np.asarray(array, dtype=np.float64)
        assert on_shape_change in ["raise", "pass", "recompile"]
        self.on_shape_change = on_shape_change

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape

    @property
    def _tf_array(self):
        return tf.constant(self._array)

    def set_data(self, array):
        self._array = np.asarray(array, dtype=np.float64)


class Parameterized(Parentable):
    def __init__(self):
        Parentable.__init__(self)
        self.scoped_keys = []
        self._tf_mode = False

    def __getattribute__(self, key):
        o = object.__getattribute__(self, key)
        try:
            if not object.__getattribute__(self, "_tf_mode"):
                return o
        except AttributeError:
            return o
        if isinstance(o, (Param, DataHolder)) and key != "_parent":
            return o._tf_array
        return o

    def __setattr__(self, key, value):
        if key in self.__dict__ and key != "_parent":
            p = self.__dict__[key]
            if isinstance(p, Param) and not isinstance(value, (Param, Parameterized)):
                p._array[...] = np.asarray(value, dtype=np.float64)
                return
            if isinstance(p, DataHolder) and not isinstance(value, DataHolder):
                p.set_data(value)
                return
            if isinstance(p, (Param, Parameterized, DataHolder)):
                p._parent = None
        if isinstance(value, (Param, Parameterized, DataHolder)) and key != "_parent":
            value._parent = self
        object.__setattr__(self, key, value)

    @property
    def sorted_params(self):
        params = [(k, c) for k, c in self.__dict__.items()
                  if isinstance(c, (Param, Parameterized)) and k != "_parent"]
        return [c for _, c in sorted(params, key=lambda kc: kc[0])]

    @property
    def data_holders(self):
        return [c for k, c in self.__dict__.items() if isinstance(c, DataHolder) and k != "_parent"]

    @property
    def fixed(self):
        return all(p.fixed for p in self.sorted_params)

    def get_free_state(self):
        parts = [p.get_free_state() for p in self.sorted_params]
        return np.hstack(parts) if parts else np.empty((0,), np.float64)

    def set_state(self, x):
        count = 0
        for p in self.sorted_params:
            count += p.set_state(x[count:])
        return count

    def make_tf_array(self, X):
        if isinstance(X, np.ndarray):      # the reference's tests pass the numpy free state directly
            import torch
            X = torch.as_tensor(X, dtype=torch.float64)
        count = 0
        for p in self.sorted_params:
            count += p.make_tf_array(X[count:])
        return count

    def build_prior(self):
        total = tf.constant(0.0)
        for p in self.sorted_params:
            total = total + p.build_prior()
        return total

    def get_feed_dict(self):
        """GPflow fed DataHolder placeholders through this dict; data are constants here."""
        return {}

    @contextmanager
    def tf_mode(self):
        self._begin_tf_mode()
        try:
            yield
        finally:
            self._end_tf_mode()

    def _begin_tf_mode(self):
        for p in self.sorted_params:
            if isinstance(p, Parameterized):
                p._begin_tf_mode()
        self._tf_mode = True

    def _end_tf_mode(self):
        self._tf_mode = False
        for p in self.sorted_params:
            if isinstance(p, Parameterized):
                p._end_tf_mode()


class ParamList(Parameterized):
    def __init__(self, list_of_params=()):
        Parameterized.__init__(self)
        for item in list_of_params:
            assert isinstance(item, (Param, Parameterized))
            item._parent = self
        self._list = list(list_of_params)

    @property
    def sorted_params(self):
        return self._list

    def __getitem__(self, key):
        o = self.sorted_params[key]
        if isinstance(o, Param) and self._tf_mode:
            return o._tf_array
        return o

    def __iter__(self):
        return iter(self[i] for i in range(len(self._list)))

    def __len__(self):
        return len(self._list)

    def append(self, item):
        item._parent = self
        self._list.append(item)


class AutoFlow(object):
    """Decorator: run a graph-building method at the model's current state and return numpy."""

    def __init__(self, *tf_arg_tuples):
        self.tf_arg_tuples = tf_arg_tuples

    def __call__(self, tf_method):
        def runner(instance, *np_args):
            import torch
            x = torch.as_tensor(instance.get_free_state(), dtype=torch.float64)
            instance.make_tf_array(x)
            args = []
            for a, (dtype, _shape) in zip(np_args, self.tf_arg_tuples):
                args.append(np.atleast_1d(np.asarray(a)) if dtype == tf.int32 else tf.constant(a))
            with instance.tf_mode(), torch.no_grad():
                out = tf_method(instance, *args)
            if isinstance(out, (tuple, list)):
                return [o.numpy() for o in out]
            return out.numpy()
        runner.__name__ = tf_method.__name__
        return runner
