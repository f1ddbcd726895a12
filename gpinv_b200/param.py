"""Parameter tree: ``Param``, ``DataHolder``, ``Parameterized``, ``ParamList``.

The reference only subclasses GPflow's classes here (GPinv/param.py:14-24); the behaviour
those four names need on the StVGP / GPMC path is rebuilt from scratch:

* ``Param(array, transform)`` with ``.value`` (constrained numpy copy), ``.fixed``, ``.prior``;
* ``Parameterized``: attribute tree; ``get_free_state()`` / ``set_state(x)`` pack the
  unconstrained values of all non-fixed params, children visited in attribute-name order
  (GPflow ``sorted_params``; ASCII order, so GPMC's ``V`` precedes ``kern``);
* ``tensor_mode``: while evaluating the objective, attribute access returns the device tensor
  of each ``Param`` / ``DataHolder`` (GPflow's ``tf_mode``), so model code reads like the
  reference's graph-building code;
* assignment ``kern.lengthscales = 0.3`` updates the existing Param in place
  (notebooks/Spectroscopic_Abel_inversion.ipynb:394).

The local/global ``HierarchicParameterized`` machinery of GPinv/param.py:30-420 is out of
scope: StVGP / GPMC never touch it.
"""
from __future__ import annotations

from contextlib import contextmanager

import numpy as np
import torch

from . import transforms

_DEVICE = None

# Bumped whenever something a captured CUDA graph may have baked in changes from the host side:
# a DataHolder gets new data, a Param is assigned a value, a Param is fixed / released.  Model
# compares it with the value recorded at capture time and re-captures (gpinv_b200/model.py).
_STATE_EPOCH = [0]


def state_epoch() -> int:
    return _STATE_EPOCH[0]


def default_device() -> torch.device:
    """The device model tensors live on.  CUDA only: there is no CPU compute path."""
    global _DEVICE
    if _DEVICE is None:
        if not torch.cuda.is_available():
            raise RuntimeError("gpinv_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _DEVICE = torch.device("cuda", torch.cuda.current_device())
    return _DEVICE


def set_default_device(dev):
    global _DEVICE
    _DEVICE = torch.device(dev)


class Parentable(object):
    _parent = None

    @property
    def name(self):
        if self._parent is None:
            return self.__dict__.get("_name", "unnamed")
        if isinstance(self._parent, ParamList):
            return "item%i" % self._parent._list.index(self)
        for k, v in self._parent.__dict__.items():
            if v is self and k != "_parent":
                return k
        return "unnamed"

    @property
    def long_name(self):
        if self._parent is None:
            return self.name
        return self._parent.long_name + "." + self.name

    @property
    def highest_parent(self):
        return self if self._parent is None else self._parent.highest_parent


class Param(Parentable):
    def __init__(self, array, transform=None):
        self._array = np.array(np.atleast_1d(array), dtype=np.float64)
        self.transform = transform if transform is not None else transforms.Identity()
        self._fixed = False
        self._fixed_dev = None      # device copy of a fixed Param (uploaded outside graph capture)
        self.prior = None
        self._tensor = None
        self._free_leaf = None
        self._free_slice = None     # (x leaf, begin, end) when the tensor came from the fused unpack
        self._before_use = None     # optional callable run once before the tensor is first read
        self._tril_rows = None      # n if the value is one row-major [n,n] matrix read as tril() only

    @property
    def value(self):
        return self._array.copy()

    @property
    def fixed(self):
        return self._fixed

    @fixed.setter
    def fixed(self, flag):
        if bool(flag) != self._fixed:
            self._fixed = bool(flag)
            self._fixed_dev = None
            _STATE_EPOCH[0] += 1

    def fixed_tensor(self, device):
        """Device tensor of a fixed Param; uploaded once (and again after ``assign``), never
        inside a CUDA-graph capture (a pageable host-to-device copy is illegal there)."""
        if self._fixed_dev is None or self._fixed_dev.device != torch.device(device):
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("fixed Param %s has no device copy yet and a CUDA graph is being captured"
                                   % self.long_name)
            self._fixed_dev = torch.as_tensor(self._array, dtype=torch.float64).to(device)
        return self._fixed_dev

    @property
    def shape(self):
        return self._array.shape

    @property
    def size(self):
        return self._array.size

    def get_free_state(self):
        if self.fixed:
            return np.empty((0,), np.float64)
        return np.asarray(self.transform.backward(self._array), dtype=np.float64).ravel()

    def free_size(self):
        return 0 if self.fixed else self._array.size

    def set_state(self, x):
        if self.fixed:
            return 0
        n = self._array.size
        self._array[...] = np.asarray(self.transform.forward(np.asarray(x[:n], dtype=np.float64))).reshape(self._array.shape)
        return n

    def assign(self, value):
        v = np.asarray(value, dtype=np.float64)
        self._array[...] = v
        self._tensor = None
        self._fixed_dev = None
        _STATE_EPOCH[0] += 1

    def __str__(self):
        return "%s transform:%s prior:%s%s\n%s" % (self.long_name, self.transform, self.prior,
                                                   " [FIXED]" if self.fixed else "", self._array)


class DataHolder(Parentable):
    """Non-trainable array that lives on the device next to the params
    (GPinv/param.py:17-18; ``on_shape_change`` kept for signature compatibility)."""

    def __init__(self, array, on_shape_change="raise"):
        self._array = np.ascontiguousarray(array, dtype=np.float64)
        assert on_shape_change in ("raise", "pass", "recompile")
        self.on_shape_change = on_shape_change
        self._tensor = None
        self._on_change = None      # optional callable(array) run after set_data (derived layouts)

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape

    def set_data(self, array):
        array = np.ascontiguousarray(array, dtype=np.float64)
        if array.shape != self._array.shape and self.on_shape_change == "raise":
            raise ValueError("The shape of this data must not change (%s -> %s)" % (self._array.shape, array.shape))
        if (self._tensor is not None and array.shape == self._array.shape
                and not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing())):
            # same shape: refresh the existing device tensor IN PLACE so that anything holding its
            # address (a captured CUDA graph) reads the new data
            self._array = array
            self._tensor.copy_(torch.as_tensor(array, dtype=torch.float64))
        else:
            self._array = array
            self._tensor = None
            _STATE_EPOCH[0] += 1
        if self._on_change is not None:
            self._on_change(array)

    def tensor(self):
        if self._tensor is None:
            self._tensor = torch.as_tensor(self._array, dtype=torch.float64).to(default_device())
        return self._tensor

    def __str__(self):
        return "%s [data %s]" % (self.long_name, self._array.shape)


class Parameterized(Parentable):
    def __init__(self):
        self._tensor_mode = False

    # ---- attribute plumbing -------------------------------------------------------
    def __getattribute__(self, key):
        o = object.__getattribute__(self, key)
        if isinstance(o, (Param, DataHolder)) and key != "_parent":
            try:
                tm = object.__getattribute__(self, "_tensor_mode")
            except AttributeError:
                tm = False
            if tm:
                if isinstance(o, DataHolder):
                    return o.tensor()
                if o._tensor is None:
                    raise RuntimeError("Param %s has no tensor (not inside an objective evaluation)" % key)
                if o._before_use is not None:
                    cb, o._before_use = o._before_use, None
                    cb()
                return o._tensor
        return o

    def __setattr__(self, key, value):
        if key in self.__dict__ and key != "_parent":
            p = self.__dict__[key]
            if isinstance(p, Param) and not isinstance(value, Param):
                p.assign(value)
                return
            if isinstance(p, DataHolder) and not isinstance(value, DataHolder):
                p.set_data(value)
                return
            if isinstance(p, (Param, Parameterized, DataHolder)):
                p._parent = None
        if isinstance(value, (Param, Parameterized, DataHolder)) and key != "_parent":
            value._parent = self
        object.__setattr__(self, key, value)

    # ---- tree walking -------------------------------------------------------------
    @property
    def sorted_params(self):
        items = [(k, v) for k, v in self.__dict__.items()
                 if isinstance(v, (Param, Parameterized)) and k != "_parent"]
        return [v for _, v in sorted(items, key=lambda kv: kv[0])]

    @property
    def data_holders(self):
        return [v for k, v in self.__dict__.items() if isinstance(v, DataHolder) and k != "_parent"]

    def all_params(self):
        out = []
        for p in self.sorted_params:
            if isinstance(p, Param):
                out.append(p)
            else:
                out.extend(p.all_params())
        return out

    def get_free_state(self):
        parts = [p.get_free_state() for p in self.all_params()]
        return np.concatenate(parts) if parts else np.empty((0,), np.float64)

    def free_size(self):
        return int(sum(p.free_size() for p in self.all_params()))

    def set_state(self, x):
        x = np.asarray(x, dtype=np.float64).ravel()
        off = 0
        for p in self.all_params():
            off += p.set_state(x[off:])
        if off != x.size:
            raise ValueError("state vector has %d entries, the model has %d free parameters" % (x.size, off))
        return off

    def get_parameter_dict(self):
        return {p.long_name: p.value for p in self.all_params()}

    def _all_parameterized(self):
        out = [self]
        for p in self.sorted_params:
            if isinstance(p, Parameterized):
                out.extend(p._all_parameterized())
        return out

    @contextmanager
    def tensor_mode(self):
        nodes = self._all_parameterized()
        for nd in nodes:
            object.__setattr__(nd, "_tensor_mode", True)
        try:
            yield
        finally:
            for nd in nodes:
                object.__setattr__(nd, "_tensor_mode", False)

    # alias so that code written against the reference keeps working
    tf_mode = tensor_mode

    def make_tensors(self, x: torch.Tensor, override=None):
        """Bind every Param to a tensor derived from the free-state vector ``x`` (device,
        fp64).  Returns the list of leaf tensors (one per non-fixed Param, views of ``x``'s
        storage) whose gradients, concatenated, form d/dx.  ``override`` maps the index of a
        free Param (position in the returned list) to a separately allocated flat tensor that
        holds its block; ``x`` then packs only the remaining blocks (used by the pipelined host
        path, whose upload thread fills that block while other leaves are already in use)."""
        leaves = []
        off = 0
        for p in self.all_params():
            if p.fixed:
                p._tensor = p.fixed_tensor(x.device)
                p._free_leaf = None
                continue
            n = p.size
            if override and len(leaves) in override:
                src = override[len(leaves)]
                if src.numel() != n:
                    raise ValueError("override block has %d entries, the Param has %d" % (src.numel(), n))
            else:
                src = x[off:off + n]
                off += n
            leaf = src.detach().view(p.shape).requires_grad_(True)
            p._free_leaf = leaf
            p._tensor = p.transform.torch_forward(leaf)
            leaves.append(leaf)
        if off != x.numel():
            raise ValueError("state vector has %d entries, the model has %d free parameters" % (x.numel(), off))
        return leaves

    def fused_layout(self):
        """(free params, offsets, transform codes) for the one-kernel unpack / pack of the free state, or
        None when a free Param uses a transform the kernels do not implement (identity and
        ``transforms.positive`` are built in; anything else takes the per-Param torch path)."""
        params = [p for p in self.all_params() if not p.fixed]
        offs, codes = [0], []
        for p in params:
            tr = p.transform
            if isinstance(tr, transforms.Identity):
                codes.append(0)
            elif isinstance(tr, transforms.Log1pe) and tr._lower == 1e-6:
                codes.append(1)
            else:
                return None
            offs.append(offs[-1] + p.size)
        if not params or len(params) > 32:
            return None
        return params, offs, codes

    def make_tensors_fused(self, x: torch.Tensor):
        """Like ``make_tensors`` with ONE kernel for every block (``gpinv::unpack_state``; its reverse packs
        the gradient, chain rule included, in one kernel too).  Returns the single leaf (``x`` itself,
        requiring grad) or None when the layout is not supported."""
        lay = self.fused_layout()
        if lay is None:
            return None
        from . import ops
        params, offs, codes = lay
        if offs[-1] != x.numel():
            raise ValueError("state vector has %d entries, the model has %d free parameters" % (x.numel(), offs[-1]))
        for p in self.all_params():
            if p.fixed:
                p._tensor = p.fixed_tensor(x.device)
                p._free_leaf = None
        leaf = x.detach().requires_grad_(True)
        vals = ops.unpack_state(leaf, offs, codes)
        for i, (p, v) in enumerate(zip(params, vals)):
            p._tensor = v.view(p.shape)
            p._free_leaf = None
            p._free_slice = (leaf, offs[i], offs[i + 1])
        return leaf

    def clear_tensors(self):
        for p in self.all_params():
            p._tensor = None
            p._free_leaf = None
            p._free_slice = None

    def build_prior(self):
        """sum of log priors (+ log-Jacobian of the transform) over params that carry one."""
        total = None
        for p in self.all_params():
            if p.prior is None or p.fixed:
                continue
            leaf = p._free_leaf
            if leaf is None:                      # fused unpack: the free values are a slice of x
                xl, a, b = p._free_slice
                leaf = xl[a:b].view(p.shape)
            t = p.prior.logp(p._tensor) + p.transform.torch_log_jacobian(leaf)
            total = t if total is None else total + t
        return total

    def __str__(self):
        return "\n".join(str(p) for p in self.all_params())


class ParamList(Parameterized):
    """List container whose items are visited in list order (GPflow ``ParamList``; used by
    ``kernels.Stack`` GPinv/kernels.py:180 and ``mean_functions.Stack`` GPinv/mean_functions.py:65)."""

    def __init__(self, list_of_params):
        Parameterized.__init__(self)
        assert isinstance(list_of_params, list)
        for item in list_of_params:
            assert isinstance(item, (Param, Parameterized))
            item._parent = self
        self._list = list_of_params

    @property
    def sorted_params(self):
        return self._list

    def __getitem__(self, key):
        o = self._list[key]
        if isinstance(o, Param) and self._tensor_mode:
            return o._tensor
        return o

    def __len__(self):
        return len(self._list)

    def __iter__(self):
        for i in range(len(self._list)):
            yield self[i]

    def append(self, item):
        assert isinstance(item, (Param, Parameterized))
        item._parent = self
        self._list.append(item)
