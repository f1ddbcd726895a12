"""Stand-in for the subset of GPflow 0.3 (2016, CamelCase package) that /root/reference/GPinv
subclasses or calls.  TEST INFRASTRUCTURE ONLY -- see oracle/refshim/README.md.  GPflow is an
un-vendored, unpinned dependency of the reference (setup.py:33), so these modules restate its
published behaviour; they contain no GPinv logic."""
from . import _settings, scoping, transforms, densities, tf_wraps, priors, param  # noqa: F401
from . import kullback_leiblers, kernels, likelihoods, mean_functions, model, gpmc, gpr, hmc  # noqa: F401
