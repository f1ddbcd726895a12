"""Numerical settings (GPflow ``_settings.settings`` as used at GPinv/kernels.py:8,58,
GPinv/kronecker.py:56,59, GPinv/stvgp.py:31)."""
import numpy as np
import torch


class _Numerics:
    jitter_level = 1e-6


class _Dtypes:
    float_type = torch.float64
    np_float_type = np.float64


class _Verbosity:
    optimisation_verb = False
    hmc_verb = True


class _Settings:
    numerics = _Numerics()
    dtypes = _Dtypes()
    verbosity = _Verbosity()


settings = _Settings()
float_type = torch.float64
np_float_type = np.float64
