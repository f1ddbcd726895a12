"""Log densities on torch tensors (GPflow ``densities`` as re-exported at GPinv/__init__.py:1;
``gaussian`` is used at GPinv/likelihoods.py:76 and in the notebooks' likelihoods)."""
import math

import torch

LOG2PI = math.log(2.0 * math.pi)


def gaussian(x, mu, var):
    return -0.5 * LOG2PI - 0.5 * torch.log(var) - 0.5 * torch.square(mu - x) / var


def lognormal(x, mu, var):
    lnx = torch.log(x)
    return gaussian(lnx, mu, var) - lnx


def student_t(x, mean, scale, deg_free):
    deg_free = torch.as_tensor(deg_free, dtype=x.dtype, device=x.device)
    const = torch.lgamma((deg_free + 1.0) * 0.5) - torch.lgamma(deg_free * 0.5) \
        - 0.5 * (torch.log(torch.square(scale)) + torch.log(deg_free) + math.log(math.pi))
    return const - 0.5 * (deg_free + 1.0) * torch.log(1.0 + (1.0 / deg_free) * torch.square((x - mean) / scale))


def exponential(lamb, y):
    return -y / lamb - torch.log(lamb)


def laplace(mu, sigma, y):
    return -torch.abs(mu - y) / sigma - torch.log(2.0 * sigma)
