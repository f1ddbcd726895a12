"""The one helper the reference takes from GPflow ``tf_wraps`` (GPinv/kernels.py:5,58)."""
import torch


def eye(N, device=None):
    return torch.eye(int(N), dtype=torch.float64, device=device)
