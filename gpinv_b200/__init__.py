"""gpinv_b200 — B200-native (sm_100a) implementation of the GPinv StVGP / GPMC hot path.

Module names mirror the reference package (``GPinv.stvgp``, ``GPinv.kernels`` ...)."""
__version__ = "0.1.0"
