"""Line-of-sight geometry for the Abel-inversion examples: vectorised, formula-identical
counterparts of ``testing/make_LosMatrix.py`` (``make_LosMatrix`` :6-46, ``make_cosTheta``
:48-63), which the reference's notebooks import with ``sys.path.append('../testing/')``.
Host-side input generators (numpy); the reference's O(M n) Python loops take ~85 s at the
headline size 8192 x 4096."""
import numpy as np


def make_LosMatrix(r, z):
    """A[i,j] = chord length of line of sight i (height z[i]) inside radial shell j.
    Shell j in [1, n-2] spans the mid-points around r[j]; the last shell ends at r[n-1];
    column 0 is identically zero and an exact tie |z| == r_in gives 0 (both quirks of the
    reference's strict inequalities are kept)."""
    r = np.asarray(r, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    n = r.shape[0]
    r_in = np.zeros(n)
    r_out = np.zeros(n)
    r_in[1:] = 0.5 * (r[:-1] + r[1:])
    r_out[1:n - 1] = 0.5 * (r[2:] + r[1:n - 1])
    r_out[n - 1] = r[n - 1]
    az = np.abs(z)[:, None]
    z2 = (z ** 2)[:, None]
    with np.errstate(invalid="ignore"):
        so = 2. * np.sqrt(r_out[None, :] ** 2. - z2)
        si = 2. * np.sqrt(r_in[None, :] ** 2. - z2)
    both = (az < r_out[None, :]) & (az < r_in[None, :])
    outer = (az < r_out[None, :]) & (az > r_in[None, :])
    A = np.where(both, so - si, np.where(outer, so, 0.0))
    A[:, 0] = 0.0
    return np.ascontiguousarray(A)


def make_cosTheta(r, z):
    """cosTheta[i,j] = z[i]/r[j] where the chord reaches radius r[j], else 0."""
    r = np.asarray(r, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = z[:, None] / r[None, :]
    return np.ascontiguousarray(np.where(np.abs(z)[:, None] < r[None, :], c, 0.0))
