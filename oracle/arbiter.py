"""Extended-precision arbiter for the ill-conditioned hyper-gradients.
TEST INFRASTRUCTURE ONLY (same rules as gpinv_oracle.py).

The lengthscale gradient of the StVGP ELBO is a cancelling sum through L^-T ... L^-1 with
cond(K) ~ 1e7..2e9, so two correct fp64 implementations (this repo's CUDA path, the torch-CPU
oracle, TensorFlow on two machines) differ by 1e-9..1e-6 relative on that one number.  To say
which of them is *right*, the forward ELBO of GPinv/stvgp.py:96-107,144-185 is restated here in
``np.longdouble`` (x87 80-bit, eps = 1.1e-19: ~3.3 more decimal digits than fp64) and the
hyper-gradients are taken by 6th-order central differences of that ELBO in the *free* space
(7-point stencil, h = 2e-3): truncation O(h^6 f^(7)) ~ 1e-13, rounding eps_ld * cond * |f| / h,
i.e. ~1e-9 relative at the worst conditioning used here.  A second-order difference with
h = 1e-6 is NOT good enough (its rounding term is 1e-6 relative); ``--selfcheck`` prints the
agreement of h = 2e-3 with h = 4e-3, which is the arbiter's own error bar.

Everything is plain numpy on longdouble arrays (no BLAS: numpy falls back to its own loops),
a column-by-column Cholesky, and the same operation ORDER as the reference where it matters
(expansion-form square_dist, jitter on the core before variance scaling).

    python -m oracle.arbiter            # regenerate tests/golden/arb_*.npz  (~5 min)
    python -m oracle.arbiter --selfcheck
"""
from __future__ import annotations

import os
import sys

import numpy as np

LD = np.longdouble
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
JITTER = LD("1e-6")
POS_EPS = LD("1e-6")
LOG2PI = np.log(LD(2) * LD(np.pi)) if False else np.log(LD(2) * np.arccos(LD(-1)))


def _ld(a):
    return np.asarray(a, dtype=LD)


def softplus(x):
    x = _ld(x)
    return np.where(x > 40, x, np.log1p(np.exp(np.minimum(x, LD(40))))) + POS_EPS


def square_dist(X, X2, ls):
    """GPflow Stationary.square_dist, expansion form, same order of additions."""
    X = X / ls
    Xs = np.sum(X * X, 1)
    if X2 is None:
        return LD(-2) * (X @ X.T) + Xs[:, None] + Xs[None, :]
    X2 = X2 / ls
    X2s = np.sum(X2 * X2, 1)
    return LD(-2) * (X @ X2.T) + Xs[:, None] + X2s[None, :]


def kcore(kind, X, ls):
    if kind == "rbf":
        return np.exp(-square_dist(X, None, ls) / 2)
    Xa = np.abs(X)
    e1 = np.exp(-square_dist(Xa, Xa, ls) / 2)
    e2 = np.exp(-square_dist(Xa, -Xa, ls) / 2)
    return e1 + e2 if kind == "rbf_csym" else e1 - e2


def cholesky(A):
    """Left-looking, one column per step, longdouble throughout."""
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        c = A[j:, j] - L[j:, :j] @ L[j, :j]
        d = np.sqrt(c[0])
        L[j, j] = d
        L[j + 1:, j] = c[1:] / d
    return L


def elbo(case, x_hyper_free, chol_cache=None):
    """ELBO of an Abel / Gaussian StVGP with ONE latent function in longdouble.
    ``case``: dict(kind, X, Y, lik ('gaussian'|'abel'), A, has_mean, q_mu [n], q_sqrt [n,n], eps [S,n]).
    ``x_hyper_free``: free-space [lengthscale, kernel variance, noise variance(, mean c)]."""
    xh = _ld(x_hyper_free)
    ls, v, s2 = softplus(xh[0]), softplus(xh[1]), softplus(xh[2])
    c = xh[3] if case["has_mean"] else LD(0)
    X = _ld(case["X"])
    n = X.shape[0]
    key = repr(xh[0])
    if chol_cache is not None and key in chol_cache:
        Lc = chol_cache[key]
    else:
        core = kcore(case["kind"], X, ls) + JITTER * np.eye(n, dtype=LD)
        Lc = cholesky(core)
        if chol_cache is not None:
            chol_cache[key] = Lc
    E = _ld(case["eps"])                       # [S,n]
    S = E.shape[0]
    Q = np.tril(_ld(case["q_sqrt"]))
    U = _ld(case["q_mu"])[None, :] + E @ Q.T   # [S,n]
    logdet = LD(S) * np.sum(np.log(np.diag(Q) ** 2))
    KL = -LD(0.5) * logdet - LD(0.5) * np.sum(E * E) + LD(0.5) * np.sum(U * U)
    F = np.sqrt(v) * (U @ Lc.T) + c
    Y = _ld(case["Y"]).reshape(-1)
    if case["lik"] == "gaussian":
        res = F - Y[None, :]
    else:
        res = np.exp(F) @ _ld(case["A"]).T - Y[None, :]
    lik = np.sum(-LD(0.5) * LOG2PI - LD(0.5) * np.log(s2) - LD(0.5) * res * res / s2)
    return (lik - KL) / LD(S)


def hyper_gradient(case, x_hyper_free, h=2e-3, verbose=False):
    """(-ELBO, -dELBO/dx_hyper) by 6th-order central differences in longdouble (the sign
    convention of ``_objective``):  f' ~ [-f(-3h) + 9 f(-2h) - 45 f(-h) + 45 f(h) - 9 f(2h) + f(3h)] / 60h."""
    xh = _ld(x_hyper_free)
    cache = {}
    f0 = elbo(case, xh, cache)
    g = np.zeros(xh.size, dtype=LD)
    w = {-3: LD(-1), -2: LD(9), -1: LD(-45), 1: LD(45), 2: LD(-9), 3: LD(1)}
    for i in range(xh.size):
        acc = LD(0)
        for k, wk in w.items():
            e = np.zeros(xh.size, dtype=LD)
            e[i] = LD(k) * LD(h)
            acc += wk * elbo(case, xh + e, cache)
        g[i] = acc / (LD(60) * LD(h))
        if verbose:
            print("  d/dx[%d] = %s" % (i, repr(-g[i])), flush=True)
    return -f0, -g


def solve_lower(L, B):
    """X = L^-1 B by forward substitution (row at a time), longdouble."""
    n = L.shape[0]
    X = np.zeros_like(B)
    for i in range(n):
        X[i] = (B[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def lengthscale_gradient_analytic(case, x_hyper_free):
    """-dELBO/dx_0 (free lengthscale) by FORWARD-mode differentiation in longdouble:
    dK = dcore/dl, dL = L Phi(L^-1 dK L^-T), dF = sqrt(v) U dL^T, dELBO = sum(dlik/dF o dF)/S.
    No differencing, so the only error is eps_ld * cond(K) ~ 1e-11."""
    xh = _ld(x_hyper_free)
    ls, v, s2 = softplus(xh[0]), softplus(xh[1]), softplus(xh[2])
    c = xh[3] if case["has_mean"] else LD(0)
    dls = LD(1) / (LD(1) + np.exp(-xh[0]))                 # d softplus / dx
    X = _ld(case["X"])
    n = X.shape[0]
    if case["kind"] == "rbf":
        D = square_dist(X, None, ls)
        core = np.exp(-D / 2)
        dcore = core * D / ls
    else:
        Xa = np.abs(X)
        D1, D2 = square_dist(Xa, Xa, ls), square_dist(Xa, -Xa, ls)
        e1, e2 = np.exp(-D1 / 2), np.exp(-D2 / 2)
        sgn = LD(1) if case["kind"] == "rbf_csym" else LD(-1)
        core = e1 + sgn * e2
        dcore = (e1 * D1 + sgn * e2 * D2) / ls
    Lc = cholesky(core + JITTER * np.eye(n, dtype=LD))
    W = solve_lower(Lc, dcore)                  # L^-1 dK
    W = solve_lower(Lc, W.T).T                  # L^-1 dK L^-T
    Phi = np.tril(W)
    Phi[np.diag_indices(n)] *= LD(0.5)
    dL = Lc @ Phi
    E = _ld(case["eps"])
    S = E.shape[0]
    Q = np.tril(_ld(case["q_sqrt"]))
    U = _ld(case["q_mu"])[None, :] + E @ Q.T
    F = np.sqrt(v) * (U @ Lc.T) + c
    dF = np.sqrt(v) * (U @ dL.T)
    Y = _ld(case["Y"]).reshape(-1)
    if case["lik"] == "gaussian":
        res, dres = F - Y[None, :], dF
    else:
        A = _ld(case["A"])
        eF = np.exp(F)
        res, dres = eF @ A.T - Y[None, :], (eF * dF) @ A.T
    dlik = np.sum(-res * dres / s2)
    return -(dlik / LD(S)) * dls


# ---- the arbitrated cases: cfg2 family (Gaussian regression) and an Abel case ----
def arb_case(name):
    """(case for the arbiter, oracle spec, full free state x, eps [1,S,n])."""
    from tests import baseline_configs as B
    from oracle import gpinv_oracle as O
    kind, n = name.rsplit("_n", 1)
    n = int(n)
    rng = np.random.RandomState(1)
    if kind == "gauss":                     # BASELINE configs[1] scaled: X on [0,6], l=0.5, S=64
        S = 64
        X = np.linspace(0, 6, n)[:, None]
        Y = np.sin(X) + 0.3 * np.random.RandomState(0).randn(n, 1)
        spec = dict(model="stvgp", X=X, Y=Y, kern=dict(type="rbf", input_dim=1, output_dim=1, ARD=False, active_dims=None),
                    mean=dict(type="zero", output_dim=1), lik=dict(type="gaussian"), num_latent=1)
        hyper = {"kern.lengthscales": [0.5], "kern.variance": [1.0], "likelihood.variance": [0.09]}
        A, has_mean, kk = None, False, "rbf"
    elif kind == "abel":                    # headline family scaled: RBF_csym, l=0.3, M = n chords
        S = 32
        spec = B.abel_spec(n, n, O.make_los_matrix)
        hyper = {"kern.lengthscales": [0.3], "kern.variance": [1.0], "likelihood.variance": [1.0], "mean_function.c": [0.0]}
        A, has_mean, kk = spec["lik"]["A"], True, "rbf_csym"
    else:
        raise KeyError(name)
    q_mu = 0.3 * rng.randn(n, 1)
    q_sqrt = (0.5 * np.eye(n) + 0.05 / np.sqrt(n) * np.tril(rng.randn(n, n)))[:, :, None]
    eps = np.random.RandomState(2).randn(1, S, n)
    vals = dict(hyper)
    vals.update(q_mu=q_mu, q_sqrt=q_sqrt)
    x = O.pack_constrained(spec, vals)
    nh = len(hyper)
    case = dict(kind=kk, X=spec["X"], Y=spec["Y"], lik=spec["lik"]["type"], A=A, has_mean=has_mean,
                q_mu=q_mu[:, 0], q_sqrt=q_sqrt[:, :, 0], eps=eps[0])
    return case, spec, x, eps, nh


def baseline_case(name):
    """Arbiter inputs of a full-size BASELINE.json configuration with ONE latent function and an
    RBF-family kernel: cfg2 (Gaussian regression), cfg4 (GPMC: u = V, no draws), cfg5 (headline)."""
    from tests import baseline_configs as B
    from oracle import gpinv_oracle as O
    c = B.case(name, O.make_los_matrix, O.make_cos_theta)
    spec, con = c["spec"], c["constrained"]
    n = spec["X"].shape[0]
    if spec["model"] == "gpmc":
        q_mu, q_sqrt, eps = con["V"][:, 0], np.eye(n), np.zeros((1, n))
    else:
        q_mu, q_sqrt, eps = con["q_mu"][:, 0], con["q_sqrt"][:, :, 0], c["eps"][0]
    has_mean = "mean_function.c" in con
    case = dict(kind=spec["kern"]["type"], X=spec["X"], Y=spec["Y"], lik=spec["lik"]["type"],
                A=spec["lik"].get("A"), has_mean=has_mean, q_mu=q_mu, q_sqrt=q_sqrt, eps=eps)
    keys = ["kern.lengthscales", "kern.variance", "likelihood.variance"] + (["mean_function.c"] if has_mean else [])
    pos = {"kern.lengthscales", "kern.variance", "likelihood.variance"}
    xh = np.concatenate([np.log(np.expm1(np.asarray(con[k], dtype=np.float64).ravel() - 1e-6)) if k in pos
                         else np.asarray(con[k], dtype=np.float64).ravel() for k in keys])
    return case, xh, c


def baseline_main(name):
    """arb_<cfg>.npz: longdouble forward-mode lengthscale gradient (and the value) of a full-size
    BASELINE configuration.  cfg5 takes ~25 minutes of one core (no BLAS for longdouble)."""
    import time
    t0 = time.time()
    case, xh, c = baseline_case(name)
    g_ls = lengthscale_gradient_analytic(case, xh)
    f = -elbo(case, xh)
    if c["spec"]["model"] == "gpmc":          # ELBO(q = delta at V) differs from the log joint by the prior's constant
        f = f + LD(case["X"].shape[0]) * LD(0.5) * LOG2PI
    np.savez(os.path.join(GOLDEN, "arb_%s.npz" % name), f=np.float64(f), g_ls=np.float64(g_ls),
             f_hi=np.array([repr(f)]), g_hi=np.array([repr(g_ls)]), x_hyper=xh)
    print("wrote arb_%s f=%s dls=%s (%.0f s)" % (name, repr(f), repr(g_ls), time.time() - t0), flush=True)


ARB_CASES = ["gauss_n256", "gauss_n512", "gauss_n1024", "abel_n256", "abel_n512", "abel_n1024"]


def main():
    import time
    if "--selfcheck" in sys.argv:
        case, spec, x, eps, nh = arb_case("gauss_n256")
        f1, g1 = hyper_gradient(case, x[:nh], 2e-3)
        f2, g2 = hyper_gradient(case, x[:nh], 4e-3)
        print("h=2e-3 vs h=4e-3: max rel diff %.2e" % float(np.max(np.abs(g1 - g2) / np.abs(g1))))
        from oracle import gpinv_oracle as O
        fo, go = O.objective(spec, x, eps)
        print("oracle vs arbiter: f %.2e  hyper %s" % (abs(fo - float(f1)) / abs(float(f1)),
                                                        np.abs(go[:nh] - g1.astype(np.float64)) / np.abs(g1.astype(np.float64))))
        return
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ARB_CASES
    for name in [a for a in names if a.startswith("cfg")]:
        baseline_main(name)
    names = [a for a in names if not a.startswith("cfg")]
    for name in names:
        t0 = time.time()
        case, spec, x, eps, nh = arb_case(name)
        f, g = hyper_gradient(case, x[:nh])
        g_fd0 = g[0]
        g[0] = lengthscale_gradient_analytic(case, x[:nh])     # forward mode: no differencing noise
        print("   lengthscale: forward-mode %s  vs 6th-order difference %s (rel %.1e)"
              % (repr(g[0]), repr(g_fd0), float(abs(g_fd0 - g[0]) / abs(g[0]))), flush=True)
        # the fp64 torch-CPU oracle on the same inputs: its distance from the arbiter is the measured
        # conditioning noise floor of a correct fp64 implementation (used as the tolerance scale)
        from oracle import gpinv_oracle as O
        fo, go = O.objective(spec, x, eps)
        g64 = g.astype(np.float64)
        oracle_err = np.abs(go[:nh] - g64) / np.abs(g64)
        print("   oracle vs arbiter: f %.1e  hyper %s" % (abs(fo - float(f)) / abs(float(f)), oracle_err), flush=True)
        np.savez(os.path.join(GOLDEN, "arb_%s.npz" % name), f=np.float64(f), g_hyper=g.astype(np.float64),
                 oracle_err=oracle_err, oracle_f_err=np.float64(abs(fo - float(f)) / abs(float(f))),
                 f_hi=np.array([repr(f)]), g_hi=np.array([repr(v) for v in g]), x_hyper=x[:nh])
        print("wrote arb_%-12s f=%s g_hyper=%s (%.0f s)" % (name, repr(f), g.astype(np.float64), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
