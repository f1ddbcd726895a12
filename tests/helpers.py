"""Test helpers: build the oracle's plain-dict description of a gpinv_b200 model, and the
synthetic configurations of SURVEY.md 8(d)."""
import numpy as np

from gpinv_b200 import kernels, kronecker, likelihoods, mean_functions
from oracle import gpinv_oracle as O


def _kern_spec(k):
    if isinstance(k, kernels.Stack):
        return {"type": "stack", "kern_list": [_kern_spec(s) for s in k.kern_list._list]}
    base = {"input_dim": k.input_dim, "output_dim": k.output_dim, "ARD": k.ARD,
            "active_dims": None if isinstance(k.active_dims, slice) else list(k.active_dims)}
    if isinstance(k, kronecker.RBF2D):
        base.update(type="rbf2d", dim1=k.dim1.value, dim2=k.dim2.value)
    elif isinstance(k, kernels.RBF_csym):
        base["type"] = "rbf_csym"
    elif isinstance(k, kernels.RBF_casym):
        base["type"] = "rbf_casym"
    else:
        base["type"] = "rbf"
    return base


def _mean_spec(m):
    if isinstance(m, mean_functions.Stack):
        return {"type": "stack", "mean_list": [_mean_spec(s) for s in m.mean_list._list]}
    if isinstance(m, mean_functions.Constant):
        return {"type": "constant", "output_dim": m.output_dim}
    return {"type": "zero", "output_dim": m.output_dim}


def _lik_spec(l):
    if isinstance(l, likelihoods.AbelLikelihood):
        return {"type": "abel", "A": l.Amat.value}
    if isinstance(l, likelihoods.SpecAbelLikelihood):
        return {"type": "spec_abel", "A": l.Amat.value, "cosT": l.cosT.value, "lam": l.lam.value}
    return {"type": "gaussian"}


def spec_from_model(m):
    is_st = hasattr(m, "q_mu")
    return {
        "model": "stvgp" if is_st else "gpmc",
        "X": m.X.value, "Y": m.Y.value,
        "kern": _kern_spec(m.kern), "mean": _mean_spec(m.mean_function), "lik": _lik_spec(m.likelihood),
        "num_latent": m.num_latent,
        "q_diag": getattr(m, "q_diag", False), "KL_analytic": getattr(m, "KL_analytic", False),
    }


def model_from_spec(spec, S=6):
    """The gpinv_b200 model for an oracle-style plain-dict description (inverse of
    ``spec_from_model``; used with the reference-executed fixtures tests/golden/ref_*.npz)."""
    from gpinv_b200 import gpmc, stvgp

    def kern(k):
        if k["type"] == "stack":
            return kernels.Stack([kern(s) for s in k["kern_list"]])
        if k["type"] == "rbf2d":
            return kronecker.RBF2D(k["input_dim"], k["output_dim"], k["dim1"], k["dim2"], ARD=k["ARD"])
        cls = {"rbf": kernels.RBF, "rbf_csym": kernels.RBF_csym, "rbf_casym": kernels.RBF_casym}[k["type"]]
        return cls(k["input_dim"], k["output_dim"], ARD=k["ARD"])

    def mean(m):
        if m["type"] == "stack":
            return mean_functions.Stack([mean(s) for s in m["mean_list"]])
        if m["type"] == "constant":
            return mean_functions.Constant(m["output_dim"])
        return mean_functions.Zero(m["output_dim"])

    lk = spec["lik"]
    if lk["type"] == "abel":
        lik = likelihoods.AbelLikelihood(lk["A"])
    elif lk["type"] == "spec_abel":
        lik = likelihoods.SpecAbelLikelihood(lk["A"], lk["cosT"], lk["lam"])
    else:
        lik = likelihoods.Gaussian()
    if spec["model"] == "gpmc":
        return gpmc.GPMC(spec["X"], spec["Y"], kern(spec["kern"]), lik, mean(spec["mean"]),
                         num_latent=spec["num_latent"])
    return stvgp.StVGP(spec["X"], spec["Y"], kern(spec["kern"]), lik, mean(spec["mean"]),
                       num_latent=spec["num_latent"], q_diag=spec.get("q_diag", False),
                       KL_analytic=spec.get("KL_analytic", False), num_samples=S)


def abel_data(n, M, seed=0):
    """SURVEY.md 8(d) cfg1/cfg5 generator (notebooks/Abel_inversion.ipynb:115-120,183-185)."""
    r = np.linspace(0, 1., n)
    z = np.linspace(-0.9, 0.9, M)
    A = O.make_los_matrix(r, z)
    g = np.exp(-(r - 0.3) ** 2 / 0.1) + np.exp(-(r + 0.3) ** 2 / 0.1)
    y = A @ g + 0.1 * np.random.RandomState(seed).randn(M)
    return r, z, A, y


def spec_data(n, M, nlam, seed=0):
    """Spectroscopic notebook generator (Spectroscopic_Abel_inversion.ipynb:112-119,224-243)."""
    r = np.linspace(0, 1., n)
    a = np.exp(-(r - 0.3) ** 2 / 0.1) + np.exp(-(r + 0.3) ** 2 / 0.1)
    v = 3. * np.exp(-(r - 0.6) ** 2 / 0.05) * (r - 0.6)
    sigma = 1. * np.exp(-r * r / 0.3) + 0.2
    z = np.linspace(-0.9, 0.9, M)
    A = O.make_los_matrix(r, z)
    cosT = O.make_cos_theta(r, z)
    lam = np.linspace(-3, 3, nlam)
    f = a[None, None, :] / (np.sqrt(2 * np.pi) * sigma[None, None, :]) * np.exp(
        -0.5 * ((lam[:, None, None] - v[None, None, :] * cosT[None]) / sigma[None, None, :]) ** 2)
    y = np.sum(A[None] * f, 2) + 0.02 * np.random.RandomState(seed).randn(nlam, M)
    return r, A, cosT, lam, y


def perturbed_state(m, scale=0.1, seed=1):
    x0 = m.get_free_state()
    return x0 + scale * np.random.RandomState(seed).randn(x0.size)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def same_evaluation(f1, g1, f0, g0, n_hyper, f_tol=1e-10, q_tol=1e-10, hyper_tol=1e-7):
    """Two runs of the SAME evaluation through different host / launch paths (eager, graph replay,
    pipelined transfers, sharded ranks).  They differ by the order of fp64 atomics in the fused
    reductions only; the `n_hyper` leading scalar hyper-parameter gradients amplify that by the
    conditioning of K, so they get their own, per-entry, bound; everything else
    (q_mu, q_sqrt, V blocks) must agree to `q_tol` in relative norm."""
    g1, g0 = np.asarray(g1), np.asarray(g0)
    if abs(f1 - f0) > f_tol * abs(f0):
        return False
    if g0.size > n_hyper and rel_err(g1[n_hyper:], g0[n_hyper:]) > q_tol:
        return False
    h1, h0 = g1[:n_hyper], g0[:n_hyper]
    return bool(np.all(np.abs(h1 - h0) <= hyper_tol * np.maximum(np.abs(h0), 1e-300)))


def check_compact(f, g, d, n_hyper, f_tol, g_tol, hyper_tol):
    """Compare a value / gradient with a compact reference-executed fixture (refc_*.npz): value,
    norm / 8 seeded projections / 2000 seeded samples of the variational blocks, and each
    hyper-parameter gradient (conditioning-aware bound)."""
    from tests.golden.make_golden import compact_summary
    s = compact_summary(g, n_hyper)
    assert abs(f - float(d["f"])) <= f_tol * abs(float(d["f"])), (f, float(d["f"]))
    scale = float(d["g_rest_norm"])
    assert abs(float(s["g_rest_norm"]) - scale) <= g_tol * scale
    # a projection on a unit-variance Gaussian direction has magnitude ~ |g|; bound the error by it
    assert np.max(np.abs(s["g_proj"] - d["g_proj"])) <= g_tol * scale * 10
    assert np.linalg.norm(s["g_samples"] - d["g_samples"]) <= g_tol * max(np.linalg.norm(d["g_samples"]), 1e-300) * 10
    assert np.all(np.abs(s["g_hyper"] - d["g_hyper"]) <= hyper_tol * np.abs(d["g_hyper"])), (s["g_hyper"], d["g_hyper"])


def lbar_pack_reference(Lbar, G):
    """numpy restatement of the reduce-scatter layout of the column-block sharded Cholesky reverse
    (include/gpinv_b200.h, gpinv_lbar_pack_f64): G chunks of (n + w) w entries, w = n / 2G; chunk r holds
    tril(Lbar)[r w:, block r] followed by tril(Lbar)[(2G-1-r) w:, block 2G-1-r], rows x w, row-major."""
    n = Lbar.shape[0]
    w = n // (2 * G)
    T = np.tril(Lbar)
    out = []
    for r in range(G):
        for b in (r, 2 * G - 1 - r):
            out.append(T[b * w:, b * w:(b + 1) * w].reshape(-1))
    return np.concatenate(out)


def shard_reverse_reference(L, chunk, c0s, w, dK):
    """numpy restatement of one rank's share of <G_K, dK>, G_K = W^T Phi(L^T Lbar) W, from its reduced
    chunk of column blocks (first columns c0s, width w): sum over its blocks J of <(W^T Phi[:,J]) W[J,:], dK>."""
    n = L.shape[0]
    W = np.linalg.inv(L)
    total, o = 0.0, 0
    for c0 in c0s:
        m = n - c0
        D = chunk[o:o + m * w].reshape(m, w)
        o += m * w
        P = L[c0:, c0:].T @ D                       # rows c0.. of L^T Lbar[:, J]
        sq = np.tril(P[:w])
        sq[np.diag_indices(w)] *= 0.5
        P[:w] = sq                                  # Phi: lower part, diagonal halved
        A = W[c0:, :].T @ P                         # n x w
        Gk = A @ W[c0:c0 + w, :]                    # n x n
        total += float(np.sum(Gk * dK))
    return total
