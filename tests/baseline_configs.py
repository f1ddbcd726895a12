"""The five BASELINE.json configurations as concrete, seeded inputs.

One definition serves every side: the oracle (plain-dict ``spec``), the reference runner
(``oracle/run_reference.py`` builds GPinv's own model from the same spec), the CUDA models
(``tests.helpers.model_from_spec``) and ``bench.py`` (cfg5).  numpy only; the line-of-sight
generator is passed in so that neither the oracle nor the product is imported here.

``case(name, los) -> dict(spec, S, constrained, x, eps)``: ``constrained`` maps parameter names
to constrained-space values (what a user assigns: ``m.kern.lengthscales = 0.3``), ``x`` is the
packed free state in GPflow ``sorted_params`` order, ``eps`` the [R,S,n] draws (None for GPMC).
"""
import numpy as np

POS_EPS = 1e-6


def _softplus_inv(y):
    y = np.asarray(y, dtype=np.float64) - POS_EPS
    return np.where(y > 35.0, y, np.log(np.expm1(np.minimum(y, 35.0))))


def _abel_profile(r):
    # notebooks/Abel_inversion.ipynb:115-120
    return np.exp(-(r - 0.3) ** 2 / 0.1) + np.exp(-(r + 0.3) ** 2 / 0.1)


def _qstate(n, R=1, seed=1):
    """q_mu = 0.3 randn, q_sqrt = 0.5 I + 0.05/sqrt(n) tril(randn) per latent."""
    rng = np.random.RandomState(seed)
    q_mu = 0.3 * rng.randn(n, R)
    q_sqrt = np.stack([0.5 * np.eye(n) + 0.05 / np.sqrt(n) * np.tril(rng.randn(n, n)) for _ in range(R)], -1)
    return q_mu, q_sqrt


def spec_inputs(n, M, nlam, cos_theta, los, seed=0):
    """Spectroscopic notebook generator (Spectroscopic_Abel_inversion.ipynb:112-119,224-243)."""
    r = np.linspace(0, 1., n)
    a = np.exp(-(r - 0.3) ** 2 / 0.1) + np.exp(-(r + 0.3) ** 2 / 0.1)
    v = 3. * np.exp(-(r - 0.6) ** 2 / 0.05) * (r - 0.6)
    sigma = 1. * np.exp(-r * r / 0.3) + 0.2
    z = np.linspace(-0.9, 0.9, M)
    A = los(r, z)
    cosT = cos_theta(r, z)
    lam = np.linspace(-3, 3, nlam)
    f = a[None, None, :] / (np.sqrt(2 * np.pi) * sigma[None, None, :]) * np.exp(
        -0.5 * ((lam[:, None, None] - v[None, None, :] * cosT[None]) / sigma[None, None, :]) ** 2)
    y = np.sum(A[None] * f, 2) + 0.02 * np.random.RandomState(seed).randn(nlam, M)
    return r, A, cosT, lam, y


def abel_spec(n, M, los, model="stvgp"):
    r = np.linspace(0, 1., n)
    z = np.linspace(-0.9, 0.9, M)
    A = los(r, z)
    y = A @ _abel_profile(r) + 0.1 * np.random.RandomState(0).randn(M)
    return dict(model=model, X=r[:, None], Y=y[:, None],
                kern=dict(type="rbf_csym", input_dim=1, output_dim=1, ARD=False, active_dims=None),
                mean=dict(type="constant", output_dim=1), lik=dict(type="abel", A=A), num_latent=1)


# name -> (n, M, S) of the Abel StVGP family; "cfg5" is the headline (BASELINE.json configs[4])
ABEL_SIZES = {"cfg1": (30, 40, 10), "cfg5": (4096, 8192, 1024), "cfg5_s4": (4096, 8192, 4),
              "mid": (2048, 4096, 512), "small": (256, 512, 64)}


def case(name, los=None, cos_theta=None):
    if los is None:
        from gpinv_b200.los import make_LosMatrix as los, make_cosTheta as cos_theta
    if name in ABEL_SIZES:
        n, M, S = ABEL_SIZES[name]
        spec = abel_spec(n, M, los)
        if name == "cfg1":                 # the notebook's initial state (perturbed by the tests)
            con = {"kern.lengthscales": np.ones(1), "kern.variance": np.ones(1), "likelihood.variance": np.ones(1),
                   "mean_function.c": np.zeros(1), "q_mu": np.zeros((n, 1)), "q_sqrt": np.eye(n)[:, :, None]}
        else:
            q_mu, q_sqrt = _qstate(n)
            con = {"kern.lengthscales": 0.3 * np.ones(1), "kern.variance": np.ones(1),
                   "likelihood.variance": np.ones(1), "mean_function.c": np.zeros(1), "q_mu": q_mu, "q_sqrt": q_sqrt}
        eps = np.random.RandomState(2).randn(1, S, n)
        order = ["kern.lengthscales", "kern.variance", "likelihood.variance", "mean_function.c", "q_mu", "q_sqrt"]
        pos = {"kern.lengthscales", "kern.variance", "likelihood.variance"}
    elif name == "cfg2":                   # GP regression n=1024, S=256, Gaussian
        n, S = 1024, 256
        X = np.linspace(0, 6, n)[:, None]
        Y = np.sin(X) + 0.3 * np.random.RandomState(0).randn(n, 1)
        spec = dict(model="stvgp", X=X, Y=Y, kern=dict(type="rbf", input_dim=1, output_dim=1, ARD=False, active_dims=None),
                    mean=dict(type="zero", output_dim=1), lik=dict(type="gaussian"), num_latent=1)
        q_mu, q_sqrt = _qstate(n)
        con = {"kern.lengthscales": 0.5 * np.ones(1), "kern.variance": np.ones(1), "likelihood.variance": 0.09 * np.ones(1),
               "q_mu": q_mu, "q_sqrt": q_sqrt}
        eps = np.random.RandomState(2).randn(1, S, n)
        order = ["kern.lengthscales", "kern.variance", "likelihood.variance", "q_mu", "q_sqrt"]
        pos = {"kern.lengthscales", "kern.variance", "likelihood.variance"}
    elif name in ("cfg3a", "cfg3a_s8"):    # spectroscopic, Stack kernel, R=3 (notebook-faithful, scaled)
        n, M, nlam, S = 64, 64, 128, (256 if name == "cfg3a" else 8)
        r, A, cosT, lam, y = spec_inputs(n, M, nlam, cos_theta, los)
        spec = dict(model="stvgp", X=r[:, None], Y=y,
                    kern=dict(type="stack", kern_list=[
                        dict(type="rbf_csym", input_dim=1, output_dim=2, ARD=False, active_dims=None),
                        dict(type="rbf_casym", input_dim=1, output_dim=1, ARD=False, active_dims=None)]),
                    mean=dict(type="stack", mean_list=[dict(type="constant", output_dim=2),
                                                       dict(type="zero", output_dim=1)]),
                    lik=dict(type="spec_abel", A=A, cosT=cosT, lam=lam), num_latent=3)
        q_mu, q_sqrt = _qstate(n, 3)
        con = {"kern.kern_list.0.lengthscales": 0.3 * np.ones(1), "kern.kern_list.0.variance": np.ones(2),
               "kern.kern_list.1.lengthscales": 0.3 * np.ones(1), "kern.kern_list.1.variance": np.ones(1),
               "likelihood.variance": np.ones(1), "mean_function.mean_list.0.c": -2.0 * np.ones(2),
               "q_mu": q_mu, "q_sqrt": q_sqrt}
        eps = np.random.RandomState(2).randn(3, S, n)
        order = ["kern.kern_list.0.lengthscales", "kern.kern_list.0.variance", "kern.kern_list.1.lengthscales",
                 "kern.kern_list.1.variance", "likelihood.variance", "mean_function.mean_list.0.c", "q_mu", "q_sqrt"]
        pos = {k for k in order if k.endswith(("lengthscales", "variance"))}
    elif name in ("cfg3b", "cfg3b_s2", "cfg3b_small", "cfg3b_small_full"):
        # Kronecker path: RBF2D on a (radial x wavelength) grid, R=3, Gaussian likelihood
        if name.startswith("cfg3b_small"):
            n1, n2, S = 12, 20, 16
        else:
            n1, n2, S = 64, 128, (256 if name == "cfg3b" else 2)
        d1, d2 = np.linspace(0, 1, n1), np.linspace(-3, 3, n2)
        X = np.array([[a, b] for a in d1 for b in d2])
        n, R = n1 * n2, 3
        rng = np.random.RandomState(1)
        Y = rng.randn(n, R)
        full = name.endswith("_full")
        spec = dict(model="stvgp", X=X, Y=Y,
                    kern=dict(type="rbf2d", input_dim=1, output_dim=R, ARD=False, active_dims=None,
                              dim1=d1[:, None], dim2=d2[:, None]),
                    mean=dict(type="zero", output_dim=R), lik=dict(type="gaussian"), num_latent=R, q_diag=not full)
        q_mu = 0.3 * rng.randn(n, R)
        if full:
            q_sqrt = np.stack([0.5 * np.eye(n) + 0.05 / np.sqrt(n) * np.tril(rng.randn(n, n)) for _ in range(R)], -1)
        else:
            q_sqrt = 0.5 + 0.2 * rng.rand(n, R)
        con = {"kern.lengthscales": 0.7 * np.ones(1), "kern.variance": np.array([1.0, 0.5, 2.0]),
               "likelihood.variance": 0.5 * np.ones(1), "q_mu": q_mu, "q_sqrt": q_sqrt}
        eps = np.random.RandomState(2).randn(R, S, n)
        order = ["kern.lengthscales", "kern.variance", "likelihood.variance", "q_mu", "q_sqrt"]
        pos = {"kern.lengthscales", "kern.variance", "likelihood.variance"} | (set() if full else {"q_sqrt"})
    elif name == "cfg4":                   # GPMC whitened Abel n=2048, one log-density + gradient
        n = 2048
        spec = abel_spec(n, n, los, model="gpmc")
        V = 0.5 * np.random.RandomState(1).randn(n, 1)
        con = {"V": V, "kern.lengthscales": 0.3 * np.ones(1), "kern.variance": np.ones(1),
               "likelihood.variance": np.ones(1), "mean_function.c": np.zeros(1)}
        eps, S = None, 1
        order = ["V", "kern.lengthscales", "kern.variance", "likelihood.variance", "mean_function.c"]
        pos = {"kern.lengthscales", "kern.variance", "likelihood.variance"}
    else:
        raise KeyError(name)
    x = np.concatenate([(_softplus_inv(con[k]) if k in pos else np.asarray(con[k], dtype=np.float64)).ravel()
                        for k in order])
    return dict(spec=spec, S=S, constrained=con, x=x, eps=eps, order=order)


def hyper_count(c):
    """Number of leading/trailing scalar hyper-parameters in x (everything but q_mu/q_sqrt/V)."""
    return int(sum(np.size(c["constrained"][k]) for k in c["order"] if k not in ("q_mu", "q_sqrt", "V")))


def hyper_mask(c):
    """Boolean mask over x selecting the scalar hyper-parameters."""
    parts = [np.full(np.size(c["constrained"][k]), k not in ("q_mu", "q_sqrt", "V")) for k in c["order"]]
    return np.concatenate(parts)


def compact_summary(g, mask):
    """What a compact fixture keeps of a gradient: the hyper part, the norm of the rest, 8 seeded
    projections and 2000 seeded samples of the rest."""
    g = np.asarray(g, dtype=np.float64)
    rest = g[~mask]
    rng = np.random.RandomState(77)
    idx = rng.randint(0, rest.size, size=2000)
    proj = np.array([rest.dot(np.random.RandomState(100 + k).randn(rest.size)) for k in range(8)])
    return dict(g_hyper=g[mask].copy(), g_rest_norm=np.float64(np.linalg.norm(rest)), g_proj=proj,
                g_samples=rest[idx].copy())


def check_compact(f, g, d, mask, f_tol, g_tol, hyper_tol):
    """Compare a value / gradient with a compact fixture (see ``compact_summary``)."""
    s = compact_summary(g, mask)
    assert abs(f - float(d["f"])) <= f_tol * abs(float(d["f"])), (f, float(d["f"]))
    scale = float(d["g_rest_norm"])
    assert abs(float(s["g_rest_norm"]) - scale) <= g_tol * scale, (float(s["g_rest_norm"]), scale)
    # a projection on a unit-variance Gaussian direction has magnitude ~ |g|; bound the error by it
    assert np.max(np.abs(s["g_proj"] - d["g_proj"])) <= g_tol * scale * 10
    assert np.linalg.norm(s["g_samples"] - d["g_samples"]) <= g_tol * max(np.linalg.norm(d["g_samples"]), 1e-300) * 10
    hy = np.abs(s["g_hyper"] - d["g_hyper"]) <= np.asarray(hyper_tol) * np.abs(d["g_hyper"])
    assert np.all(hy), (s["g_hyper"], d["g_hyper"])
