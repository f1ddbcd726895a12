"""Generates the golden fixtures tests/golden/*.npz with the CPU oracle.

The reference itself cannot run here (TF<=0.11 / GPflow 0.3 are not installable),
so these vectors are outputs of the oracle restatement, which is pinned separately by the
published golden value and identities in tests/test_oracle.py.  Re-run:
    python -m tests.golden.make_golden
Each file: x (free state), eps [R,S,n] (absent for GPMC), f = -objective, g = -gradient."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gpinv_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = ["abel_stvgp_n30", "gauss_stvgp_n40", "gauss_stvgp_qdiag_klanalytic", "abel_gpmc_n30", "spec_stack_n12"]


def _abel(n, M):
    from tests.helpers import abel_data
    return abel_data(n, M)


def build_spec(name):
    if name == "abel_stvgp_n30" or name == "abel_gpmc_n30":
        r, z, A, y = _abel(30, 40)
        return dict(model="stvgp" if "stvgp" in name else "gpmc", X=r[:, None], Y=y[:, None],
                    kern=dict(type="rbf_csym", input_dim=1, output_dim=1, ARD=False),
                    mean=dict(type="constant", output_dim=1), lik=dict(type="abel", A=A), num_latent=1)
    if name.startswith("gauss_stvgp"):
        X = np.linspace(0, 6, 40)[:, None]
        Y = np.sin(X) + 0.3 * np.random.RandomState(0).randn(40, 1)
        extra = dict(q_diag=True, KL_analytic=True) if "qdiag" in name else {}
        return dict(model="stvgp", X=X, Y=Y, kern=dict(type="rbf", input_dim=1, output_dim=1, ARD=False),
                    mean=dict(type="zero", output_dim=1), lik=dict(type="gaussian"), num_latent=1, **extra)
    if name == "spec_stack_n12":
        from tests.helpers import spec_data
        r, A, cosT, lam, y = spec_data(12, 9, 7)
        return dict(model="stvgp", X=r[:, None], Y=y,
                    kern=dict(type="stack", kern_list=[
                        dict(type="rbf_csym", input_dim=1, output_dim=2, ARD=False),
                        dict(type="rbf_casym", input_dim=1, output_dim=1, ARD=False)]),
                    mean=dict(type="stack", mean_list=[dict(type="constant", output_dim=2),
                                                       dict(type="zero", output_dim=1)]),
                    lik=dict(type="spec_abel", A=A, cosT=cosT, lam=lam), num_latent=3)
    if name == "abel_stvgp_cfg1":          # BASELINE.json configs[0]: N=30, 40 chords, S=10
        return build_spec("abel_stvgp_n30")
    if name == "abel_stvgp_n200":          # mid-size: multi-tile GEMMs and a 4-panel Cholesky on the GPU
        r, z, A, y = _abel(200, 150)
        return dict(model="stvgp", X=r[:, None], Y=y[:, None],
                    kern=dict(type="rbf_csym", input_dim=1, output_dim=1, ARD=False),
                    mean=dict(type="constant", output_dim=1), lik=dict(type="abel", A=A), num_latent=1)
    if name == "gauss_stvgp_full_klanalytic":
        spec = build_spec("gauss_stvgp_n40")
        spec["KL_analytic"] = True
        return spec
    if name == "ard_multi_gauss":
        rng = np.random.RandomState(0)
        n, R = 24, 3
        X, Y = rng.rand(n, 2), rng.randn(n, R)
        return dict(model="stvgp", X=X, Y=Y, kern=dict(type="rbf", input_dim=2, output_dim=R, ARD=True),
                    mean=dict(type="constant", output_dim=R), lik=dict(type="gaussian"), num_latent=R)
    if name == "rbf2d_kron_qdiag":
        rng = np.random.RandomState(0)
        d1, d2 = np.linspace(0, 1, 4), np.linspace(-1, 1, 5)
        X = np.array([[a, b] for a in d1 for b in d2])
        R = 2
        return dict(model="stvgp", X=X, Y=rng.randn(X.shape[0], R),
                    kern=dict(type="rbf2d", input_dim=1, output_dim=R, ARD=False, dim1=d1[:, None], dim2=d2[:, None]),
                    mean=dict(type="zero", output_dim=R), lik=dict(type="gaussian"), num_latent=R, q_diag=True)
    raise KeyError(name)


# ---- larger cases with COMPACT fixtures (the state / gradient have 1e5..1e6 entries: the fixture
# stores the hyper-parameter part of x, the value, and seeded projections / samples of the gradient;
# the big blocks of x and the draws are regenerated from seeds on every side) ----
COMPACT_CASES = ["cfg2_gauss_n1024", "abel_stvgp_n512"]


def compact_case(name):
    """(spec, S, hyper values, q_mu, q_sqrt, eps) of a compact case; BASELINE.json configs[1] for
    cfg2 (lengthscale 0.5, variance 1, noise 0.09, q seeds 1, eps seed 2)."""
    rng = np.random.RandomState(1)
    if name == "cfg2_gauss_n1024":
        n, S = 1024, 256
        X = np.linspace(0, 6, n)[:, None]
        Y = np.sin(X) + 0.3 * np.random.RandomState(0).randn(n, 1)
        spec = dict(model="stvgp", X=X, Y=Y, kern=dict(type="rbf", input_dim=1, output_dim=1, ARD=False),
                    mean=dict(type="zero", output_dim=1), lik=dict(type="gaussian"), num_latent=1)
        hyper = {"kern.lengthscales": 0.5, "kern.variance": np.ones(1), "likelihood.variance": 0.09}
    elif name == "abel_stvgp_n512":
        n, S = 512, 32
        r, z, A, y = _abel(n, 400)
        spec = dict(model="stvgp", X=r[:, None], Y=y[:, None],
                    kern=dict(type="rbf_csym", input_dim=1, output_dim=1, ARD=False),
                    mean=dict(type="constant", output_dim=1), lik=dict(type="abel", A=A), num_latent=1)
        hyper = {"kern.lengthscales": 0.3, "kern.variance": np.ones(1), "likelihood.variance": np.ones(1),
                 "mean_function.c": np.zeros(1)}
    else:
        raise KeyError(name)
    q_mu = 0.3 * rng.randn(n, 1)
    q_sqrt = (0.5 * np.eye(n) + 0.05 / np.sqrt(n) * np.tril(rng.randn(n, n)))[:, :, None]
    eps = np.random.RandomState(2).randn(1, S, n)
    return spec, S, hyper, q_mu, q_sqrt, eps


def compact_state(x_hyper, q_mu, q_sqrt):
    """free state = [hyper-parameters (stored in the fixture), q_mu, q_sqrt] (identity transforms)."""
    return np.concatenate([np.asarray(x_hyper, dtype=np.float64).ravel(), q_mu.ravel(), q_sqrt.ravel()])


def compact_summary(g, n_hyper):
    """What a compact fixture keeps of a gradient: hyper part, block norms, 8 seeded projections,
    2000 seeded samples."""
    g = np.asarray(g, dtype=np.float64)
    rest = g[n_hyper:]
    rng = np.random.RandomState(77)
    idx = rng.randint(0, rest.size, size=2000)
    proj = np.array([rest.dot(np.random.RandomState(100 + k).randn(rest.size)) for k in range(8)])
    return dict(g_hyper=g[:n_hyper].copy(), g_rest_norm=np.float64(np.linalg.norm(rest)), g_proj=proj,
                g_samples=rest[idx].copy())


# cases whose fixtures are produced by the reference's own code (oracle/run_reference.py)
REF_CASES = CASES + ["gauss_stvgp_full_klanalytic", "ard_multi_gauss", "rbf2d_kron_qdiag",
                     "abel_stvgp_cfg1", "abel_stvgp_n200"]
REF_SAMPLES = {"abel_stvgp_cfg1": 10, "abel_stvgp_n200": 8}      # default 6


# ---- BASELINE.json configurations at full size (tests/baseline_configs.py) ----
# oc_<name>.npz  : compact fixture computed by the ORACLE at the full configuration
# refb_<name>.npz: compact fixture computed by the REFERENCE'S OWN CODE (oracle/run_reference.py).
#                  The reference tiles [R,S,n,n] (GPinv/stvgp.py:162,174): cfg5 and cfg3b only fit in
#                  memory with a handful of samples, so they run at S=4 / S=2 (same n, M, kernels;
#                  the ELBO is a per-sample mean, so the S-sample cases exercise the same arithmetic).
ORACLE_BASELINE = ["cfg2", "cfg3a", "cfg3b", "cfg3b_small_full", "cfg4", "cfg5"]
REF_BASELINE = ["cfg3a", "cfg3a_s8", "cfg3b_s2", "cfg3b_small_full", "cfg4", "cfg5_s4"]


def oracle_in_chunks(spec, x, eps, chunk):
    """O.objective over sample chunks: ELBO(S) = sum_c (S_c/S) ELBO_c  (bounds the [S,k,m,n]
    temporaries of the spectroscopic likelihood)."""
    if eps is None or eps.shape[1] <= chunk:
        return O.objective(spec, x, eps)
    S = eps.shape[1]
    f, g = 0.0, 0.0
    for s0 in range(0, S, chunk):
        e = eps[:, s0:s0 + chunk]
        fc, gc = O.objective(spec, x, e)
        f += fc * e.shape[1] / S
        g = g + gc * (e.shape[1] / S)
    return f, g


def baseline_oracle_fixture(name):
    from tests import baseline_configs as B
    c = B.case(name, O.make_los_matrix, O.make_cos_theta)
    chunk = 16 if name.startswith("cfg3a") else 1 << 30
    f, g = oracle_in_chunks(c["spec"], c["x"], c["eps"], chunk)
    out = dict(f=np.float64(f), x_head=c["x"][:8].copy(), x_sum=np.float64(c["x"].sum()))
    out.update(B.compact_summary(g, B.hyper_mask(c)))
    return out


def main():
    import time
    if "--baseline" in sys.argv:
        for name in ORACLE_BASELINE:
            t0 = time.time()
            out = baseline_oracle_fixture(name)
            np.savez(os.path.join(HERE, "oc_%s.npz" % name), **out)
            print("wrote oc_%-18s f=%.15g |g_rest|=%.6g  (%.1f s)" % (name, out["f"], out["g_rest_norm"], time.time() - t0))
        return
    for i, name in enumerate(CASES):
        spec = build_spec(name)
        rng = np.random.RandomState(10 + i)
        n, R = spec["X"].shape[0], spec["num_latent"]
        size = O.state_size(spec)
        x = 0.2 * rng.randn(size)
        if spec["model"] == "stvgp" and not spec.get("q_diag", False):
            x[size - n * n * R:] += np.stack([np.eye(n)] * R, -1).ravel()
        out = dict(x=x)
        eps = None
        if spec["model"] == "stvgp":
            eps = rng.randn(R, 6, n)
            out["eps"] = eps
        f, g = O.objective(spec, x, eps)
        out.update(f=np.float64(f), g=g)
        np.savez(os.path.join(HERE, name + ".npz"), **out)
        print(name, size, f, np.linalg.norm(g))


if __name__ == "__main__":
    main()
