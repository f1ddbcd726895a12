"""Runs the reference's OWN code (/root/reference/GPinv, unmodified, imported where it lies) over
the TensorFlow/GPflow stand-ins of oracle/refshim and writes the golden vectors
``tests/golden/ref_*.npz``.  TEST INFRASTRUCTURE ONLY; needs /root/reference, so it runs in the
build container, never on the GPU box (the committed fixtures travel instead).

    python -m oracle.run_reference            # regenerate every fixture
    python -m oracle.run_reference --check    # recompute and compare with the committed files

What executes reference lines: ``StVGP.build_likelihood/_sample/_analytical_KL``
(GPinv/stvgp.py:96-107,144-195), ``Kern.K/Kdiag/Cholesky`` and the RBF / csym / casym / Stack
cores (GPinv/kernels.py:42-189), ``RBF2D.Cholesky`` + ``kronecker_product``
(GPinv/kronecker.py:10-66), ``Gaussian.logp`` (GPinv/likelihoods.py:70-76), the mean functions
(GPinv/mean_functions.py:31-72), ``GPMC.build_likelihood/_get_f`` (GPinv/gpmc.py:39-82),
``conditional`` (GPinv/conditionals.py:38-78), ``make_LosMatrix/make_cosTheta``
(testing/make_LosMatrix.py) and the ``AbelLikelihood`` / ``SpecAbelLikelihood`` classes, which are
``exec``-ed straight from the code cells of notebooks/Abel_inversion.ipynb and
notebooks/Spectroscopic_Abel_inversion.ipynb.  What is a restatement: the GPflow base classes
and formulas in oracle/refshim/GPflow (GPflow is not vendored in the reference).

Each fixture: x (free state, GPflow ``sorted_params`` order), eps [R,S,n] (fed to the graph's
``tf.random_normal([R,N,n,1])``, GPinv/stvgp.py:164), f = -objective, g = -gradient;
``*_predict``: Xnew, mu, var, var_full from ``build_predict`` at ``set_state(x)``."""
import json
import os
import sys

sys.dont_write_bytecode = True
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SHIM = os.path.join(ROOT, "oracle", "refshim")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    if not os.path.isdir(os.path.join(REF, "GPinv")):
        raise RuntimeError("/root/reference is not present: fixtures can only be generated in the build container")
    for p in (os.path.join(REF, "testing"), REF, SHIM, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tensorflow as tf            # oracle/refshim/tensorflow
    import GPflow                      # oracle/refshim/GPflow
    import GPinv                       # /root/reference/GPinv  (the real thing)
    import make_LosMatrix              # /root/reference/testing/make_LosMatrix.py
    assert GPinv.__file__.startswith(REF) and tf.__version__.endswith("refshim")
    return tf, GPflow, GPinv, make_LosMatrix


def _notebook_class(nb, class_name, env):
    """exec the notebook code cell that defines ``class_name`` inside ``env``."""
    cells = json.load(open(os.path.join(REF, "notebooks", nb)))["cells"]
    for c in cells:
        src = "".join(c["source"])
        if c["cell_type"] == "code" and ("class %s(" % class_name) in src:
            exec(compile(src, "%s:%s" % (nb, class_name), "exec"), env)
            return env[class_name]
    raise KeyError(class_name)


def build_reference_model(spec, S):
    tf, GPflow, GPinv, make_LosMatrix = _import_reference()

    def kern(k):
        if k["type"] == "stack":
            return GPinv.kernels.Stack([kern(s) for s in k["kern_list"]])
        if k["type"] == "rbf2d":
            return GPinv.kronecker.RBF2D(k["input_dim"], k["output_dim"], k["dim1"], k["dim2"], ARD=k["ARD"])
        cls = {"rbf": GPinv.kernels.RBF, "rbf_csym": GPinv.kernels.RBF_csym,
               "rbf_casym": GPinv.kernels.RBF_casym}[k["type"]]
        return cls(k["input_dim"], k["output_dim"], ARD=k["ARD"])

    def mean(m):
        if m["type"] == "stack":
            return GPinv.mean_functions.Stack([mean(s) for s in m["mean_list"]])
        if m["type"] == "constant":
            return GPinv.mean_functions.Constant(m["output_dim"])
        return GPinv.mean_functions.Zero(m["output_dim"])

    lk = spec["lik"]
    env = {"GPinv": GPinv, "tf": tf, "np": np}
    if lk["type"] == "abel":
        lik = _notebook_class("Abel_inversion.ipynb", "AbelLikelihood", env)(lk["A"])
    elif lk["type"] == "spec_abel":
        # the notebook's sample_F reads this *global* numpy array (cell line 343); TF converts an
        # ndarray operand to a constant tensor implicitly, torch does not, so hand it over converted
        env["cosTheta"] = tf.constant(lk["cosT"])
        lik = _notebook_class("Spectroscopic_Abel_inversion.ipynb", "SpecAbelLikelihood", env)(
            lk["A"], lk["cosT"], lk["lam"])
    else:
        lik = GPinv.likelihoods.Gaussian()
    if spec["model"] == "gpmc":
        return GPinv.gpmc.GPMC(spec["X"], spec["Y"], kern(spec["kern"]), lik, mean(spec["mean"]),
                               num_latent=spec["num_latent"])
    return GPinv.stvgp.StVGP(spec["X"], spec["Y"], kern(spec["kern"]), lik, mean(spec["mean"]),
                             num_latent=spec["num_latent"], q_diag=spec.get("q_diag", False),
                             KL_analytic=spec.get("KL_analytic", False), num_samples=S)


def reference_los(r, z):
    """A, cosTheta from the reference's own generator (testing/make_LosMatrix.py)."""
    _, _, _, mk = _import_reference()
    return mk.make_LosMatrix(r, z), mk.make_cosTheta(r, z)


def case_inputs(name, index):
    """Seeded state / draws of a case (independent of the model implementation)."""
    from tests.golden.make_golden import build_spec, REF_SAMPLES
    spec = build_spec(name)
    rng = np.random.RandomState(100 + index)
    n, R = spec["X"].shape[0], spec["num_latent"]
    S = REF_SAMPLES.get(name, 6)
    return spec, rng, n, R, S


def compute(name, index):
    spec, rng, n, R, S = case_inputs(name, index)
    m = build_reference_model(spec, S)
    x0 = m.get_free_state()
    x = x0 + (0.05 if n >= 100 else 0.2) * rng.randn(x0.size)
    out = dict(x=x)
    draws = []
    if spec["model"] == "stvgp":
        eps = rng.randn(R, S, n)
        out["eps"] = eps
        draws = [eps[:, :, :, None]]                 # v_samples [R,N,n,1]
    f, g = m._objective(x, draws)
    out.update(f=np.float64(f), g=g)
    return spec, m, out


def compute_predict(name, index):
    spec, m, out = compute(name, index)
    m.set_state(out["x"])
    Xnew = np.linspace(0., 1.2, 17)[:, None]
    mu, var = m.predict_f(Xnew)
    _, var_full = m.predict_f(Xnew, full_cov=True)
    return dict(x=out["x"], Xnew=Xnew, mu=mu, var=var, var_full=var_full)


def compute_compact(name):
    """Larger cases: run the reference, keep a compact summary (tests/golden/make_golden.py)."""
    from tests.golden.make_golden import compact_case, compact_state, compact_summary
    spec, S, hyper, q_mu, q_sqrt, eps = compact_case(name)
    m = build_reference_model(spec, S)
    m.kern.lengthscales = hyper["kern.lengthscales"]
    m.kern.variance = hyper["kern.variance"]
    m.likelihood.variance = hyper["likelihood.variance"]
    if "mean_function.c" in hyper:
        m.mean_function.c = hyper["mean_function.c"]
    m.q_mu = q_mu
    m.q_sqrt = q_sqrt
    x = m.get_free_state()
    n_hyper = x.size - q_mu.size - q_sqrt.size
    assert np.array_equal(x, compact_state(x[:n_hyper], q_mu, q_sqrt))
    f, g = m._objective(x, [eps[:, :, :, None]])
    out = dict(x_hyper=x[:n_hyper].copy(), f=np.float64(f))
    out.update(compact_summary(g, n_hyper))
    return out


def _assign(m, name, value):
    """m.<dotted name> = value (list containers indexed by position)."""
    obj = m
    parts = name.split(".")
    for p in parts[:-1]:
        obj = obj[int(p)] if p.isdigit() else getattr(obj, p)
    setattr(obj, parts[-1], value)


def compute_baseline(name):
    """A BASELINE.json configuration (tests/baseline_configs.py) through the reference's own code:
    the model is built from the spec, the constrained values are assigned the way a user does
    (``m.kern.lengthscales = 0.3``), and its free state must come out as the case's packed ``x``."""
    from tests import baseline_configs as B
    from oracle import gpinv_oracle as O
    c = B.case(name, O.make_los_matrix, O.make_cos_theta)
    m = build_reference_model(c["spec"], c["S"])
    for k in c["order"]:
        _assign(m, k, c["constrained"][k])
    x = m.get_free_state()
    assert x.size == c["x"].size and np.allclose(x, c["x"], rtol=1e-13, atol=1e-13), name
    draws = [] if c["eps"] is None else [c["eps"][:, :, :, None]]
    f, g = m._objective(c["x"], draws)
    out = dict(f=np.float64(f), x_head=c["x"][:8].copy(), x_sum=np.float64(c["x"].sum()))
    out.update(B.compact_summary(g, B.hyper_mask(c)))
    return out


def main():
    from tests.golden.make_golden import REF_CASES, COMPACT_CASES, REF_BASELINE
    if "--baseline" in sys.argv:
        # minutes of CPU and tens of GB of tiles: generated on demand, not part of --check
        import time
        names = [a for a in sys.argv[1:] if not a.startswith("--")] or REF_BASELINE
        for name in names:
            t0 = time.time()
            out = compute_baseline(name)
            np.savez(os.path.join(GOLDEN, "refb_%s.npz" % name), **out)
            print("wrote refb_%-18s f=%.15g |g_rest|=%.6g  (%.1f s)" % (name, out["f"], out["g_rest_norm"], time.time() - t0),
                  flush=True)
        return
    check = "--check" in sys.argv
    worst = 0.0
    for i, name in enumerate(REF_CASES):
        spec, m, out = compute(name, i)
        path = os.path.join(GOLDEN, "ref_%s.npz" % name)
        if check:
            d = np.load(path)
            e = max(abs(float(d["f"]) - float(out["f"])) / abs(float(out["f"])),
                    np.linalg.norm(d["g"] - out["g"]) / np.linalg.norm(out["g"]))
            worst = max(worst, e)
            print("check %-32s f=%.15g  max rel diff vs committed fixture %.1e" % (name, out["f"], e))
        else:
            np.savez(path, **out)
            print("wrote %-32s |x|=%d f=%.15g |g|=%.6g" % (name, out["x"].size, out["f"], np.linalg.norm(out["g"])))
    for i, name in enumerate(["abel_stvgp_n30", "gauss_stvgp_n40", "abel_gpmc_n30"]):
        out = compute_predict(name, 50 + i)
        path = os.path.join(GOLDEN, "ref_%s_predict.npz" % name)
        if check:
            d = np.load(path)
            e = max(np.abs(d[k] - out[k]).max() for k in ("mu", "var", "var_full"))
            worst = max(worst, e)
            print("check %-32s max abs diff %.1e" % (name + "_predict", e))
        else:
            np.savez(path, **out)
            print("wrote %s" % os.path.basename(path))
    for name in COMPACT_CASES:
        out = compute_compact(name)
        path = os.path.join(GOLDEN, "refc_%s.npz" % name)
        if check:
            d = np.load(path)
            e = max(abs(float(d["f"]) - float(out["f"])) / abs(float(out["f"])),
                    np.max(np.abs(d["g_proj"] - out["g_proj"]) / np.abs(out["g_proj"])))
            worst = max(worst, e)
            print("check %-32s f=%.15g  max rel diff vs committed fixture %.1e" % (name, out["f"], e))
        else:
            np.savez(path, **out)
            print("wrote %-32s f=%.15g |g_rest|=%.6g" % (name, out["f"], out["g_rest_norm"]))
    # the reference's own LOS generator on the cfg1 grid (pins the vectorised product generator)
    r, z = np.linspace(0, 1., 30), np.linspace(-0.9, 0.9, 40)
    A, cosT = reference_los(r, z)
    path = os.path.join(GOLDEN, "ref_los_30x40.npz")
    if check:
        d = np.load(path)
        worst = max(worst, np.abs(d["A"] - A).max(), np.abs(d["cosT"] - cosT).max())
    else:
        np.savez(path, r=r, z=z, A=A, cosT=cosT)
    if check:
        print("worst difference: %.2e" % worst)
        sys.exit(0 if worst < 1e-12 else 1)


if __name__ == "__main__":
    main()
