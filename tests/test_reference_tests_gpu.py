"""The reference's own model tests (testing/test_stvgp.py, testing/test_gpmc.py) ported to the
drop-in API: same data, same calls, same tolerances.  Where the reference compares with
``GPflow.gpr.GPR`` (not installable here) the exact GP regression is evaluated in numpy / scipy
inside this file."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data():
    rng = np.random.RandomState(0)                  # testing/test_stvgp.py:8-11
    X = np.linspace(0., 1., 20)
    Y = np.sin(3. * X) + rng.randn(20) * 0.1
    return X, Y


def _gpr_reference(X, Y, Xnew):
    """GPflow.gpr.GPR(X, Y, RBF(1)).optimize() and predict_f, restated: maximise the exact log
    marginal likelihood over softplus(+1e-6)-transformed [lengthscale, variance, noise]."""
    from scipy.optimize import minimize
    X = X.reshape(-1, 1)
    Y = Y.reshape(-1, 1)
    n = X.shape[0]

    def sp(x):
        return np.log1p(np.exp(x)) + 1e-6

    def K(A, B, ls, v):
        return v * np.exp(-0.5 * ((A - B.T) / ls) ** 2)

    def nlml(x):
        ls, v, s2 = sp(x)
        L = np.linalg.cholesky(K(X, X, ls, v) + s2 * np.eye(n))
        a = np.linalg.solve(L, Y)
        return float(0.5 * np.sum(a * a) + np.sum(np.log(np.diag(L))) + 0.5 * n * np.log(2 * np.pi))

    x0 = np.log(np.expm1(np.ones(3) - 1e-6))
    r = minimize(nlml, x0, method="L-BFGS-B")
    ls, v, s2 = sp(r.x)
    Kmm = K(X, X, ls, v) + s2 * np.eye(n)
    Kmn = K(X, Xnew.reshape(-1, 1), ls, v)
    L = np.linalg.cholesky(Kmm)
    A = np.linalg.solve(L, Kmn)
    mu = A.T @ np.linalg.solve(L, Y)
    var = v - np.sum(A * A, 0)
    return r.fun, ls, v, s2, mu, var.reshape(-1, 1)


def _gp():
    from gpinv_b200 import gpmc, kernels, likelihoods, stvgp, train
    return gpmc, kernels, likelihoods, stvgp, train


def test_build():                                   # testing/test_stvgp.py:13-17
    gpmc, kernels, likelihoods, stvgp, train = _gp()
    X, Y = _data()
    m = stvgp.StVGP(X.reshape(-1, 1), Y.reshape(-1, 1), kernels.RBF(1, output_dim=1), likelihoods.Gaussian())
    m._compile()


def test_optimize():                                # testing/test_stvgp.py:19-57
    gpmc, kernels, likelihoods, stvgp, train = _gp()
    X, Y = _data()
    Xnew = np.linspace(0., 1., 22)
    fun_ref, ls_ref, v_ref, s2_ref, mu_ref, var_ref = _gpr_reference(X, Y, Xnew)
    assert abs(fun_ref - (-11.456799)) < 1e-3
    m = stvgp.StVGP(X.reshape(-1, 1), Y.reshape(-1, 1), kern=kernels.RBF(1, output_dim=1),
                    likelihood=likelihoods.Gaussian(), seed=1)
    trainer = train.AdamOptimizer(learning_rate=0.002)
    m.optimize()                                    # first optimize by scipy
    rslt = m.optimize(trainer, maxiter=3000)        # stochastic optimisation
    assert np.allclose(rslt['fun'], fun_ref, atol=0.5)
    assert np.allclose(m.kern.lengthscales.value, ls_ref, rtol=0.2)
    assert np.allclose(m.kern.variance.value, v_ref, rtol=0.4)
    assert np.allclose(m.likelihood.variance.value, s2_ref, rtol=0.2)
    mu, var = m.predict_f(Xnew.reshape(-1, 1))
    assert np.allclose(mu, mu_ref, atol=0.03)
    assert np.allclose(var, var_ref, atol=0.003)


def test_samples():                                 # testing/test_stvgp.py:59-70
    gpmc, kernels, likelihoods, stvgp, train = _gp()
    X, Y = _data()
    m = stvgp.StVGP(X.reshape(-1, 1), Y.reshape(-1, 1), kern=kernels.RBF(1, output_dim=1),
                    likelihood=likelihoods.Gaussian(), seed=1)
    num_samples = 10
    assert np.allclose(m.sample_F(num_samples).shape, [num_samples, X.shape[0], 1])
    assert np.allclose(m.sample_Y(num_samples).shape, [num_samples, X.shape[0], 1])


@pytest.mark.parametrize("opts", [dict(KL_analytic=True), dict(q_diag=True)])
def test_KL_analytic_and_qdiag_train(opts):         # testing/test_stvgp.py:72-94
    gpmc, kernels, likelihoods, stvgp, train = _gp()
    X, Y = _data()
    m = stvgp.StVGP(X.reshape(-1, 1), Y.reshape(-1, 1), kernels.RBF(1, output_dim=1), likelihoods.Gaussian(),
                    seed=1, **opts)
    f0, _ = m._objective(m.get_free_state())
    rslt = m.optimize(train.AdamOptimizer(learning_rate=0.002), maxiter=3000)
    assert np.isfinite(rslt['fun']) and rslt['fun'] < f0     # it trains (the reference only runs it)


def test_gpmc_build_optimize_sample_predict():      # testing/test_gpmc.py:13-41
    gpmc, kernels, likelihoods, stvgp, train = _gp()
    X, Y = _data()
    m = gpmc.GPMC(X.reshape(-1, 1), Y.reshape(-1, 1), kern=kernels.RBF(1, output_dim=1),
                  likelihood=likelihoods.Gaussian())
    m._compile()
    m.optimize()
    s = m.sample(10, verbose=True, epsilon=0.12, Lmax=15)
    assert s.shape == (10, m.get_free_state().size) and np.all(np.isfinite(s))
    Xnew = np.linspace(0., 1., 22)
    mu, var = m.predict_f(Xnew.reshape(-1, 1))
    assert mu.shape == (22, 1) and var.shape == (22, 1)
    # (beyond the reference test, which only calls these) the joint MAP of V and the
    # hyper-parameters is not the GPR posterior mean, but it must stay in its neighbourhood
    _, _, _, _, mu_ref, _ = _gpr_reference(X, Y, Xnew)
    assert np.max(np.abs(mu - mu_ref)) < 0.5
    assert m.sample_F(1).shape == (1, 20, 1) and m.sample_Y(1).shape == (1, 20, 1)


def test_device_hmc_matches_host_hmc_on_gpmc():
    """GPMC.sample runs the device-resident chain; driven by the same RNG it must visit the same
    states as the host loop that calls _objective(numpy) per leapfrog step."""
    gpmc, kernels, likelihoods, stvgp, train = _gp()
    from gpinv_b200 import hmc
    X, Y = _data()
    m = gpmc.GPMC(X.reshape(-1, 1), Y.reshape(-1, 1), kern=kernels.RBF(1, output_dim=1),
                  likelihood=likelihoods.Gaussian())
    m.optimize(maxiter=30)
    x0 = m.get_free_state()
    a = hmc.sample_HMC(m._objective, 8, Lmax=6, epsilon=0.05, x0=x0, RNG=np.random.RandomState(3))
    b = m.sample(8, Lmax=6, epsilon=0.05, RNG=np.random.RandomState(3))
    assert np.allclose(a, b, rtol=1e-9, atol=1e-9)
    m.enable_cuda_graph()                  # replayed graph inside the leapfrog loop
    c = m.sample(8, Lmax=6, epsilon=0.05, RNG=np.random.RandomState(3))
    assert np.allclose(a, c, rtol=1e-9, atol=1e-9)
