"""CPU tests of the host-side logic: parameter tree / packing, transforms, C-ABI exports,
loud failure without CUDA, HMC loop, optimizer plumbing, sample sharding over gloo."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import gpinv_b200
from gpinv_b200 import _capi, gpmc, hmc, kernels, kronecker, likelihoods, mean_functions, stvgp, transforms
from gpinv_b200.param import DataHolder, Param, ParamList, Parameterized
from oracle import gpinv_oracle as O
from tests.helpers import abel_data, spec_from_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(_capi.LIB_PATH):
        _capi.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _capi.header_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_capi.SIGNATURES)
    assert _capi.lib().gpinv_version() >= 100


def test_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gpinv_b200 import ops
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.potrf(torch.eye(4, dtype=torch.float64))
    X = np.linspace(0, 1, 5)[:, None]
    m = stvgp.StVGP(X, np.sin(X), kernels.RBF(1, 1), likelihoods.Gaussian())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m._objective(m.get_free_state())


def test_product_does_not_import_the_oracle():
    code = "import sys; import gpinv_b200, gpinv_b200.stvgp, gpinv_b200.gpmc, gpinv_b200.kronecker; " \
           "assert not any(k.startswith('oracle') for k in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for fn in os.listdir(os.path.join(ROOT, "gpinv_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "gpinv_b200", fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_native_trapezoid_gather_packs_the_lower_row_blocks():
    """Host half of the q_sqrt upload (gpinv_upload_tril_blocks_f64 without a device target): block k
    of rb rows lands at off[k] as a packed [rows, r1] matrix holding columns [0, r1); skipped blocks
    are left untouched; ragged last block."""
    lib = _capi.lib()
    nq, rb = 300, 64
    rng = np.random.RandomState(0)
    x = rng.randn(nq * nq)
    blocks, o = [], 0
    for r0 in range(0, nq, rb):
        r1 = min(nq, r0 + rb)
        blocks.append((r0, r1, o))
        o += (r1 - r0) * r1
    off = np.array([b[2] for b in blocks], dtype=np.int64)
    take = np.array([1, 0, 1, 1, 1], dtype=np.int32)
    dst = np.full(o, -7.0)
    rc = lib.gpinv_upload_tril_blocks_f64(x.ctypes.data, nq, rb, dst.ctypes.data, None, off.ctypes.data,
                                          take.ctypes.data, len(blocks), 3, None)
    assert rc == 0
    X2 = x.reshape(nq, nq)
    for k, (r0, r1, oo) in enumerate(blocks):
        got = dst[oo:oo + (r1 - r0) * r1].reshape(r1 - r0, r1)
        if take[k]:
            assert np.array_equal(got, X2[r0:r1, :r1])
        else:
            assert np.all(got == -7.0)


def test_make_tensors_with_a_separately_held_block():
    """Parameterized.make_tensors(x, override=...): the pipelined host path keeps q_sqrt in its own
    allocation; the remaining blocks are sliced from the packed rest in order."""
    X = np.linspace(0, 1, 6)[:, None]
    m = stvgp.StVGP(X, np.sin(X), kernels.RBF(1, 1), likelihoods.Gaussian())
    x = m.get_free_state() + 0.1 * np.arange(m.get_free_state().size)
    params = [p for p in m.all_params() if not p.fixed]
    big = max(range(len(params)), key=lambda i: params[i].size)
    sizes = [p.size for p in params]
    b0 = sum(sizes[:big])
    xt = torch.as_tensor(x)
    rest = torch.cat([xt[:b0], xt[b0 + sizes[big]:]])
    leaves_a = m.make_tensors(xt)
    vals_a = [l.detach().clone().reshape(-1) for l in leaves_a]
    m.clear_tensors()
    leaves_b = m.make_tensors(rest, override={big: xt[b0:b0 + sizes[big]].clone()})
    for a, b in zip(vals_a, leaves_b):
        assert torch.equal(a, b.detach().reshape(-1))
    m.clear_tensors()
    with pytest.raises(ValueError):
        m.make_tensors(rest, override={big: xt[:3].clone()})
    m.clear_tensors()


def test_positive_transform_roundtrip_and_offset():
    t = transforms.positive
    y = np.array([1e-5, 0.007, 0.3, 1.0, 50.0])
    assert np.allclose(t.forward(t.backward(y)), y, rtol=1e-13)
    x = np.array([-3.0, 0.0, 2.0])
    assert np.allclose(t.forward(x), np.log1p(np.exp(x)) + 1e-6)
    xt = torch.tensor(x, requires_grad=True)
    assert np.allclose(t.torch_forward(xt).detach().numpy(), t.forward(x))


def test_packing_order_matches_oracle_layout_stvgp_and_gpmc():
    r, z, A, y = abel_data(12, 9)
    m = stvgp.StVGP(r[:, None], y[:, None], kernels.RBF_csym(1, 1), likelihoods.AbelLikelihood(A),
                    mean_functions.Constant(1))
    names = [p.long_name.split(".", 1)[1] for p in m.all_params()]
    assert names == [n for n, _, _ in O.param_layout(spec_from_model(m))]
    assert m.get_free_state().size == 4 + 12 + 144
    g = gpmc.GPMC(r[:, None], y[:, None],
                  kernels.Stack([kernels.RBF_csym(1, 2), kernels.RBF_casym(1, 1)]), likelihoods.Gaussian(),
                  mean_functions.Stack([mean_functions.Constant(2), mean_functions.Zero(1)]), num_latent=3)
    names = [p.long_name.split(".", 1)[1].replace("item", "") for p in g.all_params()]
    assert names == [n for n, _, _ in O.param_layout(spec_from_model(g))]
    assert names[0] == "V"


def test_state_roundtrip_assignment_and_fixed():
    X = np.linspace(0, 1, 7)[:, None]
    m = stvgp.StVGP(X, np.sin(X), kernels.RBF(1, 1), likelihoods.Gaussian())
    m.kern.lengthscales = 0.3
    assert np.allclose(m.kern.lengthscales.value, [0.3])
    x = m.get_free_state() + 0.1
    m.set_state(x)
    assert np.allclose(m.get_free_state(), x, rtol=1e-12)
    m.kern.variance.fixed = True
    assert m.get_free_state().size == x.size - 1
    with pytest.raises(ValueError):
        m.set_state(x)
    # q_sqrt initial value: identity per latent, no transform (GPinv/stvgp.py:71-73)
    m2 = stvgp.StVGP(X, np.sin(X), kernels.RBF(1, 1), likelihoods.Gaussian())
    assert np.array_equal(m2.q_sqrt.value[:, :, 0], np.eye(7))
    assert "lengthscales" in str(m.kern)


def test_stvgp_recreates_variational_params_when_X_changes():
    X = np.linspace(0, 1, 7)[:, None]
    m = stvgp.StVGP(X, np.sin(X), kernels.RBF(1, 1), likelihoods.Gaussian())
    X2 = np.linspace(0, 1, 9)[:, None]
    m.X = X2
    m.Y = np.sin(X2)
    m._prepare()
    assert m.q_mu.shape == (9, 1) and m.q_sqrt.shape == (9, 9, 1)


def test_default_num_latent_and_mean():
    X = np.linspace(0, 1, 7)[:, None]
    m = stvgp.StVGP(X, np.zeros((7, 2)), kernels.RBF(1, 2), likelihoods.Gaussian())
    assert m.num_latent == 2 and isinstance(m.mean_function, mean_functions.Zero)
    with pytest.raises(AssertionError):
        kernels.RBF(1, 2, variance=np.ones(3))
    with pytest.raises(AssertionError):
        kernels.Stack([object()])


def test_hmc_samples_a_standard_normal():
    def f(x):
        return 0.5 * float(x @ x), x.copy()
    s = hmc.sample_HMC(f, 600, Lmax=15, epsilon=0.25, x0=np.zeros(3), burn=100, RNG=np.random.RandomState(1))
    assert s.shape == (600, 3)
    assert np.all(np.abs(s.mean(0)) < 0.3) and np.all(np.abs(s.var(0) - 1.0) < 0.4)


def test_device_hmc_walks_the_same_chain_as_the_host_loop():
    """sample_HMC_device (tensors, one host read per iteration) against sample_HMC (numpy) on a
    correlated Gaussian, incl. thinning and burn-in: same RNG stream, same accept decisions."""
    import torch
    from gpinv_b200 import hmc
    A = np.array([[2.0, 0.3, 0.0], [0.3, 1.0, 0.2], [0.0, 0.2, 0.5]])

    def f_np(x):
        return 0.5 * x @ A @ x, A @ x

    At = torch.as_tensor(A)

    def f_t(x):
        return (0.5 * x @ At @ x).reshape(1), At @ x

    x0 = np.array([0.3, -0.2, 0.1])
    a = hmc.sample_HMC(f_np, 40, Lmax=8, epsilon=0.3, x0=x0, thin=2, burn=5, RNG=np.random.RandomState(4))
    b = hmc.sample_HMC_device(f_t, 40, Lmax=8, epsilon=0.3, x0=x0, device="cpu", thin=2, burn=5,
                              RNG=np.random.RandomState(4))
    assert np.allclose(a, b, rtol=0, atol=1e-12)

    def f_nan(x):                         # NaN gradient: premature reject, the state must not move
        return torch.zeros(1, dtype=torch.float64), torch.full_like(x, float("nan"))
    c = hmc.sample_HMC_device(f_nan, 5, Lmax=4, epsilon=0.1, x0=x0, device="cpu", RNG=np.random.RandomState(0))
    assert np.allclose(c, np.tile(x0, (5, 1)))


def test_hmc_rejects_nan_gradient():
    calls = {"n": 0}

    def f(x):
        calls["n"] += 1
        g = x.copy()
        if calls["n"] > 1:
            g[0] = np.nan
        return 0.5 * float(x @ x), g
    s = hmc.sample_HMC(f, 5, Lmax=5, epsilon=0.1, x0=np.ones(2), RNG=np.random.RandomState(0))
    assert np.allclose(s, 1.0)


def test_kronecker_product_dense_helper():
    rng = np.random.RandomState(0)
    A, B = rng.randn(3, 2), rng.randn(5, 4)
    out = kronecker.kronecker_product(torch.tensor(A), torch.tensor(B)).numpy()
    assert np.allclose(out, O.loop_kronecker(A, B))


_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gpinv_b200.dist import DistContext, sharded_objective
from oracle import gpinv_oracle as O
from tests.golden.make_golden import build_spec
rank, world = int(sys.argv[2]), int(sys.argv[3])
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[4]
dist.init_process_group("gloo", rank=rank, world_size=world)
spec = build_spec("abel_stvgp_n30")
d = np.load(os.path.join(sys.argv[1], "tests", "golden", "abel_stvgp_n30.npz"))
ctx = DistContext()

def local(x, eps_local, S_global, share):
    # this rank's partial ELBO: (lik_local - KL_local) / S_global, gradient by autograd
    xt = O._t(x).clone().requires_grad_(True)
    P = O.unpack(spec, xt)
    X = O._t(spec["X"])
    L = O._kern_chol(spec["kern"], "kern", P, X)
    mean = O._mean(spec["mean"], "mean_function", P, X.shape[0])
    f, KL = O.stvgp_sample(P["q_mu"], P["q_sqrt"], L, mean, eps_local)
    val = (torch.sum(O._logp(spec, P, f)) - KL) / S_global
    (g,) = torch.autograd.grad(val, xt)
    return -val.detach(), -g

f, g = sharded_objective(local, d["x"], O._t(d["eps"]), ctx)
ok = abs(float(f) - float(d["f"])) <= 1e-12 * abs(float(d["f"])) and \
     np.linalg.norm(g.numpy() - d["g"]) <= 1e-11 * np.linalg.norm(d["g"])
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


def test_sample_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port]) for r in range(2)]
    codes = [p.wait(timeout=300) for p in procs]
    assert codes == [0, 0], codes


_SHARD_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gpinv_b200.dist import DistContext
from tests.helpers import lbar_pack_reference, shard_reverse_reference
rank, world = int(sys.argv[2]), int(sys.argv[3])
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[4]
dist.init_process_group("gloo", rank=rank, world_size=world)
ctx = DistContext()
ctx.SHARD_MIN_N = 16
n = 48
assert ctx.shard_width(n) == n // (2 * world) and ctx.shard_width(n + 2) == 0
rng = np.random.RandomState(0)                       # the same on every rank: L, dK
x = np.sort(rng.rand(n))
K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / 0.1 ** 2) + 1e-3 * np.eye(n)
L = np.linalg.cholesky(K)
dK = rng.randn(n, n); dK = dK + dK.T
Lbar_r = np.random.RandomState(10 + rank).randn(n, n)      # this rank's partial dObj/dL
# exchange 1: reduce-scatter of the packed column blocks
w = ctx.shard_width(n)
inp = torch.from_numpy(lbar_pack_reference(Lbar_r, world))
out = torch.empty(inp.numel() // world, dtype=torch.float64)
ctx.reduce_scatter(out, inp)
share = shard_reverse_reference(L, out.numpy(), ctx.shard_columns(n), w, dK)
tot = torch.tensor([share], dtype=torch.float64)
ctx.all_reduce(tot)
# replicated formula on the rank-summed Lbar
Lbar = sum(np.random.RandomState(10 + r).randn(n, n) for r in range(world))
P = np.tril(L.T @ np.tril(Lbar)); P[np.diag_indices(n)] *= 0.5
W = np.linalg.inv(L)
S = W.T @ P @ W
want = float(np.sum(0.5 * (S + S.T) * dK))
ok = abs(float(tot) - want) <= 1e-10 * abs(want)
if not ok:
    print("rank %d: sharded %r replicated %r" % (rank, float(tot), want), flush=True)
# all_gather (exchange of the sharded inversion's slices): rank order, every rank gets everything
mine = torch.full((5,), float(rank), dtype=torch.float64)
got = torch.empty(5 * world, dtype=torch.float64)
ctx.all_gather(got, mine)
ok = ok and bool(torch.equal(got.view(world, 5), torch.arange(world, dtype=torch.float64)[:, None].expand(world, 5)))
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


@pytest.mark.parametrize("world", [2, 3])
def test_column_block_sharded_cholesky_reverse_gloo(tmp_path, world):
    """Host-side logic of the sharded Cholesky reverse (gpinv_b200/dist.py): block assignment, the
    reduce-scatter layout and the sum of the ranks' shares reproduce the replicated formula."""
    script = tmp_path / "shard_worker.py"
    script.write_text(_SHARD_WORKER)
    port = str(33000 + os.getpid() % 2000 + world)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), str(world), port]) for r in range(world)]
    codes = [p.wait(timeout=300) for p in procs]
    assert codes == [0] * world, codes


def test_los_generators_match_the_loop_restatement():
    from gpinv_b200 import los
    r, z = np.linspace(0, 1, 30), np.linspace(-0.9, 0.9, 40)
    assert np.array_equal(los.make_LosMatrix(r, z), O.make_los_matrix_loops(r, z))
    assert np.array_equal(los.make_cosTheta(r, z), O.make_cos_theta(r, z))
    r, z = np.linspace(0, 1, 17), np.linspace(-1.1, 1.1, 23)      # chords outside the plasma
    assert np.array_equal(los.make_LosMatrix(r, z), O.make_los_matrix_loops(r, z))


def test_shard_bounds():
    from gpinv_b200.dist import DistContext
    c = DistContext(rank=3, world_size=8)
    assert c.shard_bounds(1024) == (384, 512)
    with pytest.raises(ValueError):
        c.shard_bounds(1001)
