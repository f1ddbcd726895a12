"""Standalone op-by-op check against torch references (prints every error; exit 1 on failure).
Run on the GPU box:  python tests/gpu_selfcheck.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops  # noqa: E402
from oracle import gpinv_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
FAIL = []


def rel(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def report(name, err, tol):
    ok = err <= tol and err == err
    print("%-58s err=%.3e tol=%.1e %s" % (name, err, tol, "ok" if ok else "FAIL"), flush=True)
    if not ok:
        FAIL.append(name)


def t(a):
    return torch.as_tensor(a, dtype=torch.float64, device=dev)


def check_gemm():
    g = torch.Generator().manual_seed(0)
    for (m, n, k) in [(8, 8, 4), (37, 53, 29), (128, 128, 16), (129, 257, 100), (300, 200, 1000), (1024, 512, 384)]:
        for ta in (False, True):
            for tb in (False, True):
                A = torch.randn((k, m) if ta else (m, k), generator=g, dtype=torch.float64)
                B = torch.randn((n, k) if tb else (k, n), generator=g, dtype=torch.float64)
                ref = (A.T if ta else A) @ (B.T if tb else B)
                out = ops.gemm(A.to(dev), B.to(dev), ta, tb, False)
                report("gemm %dx%dx%d ta=%d tb=%d" % (m, n, k, ta, tb), rel(out, ref), 1e-13)
    A = torch.randn(300, 77, generator=g, dtype=torch.float64)
    out = ops.gemm(A.to(dev), A.to(dev), False, True, True)
    report("gemm lower 300x300x77", rel(torch.tril(out), torch.tril(A @ A.T)), 1e-13)
    # many tiles per CTA (persistent stream-K partition: whole tiles, shared tiles, ragged edges)
    gd = torch.Generator(device=dev).manual_seed(1)
    for (m, n, k) in [(2048, 2048, 512), (1000, 3000, 700), (4096, 1024, 4096), (1024, 8192, 2048), (2500, 2500, 256),
                      (130, 140, 20000), (64, 64, 50000)]:
        for ta in (False, True):
            for tb in (False, True):
                A = torch.randn((k, m) if ta else (m, k), generator=gd, dtype=torch.float64, device=dev)
                B = torch.randn((n, k) if tb else (k, n), generator=gd, dtype=torch.float64, device=dev)
                ref = torch.matmul(A.T if ta else A, B.T if tb else B)
                out = ops.gemm(A, B, ta, tb, False)
                report("gemm %dx%dx%d ta=%d tb=%d" % (m, n, k, ta, tb), rel(out, ref), 1e-12)
                out2 = ops.gemm(A, B, ta, tb, False)
                report("   ... repeat bitwise", 0.0 if torch.equal(out, out2) else 1.0, 0.5)
    for (m, k) in [(2500, 256), (3072, 1000), (4096, 64)]:
        A = torch.randn(m, k, generator=gd, dtype=torch.float64, device=dev)
        out = ops.gemm(A, A, False, True, True)
        report("gemm lower %dx%dx%d" % (m, m, k), rel(torch.tril(out), torch.tril(A @ A.T)), 1e-12)
        At = A.T.contiguous()
        out = ops.gemm(At, At, True, False, True)
        report("gemm lower TN %dx%dx%d" % (m, m, k), rel(torch.tril(out), torch.tril(A @ A.T)), 1e-12)


def check_kcore():
    rng = np.random.RandomState(0)
    for kind in ("rbf", "rbf_csym", "rbf_casym"):
        for (n, n2, D, ard) in [(10, 11, 2, True), (33, 0, 1, False), (257, 130, 1, False), (100, 0, 3, True)]:
            X = rng.randn(n, D)
            X2 = rng.randn(n2, D) if n2 else None
            ls = np.exp(0.3 * rng.randn(D)) if ard else np.array([0.7])
            lst = torch.tensor(ls, requires_grad=True)
            ref = O.kcore(kind, torch.tensor(X), None if X2 is None else torch.tensor(X2), lst)
            if X2 is None:
                ref = ref + 1e-6 * torch.eye(n, dtype=torch.float64)
            W = torch.tensor(rng.randn(*ref.shape))
            (gref,) = torch.autograd.grad((ref * W).sum(), lst)
            lsd = t(ls).requires_grad_(True)
            out = ops.kcore(t(X), None if X2 is None else t(X2), lsd, ops.MODE[kind], 1e-6 if X2 is None else 0.0)
            (gout,) = torch.autograd.grad((out * W.to(dev)).sum(), lsd)
            report("kcore %s n=%d n2=%d D=%d" % (kind, n, n2, D), rel(out, ref), 1e-14)
            report("kcore_bwd %s n=%d n2=%d D=%d" % (kind, n, n2, D), rel(gout, gref), 1e-11)


def spd(n, seed):
    rng = np.random.RandomState(seed)
    x = np.sort(rng.rand(n))[:, None] * 3.0
    K = np.exp(-0.5 * (x - x.T) ** 2 / 0.5 ** 2) + 1e-4 * np.eye(n) + 0.05 * np.diag(rng.rand(n))
    return K


def check_potrf():
    for n in (1, 5, 64, 65, 100, 256, 300, 515, 1024, 1100, 2048, 2500, 3072):   # > 1024: hybrid blocked reverse
        K = spd(n, n)
        Kt = torch.tensor(K, requires_grad=True)
        Lref = torch.linalg.cholesky(Kt)
        Kjunk = K.copy()
        Kd = t(Kjunk).requires_grad_(True)
        L, info = ops.potrf(Kd)
        report("potrf n=%d (info=%d)" % (n, int(info[0].item())), rel(L, Lref), 1e-12)
        rng = np.random.RandomState(1)
        Lbar = np.tril(rng.randn(n, n))
        (gref,) = torch.autograd.grad((Lref * torch.tensor(Lbar)).sum(), Kt)
        (gout,) = torch.autograd.grad((L * t(Lbar)).sum(), Kd)
        report("potrf_bwd n=%d" % n, rel(gout, gref), 1e-9)
        if n > 2048:
            # three blocks: the reverse sweep forks onto a side stream (second half of the split-K
            # workspace); repeated calls run with warm plans, i.e. with real concurrency
            for rep in range(2):
                (g2,) = torch.autograd.grad((ops.potrf(Kd)[0] * t(Lbar)).sum(), Kd)
                report("potrf_bwd n=%d repeat %d" % (n, rep), rel(g2, gref), 1e-9)
    # ill-conditioned kernel matrices at the headline size (cond ~1e9); repeated to catch races
    for n in (3000, 4096):
        x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)[:, None]
        K = ops.kcore(x, None, t([0.3]), 1, 1e-6)
        Lref = torch.linalg.cholesky(K)
        for rep in range(3):
            L, info = ops.potrf(K)
            report("potrf csym n=%d rep %d (info=%d)" % (n, rep, int(info[0].item())),
                   rel(L, Lref) if int(info[0].item()) == 0 else 1.0, 1e-7)   # vs cuSOLVER at cond(K)~1e9
        resid = float((L @ L.T - K).abs().max())
        report("potrf csym n=%d residual |LL^T-K|" % n, resid, 1e-12)
    Kbad = spd(100, 3)
    Kbad[70, 70] = -1.0
    L, info = ops.potrf(t(Kbad))
    report("potrf info on non-PD (info=%d)" % int(info[0].item()), 0.0 if int(info[0].item()) == 71 else 1.0, 0.5)


def check_trsm():
    for (m, n) in [(3, 5), (100, 64), (130, 200), (257, 515)]:
        K = spd(n, 7)
        L = np.linalg.cholesky(K)
        B = np.random.RandomState(2).randn(m, n)
        ref = torch.linalg.solve_triangular(torch.tensor(L), torch.tensor(B).T, upper=False).T
        out = ops.trsm_rlt(t(L), t(B))
        report("trsm_rlt m=%d n=%d" % (m, n), rel(out, ref), 1e-11)


def check_sample():
    rng = np.random.RandomState(3)
    for (n, S, R, q_diag) in [(30, 10, 1, False), (150, 70, 2, False), (130, 33, 3, True), (300, 129, 1, False)]:
        K = spd(n, 11)
        Lc = np.linalg.cholesky(K)
        sv = np.exp(0.2 * rng.randn(R))
        q_mu = rng.randn(n, R)
        mean = rng.randn(n, R)
        E = rng.randn(R, S, n)
        if q_diag:
            q_sqrt = np.exp(0.3 * rng.randn(n, R))
        else:
            q_sqrt = 0.5 * np.eye(n)[:, :, None] + 0.1 * rng.randn(n, n, R)
        # oracle
        tl = [torch.tensor(a, requires_grad=True) for a in (Lc, sv, q_sqrt, q_mu, mean)]
        Lc_t, sv_t, qs_t, qm_t, mn_t = tl
        Lfull = torch.tril(Lc_t)[:, :, None] * sv_t[None, None, :]
        f, KL = O.stvgp_sample(qm_t, qs_t, Lfull, mn_t, torch.tensor(E), q_diag)
        Wf = torch.tensor(rng.randn(S, n, R))
        obj = (f * Wf).sum() - 0.37 * KL
        gref = torch.autograd.grad(obj, tl)
        # device
        dl = [t(a).requires_grad_(True) for a in (Lc, sv, q_sqrt, q_mu, mean)]
        Lc_d, sv_d, qs_d, qm_d, mn_d = dl
        Q = qs_d.T.contiguous() if q_diag else ops.tril_unpack(qs_d)
        F, expF, kl, U = ops.stvgp_sample(Lc_d, sv_d, Q, qm_d.T.contiguous(), mn_d.T.contiguous(), t(E), q_diag, True)
        fd = F.permute(1, 2, 0)
        report("sample F n=%d S=%d R=%d diag=%d" % (n, S, R, q_diag), rel(fd, f), 1e-12)
        report("sample expF", rel(expF, torch.exp(F)), 1e-14)
        report("sample KL", rel(kl, KL.reshape(1)), 1e-12)
        objd = (fd * Wf.to(dev)).sum() - 0.37 * kl.sum()
        gout = torch.autograd.grad(objd, dl)
        names = ["Lc", "sqrt_v", "q_sqrt", "q_mu", "mean"]
        for nm, a, b in zip(names, gout, gref):
            if nm == "Lc":
                a, b = torch.tril(a), torch.tril(b)
            report("sample_bwd d/d%s" % nm, rel(a, b), 1e-11)


def check_liks():
    rng = np.random.RandomState(4)
    n, S, R = 70, 21, 2
    F = rng.randn(R, S, n)
    Y = rng.randn(n, R)
    Ft = torch.tensor(F, requires_grad=True)
    ref = ((Ft.permute(1, 2, 0) - torch.tensor(Y)[None]) ** 2).sum()
    (gref,) = torch.autograd.grad(ref, Ft)
    Fd = t(F).requires_grad_(True)
    out = ops.gauss_sumsq(Fd, t(Y))
    (gout,) = torch.autograd.grad(out.sum(), Fd)
    report("gauss_sumsq", rel(out, ref.reshape(1)), 1e-13)
    report("gauss_sumsq bwd", rel(gout, gref), 1e-13)
    for (n, M, S) in [(30, 40, 10), (200, 333, 65)]:
        r = np.linspace(0, 1, n)
        z = np.linspace(-0.9, 0.9, M)
        A = O.make_los_matrix(r, z)
        F = 0.3 * rng.randn(S, n)
        Y = rng.randn(M)
        Ft = torch.tensor(F, requires_grad=True)
        P = torch.exp(Ft) @ torch.tensor(A).T
        ref = ((torch.tensor(Y)[None] - P) ** 2).sum()
        (gref,) = torch.autograd.grad(ref, Ft)
        Fd = t(F).requires_grad_(True)
        ss, Rs = ops.abel_resid(Fd, torch.exp(Fd).detach(), t(A), t(Y))
        (gout,) = torch.autograd.grad(ss.sum(), Fd)
        report("abel_resid n=%d M=%d S=%d" % (n, M, S), rel(ss, ref.reshape(1)), 1e-13)
        report("abel_resid bwd", rel(gout, gref), 1e-12)
    for (n, m, k, S) in [(12, 9, 7, 3), (30, 40, 50, 5), (64, 33, 70, 9)]:
        from tests.helpers import spec_data
        r, A, cosT, lam, y = spec_data(n, m, k)
        F = np.stack([-1.0 + 0.3 * rng.randn(S, n), -0.5 + 0.2 * rng.randn(S, n), 0.5 * rng.randn(S, n)], 0)
        Ft = torch.tensor(F, requires_grad=True)
        Pref = O.spec_abel_transform(Ft.permute(1, 2, 0), torch.tensor(A), torch.tensor(cosT), torch.tensor(lam))
        ref = ((Pref - torch.tensor(y)[None]) ** 2).sum()
        (gref,) = torch.autograd.grad(ref, Ft)
        Fd = t(F).requires_grad_(True)
        ss, P = ops.spec_resid(Fd, t(A.T.copy()), t(cosT.T.copy()), t(lam), t(y))
        (gout,) = torch.autograd.grad(ss.sum(), Fd)
        report("spec_resid P n=%d m=%d k=%d S=%d" % (n, m, k, S), rel(P, Pref), 1e-13)
        report("spec_resid sumsq", rel(ss, ref.reshape(1)), 1e-13)
        report("spec_resid bwd", rel(gout, gref), 1e-12)
        report("spec_transform", rel(ops.spec_transform(t(F), t(A.T.copy()), t(cosT.T.copy()), t(lam)), Pref), 1e-13)
        # uniform wavelength grid: the recurrence kernels (restart every 16 wavelengths), normal line widths and
        # lines narrower than the grid spacing (direct evaluation per radial point), k not a multiple of 16
        for width, tag in ((-0.5, "wide lines"), (-5.0, "narrow lines -> direct path"), (-2.2, "mixed")):
            lam_u = np.linspace(-2.0, 2.5, k)
            dl = float((lam_u[-1] - lam_u[0]) / (k - 1))
            Fu = F.copy()
            Fu[1] = width + (1.5 if tag == "mixed" else 0.2) * rng.randn(S, n)
            Fut = torch.tensor(Fu, requires_grad=True)
            Pr = O.spec_abel_transform(Fut.permute(1, 2, 0), torch.tensor(A), torch.tensor(cosT), torch.tensor(lam_u))
            rr = ((Pr - torch.tensor(y)[None]) ** 2).sum()
            (gr,) = torch.autograd.grad(rr, Fut)
            Fud = t(Fu).requires_grad_(True)
            ss, P = ops.spec_resid(Fud, t(A.T.copy()), t(cosT.T.copy()), t(lam_u), t(y), dl)
            (go,) = torch.autograd.grad(ss.sum(), Fud)
            report("spec recurrence P (%s) n=%d m=%d k=%d" % (tag, n, m, k), rel(P, Pr), 1e-12)
            report("spec recurrence sumsq (%s)" % tag, rel(ss, rr.reshape(1)), 1e-12)
            report("spec recurrence bwd (%s)" % tag, rel(go, gr), 1e-11)
            report("spec recurrence transform (%s)" % tag,
                   rel(ops.spec_transform(t(Fu), t(A.T.copy()), t(cosT.T.copy()), t(lam_u), dl), Pr), 1e-12)
    x = rng.randn(100001)
    report("sumsq", rel(ops.sumsq(t(x)), torch.tensor([np.sum(x * x)])), 1e-13)
    # fused Gaussian log-density from the residual sum of squares, and its reverse
    for ssv, varv, cnt in ((123.4, 0.37, 4096.0), (5.0e7, 2.5, 8388608.0)):
        sst = torch.tensor([ssv], dtype=torch.float64, requires_grad=True)
        vt = torch.tensor([varv], dtype=torch.float64, requires_grad=True)
        ref = -0.5 * cnt * (np.log(2 * np.pi) + torch.log(vt)) - 0.5 * sst / vt
        gs, gv = torch.autograd.grad((3.0 * ref).sum(), [sst, vt])
        ssd, vd = t([ssv]).requires_grad_(True), t([varv]).requires_grad_(True)
        out = ops.gauss_logp(ssd, vd, cnt)
        g1, g2 = torch.autograd.grad((3.0 * out).sum(), [ssd, vd])
        report("gauss_logp value", rel(out, ref), 1e-15)
        report("gauss_logp d/dss, d/dvar", max(rel(g1, gs), rel(g2, gv)), 1e-14)


def check_kron():
    """Kronecker-factored apply and sampling stage against the dense Kronecker product."""
    rng = np.random.RandomState(5)
    for (n1, n2, B) in [(3, 5, 4), (6, 9, 7), (64, 128, 33), (70, 50, 130)]:
        L1 = np.tril(rng.randn(n1, n1)) + 2 * np.eye(n1)
        L2 = np.tril(rng.randn(n2, n2)) + 2 * np.eye(n2)
        U = rng.randn(B, n1 * n2)
        dense = np.kron(L1, L2)
        for trans in (False, True):
            ref = U @ (dense if trans else dense.T)
            out = ops.kron_apply(t(L1), t(L2), t(U), trans)
            report("kron_apply %dx%d B=%d trans=%d" % (n1, n2, B, trans), rel(out, torch.tensor(ref)), 1e-13)
    for (n1, n2, S, R, q_diag, want_exp) in [(4, 5, 6, 2, True, False), (6, 9, 7, 3, False, False), (12, 20, 16, 3, False, True),
                                             (16, 24, 40, 2, True, True), (64, 40, 33, 1, False, False)]:
        n = n1 * n2
        L1 = np.tril(0.3 * rng.randn(n1, n1)) + np.eye(n1)
        L2 = np.tril(0.3 * rng.randn(n2, n2)) + np.eye(n2)
        sv = np.exp(0.2 * rng.randn(R))
        q_mu, mean, E = 0.3 * rng.randn(n, R), 0.2 * rng.randn(n, R), rng.randn(R, S, n)
        q_sqrt = np.exp(0.3 * rng.randn(n, R)) if q_diag else 0.5 * np.eye(n)[:, :, None] + 0.1 * rng.randn(n, n, R)
        tl = [torch.tensor(a, requires_grad=True) for a in (L1, L2, sv, q_sqrt, q_mu, mean)]
        L1t, L2t, svt, qst, qmt, mnt = tl
        Lfull = O.kronecker_product(torch.tril(L1t), torch.tril(L2t))[:, :, None] * svt[None, None, :]
        f, KL = O.stvgp_sample(qmt, qst, Lfull, mnt, torch.tensor(E), q_diag)
        Wf = torch.tensor(rng.randn(S, n, R))
        obj = ((torch.exp(0.1 * f) if want_exp else f) * Wf).sum() - 0.37 * KL
        gref = torch.autograd.grad(obj, tl)
        dl = [t(a).requires_grad_(True) for a in (L1, L2, sv, q_sqrt, q_mu, mean)]
        L1d, L2d, svd, qsd, qmd, mnd = dl
        Q = qsd.T.contiguous() if q_diag else ops.tril_unpack(qsd)
        F, expF, kl, U, T = ops.kron_sample(L1d, L2d, svd, Q, qmd.T.contiguous(), mnd.T.contiguous(), t(E), q_diag, want_exp)
        fd = F.permute(1, 2, 0)
        tag = "kron_sample %dx%d S=%d R=%d diag=%d exp=%d" % (n1, n2, S, R, q_diag, want_exp)
        report(tag + " F", rel(fd, f), 1e-12)
        report(tag + " KL", rel(kl, KL.reshape(1)), 1e-12)
        if want_exp:
            report(tag + " expF", rel(expF.permute(1, 2, 0), torch.exp(f)), 1e-12)
        objd = ((torch.exp(0.1 * fd) if want_exp else fd) * Wf.to(dev)).sum() - 0.37 * kl.sum()
        gout = torch.autograd.grad(objd, dl)
        for nm, a, b in zip(("L1", "L2", "sqrt_v", "q_sqrt", "q_mu", "mean"), gout, gref):
            if nm in ("L1", "L2"):
                a, b = torch.tril(a), torch.tril(b)
            if nm == "q_sqrt" and not q_diag:
                a, b = torch.tril(a.permute(2, 0, 1)), torch.tril(b.permute(2, 0, 1))
            report(tag + " d/d" + nm, rel(a, b), 1e-11)


def check_pack():
    """unpack_state / pack_grad (softplus + 1e-6 and its chain rule), the analytic whitened KL and the
    TensorFlow Adam update against torch / numpy references."""
    rng = np.random.RandomState(7)
    sizes = [1, 3, 1, 130, 130 * 130, 7]
    codes = [1, 1, 1, 0, 0, 1]
    offs = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    x = rng.randn(offs[-1]) * 2.0
    x[0] = 40.0
    x[1] = -40.0
    xt = torch.tensor(x, requires_grad=True)
    ref_vals = [(torch.nn.functional.softplus(xt[a:b], threshold=1e9) + 1e-6) if c else xt[a:b]
                for a, b, c in zip(offs[:-1], offs[1:], codes)]
    Ws = [torch.tensor(rng.randn(sz)) for sz in sizes]
    ref_obj = sum((v * w).sum() for i, (v, w) in enumerate(zip(ref_vals, Ws)) if i != 2)     # block 2 unused
    (gref,) = torch.autograd.grad(ref_obj, xt)
    xd = t(x).requires_grad_(True)
    vals = ops.unpack_state(xd, offs, codes)
    for i, (v, r) in enumerate(zip(vals, ref_vals)):
        report("unpack_state block %d" % i, rel(v, r), 1e-15)
    obj = sum((v * w.to(dev)).sum() for i, (v, w) in enumerate(zip(vals, Ws)) if i != 2)
    (gout,) = torch.autograd.grad(obj, xd)
    report("pack_grad (chain rule, unused block -> 0)", rel(gout, gref), 1e-15)
    for (n, R, q_diag) in [(5, 1, False), (40, 3, False), (130, 2, True), (257, 1, False)]:
        q_mu = rng.randn(n, R)
        q_sqrt = np.exp(0.3 * rng.randn(n, R)) if q_diag else 0.5 * np.eye(n)[:, :, None] + 0.1 * rng.randn(n, n, R)
        qm, qs = torch.tensor(q_mu, requires_grad=True), torch.tensor(q_sqrt, requires_grad=True)
        ref = O.gauss_kl_white_diag(qm, qs) if q_diag else O.gauss_kl_white(qm, qs)
        gm, gq = torch.autograd.grad(3.0 * ref, [qm, qs])
        qmd, qsd = t(q_mu).requires_grad_(True), t(q_sqrt).requires_grad_(True)
        out = ops.kl_white(qmd, qsd, q_diag)
        gmd, gqd = torch.autograd.grad(3.0 * out.sum(), [qmd, qsd])
        report("kl_white n=%d R=%d diag=%d" % (n, R, q_diag), rel(out, ref.reshape(1)), 1e-13)
        report("kl_white d/dq_mu", rel(gmd, gm), 1e-14)
        report("kl_white d/dq_sqrt (zero above the diagonal)", rel(gqd, gq), 1e-13)
    # Adam (TensorFlow's formula), 7 steps, odd length
    n = 1001
    x0, lr, b1, b2, eps = rng.randn(n), 0.01, 0.9, 0.999, 1e-8
    gs = [rng.randn(n) for _ in range(7)]
    xr, mr, vr = x0.copy(), np.zeros(n), np.zeros(n)
    for k, g in enumerate(gs):
        tt = k + 1
        lr_t = lr * np.sqrt(1 - b2 ** tt) / (1 - b1 ** tt)
        mr = b1 * mr + (1 - b1) * g
        vr = b2 * vr + (1 - b2) * g * g
        xr = xr - lr_t * mr / (np.sqrt(vr) + eps)
    xd, md, vd = t(x0), torch.zeros(n, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for g in gs:
        ops.adam_step(xd, md, vd, t(g), step, lr, b1, b2, eps)
    report("adam_step x after 7 steps", rel(xd, torch.tensor(xr)), 1e-14)
    report("adam_step step counter", abs(int(step.item()) - 7), 0.5)


def check_shard():
    """Column-block sharded Cholesky reverse (multi-GPU path) on one GPU: the pack layout against its
    numpy restatement, the triangle compaction round trip, and the sum over the ranks' shares against the
    replicated reverse, for G = 2, 4, 8."""
    from tests.helpers import lbar_pack_reference
    for n, Gs in ((1024, (2, 4, 8)), (640, (2,)), (2048, (4,))):
        x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)[:, None]
        ls = t([0.3])
        L, info, W = ops.chol_core(x, ls, 1, 1e-6, True)
        rng = np.random.RandomState(n)
        Lbar = t(rng.randn(n, n))                   # the strict upper part must be ignored
        ref = ops.chol_core_bwd(x, ls, 1, L, Lbar, W)
        for G in Gs:
            w = n // (2 * G)
            packed = ops.lbar_pack(Lbar, G)
            report("lbar_pack n=%d G=%d" % (n, G),
                   rel(packed, torch.tensor(lbar_pack_reference(Lbar.cpu().numpy(), G))), 0.0)
            ops.trtri_side_join()
            tot = torch.zeros_like(ref)
            chunks = packed.view(G, -1)
            for r in range(G):
                c0s = [r * w, (2 * G - 1 - r) * w]
                tot += ops.chol_core_bwd_shard(x, ls, 1, L, W, chunks[r].contiguous(), w, c0s)
            report("chol_core_bwd_shard n=%d G=%d (sum of shares vs replicated)" % (n, G), rel(tot, ref), 1e-7)
    # sharded last level of the inversion: G emulated ranks, their slices gathered, against the whole inverse
    for n, Gs in ((1024, (2, 4, 8)), (2048, (2, 8))):
        x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)[:, None]
        ls = t([0.3])
        L, info, Wfull = ops.chol_core(x, ls, 1, 1e-6, True)
        ops.trtri_side_join()
        for G in Gs:
            assert ops.inverse_shardable(n, G)
            Ws, packs = [], []
            for r in range(G):
                W = torch.empty((2, n, n), dtype=torch.float64, device=dev)
                ops.INVERSE_SHARD = (r, G)
                try:
                    ops.start_inverse(L, W)
                finally:
                    ops.INVERSE_SHARD = None
                ops.trtri_side_join()
                Ws.append(W)
                packs.append(ops.inverse_slices_pack(W, r, G))
            gathered = torch.stack(packs).reshape(-1)
            worst = 0.0
            for r in range(G):
                ops.inverse_slices_unpack_(Ws[r], gathered, G)
                worst = max(worst, rel(Ws[r][0], Wfull[0]))
            report("sharded inversion n=%d G=%d (every rank's W vs the whole inverse)" % (n, G), worst, 1e-12)
    for n in (1, 7, 300, 1024):
        M = t(np.random.RandomState(n).randn(n, n))
        c = ops.tri_compact(M)
        iu = np.tril_indices(n)
        report("tri_compact n=%d" % n, rel(c, torch.tensor(M.cpu().numpy()[iu])), 0.0)
        Z = torch.full((n, n), 7.0, dtype=torch.float64, device=dev)
        ops.tri_expand_(c, Z)
        want = torch.where(torch.ones(n, n, device=dev).tril().bool(), M, torch.full_like(M, 7.0))
        report("tri_expand n=%d" % n, rel(Z, want), 0.0)


if __name__ == "__main__":
    t0 = time.time()
    names = sys.argv[1:] or ["check_gemm", "check_kcore", "check_potrf", "check_trsm", "check_sample", "check_kron",
                             "check_pack", "check_liks", "check_shard"]
    for fn in [globals()[nm] for nm in names]:
        try:
            fn()
            torch.cuda.synchronize()
        except Exception as e:  # keep going so one broken op does not hide the others
            import traceback
            traceback.print_exc()
            FAIL.append(fn.__name__ + ": " + repr(e)[:200])
    print("elapsed %.1fs; %d failures" % (time.time() - t0, len(FAIL)))
    for f in FAIL:
        print("FAILED:", f)
    sys.exit(1 if FAIL else 0)
