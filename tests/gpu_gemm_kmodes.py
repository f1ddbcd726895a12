"""Triangular-aware GEMM variants (kmode 1-5, lower output, all layouts) at ragged and aligned
sizes: correctness against torch on operands that are really triangular, and run-to-run bitwise
reproducibility (race detector).  python tests/gpu_gemm_kmodes.py [n ...]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import _capi, ops
dev = torch.device("cuda:0")
ops.ensure_workspace(dev)
lib = _capi.lib()
g = torch.Generator(device=dev).manual_seed(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
bad = 0
P = lambda t: C.c_void_p(t.data_ptr())


def view(X, layout):
    """storage such that element (r,k) of the GEMM view equals X[r,k]"""
    return X.contiguous() if layout == 0 else X.T.contiguous()


def case(tag, Aop, Bop, la, lb, lower, kmode, split, reps=int(os.environ.get('REPS', '4'))):
    """C = Aop @ Bop^T  with Aop [m,k], Bop [n,k] given as dense torch matrices (with their zeros)"""
    global bad
    m, k = Aop.shape
    n = Bop.shape[0]
    As, Bs = view(Aop, la), view(Bop, lb)
    ref = Aop @ Bop.T
    outs = []
    for r in range(reps):
        Cm = torch.full((m, n), 7.0, dtype=torch.float64, device=dev)
        rc = lib.gpinv_gemm_ex_f64(la, lb, m, n, k, 1.0, P(As), As.shape[1], 0, P(Bs), Bs.shape[1], 0, 0.0, P(Cm), n, 0, 1,
                                   int(lower), kmode, int(split), st)
        _capi.check(rc, "gemm_ex")
        outs.append(Cm)
    torch.cuda.synchronize()
    msk = torch.tril(torch.ones(m, n, dtype=torch.bool, device=dev)) if lower else torch.ones(m, n, dtype=torch.bool, device=dev)
    err = float(((outs[0] - ref)[msk]).abs().max() / ref.abs().max())
    same = all(torch.equal(outs[0][msk], o[msk]) for o in outs[1:])
    ok = err < 1e-12 and same
    bad += (not ok)
    print("%-44s la=%d lb=%d lower=%d kmode=%d split=%d  m=%d n=%d k=%d  err=%.2e repeat-bitwise=%s %s"
          % (tag, la, lb, lower, kmode, split, m, n, k, err, same, "ok" if ok else "FAIL"), flush=True)


sizes = [int(a) for a in sys.argv[1:]] or [2500, 3000, 2048]
for n in ([] if os.environ.get("ONLY_BATCHED") else sizes):
    L = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))      # lower
    W = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
    Lb = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
    X = torch.randn(n, n, dtype=torch.float64, device=dev, generator=g)
    for split in (1, 0):
        # P = tril(L^T Lbar): A(i,k) = L[k][i] (zero k < i), B(c,k) = Lbar[k][c]   -> LMN, LMN, kmode 4
        case("P = tril(L^T Lbar)", L.T.contiguous(), Lb.T.contiguous(), 1, 1, True, 4, split)
        # T1 = tril(P W): A = P lower (LK), B(c,k) = W[k][c] (zero k < c)        -> LK, LMN, kmode 5
        case("T1 = tril(P W)", L, W.T.contiguous(), 0, 1, True, 5, split)
        # tril(T1^T W): A(i,k) = T1[k][i] (zero k < i), B(c,k) = W[k][c]          -> LMN, LMN, kmode 4
        case("tril(T1^T W)", L.T.contiguous(), W.T.contiguous(), 1, 1, True, 4, split)
        # X L^T (k <= column): A dense LK, B(c,k) = L[c][k]                        -> LK, LK, kmode 1
        case("X L^T", X, L, 0, 0, False, 1, split)
        # X L (k >= column): B(c,k) = L[k][c]                                      -> LK, LMN, kmode 2
        case("X L", X, L.T.contiguous(), 0, 1, False, 2, split)
        # L X (k <= row): A = L lower LK, B(c,k) = X[k][c]                          -> LK, LMN, kmode 3
        case("L X", L, X.T.contiguous(), 0, 1, False, 3, split)
        case("dense NT", X, W, 0, 0, False, 0, split)
        case("dense TN lower", X.T.contiguous(), X.T.contiguous(), 1, 1, True, 0, split)


def batched(tag, n, s, split, reps=4):
    """the two products of one trtri level on the diagonal pairs of an n x n lower matrix:
    T21 = L21 W11 (kmode 2), W21 = -W22 T21 (kmode 3), batched over the pairs, windows of ld = n"""
    global bad
    Lm = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g)) + 3 * torch.eye(n, dtype=torch.float64, device=dev)
    Wm = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
    nfull = n // (2 * s)
    if nfull < 1:
        return
    strL = 2 * s * n + 2 * s
    outs = []
    for r in range(reps):
        T = torch.zeros(n, n, dtype=torch.float64, device=dev)
        Wc = Wm.clone()
        off21 = s * n            # block (s.., 0..) of the first pair
        rc = lib.gpinv_gemm_ex_f64(0, 1, s, s, s, 1.0, C.c_void_p(Lm.data_ptr() + 8 * off21), n, strL, P(Wc), n, strL, 0.0,
                                   C.c_void_p(T.data_ptr() + 8 * off21), n, strL, nfull, 0, 2, int(split), st)
        _capi.check(rc, "gemm_ex")
        rc = lib.gpinv_gemm_ex_f64(0, 1, s, s, s, -1.0, C.c_void_p(Wc.data_ptr() + 8 * (s * n + s)), n, strL,
                                   C.c_void_p(T.data_ptr() + 8 * off21), n, strL, 0.0,
                                   C.c_void_p(Wc.data_ptr() + 8 * off21), n, strL, nfull, 0, 3, int(split), st)
        _capi.check(rc, "gemm_ex")
        outs.append(Wc)
    torch.cuda.synchronize()
    ref = Wm.clone()
    for b in range(nfull):
        i0 = 2 * s * b
        T21 = Lm[i0 + s:i0 + 2 * s, i0:i0 + s] @ Wm[i0:i0 + s, i0:i0 + s]
        ref[i0 + s:i0 + 2 * s, i0:i0 + s] = -Wm[i0 + s:i0 + 2 * s, i0 + s:i0 + 2 * s] @ T21
    err = float((outs[0] - ref).abs().max() / ref.abs().max())
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    ok = err < 1e-12 and same
    bad += (not ok)
    print("%-30s n=%d s=%d batch=%d split=%d err=%.2e repeat-bitwise=%s %s" % (tag, n, s, nfull, split, err, same, "ok" if ok else "FAIL"), flush=True)


for n in sizes:
    for s in (64, 128, 256, 512, 1024):
        for split in (1, 0):
            batched("trtri level (batched windows)", n, s, split)
print("failures:", bad)
sys.exit(1 if bad else 0)
