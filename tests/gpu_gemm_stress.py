"""Bitwise run-to-run reproducibility of the GEMM shapes of one headline evaluation (race detector):
every shape is launched `reps` times, interleaved with cache-thrashing kernels, and compared with
its first result.  python tests/gpu_gemm_stress.py [reps]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import _capi, ops
dev = torch.device("cuda:0")
ops.ensure_workspace(dev)
lib = _capi.lib()
g = torch.Generator(device=dev).manual_seed(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P_ = lambda t: C.c_void_p(t.data_ptr())
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n, M, S = 4096, 8192, 1024
rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=g)
Q = torch.tril(rnd(n, n)); E = rnd(S, n); A = rnd(M, n); R = rnd(S, M); X = rnd(n, n)
junk = torch.empty(64 << 20, dtype=torch.float64, device=dev)      # 512 MB: flushes L2
shapes = [
    # name, la, lb, m, n, k, A, lda, B, ldb, lower, kmode
    ("U = E Q^T (kmode 1)", 0, 0, S, n, n, E, n, Q, n, 0, 1),
    ("abel fwd E A^T", 0, 0, S, M, n, E, n, A, n, 0, 0),
    ("abel bwd R A", 0, 1, S, n, M, R, M, A, n, 0, 0),
    ("Ubar = F Lc (kmode 2)", 0, 1, S, n, n, E, n, Q, n, 0, 2),
    ("Qbar = tril(U^T E)", 1, 1, n, n, S, E, n, E, n, 1, 0),
    ("tril(L^T Lbar) kmode 4", 1, 1, n, n, n, Q, n, X, n, 1, 4),
    ("tril(P W) kmode 5", 0, 1, n, n, n, Q, n, Q, n, 1, 5),
    ("dense NT 4096^3", 0, 0, n, n, n, X, n, X, n, 0, 0),
]
total_bad = 0
for name, la, lb, m, nn, k, Aa, lda, Bb, ldb, lower, kmode in shapes:
    for split in (1, 0):
        first, bad = None, 0
        for r in range(reps):
            out = torch.full((m, nn), float(r), dtype=torch.float64, device=dev)
            if r % 3 == 0:
                junk.fill_(1.0)
            rc = lib.gpinv_gemm_ex_f64(la, lb, m, nn, k, 1.0, P_(Aa), lda, 0, P_(Bb), ldb, 0, 0.0, P_(out), nn, 0, 1, lower, kmode, split, st)
            _capi.check(rc, "gemm_ex")
            if lower:
                out = torch.tril(out)
            if first is None:
                first = out.clone()
            elif not torch.equal(out, first):
                bad += 1
                if bad == 1:
                    D = (out - first).abs()
                    bm = -(-m // 128); bn = -(-nn // 128)
                    Bm = D[:bm * 128, :bn * 128].reshape(bm, 128, bn, 128).amax(dim=(1, 3)) if (m % 128 == 0 and nn % 128 == 0) else D
                    print("   %s split=%d rep %d: blocks %s max %.3e" % (name, split, r, (Bm > 0).nonzero().cpu().tolist()[:8], float(D.max())), flush=True)
        total_bad += bad
        print("%-28s split=%d: %d of %d repetitions differ" % (name, split, bad, reps - 1), flush=True)
print("total differing:", total_bad)
sys.exit(1 if total_bad else 0)
