"""Times the Cholesky forward / reverse sweeps (CUDA events, best of 10) and checks the reverse
against torch's solve-based gradient.  Mode / block size come from GPINV_POTRF_BWD, GPINV_BWD_BLOCK."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
for n in (2048, 4096):
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)[:, None]
    K = ops.kcore(x, None, torch.tensor([0.3], dtype=torch.float64, device=dev), 1, 1e-6)
    L, info = ops.potrf(K)
    g = torch.Generator(device=dev).manual_seed(0)
    Lbar = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
    for fn, nm in ((lambda: ops.potrf(K), "potrf"), (lambda: ops.potrf_bwd(L, Lbar), "potrf_bwd")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("%s n=%d mode=%s block=%s: %.3f ms" % (nm, n, os.environ.get("GPINV_POTRF_BWD", "hybrid"),
                                                     os.environ.get("GPINV_BWD_BLOCK", "1024"), min(ts)), flush=True)
    G = ops.potrf_bwd(L, Lbar)
    # reference: Phi form with torch triangular solves
    P = torch.tril(L.T @ Lbar); P.diagonal().mul_(0.5)
    Z = torch.linalg.solve_triangular(L.T, P, upper=True)          # L^-T P
    Z = torch.linalg.solve_triangular(L, Z.T, upper=False).T        # ... L^-1
    Gref = 0.5 * (Z + Z.T)
    print("   rel err vs solve-based: %.3e" % float((G - Gref).norm() / Gref.norm()), flush=True)
