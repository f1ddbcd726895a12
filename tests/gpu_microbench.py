"""Per-kernel timing on the GPU box (CUDA events, warm-up, L2-exceeding operands).
Not a test: prints a table used to fill DESIGN.md / profiles/."""
import os
import sys
import json

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    # cuBLAS DGEMM denominator
    for N in (4096, 8192):
        A = torch.randn(N, N, dtype=torch.float64, device=dev, generator=g)
        B = torch.randn(N, N, dtype=torch.float64, device=dev, generator=g)
        best, med = timeit(lambda: torch.matmul(A, B), iters=8)
        print("cublas dgemm %d^3: best %.3f ms (%.2f TF) median %.3f ms (%.2f TF)" % (N, best, 2 * N**3 / best / 1e9, med, 2 * N**3 / med / 1e9), flush=True)
        out["cublas_dgemm_%d_tf" % N] = 2 * N**3 / best / 1e9
        best, med = timeit(lambda: ops.gemm(A, B, False, True, False), iters=8)
        print("gpinv gemm NT %d^3: best %.3f ms (%.2f TF) median %.3f ms" % (N, best, 2 * N**3 / best / 1e9, med), flush=True)
        out["gpinv_gemm_nt_%d_tf" % N] = 2 * N**3 / best / 1e9
        best, med = timeit(lambda: ops.gemm(A, B, False, False, False), iters=8)
        print("gpinv gemm NN %d^3: best %.3f ms (%.2f TF)" % (N, best, 2 * N**3 / best / 1e9), flush=True)
        best, med = timeit(lambda: ops.gemm(A, B, True, False, False), iters=8)
        print("gpinv gemm TN %d^3: best %.3f ms (%.2f TF)" % (N, best, 2 * N**3 / best / 1e9), flush=True)
        del A, B
    # headline shapes
    n, M, S = 4096, 8192, 1024
    X = torch.randn(S, n, dtype=torch.float64, device=dev, generator=g)
    Am = torch.randn(M, n, dtype=torch.float64, device=dev, generator=g)
    Rs = torch.randn(S, M, dtype=torch.float64, device=dev, generator=g)
    best, _ = timeit(lambda: ops.gemm(X, Am, False, True, False))
    print("abel fwd shape [S,n]x[M,n]^T: %.3f ms (%.2f TF)" % (best, 2 * S * n * M / best / 1e9), flush=True)
    best, _ = timeit(lambda: ops.gemm(Rs, Am, False, False, False))
    print("abel bwd shape [S,M]x[M,n]: %.3f ms (%.2f TF)" % (best, 2 * S * n * M / best / 1e9), flush=True)
    best, _ = timeit(lambda: torch.matmul(X, Am.T))
    print("cublas abel fwd shape: %.3f ms (%.2f TF)" % (best, 2 * S * n * M / best / 1e9), flush=True)
    E = torch.randn(S, n, dtype=torch.float64, device=dev, generator=g)
    best, _ = timeit(lambda: ops.gemm(X, E, True, False, True))
    print("tril(X^T E) [n,n] K=S lower: %.3f ms (%.2f TF tri-credited)" % (best, n * n * S / best / 1e9), flush=True)
    # Cholesky
    for nn in (1024, 2048, 4096):
        x = torch.linspace(0, 1, nn, dtype=torch.float64, device=dev)[:, None]
        ls = torch.tensor([0.3], dtype=torch.float64, device=dev)
        K = ops.kcore(x, None, ls, 1, 1e-6)
        best, med = timeit(lambda: ops.potrf(K))
        print("gpinv potrf n=%d: %.3f ms (%.2f TF)  [incl. clone]" % (nn, best, nn**3 / 3 / best / 1e9), flush=True)
        out["potrf_%d_ms" % nn] = best
        bestc, _ = timeit(lambda: torch.linalg.cholesky(K))
        print("cusolver potrf n=%d: %.3f ms (%.2f TF)" % (nn, bestc, nn**3 / 3 / bestc / 1e9), flush=True)
        L, info = ops.potrf(K)
        Lbar = torch.tril(torch.randn(nn, nn, dtype=torch.float64, device=dev, generator=g))
        best, med = timeit(lambda: ops.potrf_bwd(L, Lbar))
        print("gpinv potrf_bwd n=%d: %.3f ms (%.2f TF)  [incl. clone]" % (nn, best, 2 * nn**3 / 3 / best / 1e9), flush=True)
        out["potrf_bwd_%d_ms" % nn] = best
        best, _ = timeit(lambda: ops.kcore(x, None, ls, 1, 1e-6))
        print("kcore csym n=%d: %.3f ms (%.1f GB/s written)" % (nn, best, 8 * nn * nn / best / 1e6), flush=True)
    # full sample + abel stage at cfg5
    Lc = L
    sv = torch.ones(1, dtype=torch.float64, device=dev)
    Q = torch.tril(0.5 * torch.eye(n, dtype=torch.float64, device=dev) + 0.05 / np.sqrt(n) * torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))[None].contiguous()
    qmu = 0.3 * torch.randn(1, n, dtype=torch.float64, device=dev, generator=g)
    mean = torch.zeros(1, n, dtype=torch.float64, device=dev)
    E3 = E[None].contiguous()
    best, _ = timeit(lambda: ops.stvgp_sample(Lc, sv, Q, qmu, mean, E3, False, True))
    print("stvgp_sample fwd cfg5: %.3f ms (%.2f TF tri-credited 2 n^2 S)" % (best, 2 * n * n * S / best / 1e9), flush=True)
    F, expF, kl, U = ops.stvgp_sample(Lc, sv, Q, qmu, mean, E3, False, True)
    Fbar = torch.randn_like(F)
    klbar = torch.tensor([-1.0 / S], dtype=torch.float64, device=dev)
    best, _ = timeit(lambda: ops.stvgp_sample_bwd(Lc, sv, Q, E3, U, Fbar, klbar, False))
    print("stvgp_sample bwd cfg5: %.3f ms (%.2f TF tri-credited 3 n^2 S)" % (best, 3 * n * n * S / best / 1e9), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
