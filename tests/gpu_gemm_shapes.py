"""DMMA GEMM throughput on the shapes of the headline evaluation (CUDA events, best of 8)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def T(fn, flops, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("%-42s %8.3f ms %6.2f TF/s" % (name, min(ts), flops / min(ts) / 1e9), flush=True)
n, M, S = 4096, 8192, 1024
A = torch.randn(M, n, dtype=torch.float64, device=dev, generator=g)
X = torch.randn(S, n, dtype=torch.float64, device=dev, generator=g)
R = torch.randn(S, M, dtype=torch.float64, device=dev, generator=g)
Q = torch.randn(n, n, dtype=torch.float64, device=dev, generator=g)
T(lambda: torch.matmul(Q, Q), 2 * n**3, "cublas 4096^3")
T(lambda: ops.gemm(Q, Q, False, True, False), 2 * n**3, "gpinv NT 4096^3")
T(lambda: ops.gemm(Q, Q, False, False, False), 2 * n**3, "gpinv NN 4096^3")
T(lambda: ops.gemm(Q, Q, True, False, False), 2 * n**3, "gpinv TN 4096^3")
T(lambda: ops.gemm(X, A, False, True, False), 2 * S * n * M, "abel fwd [S,n]x[M,n]^T")
T(lambda: ops.gemm(R, A, False, False, False), 2 * S * n * M, "abel bwd [S,M]x[M,n]")
T(lambda: ops.gemm(X, Q, False, True, False), 2 * S * n * n, "sample [S,n]x[n,n]^T dense")
T(lambda: ops.gemm(X, X, True, False, True), S * n * n, "tril(X^T X) K=S (tri-credited)")
Q1 = Q[:1024, :1024].contiguous()
T(lambda: ops.gemm(Q1, Q1, False, True, False), 2 * 1024**3, "gpinv NT 1024^3")
T(lambda: ops.gemm(Q1, Q1, True, False, True), 1024**3, "tril(Q1^T Q1) 1024 (tri-credited)")
T(lambda: torch.matmul(Q1, Q1), 2 * 1024**3, "cublas 1024^3")
