"""Forward Cholesky timing (CUDA events, eager launches and graph replay) at n = 4096 (or argv[1])."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.linspace(0, 1, nn, dtype=torch.float64, device=dev)[:, None]
ls = torch.tensor([0.3], dtype=torch.float64, device=dev)
K = ops.kcore(x, None, ls, 1, 1e-6)
Lref = torch.linalg.cholesky(K)
L, info = ops.potrf(K)
print("info", int(info[0].item()), "rel.err vs cuSOLVER %.2e" % float((L - Lref).abs().max() / Lref.abs().max()),
      "residual %.2e" % float((L @ L.T - K).abs().max()))


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("eager   %.3f ms" % timeit(lambda: ops.potrf(K)))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    ops.potrf(K)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = ops.potrf(K)
print("graph   %.3f ms" % timeit(g.replay))
print("cusolver %.3f ms" % timeit(lambda: torch.linalg.cholesky(K)))
