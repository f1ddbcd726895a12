"""chol_core with the interleaved inversion: correctness of L and W = L^-1, and the time of
factorisation + inversion (CUDA events; eager and graph replay) against factorisation alone."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.linspace(0, 1, nn, dtype=torch.float64, device=dev)[:, None]
ls = torch.tensor([0.3], dtype=torch.float64, device=dev)


def both():
    L, info, W = ops.chol_core(x, ls, 1, 1e-6, True)
    ops.trtri_side_join()
    return L, info, W


L, info, W = both()
torch.cuda.synchronize()
K = ops.kcore(x, None, ls, 1, 1e-6)
Lref = torch.linalg.cholesky(K)
Wref = torch.linalg.solve_triangular(Lref, torch.eye(nn, dtype=torch.float64, device=dev), upper=False)
I = torch.eye(nn, dtype=torch.float64, device=dev)
print("n=%d info %d | L rel.err %.2e | W rel.err vs solve_triangular %.2e | |W L - I| %.2e (reference: %.2e) | upper(W) max %.1e"
      % (nn, int(info[0].item()), float((L - Lref).abs().max() / Lref.abs().max()),
         float((W[0] - Wref).abs().max() / Wref.abs().max()), float((W[0] @ L - I).abs().max()),
         float((Wref @ Lref - I).abs().max()), float(torch.triu(W[0], 1).abs().max())))


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


print("factor only            eager %.3f ms" % timeit(lambda: ops.chol_core(x, ls, 1, 1e-6, False)))
print("factor + inverse+join  eager %.3f ms" % timeit(both))
for fn, name in ((lambda: ops.chol_core(x, ls, 1, 1e-6, False), "factor only"), (both, "factor + inverse+join")):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    print("%-22s graph %.3f ms" % (name, timeit(g.replay)))
