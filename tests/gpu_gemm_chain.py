"""The chain of triangular products of the inverse-form Cholesky reverse, issued back to back
through the test GEMM entry, repeated: run-to-run reproducibility per stage (race detector)."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
split = int(os.environ.get("SPLIT", "1"))
sync = os.environ.get("SYNC") is not None


def gemm(la, lb, A, B, Cm, beta, lower, kmode):
    rc = lib.gpinv_gemm_ex_f64(la, lb, n, n, n, 1.0, P_(A), n, 0, P_(B), n, 0, beta, P_(Cm), n, 0, 1, int(lower), kmode, split, st)
    _capi.check(rc, "gemm_ex")
    if sync:
        torch.cuda.synchronize()


L = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
Lb = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
W = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g)) / n ** 0.5
ws = torch.empty(3, n, n, dtype=torch.float64, device=dev)
first = None
bad = [0, 0, 0]
for r in range(reps):
    ws.fill_(float(r))                      # different stale contents every repetition
    Pm, T, D = ws[0], ws[1], ws[2]
    gemm(1, 1, L, Lb, Pm, 0.0, True, 4)      # P = tril(L^T Lbar)
    gemm(0, 1, Pm, W, T, 0.0, True, 5)       # T = tril(P W)
    gemm(1, 1, W, T, D, 0.0, True, 4)        # D = tril(W^T T)
    gemm(1, 1, T, W, D, 1.0, True, 4)        # D += tril(T^T W)
    torch.cuda.synchronize()
    cur = [torch.tril(x).clone() for x in (Pm, T, D)]
    if first is None:
        first = cur
    else:
        for i in range(3):
            if not torch.equal(cur[i], first[i]):
                bad[i] += 1
                if bad[i] == 1:
                    Dm = (cur[i] - first[i]).abs()
                    nb = -(-n // 128)
                    Dp = torch.zeros(nb * 128, nb * 128, dtype=Dm.dtype, device=dev)
                    Dp[:n, :n] = Dm
                    Bm = Dp.reshape(nb, 128, nb, 128).amax(dim=(1, 3))
                    print("stage %d rep %d differs: blocks %s max %.3e" % (i, r, (Bm > 0).nonzero().cpu().tolist()[:10], float(Dm.max())), flush=True)
print("n=%d split=%d sync=%s reps=%d: differing reps per stage [P, T, D] = %s" % (n, split, sync, reps, bad))
