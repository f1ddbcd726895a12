import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "both"
x = torch.linspace(0, 1, nn, dtype=torch.float64, device=dev)[:, None]
ls = torch.tensor([0.3], dtype=torch.float64, device=dev)
K = ops.kcore(x, None, ls, 1, 1e-6)
L, info = ops.potrf(K)
if what != "fwd":
    Lbar = torch.tril(torch.randn(nn, nn, dtype=torch.float64, device=dev))
    G = ops.potrf_bwd(L, Lbar)
torch.cuda.synchronize()
print("ok", int(info[0].item()))
