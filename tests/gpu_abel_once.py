"""One forward + backward of the Abel likelihood stage at the headline shapes (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
n, M, S = 4096, 8192, 1024
g = torch.Generator(device=dev).manual_seed(0)
F = 0.1 * torch.randn(S, n, dtype=torch.float64, device=dev, generator=g)
A = torch.rand(M, n, dtype=torch.float64, device=dev, generator=g) * 0.01
Y = torch.randn(M, dtype=torch.float64, device=dev, generator=g)
expF = torch.exp(F)
for _ in range(3):
    ss, Rs = ops.abel_resid(F, expF, A, Y)
    Fbar = ops.abel_bwd(Rs, A, expF, torch.ones(1, dtype=torch.float64, device=dev))
torch.cuda.synchronize()
print("ok", float(ss.item()))
