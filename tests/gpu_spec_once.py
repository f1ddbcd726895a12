"""Forward + reverse of the spectroscopic likelihood kernels at the cfg3a shapes (64 radial points,
64 chords, 128 wavelengths, R = 3 latent functions, S = 256 draws) -- for ncu, and a timing line with the
FP64-pipe bound: the kernels evaluate one exp per (draw, wavelength, chord, radial point)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
dev = torch.device("cuda:0")
n, m, k, S = 64, 64, 128, 256
g = torch.Generator(device=dev).manual_seed(0)
F = 0.1 * torch.randn(3, S, n, dtype=torch.float64, device=dev, generator=g)
F[0] -= 2.0
At = torch.rand(n, m, dtype=torch.float64, device=dev, generator=g) * 0.05       # geometry, transposed [n, m]
cosTt = torch.rand(n, m, dtype=torch.float64, device=dev, generator=g) * 2 - 1
lam = torch.linspace(-3, 3, k, dtype=torch.float64, device=dev)
Y = torch.rand(k, m, dtype=torch.float64, device=dev, generator=g)
one = torch.ones(1, dtype=torch.float64, device=dev)
dl = float((lam[-1] - lam[0]).item() / (k - 1))       # uniform grid: recurrence kernels (0.0 -> one exp per term)
if len(sys.argv) > 1 and sys.argv[1] == "direct":
    dl = 0.0
for _ in range(3):
    ss, P = ops.spec_resid(F, At, cosTt, lam, Y, dl)
    Fbar = ops.spec_bwd(F, At, cosTt, lam, Y, P, one, dl)
torch.cuda.synchronize()


def timeit(fn, reps=50):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


tf = timeit(lambda: ops.spec_resid(F, At, cosTt, lam, Y, dl))
tb = timeit(lambda: ops.spec_bwd(F, At, cosTt, lam, Y, P, one, dl))
nexp = float(S) * k * m * n
print("dlam = %g:  spec fwd %.1f us  (%.2f G exp/s)   spec bwd %.1f us  (%.2f G exp-terms/s)   [%d x %d x %d x %d = %.3g exponentials]"
      % (dl, tf * 1e6, nexp / tf / 1e9, tb * 1e6, nexp / tb / 1e9, S, k, m, n, nexp))
print("ok", float(ss.item()), float(Fbar.abs().sum().item()))
