"""Repeat potrf / potrf_bwd at awkward sizes and report run-to-run differences (race detector)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import ops
from tests.gpu_selfcheck import spd, t
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for n in (2500, 3000, 4096):
    K = t(spd(n, n))
    Lbar = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
    L0, _ = ops.potrf(K)
    G0 = ops.potrf_bwd(L0, Lbar)
    torch.cuda.synchronize()
    dl, dg = [], []
    shown = False
    for r in range(reps):
        L, info = ops.potrf(K)
        G = ops.potrf_bwd(L0, Lbar)
        dl.append(float((L - L0).abs().max()))
        dg.append(float((G - G0).abs().max() / G0.abs().max()))
        if dg[-1] > 1e-13 and not shown:
            shown = True
            D = (G - G0).abs()
            blk = 128
            nb = -(-n // blk)
            Dp = torch.zeros(nb * blk, nb * blk, dtype=D.dtype, device=D.device)
            Dp[:n, :n] = D
            Bm = Dp.reshape(nb, blk, nb, blk).amax(dim=(1, 3))
            badb = (Bm > 1e-13 * float(G0.abs().max())).nonzero().cpu().tolist()
            print("   n=%d rep %d: %d of %d 128-blocks differ; first: %s ... last: %s" % (n, r, len(badb), nb * nb, badb[:12], badb[-6:]), flush=True)
    print("n=%d  potrf max|L-L0| over %d reps: %.3e (nonzero in %d)   potrf_bwd rel diff: %.3e (nonzero in %d)"
          % (n, reps, max(dl), sum(d > 0 for d in dl), max(dg), sum(d > 1e-13 for d in dg)), flush=True)
