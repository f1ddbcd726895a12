"""In-place windowed GEMM updates as the blocked Cholesky issues them (C ABI gpinv_gemm_f64 on
sub-matrix windows of one n x n buffer, beta = 1, lower output), against torch."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpinv_b200 import _capi, ops
dev = torch.device("cuda:0")
ops.ensure_workspace(dev)
lib = _capi.lib()
g = torch.Generator(device=dev).manual_seed(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
bad = 0


def ptr(t, off):
    return C.c_void_p(t.data_ptr() + 8 * off)


def run(n, o, w, t0, ncols, lower, tag):
    """A[t0:, t0:t0+ncols] -= A[t0:, o:o+w] A[t0:t0+ncols, o:o+w]^T   (lower: ncols = n - t0)"""
    global bad
    A = torch.randn(n, n, dtype=torch.float64, device=dev, generator=g)
    ref = A.clone()
    A0 = A.clone()
    P = ref[t0:, o:o + w]
    upd = ref[t0:, t0:t0 + ncols] - P @ ref[t0:t0 + ncols, o:o + w].T
    if lower:
        ref[t0:, t0:] = torch.where(torch.tril(torch.ones_like(upd)) > 0, upd, ref[t0:, t0:])
    else:
        ref[t0:, t0:t0 + ncols] = upd
    rc = lib.gpinv_gemm_f64(0, 1, n - t0, ncols, w, -1.0, ptr(A, t0 * n + o), n, ptr(A, t0 * n + o), n, 1.0,
                            ptr(A, t0 * n + t0), n, int(lower), st)
    _capi.check(rc, "gemm")
    torch.cuda.synchronize()
    if lower:   # strict upper parts of diagonal tiles are zeroed by the kernel: compare the lower triangle
        d = (torch.tril(A[t0:, t0:]) - torch.tril(ref[t0:, t0:])).abs().max().item()
        d = max(d, (A[:t0] - ref[:t0]).abs().max().item() if t0 else 0.0, (A[t0:, :t0] - ref[t0:, :t0]).abs().max().item())
    else:
        d = (A - ref).abs().max().item()
    ok = d < 1e-10
    bad += (not ok)
    if not ok:
        D = (torch.tril(A[t0:, t0:]) - torch.tril(ref[t0:, t0:])).abs() if lower else (A - ref)[t0:, t0:t0 + ncols].abs()
        m = D.shape[0]
        blk = 64
        nb0, nb1 = -(-D.shape[0] // blk), -(-D.shape[1] // blk)
        Dp = torch.zeros(nb0 * blk, nb1 * blk, dtype=D.dtype, device=D.device)
        Dp[:D.shape[0], :D.shape[1]] = D
        B = Dp.reshape(nb0, blk, nb1, blk).amax(dim=(1, 3))
        badb = (B > 1e-10).nonzero().cpu().tolist()
        print("   bad 64x64 blocks (%d of %d): %s" % (len(badb), nb0 * (nb0 + 1) // 2 if lower else nb0 * nb1, badb[:40]), flush=True)
        r, c = badb[0]
        if lower:
            W0 = A0[t0:, t0:]
            PP = (A0[t0:, o:o + w] @ A0[t0:t0 + ncols, o:o + w].T)
            sub0 = D[r * blk:(r + 1) * blk, c * blk:(c + 1) * blk]
            ij = (sub0 > 1e-10).nonzero()[0].cpu().tolist()
            i, j = r * blk + ij[0], c * blk + ij[1]
            got = A[t0 + i, t0 + j].item()
            print("   element (%d,%d): got %.6f  expected %.6f  Cin %.6f  PP^T %.6f  got-Cin %.6f" %
                  (i, j, got, ref[t0 + i, t0 + j].item(), W0[i, j].item(), PP[i, j].item(), got - W0[i, j].item()))
            # is the wrong accumulator a dot product of OTHER rows?  search PP for the value
            val = W0[i, j].item() - got
            close = ((PP - val).abs() < 1e-9).nonzero().cpu().tolist()
            print("   -(got-Cin) = %.6f matches P P^T at positions %s" % (val, close[:8]))
        sub = D[r * blk:(r + 1) * blk, c * blk:(c + 1) * blk]
        print("   first bad block: nonzero rows", (sub.amax(1) > 1e-10).nonzero().flatten().cpu().tolist()[:70])
        print("   first bad block: nonzero cols", (sub.amax(0) > 1e-10).nonzero().flatten().cpu().tolist()[:70])
    print("%-40s n=%d o=%d w=%d t0=%d ncols=%d max|diff|=%.3e %s" % (tag, n, o, w, t0, ncols, d, "ok" if ok else "FAIL"), flush=True)


def fresh(m, k, tag):
    """not in place, beta = 0: tril(P P^T) into a new buffer"""
    global bad
    P = torch.randn(m, k, dtype=torch.float64, device=dev, generator=g)
    out = ops.gemm(P, P, False, True, True)
    d = (torch.tril(out) - torch.tril(P @ P.T)).abs().max().item()
    ok = d < 1e-10
    bad += (not ok)
    print("%-40s m=%d k=%d max|diff|=%.3e %s" % (tag, m, k, d, "ok" if ok else "FAIL"), flush=True)


if len(sys.argv) > 1:
    fresh(1792, 256, "fresh lower (406 small CTAs)")
    fresh(1792, 256, "fresh lower (406 small CTAs) again")
    for n in (2048,):
        run(n, 0, 256, 256, n - 256, True, "trailing lower update")
        run(n, 0, 256, 256, n - 256, True, "trailing lower update again")
else:
    for n in (1024, 2048, 2500, 3072, 4096):
        run(n, 0, 256, 256, n - 256, True, "trailing lower update")
        run(n, 0, 256, 256, 256, False, "next-panel update")
        run(n, 0, 64, 64, 192, False, "K=64 panel update")
        run(n, 512, 256, 768, n - 768, True, "trailing lower update (o=512)")
        run(n, 256, 64, 320, 192, False, "K=64 panel update (i=256)")
print("failures:", bad)
sys.exit(1 if bad else 0)
