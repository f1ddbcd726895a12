"""PyTorch custom ops (``torch.ops.gpinv.*``) over the C ABI of ``libgpinv_sm100.so``.

Each op is a thin wrapper: it allocates outputs with torch (device memory and streams are
torch's job), then calls one ``gpinv_*_f64`` entry point on the current CUDA stream.  The
backward of every op is hand-written CUDA too (registered with ``register_autograd``) — the
reference obtains it from ``tf.gradients`` (pattern at GPinv/model.py:35-37).

CUDA only.  There is no CPU implementation: calling an op with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _capi

MODE = {"rbf": 0, "rbf_csym": 1, "rbf_casym": 2}


def _p(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError("gpinv_b200 ops are CUDA-only (no CPU fallback): %s is on %s" % (name, t.device))
    if t.dtype != torch.float64:
        raise RuntimeError("gpinv_b200 ops are fp64-only: %s is %s" % (name, t.dtype))
    return t.contiguous()


_SPLIT_WS = {}

# bench.py sets this to a dict to time single ops inside an eager evaluation: name -> [(start, end)]
# CUDA events recorded on the launching stream.  None (default): no events are recorded.
TIMERS = None


class _timed(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMERS is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def __exit__(self, *exc):
        if TIMERS is not None:
            self.ev[1].record()
            TIMERS.setdefault(self.name, []).append(self.ev)
        return False


def ensure_workspace(device, nbytes: int = 320 << 20):
    """Register the (zero-initialised, persistent) split-K workspace for `device` once."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _SPLIT_WS:
        with torch.cuda.device(idx):
            buf = torch.zeros(nbytes // 8, dtype=torch.float64, device="cuda")
            torch.cuda.current_stream().synchronize()
            _capi.check(_capi.lib().gpinv_set_workspace(_p(buf), buf.numel() * 8), "set_workspace")
        _SPLIT_WS[idx] = buf
    return _SPLIT_WS[idx]


def _ws(nbytes: int, like: Tensor) -> Tensor:
    return torch.empty(max(int(nbytes) // 8 + 2, 2), dtype=torch.float64, device=like.device)


# --------------------------------------------------------------------------------------
# T1-T3 kernel core
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::kcore", mutates_args=(), device_types="cuda")
def kcore(X: Tensor, X2: Optional[Tensor], ls: Tensor, mode: int, jitter: float) -> Tensor:
    """core(X, X2) [n,n2] (+ jitter on the diagonal when X2 is None).
    GPinv/kernels.py:57-58,105-107,115-120,140-145."""
    X = _chk(X, "X")
    ls = _chk(ls, "ls").reshape(-1)
    n, D = X.shape
    n2 = n
    if X2 is not None:
        X2 = _chk(X2, "X2")
        n2 = X2.shape[0]
    ard = 1 if ls.numel() > 1 else 0
    K = torch.empty((n, n2), dtype=torch.float64, device=X.device)
    L = _capi.lib()
    ws = _ws(L.gpinv_kbuild_ws_bytes(n, n2, D), X)
    rc = L.gpinv_kbuild_f64(_p(X), _p(X2), n, n2, D, D, D, _p(ls), ard, mode, float(jitter),
                            _p(K), n2, _p(ws), _st())
    _capi.check(rc, "kbuild")
    return K


@torch.library.custom_op("gpinv::kcore_bwd", mutates_args=(), device_types="cuda")
def kcore_bwd(X: Tensor, X2: Optional[Tensor], ls: Tensor, mode: int, Kbar: Tensor) -> Tensor:
    X = _chk(X, "X")
    ls = _chk(ls, "ls").reshape(-1)
    Kbar = _chk(Kbar, "Kbar")
    n, D = X.shape
    n2 = n
    if X2 is not None:
        X2 = _chk(X2, "X2")
        n2 = X2.shape[0]
    ard = 1 if ls.numel() > 1 else 0
    out = torch.empty_like(ls)
    L = _capi.lib()
    ws = _ws(L.gpinv_kbuild_ws_bytes(n, n2, D), X)
    rc = L.gpinv_kbuild_bwd_f64(_p(X), _p(X2), n, n2, D, D, D, _p(ls), ard, mode, _p(Kbar), n2,
                                _p(out), _p(ws), _st())
    _capi.check(rc, "kbuild_bwd")
    return out


def _kcore_setup(ctx, inputs, output):
    X, X2, ls, mode, jitter = inputs
    ctx.save_for_backward(X, X2, ls) if X2 is not None else ctx.save_for_backward(X, ls)
    ctx.has_x2 = X2 is not None
    ctx.mode = mode
    ctx.ls_shape = ls.shape


def _kcore_backward(ctx, Kbar):
    if ctx.has_x2:
        X, X2, ls = ctx.saved_tensors
    else:
        (X, ls), X2 = ctx.saved_tensors, None
    g = kcore_bwd(X, X2, ls, ctx.mode, Kbar.contiguous())
    return None, None, g.reshape(ctx.ls_shape), None, None


kcore.register_autograd(_kcore_backward, setup_context=_kcore_setup)


# --------------------------------------------------------------------------------------
# T4 Cholesky
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::potrf", mutates_args=(), device_types="cuda")
def potrf(K: Tensor) -> Tuple[Tensor, Tensor]:
    """(L, info): lower Cholesky factor with zeroed strict upper triangle; info[0] (int32) is the
    1-based first failing pivot (0 = success).  GPinv/kernels.py:59."""
    K = _chk(K, "K")
    n = K.shape[0]
    Lm = K.clone()
    info = torch.zeros(2, dtype=torch.int32, device=K.device)   # [status, scratch ticket]
    rc = _capi.lib().gpinv_potrf_f64(_p(Lm), n, n, _p(info), _st())
    _capi.check(rc, "potrf")
    return Lm, info


@torch.library.custom_op("gpinv::potrf_bwd", mutates_args=(), device_types="cuda")
def potrf_bwd(L: Tensor, Lbar: Tensor) -> Tensor:
    """Symmetric dObj/dK from dObj/dL (blocked reverse sweep)."""
    L = _chk(L, "L")
    ensure_workspace(L.device)
    Ab = _chk(Lbar, "Lbar").clone()
    n = L.shape[0]
    lib = _capi.lib()
    ws = _ws(lib.gpinv_potrf_bwd_ws_bytes(n), L)
    rc = lib.gpinv_potrf_bwd_f64(_p(L), n, _p(Ab), n, n, _p(ws), _st())
    _capi.check(rc, "potrf_bwd")
    return Ab


def _potrf_setup(ctx, inputs, output):
    ctx.save_for_backward(output[0])
    ctx.set_materialize_grads(False)      # no zero tensors for the gradients of outputs nobody differentiates


def _potrf_backward(ctx, Lbar, _info_bar):
    (L,) = ctx.saved_tensors
    if Lbar is None:
        return None
    return potrf_bwd(L, Lbar.contiguous())


potrf.register_autograd(_potrf_backward, setup_context=_potrf_setup)


# --------------------------------------------------------------------------------------
# T1-T4 fused: chol(core(X,X) + jitter I) with the lower-triangle-only build and reverse
# --------------------------------------------------------------------------------------
# Explicit inverses built on the library's side stream (see chol_core): the last few are kept alive here
# so that their memory cannot be handed out again while the side stream may still be writing it (a forward
# pass whose result is dropped without a reverse pass never joins the side stream); an entry is dropped
# only after a host-side wait for the side stream.
_W_KEEP = []
# True from the launch of a side-stream inversion until a Python-level ``trtri_side_join`` (a capture that is
# about to end must join the side stream it forked, and must not wait on an inversion of an earlier capture)
SIDE_PENDING = [False]


@torch.library.custom_op("gpinv::chol_core", mutates_args=(), device_types="cuda")
def chol_core(X: Tensor, ls: Tensor, mode: int, jitter: float, with_inverse: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """(L, info, W) with L = Cholesky factor of core(X,X) + jitter*I (GPinv/kernels.py:56-59): only the
    lower triangle of the core is evaluated and it is factored in place.  ``with_inverse``: W = L^-1 is
    built on a side stream right behind the factorisation -- the reverse pass needs it, it depends on L
    only, so it overlaps the sampling / likelihood stages that follow -- else W is empty."""
    X = _chk(X, "X")
    ls = _chk(ls, "ls").reshape(-1)
    ensure_workspace(X.device)
    n, D = X.shape
    ard = 1 if ls.numel() > 1 else 0
    Lm = torch.empty((n, n), dtype=torch.float64, device=X.device)
    info = torch.zeros(2, dtype=torch.int32, device=X.device)
    lib = _capi.lib()
    ws = _ws(lib.gpinv_chol_core_ws_bytes(n, D), X)
    if with_inverse and n > 64:
        # W[0] = L^-1, W[1] = scratch of the inversion: ONE tensor, returned and saved for the reverse pass,
        # so that neither half can be recycled (by the caching allocator, or by a captured graph's pool)
        # before the reverse pass has joined the side stream
        W = torch.empty((2, n, n), dtype=torch.float64, device=X.device)
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing and len(_W_KEEP) >= 4:
            _capi.check(lib.gpinv_trtri_side_sync(), "trtri_side_sync")      # idle unless a result was dropped
            del _W_KEEP[0]
        if INVERSE_LATER is not None or INVERSE_SHARD is not None:
            rc = lib.gpinv_chol_core_f64(_p(X), n, D, D, _p(ls), ard, mode, float(jitter), _p(Lm), n, _p(info),
                                         _p(ws), _st())
            _capi.check(rc, "chol_core")
            if INVERSE_LATER is not None:
                INVERSE_LATER.append((Lm, W))    # the caller starts the inversion itself (start_inverse)
            else:
                start_inverse(Lm, W)
        else:
            # factorisation with the inversion interleaved (row blocks of W on a low-priority stream of the library)
            rc = lib.gpinv_chol_core_inv_f64(_p(X), n, D, D, _p(ls), ard, mode, float(jitter), _p(Lm), n, _p(info),
                                             _p(W), C.c_void_p(W.data_ptr() + 8 * n * n), _p(ws), _st())
            _capi.check(rc, "chol_core_inv")
            SIDE_PENDING[0] = True
        if not capturing:       # (inside a capture the tensor lives until the captured reverse pass)
            _W_KEEP.append(W)
    else:
        rc = lib.gpinv_chol_core_f64(_p(X), n, D, D, _p(ls), ard, mode, float(jitter), _p(Lm), n, _p(info),
                                     _p(ws), _st())
        _capi.check(rc, "chol_core")
        W = torch.empty(0, dtype=torch.float64, device=X.device)
    return Lm, info, W


@torch.library.custom_op("gpinv::chol_core_bwd", mutates_args=(), device_types="cuda")
def chol_core_bwd(X: Tensor, ls: Tensor, mode: int, L: Tensor, Lbar: Tensor, W: Tensor, w_ready: bool = False) -> Tensor:
    """dObj/d(lengthscales) from dObj/dL: Cholesky reverse (inverse formula, W = L^-1 from the forward
    pass when available) kept as a lower triangle and reduced straight into the lengthscale gradient
    (no symmetric n x n gradient is formed).  ``w_ready``: the caller has already ordered the stream after the
    side-stream inversion (``trtri_side_join``)."""
    X = _chk(X, "X")
    ls = _chk(ls, "ls").reshape(-1)
    L = _chk(L, "L")
    Lbar = _chk(Lbar, "Lbar")
    ensure_workspace(X.device)
    n, D = X.shape
    ard = 1 if ls.numel() > 1 else 0
    out = torch.empty_like(ls)
    lib = _capi.lib()
    ws = _ws(lib.gpinv_chol_core_bwd_ws_bytes(n, D), X)
    rc = lib.gpinv_chol_core_bwd2_f64(_p(X), n, D, D, _p(ls), ard, mode, _p(L), n, _p(Lbar), n,
                                      _p(W) if W.numel() else None, 1 if w_ready else 0, _p(out), _p(ws), _st())
    _capi.check(rc, "chol_core_bwd")
    return out


def _chol_core_setup(ctx, inputs, output):
    X, ls, mode, jitter, with_inverse = inputs
    ctx.save_for_backward(X, ls, output[0], output[2])
    ctx.mode = mode
    ctx.ls_shape = ls.shape
    # (materialised, the never-used gradient of W alone was a 268 MB zero fill per evaluation at n = 4096)
    ctx.set_materialize_grads(False)


class DeferredReverse(object):
    """Collector for the sample-sharded evaluation (gpinv_b200/dist.py): while one is installed with
    ``defer_reverse``, the reverse pass of every factorisation it accepts is NOT run where autograd reaches
    it -- (X, ls, mode, L, W, Lbar) is recorded instead and the lengthscales receive no gradient -- so that
    the caller can exchange Lbar between the ranks first and run ``chol_core_bwd_shard`` afterwards."""

    def __init__(self, accepts):
        self.accepts = accepts
        self.items = []


_DEFER = None


def defer_reverse(collector):
    """Install (or, with None, remove) the collector; returns the previous one."""
    global _DEFER
    prev, _DEFER = _DEFER, collector
    return prev


def _chol_core_backward(ctx, Lbar, _info_bar, _w_bar):
    X, ls, L, W = ctx.saved_tensors
    if Lbar is None:
        return None, None, None, None, None
    d = _DEFER
    if d is not None and W.numel() and d.accepts(L.shape[0]):
        d.items.append((X, ls, ctx.mode, L, W, Lbar.contiguous()))
        return None, None, None, None, None
    # (the C call joins the side stream unless a Python-level join has already been made -- e.g. at the end of
    # the CUDA graph that holds the forward pass, whose events a later graph cannot wait on)
    g = chol_core_bwd(X, ls, ctx.mode, L, Lbar.contiguous(), W, not SIDE_PENDING[0])
    SIDE_PENDING[0] = False
    return None, g.reshape(ctx.ls_shape), None, None, None


# None, or a list: chol_core(with_inverse=True) then only allocates W and records (L, W) here; the caller
# launches the inversion later with start_inverse (the graph-replayed host pipeline starts it in the graph
# that holds the sampling stage, so that it runs beside that stage instead of in front of the upload wait)
INVERSE_LATER = None


# None, or (rank, G) under sample sharding: the last merge level of every shardable inversion (three quarters of
# its flops) is computed for this rank's column slices only; the caller all-gathers them (Model._shard_*).
INVERSE_SHARD = None


def inverse_shardable(n: int, G: int) -> bool:
    """n = 64 * 2^k, large enough, and the slices of the last level are TMA-aligned."""
    return G > 1 and n >= 1024 and (n & (n - 1)) == 0 and (n // 2) % (4 * G) == 0


def start_inverse(Lm: Tensor, W: Tensor):
    """W[0] <- L^-1 on the library's side stream, behind the work already enqueued on the current stream."""
    n = Lm.shape[0]
    lib = _capi.lib()
    scratch = C.c_void_p(W.data_ptr() + 8 * n * n)
    if INVERSE_SHARD is not None and inverse_shardable(n, INVERSE_SHARD[1]):
        _capi.check(lib.gpinv_trtri_side_shard_f64(_p(Lm), n, _p(W), scratch, n, INVERSE_SHARD[0], INVERSE_SHARD[1],
                                                   _st()), "trtri_side_shard")
    else:
        _capi.check(lib.gpinv_trtri_side_f64(_p(Lm), n, _p(W), scratch, n, _st()), "trtri_side")
    SIDE_PENDING[0] = True


def inverse_slices_pack(W: Tensor, rank: int, G: int) -> Tensor:
    """This rank's two column slices of the bottom-left block of W[0] = L^-1 (see INVERSE_SHARD), [2, n/2, n/4G]."""
    n = W.shape[1]
    s, w = n // 2, n // (4 * G)
    blk = W[0][s:, :s]
    return torch.stack([blk[:, rank * w:(rank + 1) * w], blk[:, (2 * G - 1 - rank) * w:(2 * G - rank) * w]])


def inverse_slices_unpack_(W: Tensor, gathered: Tensor, G: int) -> None:
    """Write the all-gathered slices [G, 2, n/2, n/4G] into the bottom-left block of W[0]."""
    n = W.shape[1]
    s, w = n // 2, n // (4 * G)
    blk = W[0][s:, :s]
    src = gathered.reshape(G, 2, s, w)
    for g in range(G):                     # (plain strided copies: nothing here needs a host-built index tensor,
        for t in (0, 1):                   #  so it can sit inside a CUDA-graph capture)
            i = g if t == 0 else 2 * G - 1 - g
            blk[:, i * w:(i + 1) * w].copy_(src[g, t])


def trtri_side_join():
    """Order the current stream after the side-stream inversion started by ``chol_core(with_inverse=True)``."""
    _capi.check(_capi.lib().gpinv_trtri_side_join(_st()), "trtri_side_join")
    SIDE_PENDING[0] = False


def lbar_pack(Lbar: Tensor, G: int) -> Tensor:
    """tril(Lbar) regrouped into the reduce-scatter input of the column-block sharded reverse: G chunks of
    (n + w) w entries, w = n / 2G, chunk r = column blocks r and 2G-1-r from their diagonal down."""
    Lbar = _chk(Lbar, "Lbar")
    n = Lbar.shape[0]
    w = n // (2 * G)
    out = torch.empty(G * (n + w) * w, dtype=torch.float64, device=Lbar.device)
    _capi.check(_capi.lib().gpinv_lbar_pack_f64(_p(Lbar), n, n, w, G, _p(out), _st()), "lbar_pack")
    return out


def chol_core_bwd_shard(X: Tensor, ls: Tensor, mode: int, L: Tensor, W: Tensor, Dsh: Tensor, w: int,
                        c0s: List[int]) -> Tensor:
    """This rank's share of dObj/d(lengthscales): the reverse of its column blocks (first columns ``c0s``,
    width ``w``) of the rank-summed tril(Lbar), ``Dsh`` = its reduce-scatter output.  The caller has joined
    the side stream that built W = L^-1 (``trtri_side_join``)."""
    X = _chk(X, "X")
    ls = _chk(ls, "ls").reshape(-1)
    L = _chk(L, "L")
    Dsh = _chk(Dsh, "Dsh")
    n, D = X.shape
    ard = 1 if ls.numel() > 1 else 0
    out = torch.empty_like(ls)
    lib = _capi.lib()
    ws = _ws(lib.gpinv_chol_core_bwd_ws_bytes(n, D), X)
    c0 = (C.c_int * len(c0s))(*c0s)
    rc = lib.gpinv_chol_core_bwd_shard_f64(_p(X), n, D, D, _p(ls), ard, mode, _p(L), n, _p(W), _p(Dsh), w, c0,
                                           len(c0s), _p(out), _p(ws), _st())
    _capi.check(rc, "chol_core_bwd_shard")
    return out


def tri_compact(M: Tensor) -> Tensor:
    """Lower triangle of the n x n matrix ``M`` as n (n + 1) / 2 packed entries."""
    M = _chk(M, "M")
    n = M.shape[0]
    out = torch.empty(n * (n + 1) // 2, dtype=torch.float64, device=M.device)
    _capi.check(_capi.lib().gpinv_tri_compact_f64(_p(M), n, n, _p(out), _st()), "tri_compact")
    return out


def tri_expand_(src: Tensor, M: Tensor) -> Tensor:
    """Write the packed lower triangle ``src`` into the lower triangle of ``M`` (the rest is left alone)."""
    n = M.shape[0]
    _capi.check(_capi.lib().gpinv_tri_expand_f64(_p(src), _p(M), n, n, _st()), "tri_expand")
    return M


chol_core.register_autograd(_chol_core_backward, setup_context=_chol_core_setup)


@torch.library.custom_op("gpinv::trsm_rlt", mutates_args=(), device_types="cuda")
def trsm_rlt(L: Tensor, B: Tensor) -> Tensor:
    """X = B L^{-T} (B is [m,n]); i.e. every row x solves L x = b.  GPinv/conditionals.py:44."""
    L = _chk(L, "L")
    X = _chk(B, "B").clone()
    m, n = X.shape
    rc = _capi.lib().gpinv_trsm_rlt_f64(_p(L), n, _p(X), n, m, n, _st())
    _capi.check(rc, "trsm_rlt")
    return X


@torch.library.custom_op("gpinv::gemm", mutates_args=(), device_types="cuda")
def gemm(A: Tensor, B: Tensor, transa: bool, transb: bool, lower: bool) -> Tensor:
    """op(A) @ op(B) on the DMMA kernel (used by the prediction path, GPinv/conditionals.py)."""
    A = _chk(A, "A")
    B = _chk(B, "B")
    m, k = (A.shape[1], A.shape[0]) if transa else (A.shape[0], A.shape[1])
    n = B.shape[0] if transb else B.shape[1]
    Cm = torch.zeros((m, n), dtype=torch.float64, device=A.device)
    rc = _capi.lib().gpinv_gemm_f64(int(transa), int(transb), m, n, k, 1.0, _p(A), A.shape[1],
                                    _p(B), B.shape[1], 0.0, _p(Cm), n, int(lower), _st())
    _capi.check(rc, "gemm")
    return Cm


# --------------------------------------------------------------------------------------
# T7 tril(q_sqrt): [n,n,R] -> [R,n,n]
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::tril_unpack", mutates_args=(), device_types="cuda")
def tril_unpack(q_sqrt: Tensor) -> Tensor:
    q = _chk(q_sqrt, "q_sqrt")
    n, _, R = q.shape
    out = torch.empty((R, n, n), dtype=torch.float64, device=q.device)
    _capi.check(_capi.lib().gpinv_tril_unpack_f64(_p(q), _p(out), n, R, _st()), "tril_unpack")
    return out


@torch.library.custom_op("gpinv::tril_pack", mutates_args=(), device_types="cuda")
def tril_pack(Qbar: Tensor) -> Tensor:
    q = _chk(Qbar, "Qbar")
    R, n, _ = q.shape
    out = torch.empty((n, n, R), dtype=torch.float64, device=q.device)
    _capi.check(_capi.lib().gpinv_tril_pack_f64(_p(q), _p(out), n, R, _st()), "tril_pack")
    return out


tril_unpack.register_autograd(lambda ctx, g: tril_pack(g.contiguous()), setup_context=lambda ctx, inputs, output: None)


# --------------------------------------------------------------------------------------
# T8-T13 sampling
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::stvgp_sample", mutates_args=(), device_types="cuda")
def stvgp_sample(Lc: Tensor, sqrt_v: Tensor, Q: Tensor, q_mu: Tensor, mean: Tensor, E: Tensor,
                 q_diag: bool, want_exp: bool, U_pre: Optional[Tensor] = None,
                 kl_pre: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(F [R,S,n], expF [R,S,n] or empty, KL [1], U [R,S,n]).  GPinv/stvgp.py:144-185.
    Lc [n,n] lower core factor, sqrt_v [R], Q [R,n,n] lower (or [R,n] if q_diag), q_mu [R,n],
    mean [R,n], E [R,S,n].  ``U_pre`` / ``kl_pre``: the first half of the stage was already run by
    ``sample_u_early`` (beside the Cholesky factorisation); the U output is then empty and ``U_pre`` is what the
    reverse pass uses."""
    Lc = _chk(Lc, "Lc"); sqrt_v = _chk(sqrt_v, "sqrt_v"); Q = _chk(Q, "Q")
    q_mu = _chk(q_mu, "q_mu"); mean = _chk(mean, "mean"); E = _chk(E, "E")
    ensure_workspace(E.device)
    R, S, n = E.shape
    lib = _capi.lib()
    early = U_pre is not None and kl_pre is not None
    if early:
        _capi.check(lib.gpinv_side_early_join(_st()), "side_early_join")       # before kl_pre is read below
        U = U_pre
        kl = kl_pre.clone()
    else:
        U = torch.empty_like(E)
        kl = torch.empty(4, dtype=torch.float64, device=E.device)
    F = torch.empty_like(E)
    expF = torch.empty_like(E) if want_exp else torch.empty(0, dtype=torch.float64, device=E.device)
    rc = lib.gpinv_sample_fwd2_f64(_p(Lc), n, _p(sqrt_v), _p(Q), int(q_diag), _p(q_mu),
                                   _p(mean), _p(E), n, S, R, _p(U), _p(F),
                                   _p(expF) if want_exp else None, _p(kl), 1 if early else 0, _st())
    _capi.check(rc, "sample_fwd")
    return F, expF, kl[:1].clone(), (torch.empty(0, dtype=torch.float64, device=E.device) if early else U)


def mark_inputs_ready() -> None:
    """Mark the point of the current stream where the inputs of ``sample_u_early`` exist (call before chol_core)."""
    _capi.check(_capi.lib().gpinv_side_mark_inputs(_st()), "side_mark_inputs")


def sample_u_early(Q: Tensor, q_mu: Tensor, E: Tensor, q_diag: bool, U: Tensor, kl: Tensor, max_ctas: int = 0) -> None:
    """U [R,S,n] <- E tril(Q)^T + q_mu and kl [4] <- the KL partial sums, on the library's low-priority stream,
    behind ``mark_inputs_ready`` and the middle of the most recent Cholesky factorisation; no autograd -- the
    results are handed to ``stvgp_sample`` (U_pre, kl_pre), whose reverse pass is unchanged.  U and kl are
    allocated by the caller BEFORE ``mark_inputs_ready`` (nothing enqueued later may still own their memory)."""
    import os
    Q = _chk(Q, "Q"); q_mu = _chk(q_mu, "q_mu"); E = _chk(E, "E")
    ensure_workspace(E.device)
    R, S, n = E.shape
    if max_ctas <= 0:
        max_ctas = int(os.environ.get("GPINV_EARLY_U_CAP", "100"))
    rc = _capi.lib().gpinv_sample_u_early_f64(_p(Q), int(q_diag), _p(q_mu), _p(E), n, S, R, _p(U), _p(kl), int(max_ctas))
    _capi.check(rc, "sample_u_early")


@torch.library.custom_op("gpinv::stvgp_sample_bwd", mutates_args=(), device_types="cuda")
def stvgp_sample_bwd(Lc: Tensor, sqrt_v: Tensor, Q: Tensor, E: Tensor, U: Tensor, Fbar: Tensor,
                     klbar: Optional[Tensor], q_diag: bool) -> List[Tensor]:
    """[Lc_bar, sqrt_v_bar, Qbar, qmu_bar, mean_bar]."""
    Lc = _chk(Lc, "Lc"); sqrt_v = _chk(sqrt_v, "sqrt_v"); Q = _chk(Q, "Q")
    E = _chk(E, "E"); U = _chk(U, "U"); Fbar = _chk(Fbar, "Fbar")
    ensure_workspace(E.device)
    if klbar is not None:
        klbar = _chk(klbar, "klbar")
    R, S, n = E.shape
    dev = E.device
    Ubar = torch.empty_like(E)
    qmu_bar = torch.empty((R, n), dtype=torch.float64, device=dev)
    Qbar = torch.empty_like(Q)
    Lc_bar = torch.empty((n, n), dtype=torch.float64, device=dev)
    svdot = torch.empty(R, dtype=torch.float64, device=dev)
    mean_bar = torch.empty((R, n), dtype=torch.float64, device=dev)
    rc = _capi.lib().gpinv_sample_bwd_f64(_p(Lc), n, _p(sqrt_v), _p(Q), int(q_diag), _p(E), _p(U),
                                          _p(Fbar), _p(klbar), n, S, R, _p(Ubar), _p(qmu_bar),
                                          _p(Qbar), _p(Lc_bar), n, _p(svdot), _p(mean_bar), _st())
    _capi.check(rc, "sample_bwd")
    return [Lc_bar, svdot / sqrt_v, Qbar, qmu_bar, mean_bar]


def _sample_setup(ctx, inputs, output):
    Lc, sqrt_v, Q, q_mu, mean, E, q_diag, want_exp, U_pre, _kl_pre = inputs
    F, expF, kl, U = output
    if U_pre is not None:
        U = U_pre
    ctx.save_for_backward(Lc, sqrt_v, Q, E, U, expF)
    ctx.q_diag = q_diag
    ctx.want_exp = want_exp
    ctx.set_materialize_grads(False)


def _sample_backward(ctx, Fbar, expFbar, klbar, Ubar_unused):
    Lc, sqrt_v, Q, E, U, expF = ctx.saved_tensors
    if Fbar is None:
        Fbar = torch.zeros_like(E)
    if ctx.want_exp and expFbar is not None:
        Fbar = Fbar + expFbar * expF
    Lc_bar, sv_bar, Qbar, qmu_bar, mean_bar = stvgp_sample_bwd(
        Lc, sqrt_v, Q, E, U, Fbar.contiguous(), None if klbar is None else klbar.contiguous(),
        ctx.q_diag)
    # Qbar's strict upper triangle is unspecified; tril_unpack's backward masks it.  For the
    # q_diag layout Qbar is exact.
    return Lc_bar, sv_bar, Qbar, qmu_bar, mean_bar, None, None, None, None, None


stvgp_sample.register_autograd(_sample_backward, setup_context=_sample_setup)


# --------------------------------------------------------------------------------------
# T6 + T8-T13 with the Kronecker core kept factored (kronecker.RBF2D)
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::kron_apply", mutates_args=(), device_types="cuda")
def kron_apply(L1: Tensor, L2: Tensor, U: Tensor, trans: bool) -> Tensor:
    """F[b] = L1 U[b] L2^T (trans False: (L1 (x) L2) u, GPinv/kronecker.py:10-22 applied without forming
    the product) or L1^T U[b] L2 (trans True) for the rows of U [B, n1*n2]."""
    L1 = _chk(L1, "L1"); L2 = _chk(L2, "L2"); U = _chk(U, "U")
    n1, n2 = L1.shape[0], L2.shape[0]
    B = U.shape[0]
    T = torch.empty_like(U)
    F = torch.empty_like(U)
    rc = _capi.lib().gpinv_kron_apply_f64(_p(L1), n1, _p(L2), n2, _p(U), B, _p(T), _p(F), int(trans), _st())
    _capi.check(rc, "kron_apply")
    return F


@torch.library.custom_op("gpinv::kron_sample", mutates_args=(), device_types="cuda")
def kron_sample(L1: Tensor, L2: Tensor, sqrt_v: Tensor, Q: Tensor, q_mu: Tensor, mean: Tensor, E: Tensor,
                q_diag: bool, want_exp: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """(F [R,S,n], expF or empty, KL [1], U [R,S,n], T [R,S,n]) with the Cholesky factor
    sqrt_v_r * (L1 (x) L2) of kronecker.RBF2D kept factored (GPinv/kronecker.py:49-66 feeding
    GPinv/stvgp.py:144-185).  Q [R,n,n] lower or [R,n] (q_diag)."""
    L1 = _chk(L1, "L1"); L2 = _chk(L2, "L2"); sqrt_v = _chk(sqrt_v, "sqrt_v"); Q = _chk(Q, "Q")
    q_mu = _chk(q_mu, "q_mu"); mean = _chk(mean, "mean"); E = _chk(E, "E")
    ensure_workspace(E.device)
    R, S, n = E.shape
    n1, n2 = L1.shape[0], L2.shape[0]
    assert n1 * n2 == n
    U = torch.empty_like(E)
    T = torch.empty_like(E)
    F = torch.empty_like(E)
    expF = torch.empty_like(E) if want_exp else torch.empty(0, dtype=torch.float64, device=E.device)
    kl = torch.empty(4, dtype=torch.float64, device=E.device)
    rc = _capi.lib().gpinv_kron_sample_fwd_f64(_p(L1), n1, _p(L2), n2, _p(sqrt_v), _p(Q), int(q_diag), _p(q_mu),
                                               _p(mean), _p(E), S, R, _p(U), _p(T), _p(F),
                                               _p(expF) if want_exp else None, _p(kl), _st())
    _capi.check(rc, "kron_sample_fwd")
    return F, expF, kl[:1].clone(), U, T


@torch.library.custom_op("gpinv::kron_sample_bwd", mutates_args=(), device_types="cuda")
def kron_sample_bwd(L1: Tensor, L2: Tensor, sqrt_v: Tensor, Q: Tensor, E: Tensor, U: Tensor, T: Tensor,
                    Fbar: Tensor, klbar: Optional[Tensor], q_diag: bool) -> List[Tensor]:
    """[L1bar, L2bar, sqrt_v_bar, Qbar, qmu_bar, mean_bar]."""
    L1 = _chk(L1, "L1"); L2 = _chk(L2, "L2"); sqrt_v = _chk(sqrt_v, "sqrt_v"); Q = _chk(Q, "Q")
    E = _chk(E, "E"); U = _chk(U, "U"); T = _chk(T, "T"); Fbar = _chk(Fbar, "Fbar")
    ensure_workspace(E.device)
    if klbar is not None:
        klbar = _chk(klbar, "klbar")
    R, S, n = E.shape
    n1, n2 = L1.shape[0], L2.shape[0]
    dev = E.device
    Ubar = torch.empty_like(E)
    qmu_bar = torch.empty((R, n), dtype=torch.float64, device=dev)
    Qbar = torch.empty_like(Q)
    L1bar = torch.empty((n1, n1), dtype=torch.float64, device=dev)
    L2bar = torch.zeros((n2, n2), dtype=torch.float64, device=dev)
    svbar = torch.empty(R, dtype=torch.float64, device=dev)
    mean_bar = torch.empty((R, n), dtype=torch.float64, device=dev)
    lib = _capi.lib()
    ws = _ws(lib.gpinv_kron_bwd_ws_bytes(n1, n2, S, R), E)
    rc = lib.gpinv_kron_sample_bwd_f64(_p(L1), n1, _p(L2), n2, _p(sqrt_v), _p(Q), int(q_diag), _p(E), _p(U), _p(T),
                                       _p(Fbar), _p(klbar), S, R, _p(Ubar), _p(qmu_bar), _p(Qbar), _p(L1bar),
                                       _p(L2bar), _p(svbar), _p(mean_bar), _p(ws), _st())
    _capi.check(rc, "kron_sample_bwd")
    return [L1bar, L2bar, svbar, Qbar, qmu_bar, mean_bar]


def _kron_setup(ctx, inputs, output):
    L1, L2, sqrt_v, Q, q_mu, mean, E, q_diag, want_exp = inputs
    F, expF, kl, U, T = output
    ctx.save_for_backward(L1, L2, sqrt_v, Q, E, U, T, expF)
    ctx.q_diag = q_diag
    ctx.want_exp = want_exp
    ctx.set_materialize_grads(False)


def _kron_backward(ctx, Fbar, expFbar, klbar, Ubar_unused, Tbar_unused):
    L1, L2, sqrt_v, Q, E, U, T, expF = ctx.saved_tensors
    if Fbar is None:
        Fbar = torch.zeros_like(E)
    if ctx.want_exp and expFbar is not None:
        Fbar = Fbar + expFbar * expF
    L1bar, L2bar, sv_bar, Qbar, qmu_bar, mean_bar = kron_sample_bwd(
        L1, L2, sqrt_v, Q, E, U, T, Fbar.contiguous(), None if klbar is None else klbar.contiguous(), ctx.q_diag)
    return L1bar, L2bar, sv_bar, Qbar, qmu_bar, mean_bar, None, None, None


kron_sample.register_autograd(_kron_backward, setup_context=_kron_setup)


# --------------------------------------------------------------------------------------
# T14 Gaussian likelihood on latent samples
# --------------------------------------------------------------------------------------
LOG2PI = 1.8378770664093453


@torch.library.custom_op("gpinv::gauss_sumsq", mutates_args=(), device_types="cuda")
def gauss_sumsq(F: Tensor, Y: Tensor) -> Tensor:
    """sum_{r,s,i} (F[r,s,i] - Y[i,r])^2 -> [1]."""
    F = _chk(F, "F"); Y = _chk(Y, "Y")
    R, S, n = F.shape
    out = torch.empty(1, dtype=torch.float64, device=F.device)
    _capi.check(_capi.lib().gpinv_gauss_sumsq_f64(_p(F), _p(Y), n, S, R, _p(out), _st()), "gauss_sumsq")
    return out


@torch.library.custom_op("gpinv::gauss_resid_scaled", mutates_args=(), device_types="cuda")
def gauss_resid_scaled(F: Tensor, Y: Tensor, coef: Tensor) -> Tensor:
    """coef * (Y - F) -> [R,S,n]."""
    F = _chk(F, "F"); Y = _chk(Y, "Y"); coef = _chk(coef, "coef")
    R, S, n = F.shape
    out = torch.empty_like(F)
    _capi.check(_capi.lib().gpinv_gauss_bwd_f64(_p(F), _p(Y), n, S, R, _p(coef), _p(out), _st()), "gauss_bwd")
    return out


def _gss_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _gss_backward(ctx, g):
    F, Y = ctx.saved_tensors
    # d/dF sum (F-Y)^2 = 2 (F - Y) = -2 (Y - F)
    return gauss_resid_scaled(F, Y, (-2.0 * g).reshape(1).contiguous()), None


gauss_sumsq.register_autograd(_gss_backward, setup_context=_gss_setup)


# --------------------------------------------------------------------------------------
# T15 Abel transform + residual
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::abel_resid", mutates_args=(), device_types="cuda")
def abel_resid(F: Tensor, expF: Tensor, A: Tensor, Y: Tensor) -> Tuple[Tensor, Tensor]:
    """(sumsq [1], Rs [S,M]) with Rs = 1 y^T - exp(F) A^T.  F, expF [S,n]; A [M,n]; Y [M].
    `expF` must equal exp(F) (it is the second output of stvgp_sample); gradients flow to F."""
    expF = _chk(expF, "expF"); A = _chk(A, "A"); Y = _chk(Y, "Y")
    ensure_workspace(expF.device)
    S, n = expF.shape
    M = A.shape[0]
    Rs = torch.empty((S, M), dtype=torch.float64, device=expF.device)
    ss = torch.empty(1, dtype=torch.float64, device=expF.device)
    with _timed("abel_resid"):
        rc = _capi.lib().gpinv_abel_fwd_f64(_p(expF), _p(A), n, _p(Y), n, M, S, _p(Rs), _p(ss), _st())
    _capi.check(rc, "abel_fwd")
    return ss, Rs


@torch.library.custom_op("gpinv::abel_bwd", mutates_args=(), device_types="cuda")
def abel_bwd(Rs: Tensor, A: Tensor, expF: Tensor, coef: Tensor) -> Tensor:
    """coef * expF o (Rs A) -> [S,n]."""
    Rs = _chk(Rs, "Rs"); A = _chk(A, "A"); expF = _chk(expF, "expF"); coef = _chk(coef, "coef")
    S, M = Rs.shape
    n = A.shape[1]
    out = torch.empty((S, n), dtype=torch.float64, device=Rs.device)
    rc = _capi.lib().gpinv_abel_bwd_f64(_p(Rs), _p(A), n, _p(expF), _p(coef), n, M, S, _p(out), _st())
    _capi.check(rc, "abel_bwd")
    return out


def _abel_setup(ctx, inputs, output):
    F, expF, A, Y = inputs
    ctx.save_for_backward(output[1], A, expF)
    ctx.set_materialize_grads(False)      # (the unused gradient of Rs [S,M] was a 67 MB zero fill)


def _abel_backward(ctx, g_ss, g_rs_unused):
    Rs, A, expF = ctx.saved_tensors
    if g_ss is None:
        return None, None, None, None
    # d sumsq / dF[s,j] = 2 sum_m Rs[s,m] * (-A[m,j] expF[s,j])
    Fbar = abel_bwd(Rs, A, expF, (-2.0 * g_ss).reshape(1).contiguous())
    return Fbar, None, None, None


abel_resid.register_autograd(_abel_backward, setup_context=_abel_setup)


@torch.library.custom_op("gpinv::sumsq", mutates_args=(), device_types="cuda")
def sumsq(x: Tensor) -> Tensor:
    x = _chk(x, "x")
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    _capi.check(_capi.lib().gpinv_sumsq_f64(_p(x), x.numel(), _p(out), _st()), "sumsq")
    return out


# --------------------------------------------------------------------------------------
# T16 spectroscopic Abel transform
@torch.library.custom_op("gpinv::gauss_logp", mutates_args=(), device_types="cuda")
def gauss_logp(ss: Tensor, var: Tensor, count: float) -> Tensor:
    """-count/2 (log 2 pi + log var) - ss / (2 var) as a device scalar [1]: the Gaussian log-density summed over
    `count` observations, from the residual sum of squares (GPflow densities.gaussian)."""
    ss = _chk(ss, "ss").reshape(1)
    var = _chk(var, "var").reshape(1)
    out = torch.empty(1, dtype=torch.float64, device=ss.device)
    _capi.check(_capi.lib().gpinv_gauss_logp_f64(_p(ss), _p(var), float(count), _p(out), _st()), "gauss_logp")
    return out


@torch.library.custom_op("gpinv::gauss_logp_bwd", mutates_args=(), device_types="cuda")
def gauss_logp_bwd(ss: Tensor, var: Tensor, count: float, g: Tensor) -> List[Tensor]:
    ss = _chk(ss, "ss").reshape(1)
    var = _chk(var, "var").reshape(1)
    g = _chk(g, "g").reshape(1)
    ss_bar = torch.empty(1, dtype=torch.float64, device=ss.device)
    var_bar = torch.empty(1, dtype=torch.float64, device=ss.device)
    _capi.check(_capi.lib().gpinv_gauss_logp_bwd_f64(_p(ss), _p(var), float(count), _p(g), _p(ss_bar), _p(var_bar),
                                                     _st()), "gauss_logp_bwd")
    return [ss_bar, var_bar]


def _glp_setup(ctx, inputs, output):
    ss, var, count = inputs
    ctx.save_for_backward(ss, var)
    ctx.count = count
    ctx.shapes = (ss.shape, var.shape)


def _glp_backward(ctx, g):
    ss, var = ctx.saved_tensors
    ss_bar, var_bar = gauss_logp_bwd(ss, var, ctx.count, g.contiguous())
    return ss_bar.reshape(ctx.shapes[0]), var_bar.reshape(ctx.shapes[1]), None


gauss_logp.register_autograd(_glp_backward, setup_context=_glp_setup)


# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::spec_transform", mutates_args=(), device_types="cuda")
def spec_transform(F: Tensor, A: Tensor, cosT: Tensor, lam: Tensor, dlam: float = 0.0) -> Tensor:
    """P [S,k,m] from F [3,S,n] (notebooks/Spectroscopic_Abel_inversion.ipynb:316-346).
    A, cosT: the geometry matrices TRANSPOSED to [n,m] (see ``likelihoods.SpecAbelLikelihood``).
    ``dlam``: spacing of ``lam`` when it is a uniform grid (recurrence kernels), 0 for arbitrary wavelengths."""
    F = _chk(F, "F"); A = _chk(A, "A"); cosT = _chk(cosT, "cosT"); lam = _chk(lam, "lam")
    _, S, n = F.shape
    m, k = A.shape[1], lam.numel()
    P = torch.empty((S, k, m), dtype=torch.float64, device=F.device)
    rc = _capi.lib().gpinv_spec_fwd_f64(_p(F), _p(A), _p(cosT), _p(lam), None, n, m, k, S, float(dlam), _p(P), None,
                                        _st())
    _capi.check(rc, "spec_fwd")
    return P


@torch.library.custom_op("gpinv::spec_resid", mutates_args=(), device_types="cuda")
def spec_resid(F: Tensor, A: Tensor, cosT: Tensor, lam: Tensor, Y: Tensor, dlam: float = 0.0) -> Tuple[Tensor, Tensor]:
    """(sum (P - Y)^2 [1], P [S,k,m]); A, cosT transposed to [n,m]."""
    F = _chk(F, "F"); A = _chk(A, "A"); cosT = _chk(cosT, "cosT"); lam = _chk(lam, "lam"); Y = _chk(Y, "Y")
    _, S, n = F.shape
    m, k = A.shape[1], lam.numel()
    P = torch.empty((S, k, m), dtype=torch.float64, device=F.device)
    ss = torch.empty(1, dtype=torch.float64, device=F.device)
    rc = _capi.lib().gpinv_spec_fwd_f64(_p(F), _p(A), _p(cosT), _p(lam), _p(Y), n, m, k, S, float(dlam), _p(P),
                                        _p(ss), _st())
    _capi.check(rc, "spec_fwd")
    return ss, P


@torch.library.custom_op("gpinv::spec_bwd", mutates_args=(), device_types="cuda")
def spec_bwd(F: Tensor, A: Tensor, cosT: Tensor, lam: Tensor, Y: Tensor, P: Tensor, coef: Tensor,
             dlam: float = 0.0) -> Tensor:
    F = _chk(F, "F"); P = _chk(P, "P"); coef = _chk(coef, "coef")
    _, S, n = F.shape
    m, k = A.shape[1], lam.numel()
    out = torch.empty_like(F)
    rc = _capi.lib().gpinv_spec_bwd_f64(_p(F), _p(A), _p(cosT), _p(lam), _p(Y), _p(P), _p(coef), n, m, k, S,
                                        float(dlam), _p(out), _st())
    _capi.check(rc, "spec_bwd")
    return out


def _spec_setup(ctx, inputs, output):
    F, A, cosT, lam, Y, dlam = inputs
    ctx.save_for_backward(F, A, cosT, lam, Y, output[1])
    ctx.dlam = dlam
    ctx.set_materialize_grads(False)


def _spec_backward(ctx, g_ss, g_p_unused):
    F, A, cosT, lam, Y, P = ctx.saved_tensors
    if g_ss is None:
        return None, None, None, None, None, None
    return spec_bwd(F, A, cosT, lam, Y, P, (2.0 * g_ss).reshape(1).contiguous(), ctx.dlam), None, None, None, None, None


spec_resid.register_autograd(_spec_backward, setup_context=_spec_setup)


# --------------------------------------------------------------------------------------
# T19: packed free state <-> constrained parameter blocks, one kernel each way
# --------------------------------------------------------------------------------------
def _pack_args(offsets, transforms, tensors):
    nb = len(transforms)
    off = (C.c_longlong * (nb + 1))(*offsets)
    tr = (C.c_int * nb)(*transforms)
    ptrs = (C.c_void_p * nb)(*[None if t is None else t.data_ptr() for t in tensors])
    return nb, off, tr, ptrs


@torch.library.custom_op("gpinv::unpack_state", mutates_args=(), device_types="cuda")
def unpack_state(x: Tensor, offsets: List[int], transforms: List[int]) -> List[Tensor]:
    """Constrained values of every parameter block of the packed free state ``x`` (GPflow
    ``set_state`` with transforms.positive = softplus + 1e-6) in ONE kernel.  ``offsets`` has one
    entry more than ``transforms`` (0 identity, 1 positive); block b is a flat tensor."""
    x = _chk(x, "x")
    outs = [torch.empty(offsets[b + 1] - offsets[b], dtype=torch.float64, device=x.device)
            for b in range(len(transforms))]
    nb, off, tr, ptrs = _pack_args(offsets, transforms, outs)
    _capi.check(_capi.lib().gpinv_unpack_f64(_p(x), nb, off, tr, ptrs, _st()), "unpack")
    return outs


@torch.library.custom_op("gpinv::pack_grad", mutates_args=(), device_types="cuda")
def pack_grad(x: Tensor, offsets: List[int], transforms: List[int], grads: List[Optional[Tensor]]) -> Tensor:
    """d/dx from the gradients w.r.t. the constrained blocks (chain rule sigmoid(x) for positive
    blocks, zeros where a block received no gradient), written straight into the packed layout."""
    x = _chk(x, "x")
    gs = [None if g is None else _chk(g, "grad").reshape(-1) for g in grads]
    # one extra slot behind the gradient: Model puts the objective value there, so that [gradient, value]
    # is one contiguous buffer (one all-reduce / one host copy) without a concatenation pass
    out = torch.empty(x.numel() + 1, dtype=torch.float64, device=x.device)[:x.numel()]
    nb, off, tr, ptrs = _pack_args(offsets, transforms, gs)
    _capi.check(_capi.lib().gpinv_pack_f64(_p(x), nb, off, tr, ptrs, 1.0, _p(out), _st()), "pack")
    return out


def _unpack_setup(ctx, inputs, output):
    x, offsets, transforms = inputs
    ctx.save_for_backward(x)
    ctx.offsets, ctx.transforms = list(offsets), list(transforms)
    ctx.set_materialize_grads(False)      # blocks without a gradient arrive as None: pack_grad writes their zeros


def _unpack_backward(ctx, grads):
    (x,) = ctx.saved_tensors
    gs = [None if g is None else g.contiguous() for g in grads]
    return pack_grad(x, ctx.offsets, ctx.transforms, gs), None, None


unpack_state.register_autograd(_unpack_backward, setup_context=_unpack_setup)


# --------------------------------------------------------------------------------------
# analytic whitened KL (GPinv/stvgp.py:187-195)
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::kl_white", mutates_args=(), device_types="cuda")
def kl_white(q_mu: Tensor, q_sqrt: Tensor, q_diag: bool) -> Tensor:
    """GPflow gauss_kl_white[_diag]: q_mu [n,R]; q_sqrt [n,n,R] (lower triangle read) or [n,R]."""
    q_mu = _chk(q_mu, "q_mu"); q_sqrt = _chk(q_sqrt, "q_sqrt")
    n, R = q_mu.shape
    out = torch.empty(4, dtype=torch.float64, device=q_mu.device)
    _capi.check(_capi.lib().gpinv_kl_white_f64(_p(q_mu), _p(q_sqrt), n, R, int(q_diag), _p(out), _st()), "kl_white")
    return out[:1].clone()


@torch.library.custom_op("gpinv::kl_white_bwd", mutates_args=(), device_types="cuda")
def kl_white_bwd(q_mu: Tensor, q_sqrt: Tensor, q_diag: bool, gbar: Tensor) -> List[Tensor]:
    q_mu = _chk(q_mu, "q_mu"); q_sqrt = _chk(q_sqrt, "q_sqrt"); gbar = _chk(gbar, "gbar")
    n, R = q_mu.shape
    mu_bar, q_bar = torch.empty_like(q_mu), torch.empty_like(q_sqrt)
    _capi.check(_capi.lib().gpinv_kl_white_bwd_f64(_p(q_mu), _p(q_sqrt), n, R, int(q_diag), _p(gbar), _p(mu_bar),
                                                   _p(q_bar), _st()), "kl_white_bwd")
    return [mu_bar, q_bar]


def _klw_setup(ctx, inputs, output):
    q_mu, q_sqrt, q_diag = inputs
    ctx.save_for_backward(q_mu, q_sqrt)
    ctx.q_diag = q_diag


def _klw_backward(ctx, g):
    q_mu, q_sqrt = ctx.saved_tensors
    mu_bar, q_bar = kl_white_bwd(q_mu, q_sqrt, ctx.q_diag, g.reshape(1).contiguous())
    return mu_bar, q_bar, None


kl_white.register_autograd(_klw_backward, setup_context=_klw_setup)


# --------------------------------------------------------------------------------------
# optimiser step on the device-resident state
# --------------------------------------------------------------------------------------
@torch.library.custom_op("gpinv::adam_step", mutates_args=("x", "m", "v", "step"), device_types="cuda")
def adam_step(x: Tensor, m: Tensor, v: Tensor, g: Tensor, step: Tensor, lr: float, beta1: float, beta2: float,
              epsilon: float) -> None:
    """TensorFlow's AdamOptimizer update in place (tf.train.AdamOptimizer at testing/test_stvgp.py:31-35);
    ``step`` is a device int64 [1] advanced by the call, so the launch can sit in a CUDA graph."""
    for t, nm in ((x, "x"), (m, "m"), (v, "v"), (g, "g")):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise RuntimeError("adam_step: %s must be a contiguous fp64 CUDA tensor" % nm)
    _capi.check(_capi.lib().gpinv_adam_step_f64(_p(x), _p(m), _p(v), _p(g), x.numel(), float(lr), float(beta1),
                                                float(beta2), float(epsilon), _p(step), _st()), "adam_step")
