"""Multi-GPU: shard the Monte-Carlo samples over the ranks of one box, one all-reduce.

Every rank holds the replicated state x, data and Cholesky factor; rank g evaluates the ELBO
terms of samples [g S/G, (g+1) S/G) divided by the *global* S (sample-independent terms are
divided by G) and runs the reverse pass of the sampling stage on those partial sums.  The sum over
the ranks of [-f, -df/dx] is ONE logical all-reduce, issued in pieces so that it hides behind compute:
the q_sqrt block (half of it: the lower triangle) as soon as the sampling stage's reverse is done,
the small remainder at the end.  In between, the Cholesky reverse -- linear in Lbar -- is either run
replicated on each rank's partial Lbar, or (n >= 1024, n divisible by 4G) sharded by column blocks
after a reduce-scatter of Lbar: see ``DistContext.shard_width``.  NCCL over NVLink on the GPUs; gloo in
the CPU tests.  SURVEY.md 8(e).  The reference has no distributed path at all.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DistContext(object):
    def __init__(self, rank=None, world_size=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size

    def shard_bounds(self, S):
        if S % self.world_size != 0:
            raise ValueError("num_samples (%d) must be divisible by the number of ranks (%d)" % (S, self.world_size))
        per = S // self.world_size
        return self.rank * per, (self.rank + 1) * per

    def shard_samples(self, eps):
        """eps [R,S,n] -> this rank's [R,S/G,n] slice."""
        s0, s1 = self.shard_bounds(eps.shape[1])
        return eps[:, s0:s1, :]

    def all_reduce(self, buf):
        if self.world_size > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf


    # ---- column-block sharding of the Cholesky reverse pass ------------------------------------------
    # The reverse of the factorisation is linear in Lbar; only its contraction with dK/d(lengthscale) is
    # needed, and column block J of Phi(L^T Lbar) depends on column block J of Lbar only.  Instead of every
    # rank reversing its own partial Lbar over the whole matrix (replicated 4/3 n^3 flops), the ranks
    # reduce-scatter Lbar by column blocks and each reverses its blocks (about 3 n^3 / G flops).  2G blocks of
    # width w = n / 2G, rank r owns blocks r and 2G-1-r: the triangular work is even over the ranks.
    SHARD_MIN_N = 1024

    def shard_width(self, n):
        """Column-block width of the sharded reverse for an n x n factor, or 0 when it stays replicated."""
        import os
        G = self.world_size
        env = os.environ.get("GPINV_SHARD_REVERSE")
        if G < 2 or env == "0" or n < self.SHARD_MIN_N or n % (4 * G) != 0:
            return 0
        return n // (2 * G)

    def shard_columns(self, n):
        """First columns of this rank's blocks, in the order they lie in its reduce-scatter chunk."""
        w = n // (2 * self.world_size)
        return [self.rank * w, (2 * self.world_size - 1 - self.rank) * w]

    def reduce_scatter(self, out, inp):
        """out <- this rank's chunk of the rank-sum of inp (G equal chunks)."""
        if dist.get_backend(self.group) == "nccl":
            dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=self.group)
        else:       # gloo (the CPU tests) has no reduce-scatter: sum everything, keep the own chunk
            tmp = inp.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            out.copy_(tmp.view(self.world_size, -1)[self.rank])
        return out

    def all_gather(self, out, inp):
        """out [G * inp.numel()] <- the ranks' inp, in rank order."""
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(out, inp, group=self.group)
        else:
            dist.all_gather(list(out.view(self.world_size, -1).unbind(0)), inp.reshape(-1), group=self.group)
        return out

    def all_reduce_async(self, buf):
        """Start the sum over ranks and return the work handle (``.wait()`` orders the caller's
        stream after it)."""
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)


class SharedHostBuffers(object):
    """Page-locked host buffers shared by the ranks of ONE node (POSIX shared memory mapped and
    cudaHostRegister-ed by every rank).  The host API of the sample-sharded evaluation returns the same
    [value, gradient] on every rank: instead of G device-to-host copies of the whole gradient through the
    node's memory system, rank g downloads 1/G of it into the shared buffer and a barrier publishes it.
    All methods are collective."""

    def __init__(self, ctx, count, nbuf=2):
        self.ctx, self.count = ctx, int(count)
        self.bufs = []                       # [tensor, mmap, weakref to the array handed out]
        for _ in range(nbuf):
            self._add()

    def _add(self):
        import mmap
        import os
        ctx = self.ctx
        dev = torch.device("cuda", torch.cuda.current_device())
        tok = torch.randint(0, 2 ** 62, (1,), device=dev, dtype=torch.int64)
        dist.broadcast(tok, src=0 if ctx.group is None else dist.get_global_rank(ctx.group, 0), group=ctx.group)
        path = "/dev/shm/gpinv_b200_%x" % int(tok.item())
        nbytes = 8 * self.count
        ok = 1
        fd = -1
        try:
            if ctx.rank == 0:
                vfs = os.statvfs("/dev/shm")
                if vfs.f_bavail * vfs.f_frsize < nbytes + (64 << 20):
                    raise OSError("not enough room in /dev/shm")     # (a sparse file would SIGBUS on first touch)
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, nbytes)     # zero-filled
        except OSError:
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=ctx.group)          # (also: the file exists from here on)
        if int(flag.item()) == 0:
            raise RuntimeError("cannot create the shared result buffer %s" % path)
        try:
            if ctx.rank != 0:
                fd = os.open(path, os.O_RDWR)
            mm = mmap.mmap(fd, nbytes)
            os.close(fd)
            t = torch.frombuffer(mm, dtype=torch.float64, count=self.count)
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
            ok = 1 if int(rc) == 0 else 0
        except (OSError, ValueError, RuntimeError):
            ok, mm, t = 0, None, None
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=ctx.group)          # everyone has mapped it
        if ctx.rank == 0:
            os.unlink(path)                  # the mappings keep it alive; nothing is left behind on exit
        if int(flag.item()) == 0:
            raise RuntimeError("cannot map / page-lock the shared result buffer on every rank (ranks on different "
                               "nodes?)")
        self.bufs.append([t, mm, None])

    def acquire(self):
        """Index of a buffer that no rank's caller still holds an array of (a new one if there is none)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        busy = torch.tensor([1 if (b[2] is not None and b[2]() is not None) else 0 for b in self.bufs], device=dev)
        dist.all_reduce(busy, op=dist.ReduceOp.MAX, group=self.ctx.group)
        busy = busy.tolist()
        for i, b in enumerate(busy):
            if not b:
                return i
        self._add()
        return len(self.bufs) - 1

    def publish(self):
        """Barrier: every rank's part of the current buffer has landed (call after synchronising the streams
        that carried this rank's copies)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.zeros(1, device=dev)
        dist.all_reduce(t, group=self.ctx.group)
        t.item()


def attach(model, ctx=None):
    """Make ``model._objective`` evaluate its shard and all-reduce [-f, -grad]."""
    if not hasattr(model, "_eps_tensor"):
        # GPMC evaluates the whole likelihood on every rank (S = 1, nothing to shard): summing the
        # ranks would return G x the log-likelihood.  SURVEY.md 8(e): replicas only.
        raise TypeError("%s has no Monte-Carlo samples to shard: run independent replicas (one chain "
                        "per GPU) instead of dist.attach()" % type(model).__name__)
    model.dist = ctx if ctx is not None else DistContext()
    return model


def sharded_objective(local_fn, x, eps, ctx):
    """Engine-agnostic form used by the CPU (gloo) tests: ``local_fn(x, eps_local, S_global,
    share)`` returns this rank's partial (-f, -grad) tensors; the sum over ranks is returned."""
    S = eps.shape[1]
    e = ctx.shard_samples(eps)
    f, g = local_fn(x, e, S, 1.0 / ctx.world_size)
    buf = torch.cat([f.reshape(1), g.reshape(-1)])
    ctx.all_reduce(buf)
    return buf[:1], buf[1:]
