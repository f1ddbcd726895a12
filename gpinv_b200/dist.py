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

    def all_reduce_async(self, buf):
        """Start the sum over ranks and return the work handle (``.wait()`` orders the caller's
        stream after it)."""
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)


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
