"""Multi-GPU: shard the Monte-Carlo samples over the ranks of one box, one all-reduce.

Every rank holds the replicated state x, data and Cholesky factor; rank g evaluates the ELBO
terms of samples [g S/G, (g+1) S/G) divided by the *global* S (sample-independent terms are
divided by G), finishes its whole backward pass on those partial sums — the Cholesky and
kernel reverse passes are linear in Lbar, so no intermediate exchange is needed — and the
packed buffer [-f, -df/dx] is summed with ONE all-reduce (NCCL over NVLink on the GPUs; gloo in
the CPU tests).  SURVEY.md 8(e).  The reference has no distributed path at all.
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
