"""Model base: the upward interface that optimisers and samplers call.

In the reference this layer is GPflow's ``Model`` (``_compile`` / ``_objective`` / ``optimize``
/ ``sample``; GPinv's own clone of the pattern is GPinv/model.py:23-91,148-240).  Here
``_objective(x)`` evaluates ``build_likelihood() + build_prior()`` with the sm_100a ops and
their hand-written backward passes and returns ``(-f, -df/dx)`` exactly like GPflow's
closure (usage: notebooks/Abel_inversion.ipynb:416-418).

``x`` may be a numpy vector (reference behaviour: host in, host out — one H2D copy of ``x`` and
one D2H copy of the gradient) or a CUDA tensor (device-resident: nothing leaves HBM and
nothing synchronises).
"""
from __future__ import annotations

import os
import sys
import threading
import time
import weakref

import numpy as np
import torch

from . import hmc
from .param import Parameterized, default_device, state_epoch


class OptimizeResult(dict):
    __getattr__ = dict.__getitem__


class Model(Parameterized):
    def __init__(self, name="model"):
        Parameterized.__init__(self)
        self._name = name
        self._pending_infos = []
        self.dist = None            # set by gpinv_b200.dist.attach()
        self._pinned = {}
        self._out_pool = []
        self._use_graph = False
        self._graphs = {}

    # ---- to be provided by subclasses ----------------------------------------------
    def build_likelihood(self, **kw):
        raise NotImplementedError

    def _prepare(self):
        """Hook called before every evaluation (StVGP re-creates q_mu/q_sqrt when X changed)."""

    def _compile(self, optimizer=None, **kw):
        """GPflow ``Model._compile`` built the TF graph and session here (reference usage:
        testing/test_stvgp.py:17).  There is no graph to build: the shape re-check of
        GPinv/stvgp.py:76-94 runs, the library is loaded, the split-K workspace registered."""
        from . import ops
        self._prepare()
        ops.ensure_workspace(default_device())

    # ---- objective -------------------------------------------------------------------
    def enable_cuda_graph(self, flag=True, static_inputs=False):
        """Capture one whole evaluation (forward + hand-written backward, ~100-250 kernel launches)
        in a CUDA graph and replay it: the small configurations are launch-latency bound.  The
        returned tensors are the graph's static outputs (valid until the next evaluation).
        ``static_inputs``: the graph reads the caller's x (and eps) tensors in place instead of private
        copies refreshed before every replay (134 MB + 8 S n bytes per evaluation at the headline size);
        the caller updates those tensors in place and keeps passing the same ones -- a different tensor
        gets its own capture."""
        self._use_graph = bool(flag)
        self._static_inputs = bool(static_inputs)
        self._graphs = {}
        return self

    def _graph_key(self, x, dyn):
        key = (x.numel(),) + tuple((k, tuple(v.shape)) for k, v in sorted(dyn.items()))
        if getattr(self, "_static_inputs", False):
            key += (x.data_ptr(),) + tuple(v.data_ptr() for _, v in sorted(dyn.items()))
        return key

    def _graph_static(self, x, dyn):
        """The tensors a capture reads: the caller's own (static_inputs) or private copies."""
        if getattr(self, "_static_inputs", False):
            return x, dict(dyn)
        return x.detach().clone(), {k: v.detach().clone() for k, v in dyn.items()}

    @staticmethod
    def _graph_refresh(sx, sdyn, x, dyn):
        if sx.data_ptr() != x.data_ptr():
            sx.copy_(x)
        for k, v in dyn.items():
            if sdyn[k].data_ptr() != v.data_ptr():
                sdyn[k].copy_(v)

    @staticmethod
    def _capture(fn, pool=None):
        """Capture ``fn()`` into a new CUDA graph; returns (graph, fn's result).  Nothing but the captured
        work may touch CUDA while the capture is open: a garbage collection that frees device or pinned
        memory of unrelated objects in the middle of it invalidates the capture (seen when an earlier test
        had failed and its frames were still alive), so collect first and keep the collector off until the
        capture ends; calls of other threads (pinned-allocator event queries ...) are not checked
        (thread_local)."""
        import gc
        graph = torch.cuda.CUDAGraph()
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            kw = {} if pool is None else {"pool": pool}
            with torch.cuda.graph(graph, capture_error_mode="thread_local", **kw):
                out = fn()
        finally:
            if gc_was_on:
                gc.enable()
        return graph, out

    def _graph_inputs(self, kw):
        """Tensors that vary between evaluations besides x (StVGP: the eps draws)."""
        return {}

    def _objective_device(self, x: torch.Tensor, **kw):
        """(-f, -df/dx) as CUDA tensors ([1] and [|x|]); no host synchronisation."""
        if self.dist is not None and self.dist.world_size > 1 and self.fused_layout() is not None:
            return self._objective_sharded(x, **kw)
        if self._use_graph:
            return self._objective_graphed(x, **kw)
        return self._objective_eager(x, **kw)

    def _objective_graphed(self, x, **kw):
        self._prepare()
        dyn = self._graph_inputs(kw)
        key = self._graph_key(x, dyn)
        if getattr(self, "_graph_epoch", None) != state_epoch():
            # data, a fixed value, the jitter level ... changed on the host since the capture: the
            # graph may have baked the old value (or a freed address) in.  Re-capture.
            self._graphs = {}
            self._graph_epoch = state_epoch()
        ent = self._graphs.get(key)
        if ent is None:
            sx, sdyn = self._graph_static(x, dyn)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            # Under sample sharding the graph holds this rank's LOCAL evaluation only; the all-reduce
            # of its static output runs eagerly after every replay (a captured NCCL collective
            # gained nothing measurable and made process-group teardown hang).
            with torch.cuda.stream(side):
                for _ in range(3):      # warm-up: lazy attribute setting, workspace registration
                    self._objective_eager(sx, reduce=False, **sdyn)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph, (f, g) = self._capture(lambda: self._objective_eager(sx, reduce=False, **sdyn))
            # strong references to every non-input tensor the capture read (data holders, fixed
            # Params): the caching allocator must not hand their memory out while the graph lives
            keep = [d._tensor for nd in self._all_parameterized() for d in nd.data_holders] + \
                   [p._fixed_dev for p in self.all_params() if p.fixed]
            ent = (graph, sx, sdyn, f, g, list(self._pending_infos), keep)
            self._graphs[key] = ent
        graph, sx, sdyn, f, g, infos, _keep = ent
        self._graph_refresh(sx, sdyn, x, dyn)
        graph.replay()
        self._pending_infos = list(infos)
        if self.dist is not None:
            self.dist.all_reduce(f._base)          # f and g are views of one buffer
        return f, g

    def _objective_eager(self, x: torch.Tensor, reduce=True, **kw):
        """One evaluation, launched kernel by kernel.  ``reduce=False`` (graph capture under
        sample sharding) returns this rank's partial [-f, -grad] without the all-reduce."""
        self._prepare()
        if not (reduce and self.dist is not None and self.dist.world_size > 1):
            out = self._objective_fused(x, reduce=reduce, **kw)
            if out is not None:
                return out
        leaves = self.make_tensors(x)
        self._pending_infos = []
        early = {}
        if reduce and self.dist is not None and self.dist.world_size > 1 and leaves:
            # Sample sharding: the gradient of the largest block (q_sqrt, 134 MB at the headline
            # size) exists as soon as the sampling stage's reverse pass is done; its all-reduce
            # starts there, on NCCL's stream, and overlaps the Cholesky reverse sweep.
            big = max(range(len(leaves)), key=lambda i: leaves[i].numel())
            if leaves[big].numel() >= (1 << 16):
                def hook(g, big=big):
                    gc = g.contiguous()
                    early["grad"] = gc
                    early["work"] = self.dist.all_reduce_async(gc)
                    early["index"] = big
                leaves[big].register_hook(hook)
        try:
            with self.tensor_mode():
                f = self.build_likelihood(**kw)
                prior = self.build_prior()
                if prior is not None:
                    f = f + (prior if self.dist is None else prior / self.dist.world_size)
            nf = -f.reshape(())          # differentiate -f: the gradients come out already negated
            grads = torch.autograd.grad(nf, leaves, allow_unused=True)
        finally:
            self.clear_tensors()
        parts = []
        for g, leaf in zip(grads, leaves):
            parts.append(torch.zeros(leaf.numel(), dtype=torch.float64, device=x.device) if g is None
                         else g.reshape(-1))
        if "work" in early:
            i = early["index"]
            small = torch.cat([nf.detach().reshape(1)] + parts[:i] + parts[i + 1:])
            self.dist.all_reduce(small)
            early["work"].wait()          # stream-orders the early all-reduce before the final copy
            n0 = 1 + sum(p.numel() for p in parts[:i])
            buf = torch.cat([small[:n0], early["grad"].reshape(-1), small[n0:]])
            return buf[:1], buf[1:]
        buf = torch.cat([nf.detach().reshape(1)] + parts)
        if reduce and self.dist is not None:
            self.dist.all_reduce(buf)
        return buf[:1], buf[1:]

    def _objective_fused(self, x, reduce=True, **kw):
        """The evaluation with the free state unpacked by one kernel and the gradient packed (chain
        rule of the positive transform included) by one kernel, straight into the [gradient, value]
        buffer: no per-parameter softplus / sigmoid / zero-fill / concatenation passes.  None when
        a parameter uses a transform the kernels do not know (then the generic path runs)."""
        leaf = self.make_tensors_fused(x)
        if leaf is None:
            return None
        self._pending_infos = []
        try:
            with self.tensor_mode():
                f = self.build_likelihood(**kw)
                prior = self.build_prior()
                if prior is not None:
                    f = f + (prior if self.dist is None else prior / self.dist.world_size)
            nf = -f.reshape(())
            (gx,) = torch.autograd.grad(nf, [leaf], allow_unused=True)
        finally:
            self.clear_tensors()
        n = x.numel()
        buf = self._grad_value_buffer(gx, nf, n, x.device)
        if reduce and self.dist is not None:
            self.dist.all_reduce(buf)
        return buf[n:], buf[:n]

    @staticmethod
    def _grad_value_buffer(gx, nf, n, device):
        """[gradient (n), value] in one buffer."""
        if gx is None:
            gx = torch.zeros(n + 1, dtype=torch.float64, device=device)[:n]
        st = gx.untyped_storage()
        if gx.is_contiguous() and gx.storage_offset() == 0 and gx.numel() == n and st.nbytes() >= 8 * (n + 1):
            # pack_grad left one slot behind the gradient: [gradient, value] without a concatenation
            buf = torch.empty(0, dtype=torch.float64, device=device).set_(st, 0, (n + 1,))
            buf[n:].copy_(nf.detach().reshape(1))
        else:                                           # a prior contributed through a second path: summed copy
            buf = torch.cat([gx.reshape(-1), nf.detach().reshape(1)])
        return buf

    # ---- sample-sharded evaluation (gpinv_b200.dist.attach, G > 1) -----------------------------------
    # Three compute stages with the NCCL exchanges between them (the stages are what a CUDA graph
    # captures; the collectives stay outside the graphs):
    #   A  forward pass on this rank's samples and the reverse pass down to Lbar            [compute]
    #   x1 reduce-scatter of Lbar by column blocks; all-reduce of the q_sqrt gradient's lower
    #      triangle STARTS (NCCL stream)                                                    [NVLink]
    #   B  Cholesky reverse of this rank's column blocks -> lengthscale gradient            [compute, beside x1]
    #   x2 all-reduce of the small remainder [hyperparameters, q_mu, value]; join x1        [NVLink]
    #   C  the reduced pieces go back into the [gradient, value] buffer                     [compute]
    def _shard_stage_a(self, x, kw, st):
        from . import ops
        ctx = self.dist
        leaf = self.make_tensors_fused(x)
        self._pending_infos = []
        col = ops.DeferredReverse(lambda nn: ctx.shard_width(nn) > 0)
        prev = ops.defer_reverse(col)
        ops.INVERSE_SHARD = (ctx.rank, ctx.world_size)      # the inversion's last level: this rank's slices only
        try:
            with self.tensor_mode():
                f = self.build_likelihood(**kw)
                prior = self.build_prior()
                if prior is not None:
                    f = f + prior / ctx.world_size
            nf = -f.reshape(())
            (gx,) = torch.autograd.grad(nf, [leaf], retain_graph=True, allow_unused=True)
        finally:
            ops.defer_reverse(prev)
            ops.INVERSE_SHARD = None
            self.clear_tensors()
        n = x.numel()
        buf = self._grad_value_buffer(gx, nf, n, x.device)
        st.update(leaf=leaf, buf=buf, n=n, items=col.items, big=None, early=None, tri=0)
        if ops.SIDE_PENDING[0]:
            ops.trtri_side_join()            # (inside a capture: the side stream must be joined before it ends)
        G = ctx.world_size
        st["rs_in"] = [ops.lbar_pack(it[5], G) for it in col.items]
        st["rs_out"] = [torch.empty(r.numel() // G, dtype=torch.float64, device=x.device) for r in st["rs_in"]]
        self._inverse_slices_stage(st, col.items, ctx, x.device)
        params, offs, _codes = self.fused_layout()
        big = max(range(len(params)), key=lambda i: params[i].size)
        a, b = offs[big], offs[big + 1]
        if b - a >= (1 << 16):
            nq = getattr(params[big], "_tril_rows", None)
            st["big"] = (a, b)
            if nq is not None and nq * nq == b - a:
                # only tril(q_sqrt) enters the model: its gradient's strict upper part is exactly zero
                st["tri"] = nq
                st["early"] = ops.tri_compact(buf[a:b].view(nq, nq))
            else:
                st["early"] = buf[a:b]

    @staticmethod
    def _inverse_slices_stage(st, items, ctx, dev):
        """Pack this rank's slices of every sharded inversion (ops.INVERSE_SHARD) for the all-gather."""
        from . import ops
        G = ctx.world_size
        st["winv_in"], st["winv_out"], st["winv_W"] = [], [], []
        for it in items:
            W = it[4]
            if W.numel() and ops.inverse_shardable(it[3].shape[0], G):
                mine = ops.inverse_slices_pack(W, ctx.rank, G)
                st["winv_in"].append(mine)
                st["winv_out"].append(torch.empty(G * mine.numel(), dtype=torch.float64, device=dev))
                st["winv_W"].append(W)

    @staticmethod
    def _inverse_slices_exchange(st, ctx):
        for o, i in zip(st["winv_out"], st["winv_in"]):
            ctx.all_gather(o, i)

    @staticmethod
    def _inverse_slices_finish(st, ctx):
        from . import ops
        for W, o in zip(st["winv_W"], st["winv_out"]):
            ops.inverse_slices_unpack_(W, o, ctx.world_size)

    def _shard_exchange_1(self, st):
        ctx = self.dist
        self._inverse_slices_exchange(st, ctx)
        for o, i in zip(st["rs_out"], st["rs_in"]):
            ctx.reduce_scatter(o, i)
        st["work"] = ctx.all_reduce_async(st["early"]) if st["big"] is not None else None

    def _shard_stage_b(self, st):
        from . import ops
        ctx = self.dist
        leaf, buf, n = st["leaf"], st["buf"], st["n"]
        gx2 = None
        self._inverse_slices_finish(st, ctx)
        if st["items"]:
            outs, gs = [], []
            for (X, ls, mode, L, W, _lbar), dsh in zip(st["items"], st["rs_out"]):
                nn = L.shape[0]
                g = ops.chol_core_bwd_shard(X, ls, mode, L, W, dsh, ctx.shard_width(nn), ctx.shard_columns(nn))
                outs.append(ls)
                gs.append(g.reshape(ls.shape))
            (gx2,) = torch.autograd.grad(outs, [leaf], grad_outputs=gs, allow_unused=True)
        a, b = st["big"] if st["big"] is not None else (n, n)
        small = torch.cat([buf[:a], buf[b:]])                  # ..., value last
        if gx2 is not None:
            small[:a + n - b] += torch.cat([gx2[:a], gx2[b:n]])
        st["small"] = small

    def _shard_exchange_2(self, st):
        self.dist.all_reduce(st["small"])
        if st.get("work") is not None:
            st["work"].wait()                # stream-orders the early all-reduce before stage C
            st["work"] = None

    def _shard_stage_c(self, st):
        from . import ops
        buf, n, small = st["buf"], st["n"], st["small"]
        a, b = st["big"] if st["big"] is not None else (n, n)
        buf[:a].copy_(small[:a])
        buf[b:].copy_(small[a:])
        if st["tri"]:
            ops.tri_expand_(st["early"], buf[a:b].view(st["tri"], st["tri"]))
        return buf[n:], buf[:n]

    def _objective_sharded(self, x, **kw):
        """(-f, -df/dx) summed over the ranks; every rank gets the full result."""
        self._prepare()
        if not self._use_graph:
            st = {}
            self._shard_stage_a(x, kw, st)
            self._shard_exchange_1(st)
            self._shard_stage_b(st)
            self._shard_exchange_2(st)
            return self._shard_stage_c(st)
        dyn = self._graph_inputs(kw)
        key = ("sharded",) + self._graph_key(x, dyn)
        if getattr(self, "_graph_epoch", None) != state_epoch():
            self._graphs = {}
            self._graph_epoch = state_epoch()
        ent = self._graphs.get(key)
        if ent is None:
            sx, sdyn = self._graph_static(x, dyn)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):      # warm-up (every rank runs the same collectives in the same order)
                    w = {}
                    self._shard_stage_a(sx, sdyn, w)
                    self._shard_exchange_1(w)
                    self._shard_stage_b(w)
                    self._shard_exchange_2(w)
                    self._shard_stage_c(w)
                del w
            cur.wait_stream(side)
            torch.cuda.synchronize()
            st = {}
            gA, _ = self._capture(lambda: self._shard_stage_a(sx, sdyn, st))
            self._shard_exchange_1(st)          # (on not-yet-computed buffers: keeps the ranks' NCCL order aligned)
            gB, _ = self._capture(lambda: self._shard_stage_b(st), pool=gA.pool())
            self._shard_exchange_2(st)
            gC, (f, g) = self._capture(lambda: self._shard_stage_c(st), pool=gA.pool())
            for k in ("leaf", "items"):         # the autograd graph is not needed to replay
                st[k] = None
            keep = [d._tensor for nd in self._all_parameterized() for d in nd.data_holders] + \
                   [p._fixed_dev for p in self.all_params() if p.fixed]
            ent = (gA, gB, gC, sx, sdyn, st, f, g, list(self._pending_infos), keep)
            self._graphs[key] = ent
        gA, gB, gC, sx, sdyn, st, f, g, infos, _keep = ent
        self._graph_refresh(sx, sdyn, x, dyn)
        trace = os.environ.get("GPINV_SHARD_TRACE")
        if trace:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
        gA.replay()
        if trace:
            ev[1].record()
        self._shard_exchange_1(st)
        if trace:
            ev[2].record()
        gB.replay()
        if trace:
            ev[3].record()
        self._shard_exchange_2(st)
        if trace:
            ev[4].record()
        gC.replay()
        if trace:
            ev[5].record()
            torch.cuda.synchronize()
            acc = self.__dict__.setdefault("_shard_trace", [])
            acc.append([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
            if len(acc) == int(trace):
                m_ = np.mean(np.array(acc[2:]), axis=0)
                print("shard-trace rank %d (ms): stage A %.3f | exchange 1 (blocking part) %.3f | stage B %.3f | exchange 2 %.3f | "
                      "stage C %.3f | sum %.3f" % ((self.dist.rank,) + tuple(m_) + (float(m_.sum()),)), file=sys.stderr, flush=True)
        self._pending_infos = list(infos)
        return f, g

    def _check_infos(self):
        infos, self._pending_infos = self._pending_infos, []
        for info in infos:
            i = int(info.reshape(-1)[0].item())
            if i != 0:
                raise RuntimeError("Cholesky decomposition was not successful: the matrix is not "
                                   "positive definite at pivot %d" % i)

    def _objective(self, x, **kw):
        """numpy in -> (float, numpy) out; CUDA tensor in -> (tensor[1], tensor) out."""
        if isinstance(x, torch.Tensor):
            return self._objective_device(x, **kw)
        x = np.ascontiguousarray(x, dtype=np.float64)
        dev = default_device()
        n = x.size
        hin = self._pinned.get(n)
        if hin is None:
            hin = torch.empty(n, dtype=torch.float64).pin_memory()
            self._pinned = {n: hin}
            self._out_pool = []
        hout, finish = self._result_buffer(n)
        if n >= self.OVERLAP_MIN:
            # large states: the transfers are pipelined against the evaluation, which is replayed from
            # CUDA graphs cut where the transfers slot in (enable_cuda_graph) or launched eagerly
            if self._use_graph:
                out = self._objective_host_pipeline(x, hin, hout, finish, dev, **kw)
                if out is not None:
                    return out
            return self._objective_host_overlapped(x, hin, hout, finish, dev, **kw)
        # torch's CPU copy is multi-threaded; a plain numpy assignment of 134 MB is not
        hout._gpinv_fresh_zero = False
        hout._gpinv_upper_zero = None
        hin.copy_(torch.from_numpy(x))
        xd = hin.to(dev, non_blocking=True)
        f, g = self._objective_device(xd, **kw)
        hout[:1].copy_(f, non_blocking=True)
        hout[1:].copy_(g, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._check_infos()
        return finish()

    POOL_MIN = 1 << 16        # below this many entries the gradient is copied into a fresh array

    def _result_buffer(self, n):
        """Pinned [1+n] buffer that receives (-f, -df/dx), and the function that turns it into the
        returned ``(float, ndarray)``.  For large states the ndarray is a *view* of the pinned
        buffer (a second 134 MB host copy would cost as much as the D2H transfer itself); a buffer
        goes back into circulation only after every array derived from it has been released,
        which a weak reference to the view's base detects, so results the caller still holds are
        never overwritten."""
        if n < self.POOL_MIN:
            if getattr(self, "_small_out", None) is None or self._small_out.numel() != n + 1:
                self._small_out = torch.empty(n + 1, dtype=torch.float64).pin_memory()
            hout = self._small_out

            def finish():
                return float(hout[0].item()), hout[1:].numpy().copy()
            return hout, finish
        pool = self._out_pool
        ent = None
        for e in pool:
            if e[0].numel() == n + 1 and (e[1] is None or e[1]() is None):
                ent = e
                break
        if ent is None:
            # page-locking 134 MB takes ~100 ms: allocate two at once the first time, because an
            # optimiser typically still holds the previous gradient while it asks for the next
            for _ in range(2 if not pool else 1):
                ent = [torch.zeros(n + 1, dtype=torch.float64).pin_memory(), None]
                ent[0]._gpinv_fresh_zero = True
                pool.append(ent)
        hout = ent[0]

        def finish():
            base = hout.numpy()           # keeps ``hout``'s storage alive through its own reference
            ent[1] = weakref.ref(base)
            return float(base[0]), base[1:]
        return hout, finish

    OVERLAP_MIN = 1 << 21     # elements of x above which the host path pipelines its transfers
    DEFER_MIN_N = 1024        # factor size from which the graph-replayed host path cuts before the Cholesky reverse

    def _objective_host_overlapped(self, x, hin, hout, finish, dev, **kw):
        """Host API for large states (the 134 MB x / gradient of the headline config): the
        largest parameter block (q_sqrt) is staged and uploaded by a helper thread in chunks while
        the kernel build and the Cholesky already run, and its gradient — ready as soon as the
        sampling stage's reverse pass is done — is downloaded on a copy stream while the Cholesky
        reverse sweep still runs."""
        trace = [] if os.environ.get("GPINV_E2E_TRACE") else None

        def tr(name):
            if trace is not None:
                trace.append((name, time.perf_counter()))
        tr("start")
        self._prepare()
        n = x.size
        params = [p for p in self.all_params() if not p.fixed]
        offs, o = [], 0
        for p in params:
            offs.append((o, o + p.size))
            o += p.size
        big = max(range(len(params)), key=lambda i: params[i].size)
        b0, b1 = offs[big]
        cur = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        cs = self._copy_stream
        # the big block lives in its own allocation: the upload thread writes it in place while
        # autograd already holds (version-counted) views of the other blocks
        xd = torch.empty(n - (b1 - b0), dtype=torch.float64, device=dev)
        xt = torch.from_numpy(x)
        nq = params[big]._tril_rows
        tri = nq is not None and nq * nq == b1 - b0 and nq >= 512
        ready = threading.Event()
        state = {"h2d": 0, "d2h": 0}

        if tri:
            # Only tril(q_sqrt) is read by the model: stage, transfer and scatter the lower
            # trapezoid of each block of rows (0.53 n^2 entries instead of n^2).
            xbig = torch.zeros(b1 - b0, dtype=torch.float64, device=dev)
            nblk = 16
            rb = -(-nq // nblk)
            total = sum((min(nq, r0 + rb) - r0) * min(nq, r0 + rb) for r0 in range(0, nq, rb))
            dstage = torch.empty(total, dtype=torch.float64, device=dev)
            src2, dst2 = xt[b0:b1].view(nq, nq), xbig.view(nq, nq)

            G = 1 if self.dist is None else self.dist.world_size
            rank = 0 if self.dist is None else self.dist.rank
            blocks, o = [], 0
            for r0 in range(0, nq, rb):
                r1 = min(nq, r0 + rb)
                blocks.append((r0, r1, o))
                o += (r1 - r0) * r1
            if G > 1:
                # Sample sharding: x is the same on every rank, so the ranks split the staging
                # work (block k goes to rank k mod G, biggest blocks dealt first) and ONE
                # all-reduce over NVLink of the zero-initialised packed buffer completes it on
                # every GPU: per-rank host traffic and PCIe bytes drop by G.
                dstage.zero_()
                order = sorted(range(len(blocks)), key=lambda k: -blocks[k][2])
                mine = set(k for j, k in enumerate(order) if j % G == rank)
            else:
                mine = set(range(len(blocks)))

            offs_np = np.array([o_ for _, _, o_ in blocks], dtype=np.int64)
            take_np = np.array([1 if k in mine else 0 for k in range(len(blocks))], dtype=np.int32)

            def upload_chunks():
                # gather (pageable -> pinned, multi-threaded, GIL released) and per-block DMA are
                # pipelined inside one native call
                from . import _capi
                nthr = max(1, min(int(os.environ.get("GPINV_UPLOAD_THREADS", "8")), torch.get_num_threads()))
                rc = _capi.lib().gpinv_upload_tril_blocks_f64(
                    x.ctypes.data + 8 * b0, nq, rb, hin.data_ptr() + 8 * b0, dstage.data_ptr(),
                    offs_np.ctypes.data, take_np.ctypes.data, len(blocks), nthr,
                    torch.cuda.current_stream().cuda_stream)
                _capi.check(rc, "upload_tril_blocks")
                for k, (r0, r1, o_) in enumerate(blocks):
                    if k in mine:
                        state["h2d"] += 8 * (r1 - r0) * r1
                if G > 1:
                    self.dist.all_reduce(dstage)
                for r0, r1, o in blocks:                            # strided scatter on the device
                    dst2[r0:r1, :r1].copy_(dstage[o:o + (r1 - r0) * r1].view(r1 - r0, r1))
        else:
            xbig = torch.empty(b1 - b0, dtype=torch.float64, device=dev)

            def upload_chunks():
                nch = 4
                step = -(-(b1 - b0) // nch)
                for c in range(nch):
                    a, b = b0 + c * step, min(b1, b0 + (c + 1) * step)
                    if b > a:
                        hin[a:b].copy_(xt[a:b])
                        xbig[a - b0:b - b0].copy_(hin[a:b], non_blocking=True)
                        state["h2d"] += 8 * (b - a)

        def upload():
            try:
                with torch.cuda.stream(cs):
                    upload_chunks()
                    ev = torch.cuda.Event()
                    ev.record(cs)
                state["ev"] = ev
                tr("upload_issued")
            except BaseException as e:   # pragma: no cover
                state["err"] = e
            finally:
                ready.set()

        cs.wait_stream(cur)       # the copy stream starts after xbig's allocation / zero fill
        th = threading.Thread(target=upload, daemon=True)
        th.start()
        # head and tail (everything but the big block): small, on the compute stream
        for a, b, d in ((0, b0, 0), (b1, n, b0)):
            if b > a:
                hin[a:b].copy_(xt[a:b])
                xd[d:d + b - a].copy_(hin[a:b], non_blocking=True)
                state["h2d"] += 8 * (b - a)

        def wait_big():
            tr("wait_big")
            ready.wait()
            tr("wait_big_done")
            if "err" in state:
                raise state["err"]
            torch.cuda.current_stream().wait_event(state["ev"])

        keep = {}

        def big_grad_hook(g):
            # runs on the autograd thread as soon as d(-f)/d(big block) exists
            tr("big_grad_hook")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            gf = g.reshape(-1)
            if self.dist is not None:
                # sum this block over the ranks right away (own all-reduce, same order on every
                # rank); the small remainder is reduced after the backward pass
                gf = gf.contiguous()
                self.dist.all_reduce(gf)
                ev.record(torch.cuda.current_stream())
            keep["g"] = gf
            with torch.cuda.stream(cs):
                cs.wait_event(ev)
                if tri and getattr(hout, "_gpinv_upper_zero", None) == (b0, nq):
                    # the strict upper triangle of this gradient is exactly zero and the pinned
                    # buffer's upper part has never been written: move the lower trapezoids only
                    from . import _capi
                    for r0, r1, _ in blocks:
                        off = 8 * (r0 * nq)
                        _capi.check(_capi.lib().gpinv_memcpy2d_async(
                            hout.data_ptr() + 8 * (1 + b0) + off, 8 * nq, gf.data_ptr() + off, 8 * nq,
                            8 * r1, r1 - r0, 2, cs.cuda_stream), "memcpy2d")
                        state["d2h"] += 8 * r1 * (r1 - r0)
                else:
                    hout._gpinv_upper_zero = None
                    hout[1 + b0:1 + b1].copy_(gf, non_blocking=True)
                    state["d2h"] += 8 * (b1 - b0)
            keep["done"] = True

        if tri and getattr(hout, "_gpinv_fresh_zero", False):
            hout._gpinv_upper_zero = (b0, nq)       # zero-filled pinned buffer, first use
        hout._gpinv_fresh_zero = False
        leaves = self.make_tensors(xd, override={big: xbig})
        params[big]._before_use = wait_big
        leaves[big].register_hook(big_grad_hook)
        self._pending_infos = []
        from . import ops
        col = None
        if self.dist is not None and self.dist.world_size > 1:
            # sample sharding: the Cholesky reverse waits for the reduce-scatter of Lbar (see _shard_stage_a)
            col = ops.DeferredReverse(lambda nn: self.dist.shard_width(nn) > 0)
        prev = ops.defer_reverse(col)
        try:
            with self.tensor_mode():
                f = self.build_likelihood(**kw)
                prior = self.build_prior()
                if prior is not None:
                    f = f + (prior if self.dist is None else prior / self.dist.world_size)
            nf = -f.reshape(())
            tr("forward_issued")
            grads = torch.autograd.grad(nf, leaves, allow_unused=True, retain_graph=col is not None)
            if col is not None and col.items:
                ctx = self.dist
                ops.trtri_side_join()
                outs, gs = [], []
                for (X_, ls_, mode_, L_, W_, lbar_) in col.items:
                    nn = L_.shape[0]
                    rs_in = ops.lbar_pack(lbar_, ctx.world_size)
                    rs_out = torch.empty(rs_in.numel() // ctx.world_size, dtype=torch.float64, device=dev)
                    ctx.reduce_scatter(rs_out, rs_in)
                    g_ = ops.chol_core_bwd_shard(X_, ls_, mode_, L_, W_, rs_out, ctx.shard_width(nn), ctx.shard_columns(nn))
                    outs.append(ls_)
                    gs.append(g_.reshape(ls_.shape))
                extra = torch.autograd.grad(outs, leaves, grad_outputs=gs, allow_unused=True)
                grads = tuple(g if e is None else (e if g is None else g + e) for g, e in zip(grads, extra))
        finally:
            ops.defer_reverse(prev)
            params[big]._before_use = None
            self.clear_tensors()
            th.join()
        rest = [(i, a, b) for i, (a, b) in enumerate(offs) if not (i == big and keep.get("done"))]
        if not keep.get("done"):
            hout._gpinv_upper_zero = None
        parts = [nf.detach().reshape(1)]
        for i, a, b in rest:
            parts.append(torch.zeros(b - a, dtype=torch.float64, device=dev) if grads[i] is None
                         else grads[i].reshape(-1))
        small = torch.cat(parts)
        if self.dist is not None:
            self.dist.all_reduce(small)
        hout[:1].copy_(small[:1], non_blocking=True)
        o = 1
        for i, a, b in rest:
            hout[1 + a:1 + b].copy_(small[o:o + b - a], non_blocking=True)
            o += b - a
        tr("grads_issued")
        cur.synchronize()
        cs.synchronize()
        tr("synced")
        self._check_infos()
        out = finish()
        self._last_h2d_bytes = state["h2d"]
        self._last_d2h_bytes = state["d2h"] + 8 * (1 + n - (b1 - b0))
        tr("done")
        if trace is not None:
            print("e2e-trace " + " ".join("%s=%.2f" % (k, 1e3 * (v - trace[0][1])) for k, v in trace),
                  file=sys.stderr)
        return out

    # ---- host API, large states, graph replay ------------------------------------------------------
    # One evaluation = three CUDA graphs, cut where the host transfers (and, under sample sharding, the
    # NCCL exchanges) slot in:
    #   G1  small parameter blocks -> kernel build, Cholesky, L^-1              ||  q_sqrt upload (copy stream)
    #   G2  sampling stage, likelihood, their reverse pass  -> Qbar, Lbar
    #   G3  Cholesky reverse -> hyperparameter gradients                        ||  Qbar download (copy stream)
    # The cut between G1 and G2 is made from the callback that fires when q_sqrt is first read; the cut
    # between G2 and G3 uses the deferred reverse of ops.DeferredReverse.
    def _objective_host_pipeline(self, x, hin, hout, finish, dev, **kw):
        from . import ops, _capi
        self._prepare()
        n = x.size
        params = [p for p in self.all_params() if not p.fixed]
        offs, o = [], 0
        for p in params:
            offs.append((o, o + p.size))
            o += p.size
        big = max(range(len(params)), key=lambda i: params[i].size)
        b0, b1 = offs[big]
        nq = params[big]._tril_rows
        if nq is None or nq * nq != b1 - b0 or nq < 512 or params[big].prior is not None:
            return None
        eps_given = kw.get("eps") is not None
        key = ("host", n, tuple(np.shape(kw["eps"])) if eps_given else None)
        if getattr(self, "_graph_epoch", None) != state_epoch():
            self._graphs = {}
            self._graph_epoch = state_epoch()
        ent = self._graphs.get(key)
        if ent is None:
            warm = self.__dict__.setdefault("_host_warm", {})
            if warm.get(key, 0) < 2:          # the eager pipelined path warms up plans, workspaces, NCCL
                warm[key] = warm.get(key, 0) + 1
                return None
            ent = self._host_pipeline_capture(n, params, offs, big, nq, self._graph_inputs(kw), dev)
            self._graphs[key] = ent
        st = ent
        ctx = st["ctx"]
        cur = torch.cuda.current_stream()
        cs = st["cs"]
        xt = torch.from_numpy(x)
        h2d = d2h = 0
        shared = st.get("shared")
        if shared is not None:
            # one result buffer for the node: this rank downloads its share of it (see dist.SharedHostBuffers)
            sb = shared.bufs[shared.acquire()]
            hout = sb[0]
        for a, b, d in ((0, b0, 0), (b1, n, b0)):           # small blocks: on the compute stream
            if b > a:
                hin[a:b].copy_(xt[a:b])
                st["xd"][d:d + b - a].copy_(hin[a:b], non_blocking=True)
                h2d += 8 * (b - a)
        if eps_given:
            e = kw["eps"]
            if not isinstance(e, torch.Tensor):
                e = torch.as_tensor(np.ascontiguousarray(e, dtype=np.float64))
            st["sdyn"]["eps"].copy_(e, non_blocking=True)
        else:
            self._refill_graph_inputs(st["sdyn"])
        st["g1"].replay()
        # q_sqrt: lower trapezoids of 16 row blocks, gathered into pinned memory and sent block by block by
        # one native call while G1 runs; under sample sharding every rank sends 1/G of the blocks and one
        # all-reduce over NVLink completes the packed buffer everywhere
        with torch.cuda.stream(cs):
            if ctx is not None:
                st["dstage"].zero_()
            nthr = max(1, min(int(os.environ.get("GPINV_UPLOAD_THREADS", "8")), torch.get_num_threads()))
            rc = _capi.lib().gpinv_upload_tril_blocks_f64(
                x.ctypes.data + 8 * b0, nq, st["rb"], hin.data_ptr() + 8 * b0, st["dstage"].data_ptr(),
                st["offs_np"].ctypes.data, st["take_np"].ctypes.data, len(st["blocks"]), nthr, cs.cuda_stream)
            _capi.check(rc, "upload_tril_blocks")
            h2d += st["h2d_big"]
            if ctx is not None:
                ctx.all_reduce(st["dstage"])
            dst2 = st["xbig"].view(nq, nq)
            for r0, r1, o_ in st["blocks"]:
                dst2[r0:r1, :r1].copy_(st["dstage"][o_:o_ + (r1 - r0) * r1].view(r1 - r0, r1))
            st["ev_up"].record(cs)
        cur.wait_event(st["ev_up"])
        st["g2"].replay()
        work = None
        if ctx is not None:
            self._inverse_slices_exchange(st, ctx)
            for o_, i_ in zip(st["rs_out"], st["rs_in"]):
                ctx.reduce_scatter(o_, i_)
            work = ctx.all_reduce_async(st["qc"])
        st["ev_q"].record(cur)
        gb = st["gb"]
        with torch.cuda.stream(cs):
            cs.wait_event(st["ev_q"])
            if work is not None:
                work.wait()
                ops.tri_expand_(st["qc"], gb)
            tri_only = shared is not None or getattr(hout, "_gpinv_fresh_zero", False) or \
                getattr(hout, "_gpinv_upper_zero", None) == (b0, nq)
            if tri_only:
                for k_, (r0, r1, _) in enumerate(st["blocks"]):
                    if shared is not None and not st["take_np"][k_]:
                        continue             # another rank's share of the shared buffer
                    off = 8 * (r0 * nq)
                    _capi.check(_capi.lib().gpinv_memcpy2d_async(
                        hout.data_ptr() + 8 * (1 + b0) + off, 8 * nq, gb.data_ptr() + off, 8 * nq,
                        8 * r1, r1 - r0, 2, cs.cuda_stream), "memcpy2d")
                    d2h += 8 * r1 * (r1 - r0)
                if shared is None:
                    hout._gpinv_upper_zero = (b0, nq)
            else:
                hout._gpinv_upper_zero = None
                hout[1 + b0:1 + b1].copy_(gb.reshape(-1), non_blocking=True)
                d2h += 8 * (b1 - b0)
            if shared is None:
                hout._gpinv_fresh_zero = False
        st["g3"].replay()
        small = st["small"]
        if ctx is not None:
            ctx.all_reduce(small)
        if shared is None or ctx.rank == 0:
            hout[:1].copy_(small[:1], non_blocking=True)
            o = 1
            for i, (a, b) in enumerate(offs):
                if i != big:
                    hout[1 + a:1 + b].copy_(small[o:o + b - a], non_blocking=True)
                    o += b - a
            d2h += 8 * small.numel()
        cur.synchronize()
        cs.synchronize()
        self._pending_infos = list(st["infos"])
        self._check_infos()
        self._last_h2d_bytes, self._last_d2h_bytes = h2d, d2h
        if shared is not None:
            shared.publish()
            base = hout.numpy()
            sb[2] = weakref.ref(base)
            return float(base[0]), base[1:]
        return finish()

    def _host_pipeline_capture(self, n, params, offs, big, nq, dyn, dev):
        import gc
        from . import ops
        ctx = self.dist if (self.dist is not None and self.dist.world_size > 1) else None
        G = 1 if ctx is None else ctx.world_size
        rank = 0 if ctx is None else ctx.rank
        b0, b1 = offs[big]
        st = {"ctx": ctx}
        st["xd"] = torch.zeros(n - (b1 - b0), dtype=torch.float64, device=dev)
        st["xbig"] = torch.zeros(b1 - b0, dtype=torch.float64, device=dev)
        rb = -(-nq // 16)
        blocks, o = [], 0
        for r0 in range(0, nq, rb):
            r1 = min(nq, r0 + rb)
            blocks.append((r0, r1, o))
            o += (r1 - r0) * r1
        st["rb"], st["blocks"] = rb, blocks
        st["dstage"] = torch.zeros(o, dtype=torch.float64, device=dev)
        order = sorted(range(len(blocks)), key=lambda k: -blocks[k][2])
        mine = set(k for j, k in enumerate(order) if j % G == rank)
        st["offs_np"] = np.array([o_ for _, _, o_ in blocks], dtype=np.int64)
        st["take_np"] = np.array([1 if k in mine else 0 for k in range(len(blocks))], dtype=np.int32)
        st["h2d_big"] = sum(8 * (r1 - r0) * r1 for k, (r0, r1, _) in enumerate(blocks) if k in mine)
        st["sdyn"] = sdyn = {k: v.detach().clone() for k, v in dyn.items()}
        st["cs"] = self._copy_stream if getattr(self, "_copy_stream", None) is not None else torch.cuda.Stream()
        self._copy_stream = st["cs"]
        st["ev_up"], st["ev_q"] = torch.cuda.Event(), torch.cuda.Event()
        accept = (lambda nn: ctx.shard_width(nn) > 0) if ctx is not None else (lambda nn: nn >= self.DEFER_MIN_N)
        col = ops.DeferredReverse(accept)
        g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        mode = {"capture_error_mode": "thread_local"}

        def cut_1():
            # first read of q_sqrt: everything captured so far does not depend on the upload
            if ops.SIDE_PENDING[0]:
                ops.trtri_side_join()
            g1.capture_end()
            g2.capture_begin(pool=g1.pool(), **mode)
            active[0] = g2
            if later is not None:
                # sample sharding: the upload is short, so L^-1 is built beside the sampling stage (second graph)
                # instead of in front of the upload wait (first graph)
                for L_, W_ in later:
                    ops.start_inverse(L_, W_)
                del later[:]

        active = [None]
        ops.SIDE_PENDING[0] = False         # (a join inside the capture may only wait for work forked inside it)
        torch.cuda.synchronize()
        cap = torch.cuda.Stream()
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        prev = ops.defer_reverse(col)
        later = [] if ctx is not None else None
        ops.INVERSE_LATER = later
        ops.INVERSE_SHARD = None if ctx is None else (ctx.rank, G)
        try:
            with torch.cuda.stream(cap):
                g1.capture_begin(**mode)
                active[0] = g1
                leaves = self.make_tensors(st["xd"], override={big: st["xbig"]})
                params[big]._before_use = cut_1
                self._pending_infos = []
                with self.tensor_mode():
                    f = self.build_likelihood(**sdyn)
                    prior = self.build_prior()
                    if prior is not None:
                        f = f + (prior if ctx is None else prior / G)
                if params[big]._before_use is not None:
                    raise RuntimeError("the largest parameter block was never read by build_likelihood")
                nf = -f.reshape(())
                grads = torch.autograd.grad(nf, leaves, retain_graph=True, allow_unused=True)
                gb = grads[big].reshape(nq, nq)
                if not gb.is_contiguous():
                    gb = gb.contiguous()
                st["gb"] = gb
                if ctx is not None:
                    st["rs_in"] = [ops.lbar_pack(it[5], G) for it in col.items]
                    st["rs_out"] = [torch.empty(r.numel() // G, dtype=torch.float64, device=dev) for r in st["rs_in"]]
                    st["qc"] = ops.tri_compact(gb)
                if ops.SIDE_PENDING[0]:
                    ops.trtri_side_join()
                if ctx is not None:
                    self._inverse_slices_stage(st, col.items, ctx, dev)
                g2.capture_end()
                active[0] = None
                if ctx is not None:             # the ranks' NCCL call order stays aligned with the replay's
                    ctx.all_reduce(st["dstage"])
                    self._inverse_slices_exchange(st, ctx)
                    for o_, i_ in zip(st["rs_out"], st["rs_in"]):
                        ctx.reduce_scatter(o_, i_)
                    ctx.all_reduce(st["qc"])
                g3.capture_begin(pool=g1.pool(), **mode)
                active[0] = g3
                extra = None
                if ctx is not None:
                    self._inverse_slices_finish(st, ctx)
                if col.items:
                    outs, gs = [], []
                    for k, (X_, ls_, mode_, L_, W_, lbar_) in enumerate(col.items):
                        if ctx is not None:
                            nn = L_.shape[0]
                            g_ = ops.chol_core_bwd_shard(X_, ls_, mode_, L_, W_, st["rs_out"][k], ctx.shard_width(nn),
                                                         ctx.shard_columns(nn))
                        else:
                            g_ = ops.chol_core_bwd(X_, ls_, mode_, L_, lbar_, W_, True)
                        outs.append(ls_)
                        gs.append(g_.reshape(ls_.shape))
                    extra = torch.autograd.grad(outs, leaves, grad_outputs=gs, allow_unused=True)
                parts = [nf.detach().reshape(1)]
                for i, (a, b) in enumerate(offs):
                    if i == big:
                        continue
                    g_ = grads[i]
                    if extra is not None and extra[i] is not None:
                        g_ = extra[i] if g_ is None else g_ + extra[i]
                    parts.append(torch.zeros(b - a, dtype=torch.float64, device=dev) if g_ is None else g_.reshape(-1))
                st["small"] = torch.cat(parts)
                g3.capture_end()
                active[0] = None
                if ctx is not None:
                    ctx.all_reduce(st["small"])
        except BaseException:
            if active[0] is not None:           # do not leave the stream in capture mode behind an error
                try:
                    with torch.cuda.stream(cap):
                        active[0].capture_end()
                except Exception:
                    pass
            raise
        finally:
            ops.defer_reverse(prev)
            ops.INVERSE_LATER = None
            ops.INVERSE_SHARD = None
            params[big]._before_use = None
            self.clear_tensors()
            if gc_was_on:
                gc.enable()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        st["g1"], st["g2"], st["g3"] = g1, g2, g3
        if ctx is not None and os.environ.get("GPINV_SHARED_RESULT", "1") != "0":
            from .dist import SharedHostBuffers
            try:
                st["shared"] = SharedHostBuffers(ctx, n + 1)
            except RuntimeError as e:       # raised on every rank together: per-rank result buffers instead
                if ctx.rank == 0:
                    print("gpinv_b200: %s; every rank downloads the whole gradient" % e, file=sys.stderr)
                st["shared"] = None
        st["infos"] = list(self._pending_infos)
        st["keep"] = [d._tensor for nd in self._all_parameterized() for d in nd.data_holders] + \
                     [p._fixed_dev for p in self.all_params() if p.fixed]
        return st

    def compute_log_likelihood(self, **kw):
        x = torch.as_tensor(self.get_free_state()).to(default_device())
        self._prepare()
        self.make_tensors(x)
        try:
            with self.tensor_mode(), torch.no_grad():
                f = self.build_likelihood(**kw)
        finally:
            self.clear_tensors()
        return float(f.reshape(-1)[0].item())

    def compute_log_prior(self):
        x = torch.as_tensor(self.get_free_state()).to(default_device())
        self.make_tensors(x)
        try:
            with self.tensor_mode(), torch.no_grad():
                p = self.build_prior()
        finally:
            self.clear_tensors()
        return 0.0 if p is None else float(p.item())

    # ---- evaluate a method in tensor mode at the current state (GPflow AutoFlow) ---------
    def _run(self, fn, *args, **kw):
        self._prepare()
        x = torch.as_tensor(self.get_free_state()).to(default_device())
        self.make_tensors(x)
        try:
            with self.tensor_mode(), torch.no_grad():
                out = fn(*args, **kw)
        finally:
            self.clear_tensors()
        self._check_infos()
        if isinstance(out, tuple):
            return tuple(o.detach().cpu().numpy() for o in out)
        return out.detach().cpu().numpy()

    # ---- optimisation ------------------------------------------------------------------
    def optimize(self, method="L-BFGS-B", tol=None, callback=None, maxiter=1000, **kw):
        """``method`` is a scipy method name or an optimizer descriptor from ``gpinv_b200.train``
        (the reference takes a ``tf.train`` optimizer there).  Signature mirrored from
        GPinv/model.py:94-95."""
        self._prepare()
        if isinstance(method, str):
            return self._optimize_np(method, tol, callback, maxiter, **kw)
        return self._optimize_torch(method, callback, maxiter, **kw)

    def _optimize_np(self, method, tol, callback, maxiter, **kw):
        from scipy.optimize import minimize
        options = dict(maxiter=maxiter)
        if "max_iters" in kw:
            options["maxiter"] = kw.pop("max_iters")
        options.update(kw)
        state = {"x": self.get_free_state()}

        def obj(x):
            f, g = self._objective(x)
            if np.isfinite(f) and np.all(np.isfinite(g)):
                state["x"] = x.copy()
            else:  # GPflow's ObjectiveWrapper: keep the optimiser alive on a failed evaluation
                return 1e100 if not np.isfinite(f) else f, np.where(np.isfinite(g), g, 0.0)
            return f, g

        try:
            result = minimize(fun=obj, x0=self.get_free_state(), method=method, jac=True, tol=tol,
                              callback=callback, options=options)
            self.set_state(result.x)
        except KeyboardInterrupt:
            print("Caught KeyboardInterrupt, setting model with most recent state.")
            self.set_state(state["x"])
            return None
        return result

    def _optimize_adam_device(self, method, callback, maxiter, callback_every=1, callback_device=False, **kw):
        """Training loop with TensorFlow's Adam update fused on the device (the loop shape of
        GPinv/model.py:206-240 driving tf.train.AdamOptimizer as in testing/test_stvgp.py:31-35): the
        state x, both moments and the step counter live in HBM, the draws are refilled in place, and
        with ``enable_cuda_graph()`` one step = ONE graph replay (evaluation + gradient packing + Adam
        update; nothing is copied, nothing synchronises).  ``callback`` receives the state every
        ``callback_every`` steps: a numpy copy (reference behaviour), or the device tensor itself
        with ``callback_device=True`` (no transfer)."""
        from . import ops
        dev = default_device()
        x = torch.as_tensor(self.get_free_state()).to(dev).contiguous()
        m, v = torch.zeros_like(x), torch.zeros_like(x)
        step = torch.zeros(1, dtype=torch.int64, device=dev)
        hp = (float(method.learning_rate), float(method.kw["betas"][0]), float(method.kw["betas"][1]),
              float(method.kw["eps"]))
        dyn = self._graph_inputs({})                 # StVGP: {"eps": draws}; GPMC: {}
        graph = None
        if self._use_graph and self.dist is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):            # warm-up on copies: lazy initialisation, plans
                xw = x.clone()
                for _ in range(2):
                    self._objective_eager(xw, reduce=False, **dyn)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            import gc
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    f, g = self._objective_eager(x, reduce=False, **dyn)
                    ops.adam_step(x, m, v, g.contiguous(), step, *hp)
            finally:
                if gc_was_on:
                    gc.enable()
            self._adam_graph = (graph, x, m, v, step, dyn, f, g)     # keeps the captured buffers alive
        it = 0
        f = None
        try:
            while it < maxiter:
                if it > 0 or graph is not None:
                    self._refill_graph_inputs(dyn)   # fresh draws, in place
                if graph is not None:
                    graph.replay()
                else:
                    f, g = self._objective_device(x, **dyn)
                    ops.adam_step(x, m, v, g.contiguous(), step, *hp)
                it += 1
                if callback is not None and it % max(1, int(callback_every)) == 0:
                    callback(x if callback_device else x.detach().cpu().numpy())
        except KeyboardInterrupt:
            print("Caught KeyboardInterrupt, setting model with most recent state.")
        xf = x.detach().cpu().numpy()
        self._check_infos()
        self.set_state(xf)
        fun, jac = self._objective(xf)
        return OptimizeResult(fun=fun, jac=jac, x=xf, nit=it, success=True, status="Finished iterations.")

    def _refill_graph_inputs(self, dyn):
        """New values for the evaluation's varying inputs, written in place (StVGP: eps draws)."""

    def _optimize_torch(self, method, callback, maxiter, **kw):
        from . import train
        if isinstance(method, train.AdamOptimizer):
            return self._optimize_adam_device(method, callback, maxiter, **kw)
        dev = default_device()
        x = torch.as_tensor(self.get_free_state()).to(dev)
        x.requires_grad_(False)
        param = torch.nn.Parameter(x, requires_grad=True)
        if hasattr(method, "make"):
            opt = method.make([param])
        elif isinstance(method, type) and issubclass(method, torch.optim.Optimizer):
            opt = method([param], **kw)
        elif callable(method):
            opt = method([param])
        else:
            raise ValueError("unsupported optimizer %r" % (method,))
        f = None
        it = 0
        try:
            while it < maxiter:
                f, g = self._objective_device(param.data)
                param.grad = g
                opt.step()
                if callback is not None:
                    callback(param.data.detach().cpu().numpy())
                it += 1
        except KeyboardInterrupt:
            print("Caught KeyboardInterrupt, setting model with most recent state.")
        xf = param.data.detach().cpu().numpy()
        self._check_infos()
        self.set_state(xf)
        fun, jac = self._objective(xf)
        return OptimizeResult(fun=fun, jac=jac, x=xf, nit=it, success=True, status="Finished iterations.")

    # ---- HMC (GPflow Model.sample -> hmc.sample_HMC) -----------------------------------------
    def sample(self, num_samples, Lmax=20, epsilon=0.01, thin=1, burn=0, verbose=False,
               return_logprobs=False, RNG=None):
        self._prepare()
        if RNG is None:
            RNG = np.random.RandomState(0)
        # device-resident chain: state, momentum and gradient never leave HBM between leapfrog steps
        out = hmc.sample_HMC_device(self._objective_device, num_samples, Lmax=Lmax, epsilon=epsilon,
                                    thin=thin, burn=burn, x0=self.get_free_state(),
                                    device=default_device(), verbose=verbose,
                                    return_logprobs=return_logprobs, RNG=RNG)
        # a proposal that left the positive-definite region shows up as NaN and was rejected by
        # the sampler (GPflow: "numerical instability"): it is not an error of the call
        self._pending_infos = []
        return out


class GPModel(Model):
    """GPflow ``GPModel``: holds X, Y, kern, likelihood, mean_function; ``predict_f``."""

    def __init__(self, X, Y, kern, likelihood, mean_function, name="model"):
        Model.__init__(self, name)
        self.kern, self.likelihood, self.mean_function = kern, likelihood, mean_function
        self.X, self.Y = X, Y

    def build_predict(self, Xnew, full_cov=False):
        raise NotImplementedError

    def predict_f(self, Xnew):
        """(mean [n2,R], var [n2,R]) at the current state (GPflow AutoFlow'd predict_f)."""
        Xn = torch.as_tensor(np.ascontiguousarray(Xnew, dtype=np.float64)).to(default_device())
        return self._run(self.build_predict, Xn)

    def predict_f_full_cov(self, Xnew):
        Xn = torch.as_tensor(np.ascontiguousarray(Xnew, dtype=np.float64)).to(default_device())
        return self._run(self.build_predict, Xn, full_cov=True)
