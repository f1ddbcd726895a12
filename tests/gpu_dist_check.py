"""Multi-GPU parity at the headline size (cfg5), launched under torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/gpu_dist_check.py [cfg5|cfg2|mid] [all|host]

Every rank evaluates its share of the Monte-Carlo samples; the result (eager, CUDA-graph replay and the
host API) is compared on rank 0 with the committed compact oracle fixture of the configuration
(tests/golden/oc_*.npz) and, for the lengthscale gradient, with the longdouble arbiter
(tests/golden/arb_*.npz).  Exit code 1 on any mismatch.  Nothing here reads /root/reference."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    what = sys.argv[2] if len(sys.argv) > 2 else "all"          # "host": only the graph-replayed host API
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from gpinv_b200 import dist as gdist
    from gpinv_b200.param import set_default_device
    set_default_device(dev)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=300))
    from oracle import gpinv_oracle as O
    from tests import baseline_configs as B
    from tests.test_parity_gpu import _baseline_model, _hyper_tols, _arbitrated, LS_VS_ARBITER
    c = B.case(name, O.make_los_matrix, O.make_cos_theta)
    m = _baseline_model(c)
    if world > 1:
        gdist.attach(m)
    n_core = c["spec"]["X"].shape[0]
    if rank == 0:
        print("config %s, %d ranks, reverse %s" % (name, world, "sharded w=%d" % m.dist.shard_width(n_core)
                                                   if world > 1 and m.dist.shard_width(n_core) else "replicated"), flush=True)
    x = torch.as_tensor(c["x"]).to(dev)
    eps = torch.as_tensor(c["eps"]).to(dev)
    d = np.load(os.path.join(ROOT, "tests", "golden", "oc_%s.npz" % name))
    mask = B.hyper_mask(c)
    fails = []

    def check(label, f, g):
        f, g = float(f), np.array(g)
        try:
            B.check_compact(f, g, d, mask, f_tol=1e-8, g_tol=1e-8, hyper_tol=_hyper_tols(c))
            g_arb, i_ls, da = _arbitrated(name, c, g)
            msg = ""
            if i_ls is not None:
                e = abs(g[i_ls] - float(da["g_ls"])) / abs(float(da["g_ls"]))
                msg = "lengthscale gradient %.10g vs arbiter %.10g (rel.err %.1e)" % (g[i_ls], float(da["g_ls"]), e)
                assert e <= LS_VS_ARBITER, msg
                assert abs(f - float(da["f"])) <= 2e-9 * abs(float(da["f"]))
            if rank == 0:
                print("%-28s ok  f=%.12g  %s" % (label, f, msg), flush=True)
        except AssertionError as e:
            fails.append(label)
            if rank == 0:
                print("%-28s FAIL %s" % (label, str(e)[:300]), flush=True)

    for graph in ((False, True) if what == "all" else ()):
        m.enable_cuda_graph(graph, static_inputs=graph)
        for rep in range(2):
            f, g = m._objective_device(x, eps=eps)
            torch.cuda.synchronize()
            check("device graph=%s rep %d" % (graph, rep), f.item(), g.cpu().numpy())
        # timing, max over ranks
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m._objective_device(x, eps=eps)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("device graph=%s: %.3f ms per evaluation (%.1f evals/s)" % (graph, tt.item(), 1e3 / tt.item()), flush=True)
    eps_dev = torch.as_tensor(c["eps"]).to(dev)
    for graph in ((False, True) if what == "all" else (True,)):
        # graph=True: calls 0 and 1 take the eager pipelined path, call 2 captures, calls 3.. replay
        m.enable_cuda_graph(graph)
        for rep in range(5 if graph else 2):
            fh, gh = m._objective(c["x"], eps=eps_dev)
            check("host API graph=%s rep %d" % (graph, rep), fh, gh)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            m._objective(c["x"], eps=eps_dev)
        dist.barrier()
        if rank == 0:
            print("host API graph=%s: %.3f ms per evaluation" % (graph, 1e3 * (time.perf_counter() - t0) / 5), flush=True)
    bad = torch.tensor([len(fails)], device=dev)
    dist.all_reduce(bad)
    dist.destroy_process_group()
    sys.exit(1 if int(bad.item()) else 0)


if __name__ == "__main__":
    main()
