"""Probe of the host-side costs of the public `_objective(numpy) -> (float, numpy)` API at the
headline state size (134 MB each way): staging memcpy, pinned DMA, pageable DMA, result copy."""
import os, sys, time, json
import numpy as np
import torch
N = 16781316
x = np.random.randn(N)
dev = torch.device("cuda:0")
hin = torch.empty(N, dtype=torch.float64).pin_memory()
hout = torch.empty(N, dtype=torch.float64).pin_memory()
xd = torch.empty(N, dtype=torch.float64, device=dev)
res = {"threads": torch.get_num_threads(), "cpus": os.cpu_count()}

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3

xt = torch.from_numpy(x)
res["stage_torch_copy_ms"] = t(lambda: hin.copy_(xt))
res["stage_np_copyto_ms"] = t(lambda: np.copyto(hin.numpy(), x))
res["h2d_pinned_ms"] = t(lambda: xd.copy_(hin, non_blocking=True))
res["h2d_pageable_ms"] = t(lambda: xd.copy_(xt))
res["d2h_pinned_ms"] = t(lambda: hout.copy_(xd, non_blocking=True))
g = np.empty(N)
gt = torch.from_numpy(g)
res["d2h_pageable_ms"] = t(lambda: gt.copy_(xd))
def fresh():
    gg = np.empty(N); torch.from_numpy(gg).copy_(hout); return gg
res["result_fresh_alloc_copy_ms"] = t(fresh)
res["result_copy_prealloc_ms"] = t(lambda: gt.copy_(hout))
for nt in (4, 8, 16):
    torch.set_num_threads(nt)
    res["stage_torch_copy_%dthr_ms" % nt] = t(lambda: hin.copy_(xt))
# host register of the user's own array
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
def reg():
    r = rt.cudaHostRegister(ctypes.c_void_p(x.ctypes.data), ctypes.c_size_t(x.nbytes), 0)
    assert r == 0, r
    rt.cudaHostUnregister(ctypes.c_void_p(x.ctypes.data))
try:
    res["host_register_unregister_ms"] = t(reg, 3)
except Exception as e:
    res["host_register_error"] = repr(e)
print(json.dumps(res, indent=1))
