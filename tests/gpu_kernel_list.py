"""Kernel list of ONE evaluation of a small configuration (torch profiler, eager launches): which launches are
the library's own kernels and which are torch elementwise / fill kernels.  usage: gpu_kernel_list.py [n M S]"""
import os, sys, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from tests.test_parity_gpu import make_abel_stvgp
from tests.helpers import perturbed_state

n, M, S = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (30, 40, 10)
m = make_abel_stvgp(n, M, S)
dev = torch.device("cuda:0")
x = torch.as_tensor(perturbed_state(m)).to(dev)
eps = torch.randn((1, S, n), dtype=torch.float64, device=dev)
for _ in range(3):
    m._objective_device(x, eps=eps)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m._objective_device(x, eps=eps)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
print("n=%d M=%d S=%d: %d device activities" % (n, M, S, len(ev)))
for e in ev:
    nm = e.name
    tag = "gpinv" if "gpinv::" in nm else ("memset/memcpy" if nm.startswith("Mem") else "torch")
    print("%-14s %6.1f us  %s" % (tag, e.device_time if hasattr(e, "device_time") else e.cuda_time, nm[:150]))
