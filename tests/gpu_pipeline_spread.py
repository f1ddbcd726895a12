"""Run-to-run spread of the hyperparameter gradients through the plain host path and the graph-replayed
host pipeline (same x, same draws): separates summation-order noise from a race."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_parity_gpu import make_abel_stvgp
from tests.helpers import perturbed_state

n = int(sys.argv[1]) if len(sys.argv) > 1 else 640


def run(defer_min, graph, reps):
    m = make_abel_stvgp(n, 300, 16)
    m.kern.lengthscales = 0.3
    x = perturbed_state(m, 0.02, seed=6)
    eps = np.random.RandomState(3).randn(1, 16, n)
    if graph:
        m.OVERLAP_MIN = 1000
        m.DEFER_MIN_N = defer_min
        m.enable_cuda_graph()
    out = []
    for k in range(reps):
        f, g = m._objective(x, eps=eps)
        out.append(np.concatenate([[f], np.array(g[:4])]))
    return np.array(out)


for label, args in (("plain eager", (0, False, 20)), ("pipeline, reverse in G2", (1 << 30, True, 40)),
                    ("pipeline, reverse in G3", (512, True, 40))):
    a = run(*args)
    a = a[3:] if args[1] else a
    print("%-26s f %.12g spread %.1e | dls mean %.10g min %.10g max %.10g (rel spread %.1e) | other hypers rel spread %s"
          % (label, a[:, 0].mean(), np.ptp(a[:, 0]) / abs(a[:, 0].mean()), a[:, 1].mean(), a[:, 1].min(), a[:, 1].max(),
             np.ptp(a[:, 1]) / abs(a[:, 1].mean()), np.array2string(np.ptp(a[:, 2:], axis=0) / np.abs(a[:, 2:].mean(axis=0)), precision=1)))
