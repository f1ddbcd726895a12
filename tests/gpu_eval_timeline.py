"""In-situ kernel timeline of ONE ELBO+grad evaluation at the headline configuration (torch
profiler, CUDA activities: real concurrent execution, warm caches), with the geometry of every
DMMA GEMM (GPINV_GEMM_LOG) matched to its kernel by launch order.  Not a test: writes
gpurun_out/eval_timeline.txt, which profiles/ summarises."""
import os, sys, re, io, json, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if os.environ.get("GPINV_TIMELINE_CHILD") != "1":
    env = dict(os.environ, GPINV_TIMELINE_CHILD="1", GPINV_GEMM_LOG="1")
    r = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, stderr=subprocess.PIPE, text=True)
    logs = [l for l in r.stderr.splitlines() if l.startswith("gemm-log")]
    other = [l for l in r.stderr.splitlines() if not l.startswith("gemm-log")]
    print("\n".join(other[-20:]), file=sys.stderr)
    if r.returncode:
        sys.exit(r.returncode)
    marks = [i for i, l in enumerate(logs) if l.endswith("MARK")]
    logs = logs[marks[0] + 1:marks[1]]
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    rows = json.load(open("gpurun_out/eval_timeline_%s.json" % cfg))
    gi = 0
    with open("gpurun_out/eval_timeline_%s.txt" % cfg, "w") as f:
        span = rows[-1][0] + rows[-1][1]
        f.write("# one ELBO+grad evaluation, %s: %d kernels, GPU span %.1f us, sum of kernel durations %.1f us\n"
                % (cfg, len(rows), span, sum(r_[1] for r_ in rows)))
        f.write("# start_us  dur_us  [TF/s executed]  kernel | gemm geometry (launch order)\n")
        for st, du, nm in rows:
            extra = ""
            if ("gemm_dmma" in nm or "gemm_tma" in nm) and gi < len(logs):
                g = dict(kv.split("=") for kv in logs[gi].split()[1:] if "=" in kv)
                gi += 1
                Mm, Nn, Kk, b = int(g["M"]), int(g["N"]), int(g["K"]), int(g["batch"])
                ef, km = int(g["ef"]), int(g["kmode"])
                fl = 2.0 * Mm * Nn * Kk * b
                if ef & 128:
                    fl *= 0.5
                if km in (1, 2, 3, 4):
                    fl *= 0.5 if not (ef & 128) else 2.0 / 3.0
                if km == 5:
                    fl *= 1.0 / 3.0 if (ef & 128) else 0.5
                extra = " | %s ~%.1f TF/s" % (logs[gi - 1][9:], fl / du / 1e6)
            f.write("%10.1f %9.1f  %s%s\n" % (st, du, nm[:110], extra))
    print("gemm kernels matched: %d of %d logs" % (gi, len(logs)))
    sys.exit(0)

import numpy as np
import torch
import bench
from torch.profiler import profile, ProfilerActivity

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n, M, S = bench.CONFIGS[cfg]
dev = torch.device("cuda:0")
m = bench.build_model(n, M, S)
x = torch.as_tensor(m.get_free_state()).to(dev)
eps = torch.randn((1, S, n), dtype=torch.float64, device=dev)
for _ in range(3):
    m._objective_device(x, eps=eps)
torch.cuda.synchronize()
sys.stderr.write("gemm-log MARK\n")
sys.stderr.flush()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m._objective_device(x, eps=eps)
    torch.cuda.synchronize()
sys.stderr.write("gemm-log MARK\n")
sys.stderr.flush()
evs = [e for e in prof.events() if e.device_type.name == "CUDA" and "emcpy" not in e.name and "emset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
out = io.StringIO()
total = evs[-1].time_range.end - t0
out.write("# one evaluation cfg=%s: %d kernels, GPU span %.1f us, sum of kernel durations %.1f us\n" % (cfg, len(evs), total, sum(e.time_range.end - e.time_range.start for e in evs)))
rows = []
for e in evs:
    rows.append((e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name))
json.dump(rows, open("gpurun_out/eval_timeline_%s.json" % cfg, "w"))
for st, du, nm in rows:
    out.write("%10.1f %9.1f  %s\n" % (st, du, nm[:150]))
open("gpurun_out/eval_timeline_%s.txt" % cfg, "w").write(out.getvalue())
print("wrote timeline: span %.1f us" % total)
