"""Lengthscale-gradient error of the CUDA path against the longdouble arbiter at a full-size
BASELINE configuration (tests/golden/arb_cfg*.npz), for the reverse-sweep variant selected by the
environment.  python tests/gpu_ls_probe.py cfg4 [cfg2 cfg5]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpinv_oracle as O
from tests import baseline_configs as B
from tests.helpers import model_from_spec

tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("GPINV_"))
for name in sys.argv[1:]:
    path = os.path.join(os.path.dirname(__file__), "golden", "arb_%s.npz" % name)
    if not os.path.exists(path):
        continue
    d = np.load(path)
    c = B.case(name, O.make_los_matrix, O.make_cos_theta)
    m = model_from_spec(c["spec"], S=c["S"])
    kw = {} if c["eps"] is None else {"eps": c["eps"]}
    f, g = m._objective(c["x"], **kw)
    off = 0
    for k in c["order"]:
        if k == "kern.lengthscales":
            break
        off += int(np.size(c["constrained"][k]))
    print(json.dumps({"variant": tag or "default", "case": name, "ls_grad": float(g[off]), "arbiter": float(d["g_ls"]),
                      "rel_err": abs(float(g[off]) - float(d["g_ls"])) / abs(float(d["g_ls"])),
                      "f_rel_err": abs(f - float(d["f"])) / abs(float(d["f"]))}), flush=True)
