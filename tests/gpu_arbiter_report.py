"""CUDA hyper-gradients against the np.longdouble arbiter (tests/golden/arb_*.npz), printed per
case for the reverse-sweep variant selected by the environment (GPINV_POTRF_BWD / GPINV_BWD_BLOCK
are read once per process, so run this script once per variant).  Output: one JSON line per case."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.arbiter import ARB_CASES, arb_case
from tests.helpers import model_from_spec

tag = "bwd=%s block=%s" % (os.environ.get("GPINV_POTRF_BWD", "default"), os.environ.get("GPINV_BWD_BLOCK", "default"))
for name in ARB_CASES:
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "arb_%s.npz" % name))
    case, spec, x, eps, nh = arb_case(name)
    m = model_from_spec(spec, S=eps.shape[1])
    f, g = m._objective(x, eps=eps)
    err = np.abs(np.array(g[:nh]) - d["g_hyper"]) / np.abs(d["g_hyper"])
    print(json.dumps({"variant": tag, "case": name, "f_rel_err": abs(f - float(d["f"])) / abs(float(d["f"])),
                      "cuda_hyper_rel_err": err.tolist(), "oracle_hyper_rel_err": d["oracle_err"].tolist(),
                      "oracle_f_rel_err": float(d["oracle_f_err"])}), flush=True)
