"""Runs the reference's OWN unit tests (/root/reference/testing/test_kernels.py, test_kronecker.py,
test_mean.py, test_stvgp.py, test_gpmc.py -- every reference test that touches the StVGP / GPMC
path) with the TensorFlow / GPflow stand-ins of oracle/refshim on the import path.  TEST
INFRASTRUCTURE ONLY; needs /root/reference (build container).

    python -m oracle.run_reference_tests [module ...]

That these pass is the evidence that the stand-ins compute what the reference expects of
TensorFlow and GPflow: the kernel / Kronecker / mean tests compare against the numpy loops inside
the reference's test files, test_stvgp.py::test_optimize trains the reference's StVGP to the exact
GP-regression optimum (objective -11.4568) and compares its predictions with it.
test_model.py / test_param.py exercise HierarchicModel / LocalParam, which are out of scope
 and are not run."""
import os
import sys
import time
import unittest
import warnings

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
MODULES = ["test_kernels", "test_kronecker", "test_mean", "test_stvgp", "test_gpmc"]


def main(argv):
    if not os.path.isdir(os.path.join(REF, "testing")):
        print("reference sources not present")
        return 2
    sys.path[:0] = [os.path.join(REF, "testing"), REF, os.path.join(ROOT, "oracle", "refshim")]
    warnings.simplefilter("ignore")
    bad = 0
    for mod in (argv or MODULES):
        t0 = time.time()
        suite = unittest.defaultTestLoader.loadTestsFromName(mod)
        with open(os.devnull, "w") as sink:
            old = sys.stdout
            sys.stdout = sink                      # the reference tests print their intermediate values
            try:
                res = unittest.TextTestRunner(verbosity=0, stream=sink).run(suite)
            finally:
                sys.stdout = old
        print("%-16s ran %2d  failures %d  errors %d  (%.1f s)" % (mod, res.testsRun, len(res.failures),
                                                                  len(res.errors), time.time() - t0))
        for _, tb in res.failures + res.errors:
            print(tb[-1500:])
        bad += len(res.failures) + len(res.errors) + (0 if res.testsRun else 1)
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
