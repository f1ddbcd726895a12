for v in 0 1 2 3; do GPINV_LEAF_REFINE=$v python tests/gpu_ls_probe.py cfg4 cfg2; done
GPINV_LEAF_REFINE=3 GPINV_BWD_NO_REFINE=1 python tests/gpu_ls_probe.py cfg4 cfg2
GPINV_LEAF_REFINE=2 GPINV_BWD_NO_REFINE=1 python tests/gpu_ls_probe.py cfg4 cfg2
GPINV_BWD_LEAF_MURRAY=1 GPINV_BWD_NO_REFINE=1 python tests/gpu_ls_probe.py cfg4
python tests/gpu_arbiter_report.py | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], ['%.1e'%e for e in d['cuda_hyper_rel_err']], 'oracle', ['%.1e'%e for e in d['oracle_hyper_rel_err']])"
timeout 300 python tests/gpu_selfcheck.py check_potrf 2>&1 | grep -E "FAIL|elapsed"
python tests/gpu_potrf_bwd_bench.py 2>&1 | tail -5
