set -x
mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py > gpurun_out/r2b_selfcheck.log 2>&1; echo "selfcheck exit $?"
grep -c " ok" gpurun_out/r2b_selfcheck.log; grep "FAIL" gpurun_out/r2b_selfcheck.log | head -40; tail -5 gpurun_out/r2b_selfcheck.log
timeout 300 python tests/gpu_gemm_shapes.py > gpurun_out/r2b_gemm_tma.log 2>&1; echo "shapes exit $?"; cat gpurun_out/r2b_gemm_tma.log
GPINV_GEMM_LEGACY=1 timeout 300 python tests/gpu_gemm_shapes.py > gpurun_out/r2b_gemm_legacy.log 2>&1; cat gpurun_out/r2b_gemm_legacy.log
timeout 900 python -m pytest tests -m gpu -q -x -k "not cfg3b_small_full" -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'], d['denominators'])"
