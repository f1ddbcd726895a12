mkdir -p gpurun_out
timeout 300 python tests/gpu_gemm_inplace.py 2>&1 | grep -E " FAIL|failures"
for i in 1 2 3; do timeout 600 python tests/gpu_selfcheck.py > gpurun_out/r2d_selfcheck_$i.log 2>&1; echo "selfcheck $i exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2d_selfcheck_$i.log | head -10; done
timeout 300 python tests/gpu_gemm_shapes.py > gpurun_out/r2d_gemm_tma.log 2>&1; cat gpurun_out/r2d_gemm_tma.log
timeout 1200 python -m pytest tests -m gpu -q -k "not cfg3b_small_full" -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r2d_pytest.log
grep -E "lengthscale gradient|arbiter " gpurun_out/r2d_pytest.log | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo']); print(d['denominators'])"
