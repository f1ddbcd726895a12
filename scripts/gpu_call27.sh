mkdir -p gpurun_out
for i in 1 2; do timeout 600 python tests/gpu_selfcheck.py > gpurun_out/r2g_selfcheck_$i.log 2>&1; echo "selfcheck $i exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2g_selfcheck_$i.log | head -5; done
timeout 600 python tests/gpu_flake_probe.py 15
timeout 300 python tests/gpu_gemm_inplace.py | tail -2
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2g_pytest.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cfg5 or full_size" > gpurun_out/r2g_pytest2.log 2>&1; echo "pytest2 exit $?"; tail -3 gpurun_out/r2g_pytest2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'])"
