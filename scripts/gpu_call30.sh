mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py > gpurun_out/r2j_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed|Error" gpurun_out/r2j_selfcheck.log | head -12
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2j_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'], d['config']['launches_per_eval']); print(d['hbm_stages'])"
timeout 900 python tests/gpu_configs_bench.py > gpurun_out/r2j_configs.log 2>&1; cat gpurun_out/r2j_configs.log | cut -c1-330
