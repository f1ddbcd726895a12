mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_liks > gpurun_out/r3d_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed|gauss_logp" gpurun_out/r3d_selfcheck.log | cut -c1-150 | head
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r3d_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r3d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r3d_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['neg_elbo'], d['config']['launches_per_eval'])"
timeout 900 python tests/gpu_configs_bench.py > gpurun_out/r3d_configs.log 2>&1; cut -c1-400 gpurun_out/r3d_configs.log
