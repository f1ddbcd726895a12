mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2u_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2u_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'], d['config']['launches_per_eval'])"
timeout 900 python tests/gpu_configs_bench.py > gpurun_out/r2u_configs.log 2>&1; cut -c1-330 gpurun_out/r2u_configs.log
