mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_shard > gpurun_out/r2k_shard.log 2>&1; echo "check_shard exit $?"; grep -E "FAIL|elapsed|Error|shard n=" gpurun_out/r2k_shard.log | head -30
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "two_rank or graph or fused or dist" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2k_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['neg_elbo'], d['config']['launches_per_eval'], d['gpu_launches'])"
