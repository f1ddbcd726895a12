#!/bin/bash
# One-GPU validation on a B200 box:  gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh'
# op-level checks, the GPU test suite, smoke(), the bench line (with the CPU baseline) and the other configurations.
mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py > gpurun_out/selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed" gpurun_out/selfcheck.log | head
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'cpu', d.get('cpu_baseline', {}).get('value'), 'neg_elbo', d['config']['neg_elbo'], d['config']['launches_per_eval'])"
timeout 900 python tests/gpu_configs_bench.py > gpurun_out/configs.log 2>&1; cut -c1-400 gpurun_out/configs.log
