mkdir -p gpurun_out
timeout 300 python tests/gpu_selfcheck.py check_kron check_potrf check_sample 2>&1 | grep -E "FAIL|elapsed|Error|error" | head -30
python tests/gpu_ls_probe.py cfg4 cfg2 cfg5
python tests/gpu_potrf_bwd_bench.py 2>&1 | grep -v "rel err"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/r2e_pytest.log
grep -E "lengthscale gradient:" gpurun_out/r2e_pytest.log | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'])"
python tests/gpu_eval_timeline.py cfg5 > gpurun_out/r2e_timeline.txt 2>&1; tail -60 gpurun_out/r2e_timeline.txt
