mkdir -p gpurun_out
for n in 3000 2500 3064 3072; do SYNC=1 timeout 300 python tests/gpu_gemm_chain.py $n 15 | tail -1; done
timeout 600 python tests/gpu_flake_probe.py 10
timeout 600 python tests/gpu_selfcheck.py > gpurun_out/r2f_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2f_selfcheck.log | head
python tests/gpu_potrf_bwd_bench.py 2>&1 | grep -v "rel err"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'])"
python tests/gpu_eval_timeline.py cfg5 > gpurun_out/r2f_timeline.log 2>&1; tail -3 gpurun_out/r2f_timeline.log
