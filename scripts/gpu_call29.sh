mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_gemm check_sample check_potrf check_kron > gpurun_out/r2i_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2i_selfcheck.log | head -5
timeout 600 python tests/gpu_gemm_kmodes.py 2500 4096 2>&1 | grep -E "FAIL|failures"
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'])"
python tests/gpu_eval_timeline.py cfg5 > gpurun_out/r2i_timeline.log 2>&1; tail -2 gpurun_out/r2i_timeline.log
