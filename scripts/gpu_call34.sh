mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "pipeline or two_rank or overlapped" > gpurun_out/r2m_pytest_a.log 2>&1; echo "pytest (new) exit $?"; tail -15 gpurun_out/r2m_pytest_a.log
GPINV_E2E_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(d['value'], d['e2e'], d['roofline']['frac'], d['config']['neg_elbo'], d['config']['launches_per_eval'], d['gpu_launches'])"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2m_pytest.log
