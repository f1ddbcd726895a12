mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2y_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2y_pytest.log
python tests/gpu_spec_once.py > gpurun_out/r2_plain_spec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spec_ -s 4 -c 2 -o gpurun_out/r2_spec_rec -f python tests/gpu_spec_once.py > gpurun_out/r2_ncu_spec.log 2>&1
echo "spec full exit $?"; tail -2 gpurun_out/r2_plain_spec.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['config']['neg_elbo'])"
timeout 900 python tests/gpu_configs_bench.py > gpurun_out/r2y_configs.log 2>&1; cut -c1-330 gpurun_out/r2y_configs.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
