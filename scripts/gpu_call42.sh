mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_gemm check_sample check_liks > gpurun_out/r2s_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2s_selfcheck.log | head
python tests/gpu_abel_once.py > gpurun_out/r2s_plain_abel.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tma -s 4 -c 2 --csv --log-file gpurun_out/r2s_abel_traffic.csv python tests/gpu_abel_once.py > gpurun_out/r2s_ncu_abel.log 2>&1
echo "traffic exit $?"; grep -v "^==" gpurun_out/r2s_abel_traffic.csv | cut -d, -f5,13- | head -12
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_alone'], d['config']['neg_elbo'])"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2s_pytest.log
