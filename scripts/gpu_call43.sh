mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_gemm check_sample check_liks check_kron > gpurun_out/r2t_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2t_selfcheck.log | head
timeout 300 python tests/gpu_gemm_stress.py > gpurun_out/r2t_stress.log 2>&1; echo "stress exit $?"; tail -3 gpurun_out/r2t_stress.log
python tests/gpu_abel_once.py > gpurun_out/r2t_plain_abel.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tma -s 4 -c 2 --csv --log-file gpurun_out/r2t_abel_traffic.csv python tests/gpu_abel_once.py > gpurun_out/r2t_ncu_abel.log 2>&1
echo "traffic exit $?"; grep -v "^==" gpurun_out/r2t_abel_traffic.csv | awk -F'","' '{print $1, $(NF-2), $NF}' | cut -c1-200 | head -12
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_alone'], d['config']['neg_elbo'])"
