mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2v_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench exit $?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2_bench_launches.csv
python tests/gpu_abel_once.py > gpurun_out/r2_plain_abel.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 4 -c 2 -o gpurun_out/r2_abel_tma -f python tests/gpu_abel_once.py > gpurun_out/r2_ncu_abel.log 2>&1
echo "abel full exit $?"
timeout 300 python tests/gpu_eval_timeline.py cfg5 > gpurun_out/r2_timeline.log 2>&1; echo "timeline exit $?"; head -1 gpurun_out/eval_timeline_cfg5.txt
timeout 300 python tests/gpu_potrf_time.py 4096 2>&1 | tail -4
timeout 280 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference arm exit $?"; cat gpurun_out/r2_bench_reference_arm.json | cut -c1-400
