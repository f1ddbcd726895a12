mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2_bench_launches.csv
python tests/gpu_abel_once.py > gpurun_out/r2_plain_abel.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 4 -c 2 -o gpurun_out/r2_abel_tma python tests/gpu_abel_once.py > gpurun_out/r2_ncu_abel.log 2>&1
echo "abel full exit $?"
python tests/gpu_spec_once.py > gpurun_out/r2_plain_spec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spec_ -s 4 -c 2 -o gpurun_out/r2_spec python tests/gpu_spec_once.py > gpurun_out/r2_ncu_spec.log 2>&1
echo "spec full exit $?"; cat gpurun_out/r2_plain_spec.log | tail -2
python tests/gpu_potrf_once.py 4096 fwd > gpurun_out/r2_plain_potrf.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 10 -c 1 -o gpurun_out/r2_panel python tests/gpu_potrf_once.py 4096 fwd > gpurun_out/r2_ncu_panel.log 2>&1
echo "panel full exit $?"
timeout 300 python tests/gpu_eval_timeline.py cfg5 > gpurun_out/r2_timeline.log 2>&1; echo "timeline exit $?"; head -2 gpurun_out/eval_timeline_cfg5.txt
ls -la gpurun_out/*.ncu-rep
