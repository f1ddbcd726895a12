#!/bin/bash
# ncu evidence on ONE GPU (every ncu command runs only after the same command exited 0 without ncu):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh'     then summarise with profiles/summarize_launches.py and
#   ncu -i gpurun_out/<name>.ncu-rep --page raw --csv / --page details here.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/bench_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
python tests/gpu_abel_once.py > gpurun_out/plain_abel.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 4 -c 2 -o gpurun_out/abel_tma -f python tests/gpu_abel_once.py > gpurun_out/ncu_abel.log 2>&1
echo "abel full exit $?"
python tests/gpu_spec_once.py > gpurun_out/plain_spec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spec_ -s 4 -c 2 -o gpurun_out/spec -f python tests/gpu_spec_once.py > gpurun_out/ncu_spec.log 2>&1
echo "spec full exit $?"
python tests/gpu_potrf_once.py 4096 fwd > gpurun_out/plain_potrf.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 10 -c 1 -o gpurun_out/panel -f python tests/gpu_potrf_once.py 4096 fwd > gpurun_out/ncu_panel.log 2>&1
echo "panel full exit $?"
timeout 300 python tests/gpu_eval_timeline.py cfg5 > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
