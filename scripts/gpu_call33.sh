mkdir -p gpurun_out
echo "--- right-looking inner (default)"; timeout 300 python tests/gpu_potrf_time.py 4096 2>&1 | tail -5
echo "--- left-looking inner"; GPINV_POTRF_INNER_LEFT=1 timeout 300 python tests/gpu_potrf_time.py 4096 2>&1 | tail -5
GPINV_POTRF_INNER_LEFT=1 timeout 300 python tests/gpu_selfcheck.py check_potrf 2>&1 | grep -E "FAIL|elapsed"
echo "--- e2e trace"; GPINV_E2E_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep e2e-trace | tail -4
