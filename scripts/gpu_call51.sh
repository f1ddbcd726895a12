mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_shard > gpurun_out/r2z_shard.log 2>&1; echo "check_shard exit $?"; grep -E "FAIL|elapsed|Error|sharded inv" gpurun_out/r2z_shard.log | cut -c1-200 | head -20; tail -5 gpurun_out/r2z_shard.log | cut -c1-200
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "pipeline or two_rank or overlapped" > gpurun_out/r2z_pytest_a.log 2>&1; echo "pytest (subset) exit $?"; tail -5 gpurun_out/r2z_pytest_a.log | cut -c1-300
