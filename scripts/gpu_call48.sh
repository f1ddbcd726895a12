mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py check_liks > gpurun_out/r2x_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed|recurrence" gpurun_out/r2x_selfcheck.log | cut -c1-150 | head -50
timeout 300 python tests/gpu_spec_once.py 2>&1 | tail -2
timeout 300 python tests/gpu_spec_once.py direct 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "spec or cfg3a or reference_executed or sample_F" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2x_pytest.log
