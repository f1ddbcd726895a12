mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "pipeline or two_rank or overlapped" > gpurun_out/r2q_pytest_a.log 2>&1; echo "pytest (subset) exit $?"; tail -12 gpurun_out/r2q_pytest_a.log | cut -c1-300
