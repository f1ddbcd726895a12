mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2n_pytest.log
