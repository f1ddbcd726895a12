set -x
mkdir -p gpurun_out
timeout 300 python tests/gpu_selfcheck.py check_gemm > gpurun_out/r2c_gemm.log 2>&1; echo "gemm exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2c_gemm.log | head -40
GPINV_NO_SPLIT=1 timeout 300 python tests/gpu_selfcheck.py check_gemm > gpurun_out/r2c_gemm_nosplit.log 2>&1; echo "gemm nosplit exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2c_gemm_nosplit.log | head -40
GPINV_NO_LOOKAHEAD=1 timeout 300 python tests/gpu_selfcheck.py check_potrf > gpurun_out/r2c_potrf_nola.log 2>&1; echo "potrf nola exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2c_potrf_nola.log | head -20
GPINV_POTRF_B_CAP=0 timeout 300 python tests/gpu_selfcheck.py check_potrf > gpurun_out/r2c_potrf_nocap.log 2>&1; echo "potrf nocap exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2c_potrf_nocap.log | head -20
