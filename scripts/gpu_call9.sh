echo "=== A: no setmaxnreg (small)"; GPINV_LIB_PATH=$PWD/scratch_libs/libA.so timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | grep -E " ok| FAIL|failures"
echo "=== B: 3 stages (small)"; GPINV_LIB_PATH=$PWD/scratch_libs/libB.so timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | grep -E " ok| FAIL|failures"
