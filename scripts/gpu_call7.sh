timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | grep -E "ok|FAIL|failures"
echo ==== occ2; GPINV_TMA_SMALL_OCC=2 timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | grep -E "ok|FAIL|failures"
