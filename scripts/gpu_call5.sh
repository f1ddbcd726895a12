mkdir -p gpurun_out
timeout 300 python tests/gpu_gemm_inplace.py 2>&1 | tee gpurun_out/r2c_inplace.log | tail -30
GPINV_GEMM_LEGACY=1 timeout 300 python tests/gpu_selfcheck.py check_potrf 2>&1 | grep -E "FAIL|elapsed" | head
