GPINV_GEMM_LOG=1 timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | tail -12
echo ==== force big; GPINV_TMA_FORCE=big GPINV_GEMM_LOG=1 timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | tail -12
echo ==== nosplit small; GPINV_NO_SPLIT=1 timeout 300 python tests/gpu_gemm_inplace.py x 2>&1 | tail -12
