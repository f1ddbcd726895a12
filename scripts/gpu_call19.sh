GPINV_NO_SPLIT=1 CUDA_LAUNCH_BLOCKING=1 timeout 600 python tests/gpu_flake_probe.py 12 2>&1 | head -6
REPS=25 timeout 600 python tests/gpu_gemm_kmodes.py 3000 2>&1 | grep -E "split=0" | head -10
