timeout 900 python tests/gpu_gemm_stress.py 60 2>&1 | tail -30
