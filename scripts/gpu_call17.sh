ONLY_BATCHED=1 timeout 600 python tests/gpu_gemm_kmodes.py 2500 3000 2>&1 | tail -24
