timeout 300 python tests/gpu_gemm_chain.py 3000 20
SPLIT=0 timeout 300 python tests/gpu_gemm_chain.py 3000 20
SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 20
timeout 300 python tests/gpu_gemm_chain.py 3072 20
timeout 300 python tests/gpu_gemm_chain.py 2500 20
