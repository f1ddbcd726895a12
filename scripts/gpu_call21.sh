for n in 3008 3064 2960 3040 2504; do SYNC=1 timeout 300 python tests/gpu_gemm_chain.py $n 15 | tail -1; done
