export GPINV_TMA_RAGGED=1
for round in 1 2; do SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 40 | tail -1; timeout 300 python tests/gpu_gemm_chain.py 2500 40 | tail -1; SPLIT=0 timeout 300 python tests/gpu_gemm_chain.py 3064 40 | tail -1; done
timeout 600 python tests/gpu_flake_probe.py 20
timeout 900 python tests/gpu_gemm_stress.py 80 2>&1 | tail -4
timeout 300 python tests/gpu_gemm_shapes.py | head -8
