export GPINV_TMA_RAGGED=1
echo "=== base"; SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 15 | tail -1
for v in NOSETMAXNREG STAGES4 PROXYFENCE NOPROMO; do echo "=== $v"; GPINV_LIB_PATH=$PWD/scratch_libs/lib_$v.so SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 15 | tail -1; done
echo "=== base nosplit"; SPLIT=0 SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 15 | tail -1
