export GPINV_TMA_RAGGED=1
for round in 1 2; do
echo "=== base"; SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 40 | tail -1
for v in NOSETMAXNREG STAGES4 NOPROMO; do echo "=== $v"; GPINV_LIB_PATH=$PWD/scratch_libs/lib_$v.so SYNC=1 timeout 300 python tests/gpu_gemm_chain.py 3000 40 | tail -1; done
done
