echo default; timeout 600 python tests/gpu_flake_probe.py 20
echo no-lookahead; GPINV_NO_LOOKAHEAD=1 timeout 600 python tests/gpu_flake_probe.py 20
echo legacy-gemm; GPINV_GEMM_LEGACY=1 timeout 600 python tests/gpu_flake_probe.py 20
python tests/gpu_potrf_bwd_bench.py 2>&1 | grep -v "rel err"
