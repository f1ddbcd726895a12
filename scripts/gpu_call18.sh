echo nosplit; GPINV_NO_SPLIT=1 timeout 600 python tests/gpu_flake_probe.py 12 2>&1 | head -2
echo blocking; CUDA_LAUNCH_BLOCKING=1 timeout 600 python tests/gpu_flake_probe.py 12 2>&1 | head -2
echo default; timeout 600 python tests/gpu_flake_probe.py 12 2>&1 | head -2
