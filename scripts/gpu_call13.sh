GPINV_BWD_BLOCK=2048 GPINV_LEAF_REFINE=0 python tests/gpu_ls_probe.py cfg4 cfg2
GPINV_BWD_BLOCK=512 GPINV_LEAF_REFINE=0 python tests/gpu_ls_probe.py cfg4 cfg2
GPINV_BWD_BLOCK=512 GPINV_BWD_LEAF_MURRAY=1 python tests/gpu_ls_probe.py cfg4
GPINV_BWD_BLOCK=256 GPINV_BWD_LEAF_MURRAY=1 python tests/gpu_ls_probe.py cfg4
GPINV_BWD_BLOCK=128 GPINV_LEAF_REFINE=0 python tests/gpu_ls_probe.py cfg4
GPINV_BWD_BLOCK=64 python tests/gpu_ls_probe.py cfg4
GPINV_BWD_BLOCK=64 GPINV_BWD_NO_REFINE=1 python tests/gpu_ls_probe.py cfg4
