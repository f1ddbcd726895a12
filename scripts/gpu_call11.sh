python tests/gpu_ls_probe.py cfg4
GPINV_BWD_TRSM_SUBST=1 python tests/gpu_ls_probe.py cfg4
GPINV_BWD_LEAF_MURRAY=1 python tests/gpu_ls_probe.py cfg4
GPINV_BWD_LEAF_MURRAY=1 GPINV_BWD_TRSM_SUBST=1 python tests/gpu_ls_probe.py cfg4
GPINV_POTRF_BWD=murray python tests/gpu_ls_probe.py cfg4
GPINV_POTRF_BWD=inv python tests/gpu_ls_probe.py cfg4
GPINV_GEMM_LEGACY=1 python tests/gpu_ls_probe.py cfg4
