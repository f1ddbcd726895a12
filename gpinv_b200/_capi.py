"""ctypes binding of ``libgpinv_sm100.so`` (C ABI declared in ``include/gpinv_b200.h``).

There is deliberately no CPU fallback: if the library is missing, or a compute entry point
is called without a CUDA device, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPINV_LIB_PATH: load an alternative build of the library (kernel experiments only)
LIB_PATH = os.environ.get("GPINV_LIB_PATH") or os.path.join(_HERE, "lib", "libgpinv_sm100.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gpinv_b200.h")
CSRC_DIR = os.path.join(_HERE, "csrc")

_lib = None

_P = C.c_void_p
_I = C.c_int
_D = C.c_double
_Z = C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/gpinv_b200.h
SIGNATURES = {
    "gpinv_version": (_I, []),
    "gpinv_last_error": (C.c_char_p, []),
    "gpinv_set_workspace": (_I, [_P, _Z]),
    "gpinv_memcpy2d_async": (_I, [_P, _Z, _P, _Z, _Z, _Z, _I, _P]),
    "gpinv_upload_tril_blocks_f64": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _I, _P]),
    "gpinv_kbuild_ws_bytes": (_Z, [_I, _I, _I]),
    "gpinv_kbuild_f64": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _I, _D, _P, _I, _P, _P]),
    "gpinv_kbuild_bwd_f64": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P, _I, _P, _P, _P]),
    "gpinv_potrf_f64": (_I, [_P, _I, _I, _P, _P]),
    "gpinv_potrf_bwd_ws_bytes": (_Z, [_I]),
    "gpinv_potrf_bwd_f64": (_I, [_P, _I, _P, _I, _I, _P, _P]),
    "gpinv_chol_core_ws_bytes": (_Z, [_I, _I]),
    "gpinv_chol_core_f64": (_I, [_P, _I, _I, _I, _P, _I, _I, _D, _P, _I, _P, _P, _P]),
    "gpinv_chol_core_bwd_ws_bytes": (_Z, [_I, _I]),
    "gpinv_chol_core_bwd_f64": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _I, _P, _I, _P, _P, _P]),
    "gpinv_trtri_side_shard_f64": (_I, [_P, _I, _P, _P, _I, _I, _I, _P]),
    "gpinv_chol_core_inv_f64": (_I, [_P, _I, _I, _I, _P, _I, _I, _D, _P, _I, _P, _P, _P, _P, _P]),
    "gpinv_chol_core_bwd2_f64": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P]),
    "gpinv_trtri_side_f64": (_I, [_P, _I, _P, _P, _I, _P]),
    "gpinv_trtri_side_join": (_I, [_P]),
    "gpinv_trtri_side_sync": (_I, []),
    "gpinv_trsm_rlt_f64": (_I, [_P, _I, _P, _I, _I, _I, _P]),
    "gpinv_gemm_f64": (_I, [_I, _I, _I, _I, _I, _D, _P, _I, _P, _I, _D, _P, _I, _I, _P]),
    "gpinv_gemm_ex_f64": (_I, [_I, _I, _I, _I, _I, _D, _P, _I, C.c_longlong, _P, _I, C.c_longlong, _D, _P, _I,
                               C.c_longlong, _I, _I, _I, _I, _P]),
    "gpinv_tril_unpack_f64": (_I, [_P, _P, _I, _I, _P]),
    "gpinv_tril_pack_f64": (_I, [_P, _P, _I, _I, _P]),
    "gpinv_sample_fwd_f64": (_I, [_P, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gpinv_sample_bwd_f64": (_I, [_P, _I, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I,
                                  _P, _P, _P, _P, _I, _P, _P, _P]),
    "gpinv_kron_apply_f64": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _I, _P]),
    "gpinv_kron_sample_fwd_f64": (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "gpinv_kron_bwd_ws_bytes": (_Z, [_I, _I, _I, _I]),
    "gpinv_kron_sample_bwd_f64": (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I,
                                       _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gpinv_gauss_sumsq_f64": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "gpinv_gauss_bwd_f64": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "gpinv_abel_fwd_f64": (_I, [_P, _P, _I, _P, _I, _I, _I, _P, _P, _P]),
    "gpinv_abel_bwd_f64": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _P, _P]),
    "gpinv_spec_fwd_f64": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _D, _P, _P, _P]),
    "gpinv_spec_bwd_f64": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _D, _P, _P]),
    "gpinv_unpack_f64": (_I, [_P, _I, _P, _P, _P, _P]),
    "gpinv_pack_f64": (_I, [_P, _I, _P, _P, _P, _D, _P, _P]),
    "gpinv_kl_white_f64": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "gpinv_kl_white_bwd_f64": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "gpinv_adam_step_f64": (_I, [_P, _P, _P, _P, C.c_longlong, _D, _D, _D, _D, _P, _P]),
    "gpinv_side_mark_inputs": (_I, [_P]),
    "gpinv_side_early_join": (_I, [_P]),
    "gpinv_sample_u_early_f64": (_I, [_P, _I, _P, _P, _I, _I, _I, _P, _P, _I]),
    "gpinv_sample_fwd2_f64": (_I, [_P, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "gpinv_gauss_logp_f64": (_I, [_P, _P, _D, _P, _P]),
    "gpinv_gauss_logp_bwd_f64": (_I, [_P, _P, _D, _P, _P, _P, _P]),
    "gpinv_lbar_pack_f64": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "gpinv_chol_core_bwd_shard_f64": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _I, _P, _P, _I, _P, _I, _P, _P, _P]),
    "gpinv_tri_compact_f64": (_I, [_P, _I, _I, _P, _P]),
    "gpinv_tri_expand_f64": (_I, [_P, _P, _I, _I, _P]),
    "gpinv_sumsq_f64": (_I, [_P, _Z, _P, _P]),
}


def header_symbols():
    """Function names declared in include/gpinv_b200.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(gpinv_[a-z0-9_]+)\s*\(", text)))


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libgpinv_sm100.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "gpinv_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().gpinv_last_error()
        raise RuntimeError("gpinv_b200: %s failed (rc=%d): %s" % (what, rc, (msg or b"").decode()))
