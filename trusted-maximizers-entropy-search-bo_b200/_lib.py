"""ctypes loader for libtes_b200.so -- the C-ABI shared library (include/tes_b200.h).

There is NO CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libtes_b200.so")
_lib = None


class TesError(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree libtes_b200.so (nvcc, no GPU needed)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    if not os.path.exists(SO_PATH):
        raise TesError("build finished but %s is missing" % SO_PATH)
    return SO_PATH


c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p
c_u64 = ctypes.c_uint64

# name -> (restype, argtypes); every symbol include/tes_b200.h declares
SIGNATURES = {
    "tes_abi_version": (c_int, []),
    "tes_error_string": (ctypes.c_char_p, [c_int]),
    "tes_rff_topk_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_int]),
    "tes_rff_topk_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_dbl, c_int,
                                 c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_rff_topk_method_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_dbl, c_int,
                                        c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_int]),
    "tes_product_runlen_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "tes_product_check_f64": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "tes_product_prefix_f64": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "tes_rff_topk_product_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int]),
    "tes_rff_topk_product_f64": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int,
                                         c_dbl, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr]),
    "tes_rff_topk_product_resolve_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_i64, c_int]),
    "tes_rff_topk_product_resolve_f64": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64,
                                                 c_int, c_dbl, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_ep_acq_screen_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_i64]),
    "tes_ep_acq_screen_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr,
                                      c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_rff_eval_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_dbl, c_ptr,
                                 c_ptr]),
    "tes_topk_merge_f64": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr]),
    "tes_argmax_workspace_bytes": (c_i64, [c_i64]),
    "tes_argmax_f64": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_se_ard_cross_f64": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_dbl, c_ptr, c_ptr]),
    "tes_gemm_f64": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr]),
    "tes_pnk_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64]),
    "tes_pnk_f64": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_posterior_stats_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64]),
    "tes_posterior_stats_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl,
                                        c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_gp_mean_var_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "tes_gp_mean_var_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr,
                                    c_ptr, c_i64, c_ptr]),
    "tes_batched_chol_workspace_bytes": (c_i64, [c_i64, c_i64, c_int]),
    "tes_batched_potrf_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_batched_chol2inv_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_ep_acq_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_i64]),
    "tes_ep_acq_f64": (c_int, [c_int, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_dbl,
                               c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_int, c_int, c_ptr, c_u64, c_i64, c_i64,
                               c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_sp_acq_workspace_bytes": (c_i64, [c_i64, c_int, c_i64, c_i64, c_i64, c_i64, c_i64]),
    "tes_sp_acq_f64": (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_dbl,
                               c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_u64, c_i64, c_i64,
                               c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_batch_ep_acq_f64": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "tes_rff_draw_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64]),
    "tes_rff_draw_f64": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_i64, c_i64,
                                 c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_uniform_fill_f64": (c_int, [c_u64, ctypes.c_uint32, ctypes.c_uint32, c_u64, c_i64, c_ptr, c_ptr]),
    "tes_batched_gj_inverse_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "tes_ep_moments_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "tes_ep_moments_f64": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_rff_adam_refine_f64": (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_dbl, c_ptr,
                                        c_ptr, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr]),
    "tes_normal_fill_f64": (c_int, [c_u64, ctypes.c_uint32, ctypes.c_uint32, c_u64, c_i64, c_ptr, c_ptr]),
    "tes_dedup_f64": (c_int, [c_ptr, c_i64, c_i64, c_dbl, c_ptr, c_ptr, c_ptr]),
    "tes_sym_eig_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "tes_sym_eig_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_color_factor_f64": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "tes_joint_maxidx_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "tes_joint_maxidx_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "tes_bucket_moments_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "tes_bucket_samples_f64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr,
                                       c_ptr, c_ptr]),
    "tes_cosf_error_probe": (c_int, [ctypes.c_float, ctypes.c_float, c_dbl, c_int, c_ptr, c_ptr]),
    "tes_mma_arg_error_probe": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    "tes_tc5_arg_error_probe": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    "tes_launch_count": (c_i64, []),
    "tes_mufu_cos_burn": (c_int, [c_int, c_i64, c_ptr, ctypes.POINTER(c_dbl), c_ptr]),
    "tes_fp64_mma_burn": (c_int, [c_int, c_i64, c_ptr, ctypes.POINTER(c_dbl), c_ptr]),
    "tes_fp32_fma_burn": (c_int, [c_int, c_i64, c_ptr, ctypes.POINTER(c_dbl), c_ptr]),
    "tes_fp64_fma_burn": (c_int, [c_int, c_i64, c_ptr, ctypes.POINTER(c_dbl), c_ptr]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise TesError(
                "libtes_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.tes_abi_version() != 1:
            raise TesError("libtes_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().tes_error_string(int(rc)).decode()
        raise TesError("%s failed: code %d (%s)" % (what, rc, msg))
