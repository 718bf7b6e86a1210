"""ctypes binding of libacetn_b200.so (C ABI declared in include/acetn_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C acetn_b200/csrc`.  There is no CPU
fallback: if the shared object is missing, or no sm_100a device is present when a compute entry point is
called, this module raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacetn_b200.so")

c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_dbl = ctypes.c_double
P_i64 = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); must list every symbol of include/acetn_b200.h
SIGNATURES = {
    "acetn_b200_init": (c_int, [c_int]),
    "acetn_b200_destroy": (c_int, []),
    "acetn_b200_last_error": (ctypes.c_char_p, []),
    "acetn_b200_version": (ctypes.c_char_p, []),
    "acetn_b200_launch_count": (c_i64, []),
    "acetn_b200_reset_launch_count": (None, []),
    "acetn_b200_gemm_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64, c_i64, P_i64, c_int, c_int]),
    "acetn_b200_gemm": (c_int, [c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, P_i64, c_dbl, c_dbl, c_int, c_int, c_vp, c_sz, c_vp]),
    "acetn_b200_quarter_tensor_workspace_bytes": (c_sz, [c_i64] * 6),
    "acetn_b200_quarter_tensor": (c_int, [c_vp, c_vp, c_vp, c_vp, P_i64] + [c_i64] * 6 + [c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_quarter_tensor_enc": (c_int, [c_vp, c_vp, c_vp, c_vp, P_i64] + [c_i64] * 6 + [c_vp, c_vp, c_vp, c_sz, c_vp, c_sz, c_vp]),
    "acetn_b200_rsvd_workspace_bytes": (c_sz, [c_int, P_i64, P_i64, c_i64]),
    "acetn_b200_rsvd": (c_int, [c_int, ctypes.POINTER(c_vp), P_i64, P_i64, c_vp, c_i64, c_int, c_int, c_i64, c_dbl,
                                c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_orthonormalize_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "acetn_b200_orthonormalize": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "acetn_b200_jacobi_svd_workspace_bytes": (c_sz, [c_i64]),
    "acetn_b200_jacobi_svd": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_dbl, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_projectors_workspace_bytes": (c_sz, [c_i64] * 5),
    "acetn_b200_projectors_from_usv": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                               c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_absorb_corner_workspace_bytes": (c_sz, [c_i64] * 5),
    "acetn_b200_absorb_corner1": (c_int, [c_vp, c_vp, c_vp] + [c_i64] * 5 + [c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_absorb_corner2": (c_int, [c_vp, c_vp, c_vp] + [c_i64] * 5 + [c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_absorb_edge_workspace_bytes": (c_sz, [c_i64] * 6),
    "acetn_b200_absorb_edge": (c_int, [c_vp, c_vp, P_i64, c_vp, c_vp] + [c_i64] * 6 + [c_int, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_absorb_edge_begin_workspace_bytes": (c_sz, [c_i64] * 5),
    "acetn_b200_absorb_edge_begin": (c_int, [c_vp, c_vp, P_i64, c_vp] + [c_i64] * 5 + [c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_absorb_edge_finish_workspace_bytes": (c_sz, [c_i64] * 4),
    "acetn_b200_absorb_edge_finish": (c_int, [c_vp, c_vp] + [c_i64] * 4 + [c_int, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_double_layer_workspace_bytes": (c_sz, [c_i64] * 4),
    "acetn_b200_double_layer": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, P_i64, c_int, c_vp, P_i64, c_i64, c_i64, c_vp, c_i64,
                                        c_i64, P_i64, c_vp, c_sz, c_vp]),
    "acetn_b200_i8_supported": (c_int, [c_i64, c_i64, c_i64]),
    "acetn_b200_i8_encoded_bytes": (c_sz, [c_i64, c_i64]),
    "acetn_b200_i8_encode": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "acetn_b200_i8_matmul_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "acetn_b200_i8_matmul": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "acetn_b200_rsvd_enc_workspace_bytes": (c_sz, [c_int, P_i64, P_i64, c_i64, ctypes.POINTER(ctypes.c_int32)]),
    "acetn_b200_rsvd_enc": (c_int, [c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), P_i64, P_i64, c_vp, c_i64, c_int, c_int, c_i64, c_dbl,
                                    c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_projectors_enc_workspace_bytes": (c_sz, [c_i64] * 5 + [c_int]),
    "acetn_b200_projectors_from_usv_enc": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                                   c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_fp64_peak_probe": (c_dbl, [c_vp, c_int, c_vp]),
    "acetn_b200_als_workspace_bytes": (c_sz, [c_i64] * 3),
    "acetn_b200_als_solve": (c_int, [c_vp] * 5 + [c_i64] * 4 + [c_dbl, c_dbl, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_site_rdm_workspace_bytes": (c_sz, [P_i64, c_i64, c_i64]),
    "acetn_b200_site_rdm": (c_int, [c_vp] * 9 + [P_i64, P_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_bond_rdm_workspace_bytes": (c_sz, [P_i64, c_i64, c_i64]),
    "acetn_b200_bond_rdm": (c_int, [c_vp] * 6 + [P_i64] + [c_vp] * 6 + [P_i64, P_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_norm_tensor_workspace_bytes": (c_sz, [P_i64, c_i64, c_i64]),
    "acetn_b200_norm_tensor": (c_int, [c_vp] * 12 + [P_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "acetn_b200_permute": (c_int, [c_vp, c_vp, c_int, P_i64, P_i64, c_vp]),
    "acetn_b200_absmax": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "acetn_b200_frob_normalize": (c_int, [c_vp, c_i64, c_vp, c_sz, c_vp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the signatures.  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"acetn_b200: {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C acetn_b200/csrc`. There is no CPU fallback for backend='b200'.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def available():
    return os.path.exists(LIB_PATH)


def check(status, what):
    if status != 0:
        msg = load().acetn_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"acetn_b200.{what} failed (status {status}): {msg}")


def i64_array(values):
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])
