"""ctypes binding of libacsfit.so (include/acsfit.h).  No CPU fallback: if the shared object is
missing or cannot be loaded this module raises, and every product entry point with it."""
import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libacsfit.so")

OK, E_INVALID, E_CUDA, E_DOMAIN, E_TIMEOUT, E_NOMEM, E_PEER = 0, -1, -2, -3, -4, -5, -6
_STATUS_NAMES = {E_INVALID: "ACSFIT_E_INVALID", E_CUDA: "ACSFIT_E_CUDA", E_DOMAIN: "ACSFIT_E_DOMAIN",
                 E_TIMEOUT: "ACSFIT_E_TIMEOUT", E_NOMEM: "ACSFIT_E_NOMEM", E_PEER: "ACSFIT_E_PEER"}

c_i64, c_int, c_vp, c_u64 = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64

# every symbol include/acsfit.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "acsfit_abi_version": (c_int, []),
    "acsfit_last_error": (ctypes.c_char_p, [c_vp]),
    "acsfit_ctx_create": (c_int, [c_int, ctypes.POINTER(c_vp)]),
    "acsfit_ctx_destroy": (c_int, [c_vp]),
    "acsfit_ctx_configure": (c_int, [c_vp, c_int, c_int]),
    "acsfit_ctx_set_timing": (c_int, [c_vp, c_int]),
    "acsfit_ctx_set_knob": (c_int, [c_vp, ctypes.c_char_p, c_int]),
    "acsfit_feasible_mask": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    "acsfit_occupancy": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "acsfit_first_fit_nodes": (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "acsfit_fulfill_pending": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_i64,
                                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acsfit_node_states": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int,
                                   c_vp, c_int, c_vp, c_vp]),
    "acsfit_maintain_actions": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "acsfit_scale_up": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64,
                                c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acsfit_scale_up_host": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64,
                                     c_vp, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acsfit_occupancy_host": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_vp]),
    "acsfit_maintain_host": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp,
                                     c_i64, c_int, c_int, c_i64, c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "acsfit_cluster_init": (c_int, [c_vp, c_int, c_int, c_i64, c_i64, c_int, c_vp]),
    "acsfit_cluster_connect": (c_int, [c_vp, c_vp]),
    "acsfit_cluster_barrier": (c_int, [c_vp, c_vp]),
    "acsfit_cluster_last_plan": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acsfit_launch_count": (c_u64, [c_vp]),
    "acsfit_debug_profile": (c_int, [c_vp, c_int, c_vp, c_int, c_vp]),
    "acsfit_debug_trace": (c_int, [c_vp, c_int, c_vp, c_int]),
    "acsfit_last_pipeline_stats": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
}


class AcsfitError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (_STATUS_NAMES.get(status, "status %d" % status), message))
        self.status = status


_lib = None


def load():
    """load libacsfit.so (raises if it has not been built: there is no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m kubernetes_acs_engine_autoscaler_b200.build` "
                "(nvcc, sm_100a). The pod-fit path has no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
