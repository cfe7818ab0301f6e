"""ctypes binding of libpfb200.so (the C ABI declared in include/pfb200.h).

The shared library is built in-tree (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C qnmh_sysid2018_b200/csrc`).  There is no CPU fallback: a missing library or a missing
CUDA device raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFB200_LIB") or os.path.join(HERE, "csrc", "libpfb200.so")  # env: tuning builds

MODEL_LGSS, MODEL_SV, MODEL_SVL, MODEL_LGSS_FA = 0, 1, 2, 3
F64, F32 = 0, 1
RESAMPLE_SYSTEMATIC, RESAMPLE_STRATIFIED, RESAMPLE_MULTINOMIAL = 0, 1, 2
FLAG_STORE_HISTORY, FLAG_OBS_PER_FILTER, FLAG_SVL_OBS_MODEL, FLAG_INJECT_PER_FILTER = 1, 2, 4, 8
FLAG_STORE_STATES = 16
ABI_VERSION = 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint64)
_H = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/pfb200.h declares
SIGNATURES = {
    "pfb200_model_no_params": (C.c_int, [C.c_int]),
    "pfb200_abi_version": (C.c_int, []),
    "pfb200_last_error": (C.c_char_p, []),
    "pfb200_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_H)]),
    "pfb200_destroy": (C.c_int, [_H]),
    "pfb200_set_stream": (C.c_int, [_H, C.c_void_p]),
    "pfb200_set_option": (C.c_int, [_H, C.c_char_p, C.c_longlong]),
    "pfb200_get_info": (C.c_int, [_H, C.c_char_p, C.POINTER(C.c_longlong)]),
    "pfb200_configure": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                   C.c_int, C.c_uint, _dp, _dp, C.c_uint64, _up, _dp, _dp, _dp, _dp]),
    "pfb200_update_params": (C.c_int, [_H, _dp, C.c_uint64, _up]),
    "pfb200_execute": (C.c_int, [_H]),
    "pfb200_synchronize": (C.c_int, [_H]),
    "pfb200_last_execute_ms": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "pfb200_last_kernel_ms": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "pfb200_fetch": (C.c_int, [_H, _dp, _dp, _dp, _dp, _ip, _dp]),
    "pfb200_fetch_steps": (C.c_int, [_H, _dp, _dp, _dp]),
    "pfb200_fetch_history": (C.c_int, [_H, _dp, _dp, _ip]),
    "pfb200_run": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                             C.c_uint, _dp, _dp, C.c_uint64, _up, _dp, _dp, _dp, _dp,
                             _dp, _dp, _dp, _dp, _ip, _dp]),
    "pfb200_dist_handle_bytes": (C.c_int, []),
    "pfb200_dist_export": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "pfb200_dist_connect": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "pfb200_dist_disconnect": (C.c_int, [_H]),
    "pfb200_dist_connect_local": (C.c_int, [C.POINTER(_H), C.c_int]),
    "pfb200_dist_traj_count": (C.c_int, [_H, _ip]),
    "pfb200_dist_traj_lookup": (C.c_int, [_H, C.c_int32, _ip]),
    "pfb200_debug_profile": (C.c_int, [_H, C.POINTER(C.c_longlong), C.c_int]),
    "pfb200_philox_normals": (C.c_int, [_H, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _dp]),
    "pfb200_philox_uniforms": (C.c_int, [_H, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _dp]),
}

_lib = None


class Pfb200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("pfb200 error %d: %s" % (code, message))
        self.code = code


def load():
    """Loads libpfb200.so and declares all prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libpfb200.so not found at %s -- build it with __graft_entry__.build(); "
                          "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pfb200_abi_version() != ABI_VERSION:
        raise ImportError("libpfb200.so ABI version %d != binding version %d"
                          % (lib.pfb200_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise Pfb200Error(code, load().pfb200_last_error().decode("utf-8", "replace"))
