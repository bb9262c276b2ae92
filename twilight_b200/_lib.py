"""ctypes binding of libtwilight_b200.so (the C ABI of include/twilight_b200.h). Loading fails loudly: there is no
Python or CPU fallback for any entry point."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# TWL_LIB points the binding at another build of the same library (A/B timing of kernel variants, tools/dp_ab.py)
LIB_PATH = os.environ.get("TWL_LIB") or os.path.join(PKG, "libtwilight_b200.so")

TWL_OK = 0
ERRORS = {-1: "TWL_E_NO_DEVICE", -2: "TWL_E_CUDA", -3: "TWL_E_ARG", -4: "TWL_E_NOMEM", -5: "TWL_E_STATE"}


class ProfilePair(C.Structure):
    """twl_profile_pair"""
    _fields_ = [("freq_ref", C.c_void_p), ("freq_qry", C.c_void_p), ("gap_open_ref", C.c_void_p), ("gap_ext_ref", C.c_void_p),
                ("gap_open_qry", C.c_void_p), ("gap_ext_qry", C.c_void_p), ("ref_len", C.c_int32), ("qry_len", C.c_int32),
                ("ref_num", C.c_float), ("qry_num", C.c_float), ("gap_char_score", C.c_float), ("xdrop", C.c_int32),
                ("flen", C.c_int32)]


class PairResult(C.Structure):
    """twl_pair_result"""
    _fields_ = [("status", C.c_int32), ("path_len", C.c_int32), ("tiles", C.c_int32), ("reserved", C.c_int32),
                ("cells", C.c_uint64), ("diagonals", C.c_uint64)]


class NodeSide(C.Structure):
    """twl_node_side"""
    _fields_ = [("seq_ids", C.POINTER(C.c_int32)), ("n_ids", C.c_int32), ("aln_len", C.c_int32), ("aln_num", C.c_int32),
                ("aln_weight", C.c_float), ("msa_freq", C.c_void_p)]


class LevelPair(C.Structure):
    """twl_level_pair"""
    _fields_ = [("ref", NodeSide), ("qry", NodeSide), ("flags", C.c_int32), ("reserved", C.c_int32)]


class LevelResult(C.Structure):
    """twl_level_result"""
    _fields_ = [("status", C.c_int32), ("path_len", C.c_int32), ("tiles", C.c_int32), ("cached", C.c_int32),
                ("cells", C.c_uint64), ("diagonals", C.c_uint64), ("ref_len_dp", C.c_int32), ("qry_len_dp", C.c_int32)]


# name -> (restype, argtypes); mirrors include/twilight_b200.h one to one (tests check the export list against it)
SIGNATURES = {
    "twl_device_count": (C.c_int, []),
    "twl_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "twl_destroy": (None, [C.c_void_p]),
    "twl_last_error": (C.c_char_p, [C.c_void_p]),
    "twl_set_params": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]),
    "twl_set_marker": (C.c_int, [C.c_void_p, C.c_int]),
    "twl_align_profiles": (C.c_int, [C.c_void_p, C.POINTER(ProfilePair), C.c_int, C.POINTER(C.c_void_p), C.POINTER(PairResult)]),
    "twl_batch_stage": (C.c_int, [C.c_void_p, C.POINTER(ProfilePair), C.c_int]),
    "twl_batch_run": (C.c_int, [C.c_void_p]),
    "twl_batch_fetch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(PairResult)]),
    "twl_rows_upload": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "twl_rows_download": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "twl_rows_export": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "twl_rows_import": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int64)]),
    "twl_rows_migrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    "twl_rows_drop": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    "twl_rows_length": (C.c_int, [C.c_void_p, C.c_int32]),
    "twl_rows_lengths": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "twl_rows_clear": (C.c_int, [C.c_void_p]),
    "twl_align_level": (C.c_int, [C.c_void_p, C.POINTER(LevelPair), C.c_int, C.c_int, C.c_float, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(LevelResult)]),
    "twl_level_fetch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "twl_level_phase_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "twl_level_update_split_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "twl_level_large_restores": (C.c_int, [C.c_void_p]),
    "twl_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "twl_last_launch_count": (C.c_int, [C.c_void_p]),
    "twl_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "twl_selftest_division": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "twl_version": (C.c_char_p, []),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m twilight_b200.build` "
                               "(there is no fallback implementation)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("TWL_LIB") and not hasattr(lib, name):
                continue          # an older build given for A/B timing may lack newer entry points
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
