"""ctypes bindings for the checker side (TEST INFRASTRUCTURE): the CPU restatement oracle/libtwl_oracle.so ("port")
and, when it has been built, the unmodified reference behind oracle/_ref/libtalco_ref.so ("ref").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "libtwl_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libtalco_ref.so")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "twilight_ref")
REF_DATASET = os.path.join(ORACLE_DIR, "_ref", "dataset")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class TalcoParams(C.Structure):
    _fields_ = [("P", C.c_int32), ("score", C.POINTER(C.c_float)), ("gapOpen", C.c_float), ("gapExtend", C.c_float),
                ("gapBoundary", C.c_float), ("gapCharScore", C.c_float), ("xdrop", C.c_int32), ("fLen", C.c_int32),
                ("marker", C.c_int32)]


def nt_matrix(match=18.0, mismatch=-8.0, transition=-4.0, wildcard=False):
    """msa::Params nucleotide matrix, scoring-matrix.cpp:103-112."""
    m = np.zeros((5, 5), np.float32)
    for i in range(5):
        for j in range(5):
            if i == 4 or j == 4:
                m[i, j] = match if wildcard else 0.0
            elif i == j:
                m[i, j] = match
            elif abs(i - j) == 2:
                m[i, j] = transition
            else:
                m[i, j] = mismatch
    return m


_port = None
_ref = None


def build_port():
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "twl_oracle.cpp")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])


def port():
    global _port
    if _port is None:
        build_port()
        lib = C.CDLL(PORT_SO)
        lib.twlo_talco_align.restype = C.c_int
        lib.twlo_talco_align.argtypes = [C.POINTER(TalcoParams), C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p,
                                         C.c_float, C.c_float, i8p, C.POINTER(C.c_int), C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        lib.twlo_letter_index.restype = C.c_int
        lib.twlo_letter_index.argtypes = [C.c_char, C.c_char]
        lib.twlo_profile_from_rows.restype = None
        lib.twlo_profile_from_rows.argtypes = [C.c_char, C.c_int, C.POINTER(C.c_char_p), f32p, C.c_int, C.c_int, C.c_float, f32p]
        lib.twlo_profile_from_freq.restype = None
        lib.twlo_profile_from_freq.argtypes = [C.c_int, f32p, C.c_int, C.c_int, C.c_float, f32p]
        lib.twlo_freq_from_profile.restype = None
        lib.twlo_freq_from_profile.argtypes = [C.c_int, f32p, C.c_int, C.c_int, C.c_float, f32p]
        lib.twlo_consensus.restype = None
        lib.twlo_consensus.argtypes = [C.c_int, f32p, C.c_int, C.c_char_p]
        lib.twlo_remove_gappy.restype = C.c_int
        lib.twlo_remove_gappy.argtypes = [C.c_int, f32p, C.c_int, C.c_int, C.c_float, i32p, C.POINTER(C.c_int)]
        lib.twlo_psgp.restype = None
        lib.twlo_psgp.argtypes = [C.c_int, f32p, C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p]
        lib.twlo_pairwise_global.restype = C.c_int
        lib.twlo_pairwise_global.argtypes = [C.c_char, f32p, C.c_float, C.c_float, C.c_char_p, C.c_int, C.c_char_p, C.c_int, i8p]
        lib.twlo_add_gappy_back.restype = C.c_int
        lib.twlo_add_gappy_back.argtypes = [C.c_char, f32p, C.c_float, C.c_float, i8p, C.c_int, i32p, C.c_int, i32p, C.c_int,
                                            C.c_char_p, C.c_char_p, i8p]
        lib.twlo_merge_freq.restype = None
        lib.twlo_merge_freq.argtypes = [C.c_int, f32p, f32p, i8p, C.c_int, C.c_float, C.c_float, f32p]
        lib.twlo_update_row.restype = None
        lib.twlo_update_row.argtypes = [C.c_int, C.c_char_p, i8p, C.c_int, C.c_char_p]
        _port = lib
    return _port


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_talco_align.restype = C.c_int
        lib.ref_talco_align.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, C.c_float, C.c_float,
                                        f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, i8p,
                                        C.POINTER(C.c_int)]
        lib.ref_set_threads.restype = None
        lib.ref_set_threads.argtypes = [C.c_int]
        _ref = lib
    return _ref


class TalcoCfg:
    """Scoring + Talco_xdrop::Params bundle (TALCO-XDrop.cpp:36-53) shared by the port, ref and CUDA callers."""

    def __init__(self, score=None, gap_open=-50.0, gap_extend=-5.0, gap_boundary=None, gap_char=None, xdrop=None,
                 flen=4096, marker=1024):
        self.score = np.ascontiguousarray(nt_matrix() if score is None else score, dtype=np.float32)
        self.M = self.score.shape[0]
        self.P = self.M + 1
        self.gap_open = float(gap_open)
        self.gap_extend = float(gap_extend)
        self.gap_boundary = float(gap_extend if gap_boundary is None else gap_boundary)
        self.gap_char = float(gap_extend if gap_char is None else gap_char)
        self.xdrop = int(1000 * -gap_extend) if xdrop is None else int(xdrop)
        self.flen = int(flen)
        self.marker = int(marker)


def _prep(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def port_talco(cfg, fr, fq, gor, ger, goq, geq, ref_num, qry_num):
    """Returns (path int8[n], errorType, cells, tiles, diagonals) from the CPU restatement."""
    lib = port()
    fr, fq = _prep(fr), _prep(fq)
    R, Q = fr.shape[0], fq.shape[0]
    prm = TalcoParams(cfg.P, cfg.score.ctypes.data_as(C.POINTER(C.c_float)), cfg.gap_open, cfg.gap_extend, cfg.gap_boundary,
                      cfg.gap_char, cfg.xdrop, cfg.flen, cfg.marker)
    aln = np.zeros(R + Q + 1, np.int8)
    err, cells, tiles, diags = C.c_int(0), C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    n = lib.twlo_talco_align(C.byref(prm), R, Q, fr, fq, _prep(gor), _prep(ger), _prep(goq), _prep(geq), float(ref_num),
                             float(qry_num), aln, C.byref(err), C.byref(cells), C.byref(tiles), C.byref(diags))
    return aln[:n].copy(), err.value, cells.value, tiles.value, diags.value


def ref_talco(cfg, fr, fq, gor, ger, goq, geq, ref_num, qry_num):
    """Returns (path int8[n], errorType) from the unmodified reference (Talco_xdrop::Align_freq)."""
    lib = ref()
    fr, fq = _prep(fr), _prep(fq)
    R, Q = fr.shape[0], fq.shape[0]
    aln = np.zeros(R + Q + 1, np.int8)
    err = C.c_int(0)
    n = lib.ref_talco_align(cfg.P, R, Q, fr, fq, _prep(gor), _prep(ger), _prep(goq), _prep(geq), float(ref_num), float(qry_num),
                            cfg.score, cfg.gap_open, cfg.gap_extend, cfg.gap_boundary, cfg.gap_char, cfg.xdrop, cfg.flen,
                            cfg.marker, aln, C.byref(err))
    return aln[:n].copy(), err.value
