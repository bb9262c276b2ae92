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


def ref_pipeline(type_, cfg, ref_state, qry_state, gappy=0.95, current_task=0):
    """Runs the UNMODIFIED reference helpers (alignment_helper::calculateProfile, getConsensus, removeGappyColumns,
    calculatePSGP, Talco_xdrop::Align_freq, addGappyColumnsBack, updateFrequency, updateAlignment) on one pair through
    oracle/ref_shim.cpp::ref_pair_pipeline and returns every intermediate as a dict. `*_state` are tests.ref_msa.NodeState."""
    lib = ref()
    P = cfg.P
    nR, nQ = len(ref_state.rows), len(qry_state.rows)
    RL, QL = ref_state.aln_len, qry_state.aln_len
    mem, cap = max(RL, QL), RL + QL
    rows_r = (C.c_char_p * max(nR, 1))(*ref_state.rows)
    rows_q = (C.c_char_p * max(nQ, 1))(*qry_state.rows)
    w_r = np.ascontiguousarray(ref_state.weights, np.float32)
    w_q = np.ascontiguousarray(qry_state.weights, np.float32)
    f_r = None if ref_state.msa_freq is None else np.ascontiguousarray(ref_state.msa_freq, np.float32)
    f_q = None if qry_state.msa_freq is None else np.ascontiguousarray(qry_state.msa_freq, np.float32)
    out = dict(profile_raw=np.zeros((2, mem, P), np.float32), consensus=C.create_string_buffer(2 * mem + 1),
               profile=np.zeros((2, mem, P), np.float32), lens=np.zeros(2, np.int32), runs=np.zeros((2, mem, 2), np.int32),
               n_runs=np.zeros(2, np.int32), gap_op=np.zeros((2, mem), np.float32), gap_ex=np.zeros((2, mem), np.float32),
               aln_wo=np.zeros(cap + 1, np.int8), aln_w=np.zeros(cap + 1, np.int8), new_rows=C.create_string_buffer((nR + nQ) * cap + 1),
               cached=np.zeros((2, mem, P), np.float32), cached_flag=np.zeros(2, np.int32), merged=np.zeros((cap + 1, P), np.float32))
    n_wo, n_w, err, new_len, merged_flag = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    fn = lib.ref_pair_pipeline
    fn.restype = C.c_int
    vp = C.c_void_p
    fn.argtypes = [C.c_char, C.c_int, C.c_float, f32p, C.c_float, C.c_float, C.c_float, C.c_int,
                   C.c_int, C.POINTER(C.c_char_p), f32p, C.c_int, C.c_int, C.c_float, vp,
                   C.c_int, C.POINTER(C.c_char_p), f32p, C.c_int, C.c_int, C.c_float, vp,
                   f32p, C.c_char_p, f32p, i32p, i32p, i32p, f32p, f32p, i8p, C.POINTER(C.c_int), i8p, C.POINTER(C.c_int),
                   C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_int), f32p, i32p, f32p, C.POINTER(C.c_int)]
    fn(type_.encode(), current_task, gappy, cfg.score, cfg.gap_open, cfg.gap_extend, cfg.gap_boundary, cfg.marker,
       nR, rows_r, w_r, RL, ref_state.aln_num, ref_state.aln_weight, None if f_r is None else f_r.ctypes.data,
       nQ, rows_q, w_q, QL, qry_state.aln_num, qry_state.aln_weight, None if f_q is None else f_q.ctypes.data,
       out["profile_raw"], out["consensus"], out["profile"], out["lens"], out["runs"], out["n_runs"], out["gap_op"], out["gap_ex"],
       out["aln_wo"], C.byref(n_wo), out["aln_w"], C.byref(n_w), C.byref(err), out["new_rows"], C.byref(new_len), out["cached"],
       out["cached_flag"], out["merged"], C.byref(merged_flag))
    res = dict(error=err.value, aln_wo=out["aln_wo"][:n_wo.value].copy(), aln_w=out["aln_w"][:n_w.value].copy(), lens=out["lens"].copy(),
               profile_raw=[out["profile_raw"][0, :RL].copy(), out["profile_raw"][1, :QL].copy()],
               consensus=[out["consensus"].raw[:RL], out["consensus"].raw[mem:mem + QL]],
               profile=[out["profile"][0, :out["lens"][0]].copy(), out["profile"][1, :out["lens"][1]].copy()],
               gap_op=[out["gap_op"][0, :out["lens"][0]].copy(), out["gap_op"][1, :out["lens"][1]].copy()],
               gap_ex=[out["gap_ex"][0, :out["lens"][0]].copy(), out["gap_ex"][1, :out["lens"][1]].copy()],
               runs=[out["runs"][0, :out["n_runs"][0]].copy(), out["runs"][1, :out["n_runs"][1]].copy()],
               new_len=new_len.value, merged=out["merged"][:new_len.value].copy() if merged_flag.value else None,
               cached=[out["cached"][0, :RL].copy() if out["cached_flag"][0] else None, out["cached"][1, :QL].copy() if out["cached_flag"][1] else None])
    raw = out["new_rows"].raw
    res["new_rows"] = [raw[k * cap:k * cap + new_len.value] for k in range(nR + nQ)] if n_w.value else []
    return res


def ref_level(type_, cfg, pairs, threads, gappy=0.95, current_task=0):
    """One guide-tree level through the reference's OWN level entry point (cpu::alignmentKernel_CPU ->
    parallelAlignmentCPU, src/alignment-cpu.cpp:32-183; oracle/ref_shim.cpp::ref_level_cpu). `pairs` is a list of
    (ref NodeState, qry NodeState). Returns (new alnLen per pair, seconds inside the level call, pairs deferred)."""
    lib = ref()
    fn = lib.ref_level_cpu
    fn.restype = C.c_int
    fn.argtypes = [C.c_char, C.c_int, C.c_float, f32p, C.c_float, C.c_float, C.c_float, C.c_int, i32p, C.POINTER(C.c_char_p), f32p,
                   i32p, i32p, f32p, C.c_int, i32p, C.POINTER(C.c_double)]
    n = len(pairs)
    sides = [st for pr in pairs for st in pr]
    rows = [r for st in sides for r in st.rows]
    arr = (C.c_char_p * max(len(rows), 1))(*rows)
    w = np.ascontiguousarray(np.concatenate([np.asarray(st.weights, np.float32) for st in sides]) if sides else np.zeros(1, np.float32))
    n_seq = np.array([len(st.rows) for st in sides], np.int32)
    aln_len = np.array([st.aln_len for st in sides], np.int32)
    aln_num = np.array([st.aln_num for st in sides], np.int32)
    aln_w = np.array([st.aln_weight for st in sides], np.float32)
    new_len = np.zeros(max(n, 1), np.int32)
    sec = C.c_double(0)
    deferred = fn(type_.encode(), current_task, gappy, cfg.score, cfg.gap_open, cfg.gap_extend, cfg.gap_boundary, n, n_seq, arr, w,
                  aln_len, aln_num, aln_w, int(threads), new_len, C.byref(sec))
    return new_len[:n].copy(), sec.value, deferred


_B62_NCBI_ORDER = "ARNDCQEGHILKMFPSTWYV"
_B62_NCBI = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4
"""


def protein_matrix(scale=5.0, wildcard=False):
    """21x21 protein matrix in TWILIGHT's layout (ACDEFGHIKLMNPQRSTVWY + X): scale x BLOSUM62, X row/column 0
    (scoring-matrix.cpp:113-137). The BLOSUM values here are the public NCBI table; the matrix is an INPUT to both the
    oracle and the device, so parity does not depend on them."""
    raw = np.array([[float(v) for v in line.split()] for line in _B62_NCBI.strip().splitlines()], np.float32)
    order = "ACDEFGHIKLMNPQRSTVWY"
    idx = [_B62_NCBI_ORDER.index(c) for c in order]
    m = np.zeros((21, 21), np.float32)
    m[:20, :20] = scale * raw[np.ix_(idx, idx)]
    if wildcard:
        n = scale * float(np.mean(np.diag(raw)))
        m[20, :] = n
        m[:, 20] = n
    return m
