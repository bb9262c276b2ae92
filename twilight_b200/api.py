"""Host-side mirror of the reference's per-level kernel interface on top of the C ABI.

`Context.align_profiles` is the batched counterpart of the `Talco_xdrop::Align_freq` call the reference makes per pair
(src/alignment-cpu.cpp:98-107): same inputs (column profiles after gappy-column removal, position-specific gap
penalties, sequence counts, gapCharScore / xdrop / fLen), same outputs (alignment path, errorType).
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class TwilightError(RuntimeError):
    pass


@dataclass
class ProfilePairIn:
    """One node pair at the Align_freq boundary. freq_* are [len][P] float32, gap_* are [len] float32."""
    freq_ref: np.ndarray
    freq_qry: np.ndarray
    gap_open_ref: np.ndarray
    gap_ext_ref: np.ndarray
    gap_open_qry: np.ndarray
    gap_ext_qry: np.ndarray
    ref_num: float
    qry_num: float
    gap_char_score: Optional[float] = None   # default: gapExtend (TALCO-XDrop.cpp:46)
    xdrop: int = 0                            # default: 1000*|gapExtend|
    flen: int = 0                             # default: 4096


@dataclass
class PairOut:
    status: int
    path: np.ndarray
    tiles: int
    cells: int
    diagonals: int


@dataclass
class NodeSideIn:
    """One node of a pair for align_level: member row ids in seqsIncluded order + Node::alnLen/alnNum/alnWeight/msaFreq."""
    seq_ids: Sequence[int]
    aln_len: int
    aln_num: int
    aln_weight: float
    msa_freq: Optional[np.ndarray] = None


@dataclass
class LevelPairIn:
    ref: NodeSideIn
    qry: NodeSideIn
    profile_only: bool = False


@dataclass
class LevelOut:
    status: int
    path: np.ndarray            # final path (with gappy columns); its length is the merged node's alnLen
    tiles: int
    cells: int
    cached_ref: bool
    cached_qry: bool
    merged_freq: bool
    ref_len_dp: int
    qry_len_dp: int


F_PROFILE_RAW = (0, 1)
F_CONSENSUS = (2, 3)
F_RUNS = (4, 5)
F_PATH_WO = 6
F_FREQ = (7, 8)
F_FREQ_MERGED = 9
F_DP_PROFILE = (10, 11)


def nucleotide_matrix(match=18.0, mismatch=-8.0, transition=-4.0, wildcard=False) -> np.ndarray:
    """msa::Params nucleotide scoring matrix (src/scoring-matrix.cpp:103-112): A C G T/U N."""
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


_B62_ORDER = "ARNDCQEGHILKMFPSTWYV"
_B62 = """
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


def protein_matrix(scale: float = 5.0) -> np.ndarray:
    """The 21x21 matrix msa::Params builds for --type p (scoring-matrix.cpp:113-137): scale x BLOSUM62 in TWILIGHT's letter
    order ACDEFGHIKLMNPQRSTVWY + X (X row/column 0). The table is the public NCBI one; it is an INPUT to twl_set_params."""
    raw = np.array([[float(v) for v in line.split()] for line in _B62.strip().splitlines()], np.float32)
    idx = [_B62_ORDER.index(c) for c in "ACDEFGHIKLMNPQRSTVWY"]
    m = np.zeros((21, 21), np.float32)
    m[:20, :20] = scale * raw[np.ix_(idx, idx)]
    return m


class Context:
    """One CUDA device + one set of scoring parameters (twl_ctx)."""

    def __init__(self, device: int = 0, score: Optional[np.ndarray] = None, gap_open: float = -50.0,
                 gap_extend: float = -5.0, gap_boundary: Optional[float] = None, marker: int = 1024):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.twl_init(device, C.byref(self._h))
        if rc != 0:
            raise TwilightError(f"twl_init({device}) -> {_lib.ERRORS.get(rc, rc)}: {self._lib.twl_last_error(None).decode()}")
        self.score = np.ascontiguousarray(nucleotide_matrix() if score is None else score, np.float32)
        self.M = int(self.score.shape[0])
        self.P = self.M + 1
        self.gap_open, self.gap_extend = float(gap_open), float(gap_extend)
        self.gap_boundary = float(gap_extend if gap_boundary is None else gap_boundary)
        self._check(self._lib.twl_set_params(self._h, self.score.ctypes.data, self.M, self.gap_open, self.gap_extend, self.gap_boundary))
        self._check(self._lib.twl_set_marker(self._h, int(marker)))
        self._keep = None
        self._n = 0

    def _check(self, rc):
        if rc != 0:
            raise TwilightError(f"{_lib.ERRORS.get(rc, rc)}: {self._lib.twl_last_error(self._h).decode()}")

    def large_restores(self) -> int:
        return int(self._lib.twl_level_large_restores(self._h))

    def set_option(self, name: str, value: int):
        self._check(self._lib.twl_set_option(self._h, name.encode(), int(value)))

    def selftest_division(self, num: np.ndarray, den: np.ndarray) -> int:
        num = np.ascontiguousarray(num, np.float32)
        den = np.ascontiguousarray(den, np.float32)
        bad = C.c_int(0)
        self._check(self._lib.twl_selftest_division(self._h, num.ctypes.data, den.ctypes.data, int(num.size), C.byref(bad)))
        return bad.value

    def close(self):
        if self._h:
            self._lib.twl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- staged interface -------------------------------------------------------------------------------------
    def stage(self, pairs: Sequence[ProfilePairIn]):
        n = len(pairs)
        arr = (_lib.ProfilePair * max(n, 1))()
        keep = []
        for k, p in enumerate(pairs):
            bufs = [np.ascontiguousarray(x, np.float32) for x in
                    (p.freq_ref, p.freq_qry, p.gap_open_ref, p.gap_ext_ref, p.gap_open_qry, p.gap_ext_qry)]
            keep.append(bufs)
            if bufs[0].ndim != 2 or bufs[0].shape[1] != self.P or bufs[1].ndim != 2 or bufs[1].shape[1] != self.P:
                raise TwilightError(f"pair {k}: profiles must be [len][{self.P}]")
            a = arr[k]
            a.freq_ref, a.freq_qry, a.gap_open_ref, a.gap_ext_ref, a.gap_open_qry, a.gap_ext_qry = [b.ctypes.data for b in bufs]
            a.ref_len, a.qry_len = bufs[0].shape[0], bufs[1].shape[0]
            a.ref_num, a.qry_num = float(p.ref_num), float(p.qry_num)
            a.gap_char_score = self.gap_extend if p.gap_char_score is None else float(p.gap_char_score)
            a.xdrop, a.flen = int(p.xdrop), int(p.flen)
        self._check(self._lib.twl_batch_stage(self._h, arr, n))
        self._keep = (arr, keep)
        self._n = n
        self._lens = [(arr[k].ref_len, arr[k].qry_len) for k in range(n)]

    def run(self):
        self._check(self._lib.twl_batch_run(self._h))

    def fetch(self, want_paths: bool = True) -> List[PairOut]:
        n = self._n
        res = (_lib.PairResult * max(n, 1))()
        bufs = [np.zeros(r + q, np.int8) for r, q in self._lens]
        ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs]) if want_paths else None
        self._check(self._lib.twl_batch_fetch(self._h, ptrs, res))
        return [PairOut(res[k].status, bufs[k][:res[k].path_len].copy(), res[k].tiles, int(res[k].cells), int(res[k].diagonals))
                for k in range(n)]

    def kernel_ms(self) -> float:
        return float(self._lib.twl_last_kernel_ms(self._h))

    def launch_count(self) -> int:
        return int(self._lib.twl_last_launch_count(self._h))

    # ---- device-resident rows + level pipeline -----------------------------------------------------------------------
    def rows_upload(self, ids: Sequence[int], rows: Sequence[bytes], weights: Sequence[float]):
        n = len(ids)
        a_ids = (C.c_int32 * n)(*[int(i) for i in ids])
        a_rows = (C.c_char_p * n)(*rows)
        a_lens = (C.c_int32 * n)(*[len(r) for r in rows])
        a_w = (C.c_float * n)(*[float(w) for w in weights])
        self._check(self._lib.twl_rows_upload(self._h, n, a_ids, a_rows, a_lens, a_w))

    def rows_download(self, ids: Sequence[int]) -> List[bytes]:
        n = len(ids)
        a_ids = np.asarray(ids, np.int32)
        lens32 = np.zeros(max(n, 1), np.int32)
        self._check(self._lib.twl_rows_lengths(self._h, n, a_ids.ctypes.data_as(C.POINTER(C.c_int32)), lens32.ctypes.data_as(C.POINTER(C.c_int32))))
        lens = lens32[:n].astype(np.int64)
        if n and lens.min() < 0:
            raise TwilightError(f"twl_rows_download: unknown row id {int(a_ids[int(np.argmin(lens))])}")
        offs = np.concatenate([[0], np.cumsum((lens + 15) & ~15)]).astype(np.int64)
        need = int(offs[-1]) + 16
        if getattr(self, "_dl_buf", None) is None or self._dl_buf.size < need:     # one reused buffer, one pointer table
            self._dl_buf = np.empty(need + need // 4, np.uint8)
        buf = self._dl_buf
        ptrs = (buf.ctypes.data + offs[:-1]).astype(np.uint64)
        out_l = np.zeros(max(n, 1), np.int32)
        self._check(self._lib.twl_rows_download(self._h, n, a_ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                ptrs.ctypes.data_as(C.POINTER(C.c_void_p)), out_l.ctypes.data_as(C.POINTER(C.c_int32))))
        mv = memoryview(buf)
        return [bytes(mv[o:o + l]) for o, l in zip(offs[:-1].tolist(), out_l[:n].tolist())]

    def rows_export(self, ids: Sequence[int], dev_ptr: int, cap_bytes: int):
        """Packs rows into a contiguous DEVICE buffer (twl_rows_export). Returns (lens, offsets) as numpy arrays."""
        n = len(ids)
        a_ids = np.asarray(ids, np.int32)
        lens = np.zeros(max(n, 1), np.int32)
        offs = np.zeros(max(n, 1), np.int64)
        self._check(self._lib.twl_rows_export(self._h, n, a_ids.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(dev_ptr), int(cap_bytes),
                                              lens.ctypes.data_as(C.POINTER(C.c_int32)), offs.ctypes.data_as(C.POINTER(C.c_int64))))
        return lens[:n], offs[:n]

    def rows_import(self, ids: Sequence[int], lens, weights, dev_ptr: int, offsets):
        """Creates / overwrites rows from a contiguous DEVICE buffer (twl_rows_import)."""
        n = len(ids)
        a_ids = np.asarray(ids, np.int32)
        a_l = np.ascontiguousarray(lens, np.int32)
        a_w = np.ascontiguousarray(weights, np.float32)
        a_o = np.ascontiguousarray(offsets, np.int64)
        self._check(self._lib.twl_rows_import(self._h, n, a_ids.ctypes.data_as(C.POINTER(C.c_int32)), a_l.ctypes.data_as(C.POINTER(C.c_int32)),
                                              a_w.ctypes.data_as(C.POINTER(C.c_float)), C.c_void_p(dev_ptr), a_o.ctypes.data_as(C.POINTER(C.c_int64))))

    # Prepared calls: the ctypes argument blocks are built once, so a timed region contains only the C ABI calls.
    def prepare_rows(self, ids: Sequence[int], rows: Sequence[bytes], weights: Sequence[float], download_caps: Optional[Sequence[int]] = None):
        n = len(ids)
        prep = type("PreparedRows", (), {})()
        prep.n = n
        prep.ids = (C.c_int32 * n)(*[int(i) for i in ids])
        prep.rows = (C.c_char_p * n)(*rows)
        prep.lens = (C.c_int32 * n)(*[len(r) for r in rows])
        prep.weights = (C.c_float * n)(*[float(w) for w in weights])
        caps = [2 * len(r) + 64 for r in rows] if download_caps is None else [int(c) for c in download_caps]
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        prep.buf = np.zeros(int(offs[-1]) + 16, np.uint8)
        prep.offs = offs
        prep.ptrs = (C.c_void_p * n)(*[prep.buf.ctypes.data + int(o) for o in offs[:-1]])
        prep.out_lens = (C.c_int32 * n)()
        return prep

    def upload_prepared(self, prep):
        self._check(self._lib.twl_rows_upload(self._h, prep.n, prep.ids, prep.rows, prep.lens, prep.weights))

    def download_prepared(self, prep):
        """Rows land in prep.buf at prep.offs[k], lengths in prep.out_lens."""
        self._check(self._lib.twl_rows_download(self._h, prep.n, prep.ids, prep.ptrs, prep.out_lens))

    def prepare_level(self, pairs: Sequence[LevelPairIn]):
        n = len(pairs)
        prep = type("PreparedLevel", (), {})()
        prep.n = n
        prep.arr = (_lib.LevelPair * max(n, 1))()
        prep.keep = []
        caps = []
        for k, p in enumerate(pairs):
            for side, dst in ((p.ref, prep.arr[k].ref), (p.qry, prep.arr[k].qry)):
                ids = (C.c_int32 * max(len(side.seq_ids), 1))(*[int(i) for i in side.seq_ids])
                prep.keep.append(ids)
                dst.seq_ids = ids
                dst.n_ids = len(side.seq_ids)
                dst.aln_len, dst.aln_num, dst.aln_weight = int(side.aln_len), int(side.aln_num), float(side.aln_weight)
                if side.msa_freq is not None:
                    f = np.ascontiguousarray(side.msa_freq, np.float32)
                    prep.keep.append(f)
                    dst.msa_freq = f.ctypes.data
                else:
                    dst.msa_freq = None
            prep.arr[k].flags = 1 if p.profile_only else 0
            caps.append(p.ref.aln_len + p.qry.aln_len + 16)
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        prep.path_buf = np.zeros(int(offs[-1]) + 16, np.int8)
        prep.path_offs = offs
        prep.ptrs = (C.c_void_p * max(n, 1))(*[prep.path_buf.ctypes.data + int(o) for o in offs[:-1]])
        prep.res = (_lib.LevelResult * max(n, 1))()
        return prep

    def align_level_prepared(self, prep, task: int = 0, gappy: float = 0.95, cache_threshold: int = 1000):
        self._check(self._lib.twl_align_level(self._h, prep.arr, prep.n, int(task), float(gappy), int(cache_threshold), prep.ptrs, prep.res))

    def rows_migrate_to(self, other: "Context", ids: Sequence[int]):
        """Moves rows to another context's device (twl_rows_migrate: pack, cudaMemcpyPeer over NVLink, unpack); they leave this one."""
        arr = (C.c_int32 * max(len(ids), 1))(*[int(i) for i in ids])
        self._check(self._lib.twl_rows_migrate(self._h, other._h, len(ids), arr))

    def rows_drop(self, ids: Sequence[int]):
        arr = (C.c_int32 * max(len(ids), 1))(*[int(i) for i in ids])
        self._check(self._lib.twl_rows_drop(self._h, len(ids), arr))

    def rows_clear(self):
        self._check(self._lib.twl_rows_clear(self._h))

    _SIDE_DT = np.dtype([("seq_ids", "u8"), ("n_ids", "i4"), ("aln_len", "i4"), ("aln_num", "i4"), ("aln_weight", "f4"), ("msa_freq", "u8")])
    _PAIR_DT = np.dtype([("ref", _SIDE_DT), ("qry", _SIDE_DT), ("flags", "i4"), ("reserved", "i4")])
    _RES_DT = np.dtype([("status", "i4"), ("path_len", "i4"), ("tiles", "i4"), ("cached", "i4"), ("cells", "u8"), ("diagonals", "u8"),
                        ("ref_len_dp", "i4"), ("qry_len_dp", "i4")])

    def align_level(self, pairs: Sequence[LevelPairIn], task: int = 0, gappy: float = 0.95, cache_threshold: int = 1000) -> List[LevelOut]:
        """One guide-tree level through twl_align_level. The argument block is assembled with numpy (the structs of
        include/twilight_b200.h as structured dtypes), not one ctypes object per field."""
        n = len(pairs)
        assert self._PAIR_DT.itemsize == C.sizeof(_lib.LevelPair) and self._RES_DT.itemsize == C.sizeof(_lib.LevelResult)
        arr = np.zeros(max(n, 1), self._PAIR_DT)
        sides = [sd for p in pairs for sd in (p.ref, p.qry)]
        meta = np.array([(len(sd.seq_ids), sd.aln_len, sd.aln_num) for sd in sides], np.int64).reshape(2 * n, 3)
        wts = np.array([sd.aln_weight for sd in sides], np.float32)
        counts = meta[:, 0]
        all_ids = np.fromiter((i for sd in sides for i in sd.seq_ids), np.int32, int(counts.sum())) if n else np.zeros(0, np.int32)
        all_ids = np.concatenate([all_ids, np.zeros(1, np.int32)])                  # never an empty buffer
        starts = np.concatenate([[0], np.cumsum(counts)])[:-1] if n else np.zeros(0, np.int64)
        ptr = (all_ids.ctypes.data + 4 * starts).astype(np.uint64)
        keep = []
        freq_ptr = np.zeros(2 * n, np.uint64)
        for k, sd in enumerate(sides):
            if sd.msa_freq is not None:
                f = np.ascontiguousarray(sd.msa_freq, np.float32)
                keep.append(f)
                freq_ptr[k] = f.ctypes.data
        for name, lo in (("ref", 0), ("qry", 1)):
            v = arr[name][:n]
            v["seq_ids"] = ptr[lo::2]
            v["n_ids"] = meta[lo::2, 0]
            v["aln_len"] = meta[lo::2, 1]
            v["aln_num"] = meta[lo::2, 2]
            v["aln_weight"] = wts[lo::2]
            v["msa_freq"] = freq_ptr[lo::2]
        arr["flags"][:n] = np.fromiter((1 if p.profile_only else 0 for p in pairs), np.int32, n)
        caps = (arr["ref"]["aln_len"][:n].astype(np.int64) + arr["qry"]["aln_len"][:n] + 15) & ~15
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        path_buf = np.zeros(int(offs[-1]) + 16, np.int8)
        ptrs = np.concatenate([(path_buf.ctypes.data + offs[:-1]).astype(np.uint64), np.zeros(1, np.uint64)])
        res = np.zeros(max(n, 1), self._RES_DT)
        self._check(self._lib.twl_align_level(self._h, arr.ctypes.data_as(C.POINTER(_lib.LevelPair)), n, int(task), float(gappy), int(cache_threshold),
                                              ptrs.ctypes.data_as(C.POINTER(C.c_void_p)), res.ctypes.data_as(C.POINTER(_lib.LevelResult))))
        del keep
        r = res[:n]
        st, pl, tl, ce, ca = r["status"].tolist(), r["path_len"].tolist(), r["tiles"].tolist(), r["cells"].tolist(), r["cached"].tolist()
        rl, ql, oo = r["ref_len_dp"].tolist(), r["qry_len_dp"].tolist(), offs[:-1].tolist()
        return [LevelOut(st[k], path_buf[oo[k]:oo[k] + pl[k]].copy(), tl[k], ce[k], bool(ca[k] & 1), bool(ca[k] & 2), bool(ca[k] & 4), rl[k], ql[k])
                for k in range(n)]

    def level_fetch(self, pair: int, what: int) -> np.ndarray:
        size = C.c_size_t(0)
        self._check(self._lib.twl_level_fetch(self._h, pair, what, None, 0, C.byref(size)))
        raw = np.zeros(max(size.value, 1), np.uint8)
        self._check(self._lib.twl_level_fetch(self._h, pair, what, raw.ctypes.data, raw.size, C.byref(size)))
        raw = raw[:size.value]
        if what in F_CONSENSUS:
            return raw
        if what in F_RUNS:
            return raw.view(np.int32).reshape(-1, 2)
        if what == F_PATH_WO:
            return raw.view(np.int8)
        out = raw.view(np.float32)
        width = self.P + 2 if what in F_DP_PROFILE else self.P
        return out.reshape(-1, width)

    def level_update_split_ms(self):
        """(gappy-column restore ms, path counts + row rewrite + frequency merge ms) of the last align_level call."""
        out = (C.c_float * 2)()
        self._check(self._lib.twl_level_update_split_ms(self._h, out))
        return [float(out[0]), float(out[1])]

    def level_phase_ms(self):
        out = (C.c_float * 4)()
        self._check(self._lib.twl_level_phase_ms(self._h, out))
        return [float(x) for x in out]

    # ---- one-call interface (host buffers in, host buffers out) -------------------------------------------------
    def align_profiles(self, pairs: Sequence[ProfilePairIn]) -> List[PairOut]:
        self.stage(pairs)
        self.run()
        return self.fetch()
