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

    # ---- one-call interface (host buffers in, host buffers out) -------------------------------------------------
    def align_profiles(self, pairs: Sequence[ProfilePairIn]) -> List[PairOut]:
        self.stage(pairs)
        self.run()
        return self.fetch()
