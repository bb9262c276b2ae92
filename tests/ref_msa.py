"""Checker-side progressive aligner (TEST INFRASTRUCTURE): drives the oracle port (oracle/twl_oracle.cpp) over a
guide tree exactly in the order of src/alignment-cpu.cpp:49-170, and records what crosses each function boundary so
that the CUDA path can be compared stage by stage.  Never imported by the product package.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from tests import oracle_lib as ol

CAL_PROFILE_TH = 1000  # alignment_helper::_CAL_PROFILE_TH, msa.hpp:179


@dataclass
class NodeState:
    rows: List[bytes]
    weights: np.ndarray           # per member row
    aln_len: int
    aln_num: int
    aln_weight: float
    msa_freq: Optional[np.ndarray] = None


@dataclass
class PairRecord:
    ref: NodeState
    qry: NodeState
    profile_raw: Optional[np.ndarray] = None   # [2] list of [len][P] after calculateProfile
    consensus: Optional[list] = None
    profile: Optional[list] = None             # after removeGappyColumns (trimmed to new lens)
    runs: Optional[list] = None                # [(start,len)...] per side
    gap_op: Optional[list] = None
    gap_ex: Optional[list] = None
    aln_wo: Optional[np.ndarray] = None
    aln_w: Optional[np.ndarray] = None
    error: int = 0
    cells: int = 0
    tiles: int = 0
    merged: Optional[NodeState] = None


def _rows_ptr(rows):
    arr = (C.c_char_p * len(rows))(*rows)
    return arr


def build_profile(type_, st: NodeState, P):
    lib = ol.port()
    prof = np.zeros((st.aln_len, P), np.float32)
    if st.msa_freq is not None:
        lib.twlo_profile_from_freq(P, np.ascontiguousarray(st.msa_freq, np.float32), st.aln_len, st.aln_num,
                                   C.c_float(st.aln_weight), prof)
    else:
        lib.twlo_profile_from_rows(type_.encode(), len(st.rows), _rows_ptr(st.rows),
                                   np.ascontiguousarray(st.weights, np.float32), st.aln_len, st.aln_num,
                                   C.c_float(st.aln_weight), prof)
    return prof


def align_pair(type_, cfg: ol.TalcoCfg, ref: NodeState, qry: NodeState, gappy=0.95, current_task=0, talco=None,
               cache_threshold=CAL_PROFILE_TH) -> PairRecord:
    """One pair through the whole pipeline. `talco(cfg, fr, fq, gor, ger, goq, geq, refNum, qryNum)` defaults to the port
    and may be swapped for the reference shim or the CUDA path."""
    lib = ol.port()
    P = cfg.P
    rec = PairRecord(ref, qry)
    store = (ref.aln_num >= cache_threshold or qry.aln_num >= cache_threshold or ref.msa_freq is not None
             or qry.msa_freq is not None)
    prof, cons, runs, gop, gex, lens = [], [], [], [], [], []
    raw = []
    for st in (ref, qry):
        p = build_profile(type_, st, P)
        if store and st.msa_freq is None:  # helper.cpp:35-40
            f = np.zeros_like(p)
            lib.twlo_freq_from_profile(P, p, st.aln_len, st.aln_num, C.c_float(st.aln_weight), f)
            st.msa_freq = f
        raw.append(p.copy())
        c = C.create_string_buffer(st.aln_len + 1)
        lib.twlo_consensus(P, p, st.aln_len, c)
        cons.append(c.raw[:st.aln_len])
        rr = np.zeros(2 * st.aln_len + 2, np.int32)
        nl = C.c_int(0)
        n_runs = lib.twlo_remove_gappy(P, p, st.aln_len, st.aln_num, C.c_float(gappy), rr, C.byref(nl))
        runs.append(rr[:2 * n_runs].reshape(-1, 2).copy())
        lens.append(nl.value)
        go = np.zeros(max(nl.value, 1), np.float32)
        ge = np.zeros(max(nl.value, 1), np.float32)
        lib.twlo_psgp(P, p, nl.value, st.aln_num, C.c_float(cfg.gap_open), C.c_float(cfg.gap_extend), go, ge)
        prof.append(p[:nl.value].copy())
        gop.append(go[:nl.value])
        gex.append(ge[:nl.value])
    rec.profile_raw, rec.consensus, rec.profile, rec.runs, rec.gap_op, rec.gap_ex = raw, cons, prof, runs, gop, gex

    use = cfg
    if current_task in (1, 2) or ref.aln_num > 10000 or qry.aln_num > 10000:  # alignment-cpu.cpp:88
        use = ol.TalcoCfg(cfg.score, cfg.gap_open, cfg.gap_extend, cfg.gap_boundary, 0.0, cfg.xdrop, cfg.flen, cfg.marker)
    if lens[0] == 0:
        aln = np.full(lens[1], 1, np.int8)
    elif lens[1] == 0:
        aln = np.full(lens[0], 2, np.int8)
    else:
        fn = talco or ol.port_talco
        res = fn(use, prof[0], prof[1], gop[0], gex[0], gop[1], gex[1], ref.aln_num, qry.aln_num)
        aln, rec.error = res[0], res[1]
        if len(res) > 2:
            rec.cells, rec.tiles = res[2], res[3]
    rec.aln_wo = aln
    if len(aln) == 0:
        return rec
    out = np.zeros(ref.aln_len + qry.aln_len + 1, np.int8)
    r0 = np.ascontiguousarray(runs[0].reshape(-1), np.int32) if len(runs[0]) else np.zeros(2, np.int32)
    r1 = np.ascontiguousarray(runs[1].reshape(-1), np.int32) if len(runs[1]) else np.zeros(2, np.int32)
    n = lib.twlo_add_gappy_back(type_.encode(), cfg.score, C.c_float(cfg.gap_open), C.c_float(cfg.gap_extend),
                                np.ascontiguousarray(aln, np.int8), len(aln), r0, len(runs[0]), r1, len(runs[1]),
                                cons[0] + b"\0", cons[1] + b"\0", out)
    path = out[:n].copy()
    rec.aln_w = path
    # updateFrequency (helper.cpp:506) then updateAlignment (helper.cpp:377)
    merged_freq = None
    if ref.msa_freq is not None and qry.msa_freq is not None:
        merged_freq = np.zeros((n, P), np.float32)
        lib.twlo_merge_freq(P, np.ascontiguousarray(ref.msa_freq, np.float32), np.ascontiguousarray(qry.msa_freq, np.float32),
                            path, n, C.c_float(ref.aln_weight), C.c_float(qry.aln_weight), merged_freq)
    new_rows = []
    buf = C.create_string_buffer(n + 1)
    for side, st in ((0, ref), (1, qry)):
        for row in st.rows:
            lib.twlo_update_row(side, row, path, n, buf)
            new_rows.append(buf.raw[:n])
    rec.merged = NodeState(new_rows, np.concatenate([ref.weights, qry.weights]).astype(np.float32), n,
                           ref.aln_num + qry.aln_num, float(np.float32(ref.aln_weight) + np.float32(qry.aln_weight)), merged_freq)
    return rec


def leaf_state(seq: bytes, weight: float) -> NodeState:
    return NodeState([seq], np.array([weight], np.float32), len(seq), 1, float(np.float32(weight)))


def progressive(tree, seqs, weights, type_="n", cfg=None, gappy=0.95, talco=None, keep_records=True,
                cache_threshold=CAL_PROFILE_TH):
    """Full bottom-up MSA over `tree` (twilight_b200.synth.Tree). Returns (root NodeState, [PairRecord per merge in
    level order])."""
    from twilight_b200 import synth
    cfg = cfg or ol.TalcoCfg()
    state = {i: leaf_state(seqs[i], weights[i]) for i in range(tree.n_leaves)}
    records = []
    for level in synth.levels_bottom_up(tree):
        for a, b, parent in level:
            rec = align_pair(type_, cfg, state.pop(a), state.pop(b), gappy, 0, talco, cache_threshold)
            if rec.merged is None:
                raise RuntimeError(f"pair ({a},{b}) failed with errorType {rec.error}")
            state[parent] = rec.merged
            if keep_records:
                records.append(rec)
    return state[tree.root], records


def talco_retry_ladder(log=None):
    """Talco_xdrop::Align_freq inside the retry loop the reference runs for currentTask 1 and 2 (alignment-cpu.cpp:95-130):
    errorType 2 -> fLen = min(int(fLen * 1.2) << 1, min(lens)); errorType 1 -> xdrop *= 2, fLen = min(int(xdrop * 4) << 1,
    min(lens)); until the pair aligns (or errorType 3). `log` (a list) receives the errorType of every attempt."""
    def fn(cfg, fr, fq, gor, ger, goq, geq, ref_num, qry_num):
        xdrop, flen = cfg.xdrop, cfg.flen
        min_len = min(len(fr), len(fq))
        while True:
            use = ol.TalcoCfg(cfg.score, cfg.gap_open, cfg.gap_extend, cfg.gap_boundary, cfg.gap_char, xdrop, flen, cfg.marker)
            res = ol.port_talco(use, fr, fq, gor, ger, goq, geq, ref_num, qry_num)
            if log is not None:
                log.append(res[1])
            if res[1] in (0, 3):
                return res
            if res[1] == 2:
                flen = min(int(flen * 1.2) << 1, min_len)
            else:
                xdrop = xdrop * 2
                flen = min(int(xdrop * 4) << 1, min_len)
    return fn
