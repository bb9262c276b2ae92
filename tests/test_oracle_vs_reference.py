"""Pins the CPU restatement (oracle/twl_oracle.cpp) to the unmodified reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile) and the reference build to its golden outputs. CPU only."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from tests import oracle_lib as ol, ref_msa
from tests.cli_scenarios import DATA, ROOT, run_cli
from tests.helpers import synthetic_records

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("n,L,seed,marker", [(8, 300, 0, 1024), (10, 1500, 1, 1024), (12, 700, 2, 128), (6, 2500, 3, 256), (9, 500, 4, 32)])
def test_port_talco_equals_reference(n, L, seed, marker):
    cfg, _, seqs, root, recs = synthetic_records(n, L, seed, marker)
    for r in recs:
        a, e = ol.ref_talco(cfg, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        assert e == r.error
        assert np.array_equal(a, r.aln_wo)
    # structural validity, the reference's --check (sequencedb.cpp:87-120)
    assert sorted(x.replace(b"-", b"") for x in root.rows) == sorted(seqs)
    assert all(len(x) == root.aln_len for x in root.rows)


@needs_ref
def test_port_error_codes_equal_reference():
    cfg, _, _, _, recs = synthetic_records(4, 600, 7, 1024)
    r = recs[-1]
    for xdrop, flen in ((5, 4096), (5000, 8), (40, 4096), (5000, 40), (100, 64)):
        c = ol.TalcoCfg(xdrop=xdrop, flen=flen)
        args = (c, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        a, e = ol.ref_talco(*args)
        b, f = ol.port_talco(*args)[:2]
        assert e == f and np.array_equal(a, b), (xdrop, flen, e, f)


@needs_ref
def test_reference_cli_reproduces_golden(tmp_path):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))
    for name in ("rnasim_merge_msas", "rnasim_sub_prune"):
        out, _ = run_cli(ol.REF_CLI, name, str(tmp_path))
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == gold[name]["md5"]


@needs_ref
@pytest.mark.parametrize("n,L,seed,marker,gappy", [(10, 400, 5, 1024, 0.95), (24, 600, 1, 256, 0.6), (14, 500, 8, 64, 0.8)])
def test_port_level_functions_equal_reference_helpers(n, L, seed, marker, gappy):
    """calculateProfile / getConsensus / removeGappyColumns / calculatePSGP / addGappyColumnsBack / updateAlignment of the
    port against the unmodified alignment-helper.cpp on every merge of a synthetic tree."""
    import copy
    from twilight_b200 import synth
    tree = synth.random_tree(n, seed=seed, mean_blen=0.06)
    seqs = synth.evolve(tree, L, seed=seed, indel_rate=0.08)
    w = np.random.default_rng(seed + 1).uniform(0.5, 1.5, n).astype(np.float32)
    cfg = ol.TalcoCfg(marker=marker)
    state = {i: ref_msa.leaf_state(seqs[i], w[i]) for i in range(n)}
    checked = 0
    for level in synth.levels_bottom_up(tree):
        for a, b, parent in level:
            ra, rb = state.pop(a), state.pop(b)
            want = ol.ref_pipeline("n", cfg, copy.deepcopy(ra), copy.deepcopy(rb), gappy)
            rec = ref_msa.align_pair("n", cfg, ra, rb, gappy)
            assert rec.error == want["error"] == 0
            for s in (0, 1):
                assert np.array_equal(rec.profile_raw[s], want["profile_raw"][s])
                assert rec.consensus[s] == want["consensus"][s]
                assert np.array_equal(rec.profile[s], want["profile"][s])
                assert np.array_equal(rec.gap_op[s], want["gap_op"][s]) and np.array_equal(rec.gap_ex[s], want["gap_ex"][s])
                assert np.array_equal(np.asarray(rec.runs[s]).reshape(-1, 2), want["runs"][s])
            assert np.array_equal(rec.aln_wo, want["aln_wo"])
            assert np.array_equal(rec.aln_w, want["aln_w"])
            assert rec.merged.rows == want["new_rows"]
            state[parent] = rec.merged
            checked += 1
    assert checked == n - 1


@needs_ref
def test_port_frequency_cache_and_merge_equal_reference():
    """The >=1000-sequence branch (msaFreq cache, cached-profile path, updateFrequency) on two synthetic 1000-row nodes."""
    rng = np.random.default_rng(3)
    L = 120
    anc = rng.choice(np.frombuffer(b"ACGU", np.uint8), L)

    def family(nrows, shift):
        rows = []
        for _ in range(nrows):
            s = anc.copy()
            hit = rng.random(L) < 0.1
            s[hit] = rng.choice(np.frombuffer(b"ACGU-", np.uint8), int(hit.sum()))
            rows.append(np.roll(s, shift).tobytes())
        wts = rng.uniform(0.5, 1.5, nrows).astype(np.float32)
        return ref_msa.NodeState(rows, wts, L, nrows, float(np.sum(wts, dtype=np.float32)))
    a, b = family(1000, 0), family(3, 1)
    cfg = ol.TalcoCfg()
    import copy
    want = ol.ref_pipeline("n", cfg, copy.deepcopy(a), copy.deepcopy(b), 0.95)
    rec = ref_msa.align_pair("n", cfg, a, b, 0.95)
    assert want["cached"][0] is not None and want["cached"][1] is not None and want["merged"] is not None
    assert np.array_equal(rec.profile_raw[0], want["profile_raw"][0])
    assert np.array_equal(rec.ref.msa_freq, want["cached"][0]) and np.array_equal(rec.qry.msa_freq, want["cached"][1])
    assert np.array_equal(rec.aln_w, want["aln_w"])
    assert np.array_equal(rec.merged.msa_freq, want["merged"])
    # second merge: cached-profile branch on the ref side
    c = family(2, 0)
    m = rec.merged
    want2 = ol.ref_pipeline("n", cfg, copy.deepcopy(m), copy.deepcopy(c), 0.95)
    rec2 = ref_msa.align_pair("n", cfg, m, c, 0.95)
    assert np.array_equal(rec2.profile_raw[0], want2["profile_raw"][0])
    assert np.array_equal(rec2.aln_w, want2["aln_w"])
    assert np.array_equal(rec2.merged.msa_freq, want2["merged"])


@needs_ref
@pytest.mark.parametrize("ins_len", [3, 25, 150])
def test_port_coinciding_gappy_runs_equal_reference(ins_len):
    """Removed runs of both nodes that start at the same path position (addGappyColumnsBack + pairwiseGlobal,
    alignment-helper.cpp:324-375, 243-322): the crafted case of tests/test_level_gpu.py, port against the unmodified
    reference helpers."""
    import copy
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACGU", np.uint8)
    anc = rng.choice(letters, 500)

    def family(seed, members):
        r = np.random.default_rng(seed)
        ins = r.choice(letters, ins_len)
        rows = []
        for m in range(members):
            row = anc.copy()
            flip = r.random(row.size) < 0.03
            row[flip] = r.choice(letters, int(flip.sum()))
            mid = ins if m == 0 else np.full(ins_len, ord("-"), np.uint8)
            rows.append(np.concatenate([row[:250], mid, row[250:]]).tobytes())
        return rows

    ra, rb = family(1, 24), family(2, 24)
    L = len(ra[0])
    w = np.ones(24, np.float32)
    cfg = ol.TalcoCfg()
    sa, sb = ref_msa.NodeState(ra, w, L, 24, 24.0), ref_msa.NodeState(rb, w.copy(), L, 24, 24.0)
    want = ol.ref_pipeline("n", cfg, copy.deepcopy(sa), copy.deepcopy(sb), 0.9)
    rec = ref_msa.align_pair("n", cfg, sa, sb, 0.9)
    assert len(rec.runs[0]) >= 1 and len(rec.runs[1]) >= 1
    assert rec.error == want["error"] == 0
    assert np.array_equal(rec.aln_wo, want["aln_wo"])
    assert np.array_equal(rec.aln_w, want["aln_w"])
    assert rec.merged.rows == want["new_rows"]


@needs_ref
@pytest.mark.parametrize("seed", list(range(8)))
def test_port_pipeline_fuzz_equals_reference(seed):
    """Random small trees with random marker / gappy threshold / indel rate: every stage of every merge, port against the
    unmodified reference helpers (the same generator the GPU fuzz test uses, so the oracle the GPU is checked against is
    itself pinned on those inputs)."""
    import copy
    from twilight_b200 import synth
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(5, 14))
    L = int(rng.integers(60, 500))
    marker = int(rng.choice([32, 64, 128, 1024]))
    gappy = float(rng.choice([0.3, 0.5, 0.7, 0.9, 0.95]))
    _cache = int(rng.choice([2, 4, 1000]))
    tree = synth.random_tree(n, seed=seed, mean_blen=float(rng.uniform(0.03, 0.15)))
    seqs = synth.evolve(tree, L, seed=seed, indel_rate=float(rng.uniform(0.02, 0.25)))
    w = rng.uniform(0.5, 1.5, n).astype(np.float32)
    cfg = ol.TalcoCfg(marker=marker)
    state = {i: ref_msa.leaf_state(seqs[i], w[i]) for i in range(n)}
    for level in synth.levels_bottom_up(tree):
        for a, b, parent in level:
            ra, rb = state.pop(a), state.pop(b)
            want = ol.ref_pipeline("n", cfg, copy.deepcopy(ra), copy.deepcopy(rb), gappy)
            rec = ref_msa.align_pair("n", cfg, ra, rb, gappy)
            assert rec.error == want["error"] == 0
            for s in (0, 1):
                assert np.array_equal(rec.profile[s], want["profile"][s])
                assert np.array_equal(np.asarray(rec.runs[s]).reshape(-1, 2), want["runs"][s])
            assert np.array_equal(rec.aln_wo, want["aln_wo"])
            assert np.array_equal(rec.aln_w, want["aln_w"])
            assert rec.merged.rows == want["new_rows"]
            state[parent] = rec.merged


def test_reference_level_entry_matches_port():
    """The stock level call (cpu::alignmentKernel_CPU -> parallelAlignmentCPU on a real NodePairVec, oracle/ref_shim.cpp::
    ref_level_cpu — what `bench.py --impl reference` times) yields the same merged lengths as the port's per-pair pipeline."""
    if not ol.have_ref():
        pytest.skip("oracle/_ref/libtalco_ref.so not built")
    from tests import ref_msa
    from twilight_b200 import synth
    fam = synth.level_rows_batch(12, 500, seed=5, kind="rna")
    cfg = ol.TalcoCfg()

    def st(rows):
        return ref_msa.NodeState(rows, np.ones(len(rows), np.float32), len(rows[0]), len(rows), float(len(rows)))
    new_len, seconds, deferred = ol.ref_level("n", cfg, [(st(a), st(b)) for a, b in fam], threads=4)
    want = [len(ref_msa.align_pair("n", cfg, st(a), st(b)).aln_w) for a, b in fam]
    assert list(new_len) == want and deferred == 0 and seconds > 0
