"""twl_align_level with currentTask 1 (deferred re-alignment) and 2 (merge of sub-alignments): gapCharScore = 0
(alignment-cpu.cpp:88) and the retry ladder of alignment-cpu.cpp:116-129 run inside the call; task 0 reports the
errorType instead (the caller defers the pair). Also the empty-node case of alignment-cpu.cpp:89-90. All against the CPU
oracle, bit-exact."""
import numpy as np
import pytest

from tests import oracle_lib as ol, ref_msa
from twilight_b200 import synth

pytestmark = pytest.mark.gpu
LETTERS = np.frombuffer(b"ACGU", np.uint8)


def _family(anc, members, rng, sub=0.05):
    rows = []
    for _ in range(members):
        r = anc.copy()
        hit = rng.random(r.size) < sub
        r[hit] = rng.choice(LETTERS, int(hit.sum()))
        rows.append(r.tobytes())
    return rows


def _run_pair(ctx, cfg, ra, rb, task, gappy=0.95, talco=None):
    import twilight_b200
    na, nb = len(ra), len(rb)
    w = np.ones(na + nb, np.float32)
    ctx.rows_clear()
    ctx.rows_upload(list(range(na + nb)), ra + rb, w)
    pair = twilight_b200.LevelPairIn(twilight_b200.NodeSideIn(list(range(na)), len(ra[0]), na, float(na)),
                                     twilight_b200.NodeSideIn(list(range(na, na + nb)), len(rb[0]), nb, float(nb)))
    out = ctx.align_level([pair], task=task, gappy=gappy)[0]
    sa = ref_msa.NodeState(ra, w[:na], len(ra[0]), na, float(na))
    sb = ref_msa.NodeState(rb, w[na:], len(rb[0]), nb, float(nb))
    rec = ref_msa.align_pair("n", cfg, sa, sb, gappy, task, talco, 1000)
    return out, rec


@pytest.mark.parametrize("task", [1, 2])
def test_xdrop_ladder(task):
    """Unrelated sequences, a harsh mismatch score (a user matrix: the general 5x5 score path) and a tiny x-drop
    (gapExtend = -0.2 -> xdrop 200): the band dies (errorType 1). Task 0 reports it; tasks 1/2 double the x-drop (and reset
    fLen) until the pair aligns, with gapCharScore = 0."""
    import twilight_b200
    rng = np.random.default_rng(7)
    ra = _family(rng.choice(LETTERS, 700), 3, rng)
    rb = _family(rng.choice(LETTERS, 820), 2, rng)
    score = ol.nt_matrix(match=2.0, mismatch=-30.0, transition=-30.0)
    cfg = ol.TalcoCfg(score=score, gap_open=-50.0, gap_extend=-0.2)
    ctx = twilight_b200.Context(score=score, gap_open=-50.0, gap_extend=-0.2)
    out0, rec0 = _run_pair(ctx, cfg, ra, rb, 0)
    assert rec0.error == 1 and out0.status == 1 and len(out0.path) == 0
    assert ctx.rows_download(list(range(5))) == ra + rb          # a failed pair leaves its rows untouched
    attempts = []
    out, rec = _run_pair(ctx, cfg, ra, rb, task, talco=ref_msa.talco_retry_ladder(attempts))
    assert attempts == [1, 1, 0], attempts
    assert out.status == rec.error == 0
    assert np.array_equal(out.path, rec.aln_w)
    assert out.cells == rec.cells and out.tiles == rec.tiles
    if task == 1:
        assert ctx.rows_download(list(range(5))) == rec.merged.rows
    else:
        assert ctx.rows_download(list(range(5))) == ra + rb      # currentTask 2 composes paths only (helper.cpp:384)
    ctx.close()


def test_band_limit_ladder():
    """Gap-rich profiles (every column 80 % gaps, gappy-column removal off) make gap extension almost free (calculatePSGP:
    0.2 * gapExtend), so on unrelated 5000-column nodes the band outgrows fLen = 4096 (errorType 2); task 1 widens fLen to
    min(int(4096 * 1.2) << 1, min(lens)) and the pair aligns on the wide-band kernel."""
    import twilight_b200
    rng = np.random.default_rng(8)

    def stagger(L, members):
        base = rng.choice(LETTERS, L)
        rows = []
        for m in range(members):
            r = np.full(L, ord("-"), np.uint8)
            idx = np.arange(m, L, members)
            r[idx] = base[idx]
            rows.append(r.tobytes())
        return rows

    ra, rb = stagger(5200, 5), stagger(5000, 5)
    cfg = ol.TalcoCfg()
    ctx = twilight_b200.Context()
    out0, rec0 = _run_pair(ctx, cfg, ra, rb, 0, gappy=1.0)
    assert rec0.error == 2 and out0.status == 2
    attempts = []
    out, rec = _run_pair(ctx, cfg, ra, rb, 1, gappy=1.0, talco=ref_msa.talco_retry_ladder(attempts))
    assert attempts == [2, 0], attempts
    assert out.status == rec.error == 0
    assert np.array_equal(out.path, rec.aln_w)
    assert out.cells == rec.cells
    assert ctx.rows_download(list(range(10))) == rec.merged.rows
    ctx.close()


def test_ladder_only_retries_failed_pairs():
    """A level with one failing and several ordinary pairs in task 1: only the failing pair goes through the ladder, the
    others keep their first result."""
    import twilight_b200
    rng = np.random.default_rng(9)
    score = ol.nt_matrix(match=6.0, mismatch=-12.0, transition=-12.0)
    cfg = ol.TalcoCfg(score=score, gap_open=-50.0, gap_extend=-0.2)
    ctx = twilight_b200.Context(score=score, gap_open=-50.0, gap_extend=-0.2)
    fams, ids, rows = [], [], []
    for k in range(5):
        anc = rng.choice(LETTERS, 600)
        other = rng.choice(LETTERS, 650) if k == 2 else synth._mutate(anc, 0.05, rng, LETTERS, 0.03)
        fa, fb = _family(anc, 2, rng), _family(other, 3, rng)
        fams.append((fa, fb))
    pairs, states = [], []
    for fa, fb in fams:
        sides = []
        for fr in (fa, fb):
            mine = list(range(len(ids), len(ids) + len(fr)))
            ids += mine
            rows += fr
            sides.append(twilight_b200.NodeSideIn(mine, len(fr[0]), len(fr), float(len(fr))))
            states.append(ref_msa.NodeState(fr, np.ones(len(fr), np.float32), len(fr[0]), len(fr), float(len(fr))))
        pairs.append(twilight_b200.LevelPairIn(sides[0], sides[1]))
    ctx.rows_upload(ids, rows, np.ones(len(ids), np.float32))
    outs = ctx.align_level(pairs, task=1)
    for k, o in enumerate(outs):
        attempts = []
        rec = ref_msa.align_pair("n", cfg, states[2 * k], states[2 * k + 1], 0.95, 1, ref_msa.talco_retry_ladder(attempts), 1000)
        assert (len(attempts) > 1) == (k == 2), (k, attempts)
        assert o.status == rec.error == 0 and np.array_equal(o.path, rec.aln_w) and o.cells == rec.cells, k
    ctx.close()


def test_merge_task_with_cached_profiles():
    """currentTask 2 as the divide-and-conquer merge uses it: both nodes are known only by their msaFreq (no rows), the
    result is the path and the merged msaFreq (updateFrequency, helper.cpp:506-539), gapCharScore = 0."""
    import twilight_b200
    from twilight_b200 import api
    rng = np.random.default_rng(10)
    anc = rng.choice(LETTERS, 900)
    cfg = ol.TalcoCfg()
    states = []
    for k in range(2):
        a = synth._mutate(anc, 0.06, rng, LETTERS, 0.05)
        rows = synth.family_rows(a, 6, rng, "rna")
        st = ref_msa.NodeState(rows, rng.uniform(0.5, 1.5, 6).astype(np.float32), len(rows[0]), 6, 0.0)
        st.aln_weight = float(np.float32(st.weights.sum()))
        prof = ref_msa.build_profile("n", st, cfg.P)
        f = np.zeros_like(prof)
        ol.port().twlo_freq_from_profile(cfg.P, prof, st.aln_len, st.aln_num, st.aln_weight, f)
        states.append(ref_msa.NodeState([], np.zeros(0, np.float32), st.aln_len, st.aln_num, st.aln_weight, f))
    ctx = twilight_b200.Context()
    pair = twilight_b200.LevelPairIn(twilight_b200.NodeSideIn([], states[0].aln_len, 6, states[0].aln_weight, states[0].msa_freq),
                                     twilight_b200.NodeSideIn([], states[1].aln_len, 6, states[1].aln_weight, states[1].msa_freq))
    out = ctx.align_level([pair], task=2)[0]
    rec = ref_msa.align_pair("n", cfg, states[0], states[1], 0.95, 2, None, 1000)
    assert out.status == rec.error == 0
    assert np.array_equal(out.path, rec.aln_w) and out.cells == rec.cells
    assert out.merged_freq and np.array_equal(ctx.level_fetch(0, api.F_FREQ_MERGED), rec.merged.msa_freq)
    # the same pair with the default gap-character score (task 0) takes a different score path: make sure task is honoured
    rec0 = ref_msa.align_pair("n", cfg, states[0], states[1], 0.95, 0, None, 1000)
    out0 = ctx.align_level([pair], task=0)[0]
    assert np.array_equal(out0.path, rec0.aln_w) and out0.cells == rec0.cells
    ctx.close()


@pytest.mark.parametrize("empty_side", [0, 1])
def test_empty_node(empty_side):
    """A node of length 0 (a sequence the low-quality filter emptied, io.cpp:158): the path is the other node's columns
    against nothing (alignment-cpu.cpp:89-90) and the rows are rewritten accordingly — on the device, like every other pair."""
    import twilight_b200
    rng = np.random.default_rng(11)
    full = _family(rng.choice(LETTERS, 400), 3, rng)
    empty = [b""]
    ra, rb = (empty, full) if empty_side == 0 else (full, empty)
    cfg = ol.TalcoCfg()
    ctx = twilight_b200.Context()
    out, rec = _run_pair(ctx, cfg, ra, rb, 0)
    assert out.status == rec.error == 0
    assert np.array_equal(out.path, rec.aln_w) and len(out.path) == 400
    assert set(out.path.tolist()) == ({1} if empty_side == 0 else {2})
    assert ctx.rows_download(list(range(4))) == rec.merged.rows
    ctx.close()


def test_failed_level_leaves_the_row_store_untouched():
    """twl_align_level is all or nothing: a call that fails after its kernels ran (fault injection: twl_set_option
    inject_nomem) reports TWL_E_NOMEM, every row still reads as before, and the same level then aligns to the oracle's result."""
    import twilight_b200
    rng = np.random.default_rng(12)
    cfg = ol.TalcoCfg()
    ctx = twilight_b200.Context()
    fams = []
    ids, rows, pairs, states = [], [], [], []
    for k in range(6):
        anc = rng.choice(LETTERS, 500 + 40 * k)
        fa, fb = _family(anc, 3, rng), _family(synth._mutate(anc, 0.08, rng, LETTERS, 0.05), 2, rng)
        sides = []
        for fr in (fa, fb):
            mine = list(range(len(ids), len(ids) + len(fr)))
            ids += mine
            rows += fr
            sides.append(twilight_b200.NodeSideIn(mine, len(fr[0]), len(fr), float(len(fr))))
            states.append(ref_msa.NodeState(fr, np.ones(len(fr), np.float32), len(fr[0]), len(fr), float(len(fr))))
        pairs.append(twilight_b200.LevelPairIn(sides[0], sides[1]))
    ctx.rows_upload(ids, rows, np.ones(len(ids), np.float32))
    ctx.set_option("inject_nomem", 1)
    with pytest.raises(twilight_b200.TwilightError, match="NOMEM|out-of-memory"):
        ctx.align_level(pairs)
    assert ctx.rows_download(ids) == rows
    outs = ctx.align_level(pairs)
    want_rows = []
    for k, o in enumerate(outs):
        rec = ref_msa.align_pair("n", cfg, states[2 * k], states[2 * k + 1])
        assert o.status == rec.error == 0 and np.array_equal(o.path, rec.aln_w)
        want_rows += rec.merged.rows
    assert ctx.rows_download(ids) == want_rows
    ctx.close()
