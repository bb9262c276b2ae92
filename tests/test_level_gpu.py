"""GPU parity of the level pipeline (twl_rows_* + twl_align_level) against the CPU oracle, stage by stage:
calculateProfile / getConsensus / removeGappyColumns / calculatePSGP / DP / addGappyColumnsBack / updateFrequency /
updateAlignment. Every comparison is bit-exact."""
import numpy as np
import pytest

from tests import oracle_lib as ol, ref_msa
from twilight_b200 import synth

pytestmark = pytest.mark.gpu


def run_tree_on_gpu(ctx, tree, seqs, weights, cfg, gappy, cache_threshold, check_stages=True):
    """Drives align_level bottom-up and compares each pair with the oracle's record of the same pair."""
    import twilight_b200
    from twilight_b200 import api
    ctx.rows_clear()
    ctx.rows_upload(list(range(tree.n_leaves)), seqs, weights)
    gpu = {i: dict(ids=[i], aln_len=len(seqs[i]), aln_num=1, aln_weight=float(np.float32(weights[i])), freq=None) for i in range(tree.n_leaves)}
    cpu = {i: ref_msa.leaf_state(seqs[i], weights[i]) for i in range(tree.n_leaves)}
    cpu_ids = {i: [i] for i in range(tree.n_leaves)}
    n_pairs = 0
    for level in synth.levels_bottom_up(tree):
        pairs, recs = [], []
        for a, b, parent in level:
            ga, gb = gpu[a], gpu[b]
            pairs.append(twilight_b200.LevelPairIn(
                twilight_b200.NodeSideIn(ga["ids"], ga["aln_len"], ga["aln_num"], ga["aln_weight"], ga["freq"]),
                twilight_b200.NodeSideIn(gb["ids"], gb["aln_len"], gb["aln_num"], gb["aln_weight"], gb["freq"])))
            recs.append(ref_msa.align_pair("n", cfg, cpu.pop(a), cpu.pop(b), gappy, 0, None, cache_threshold))
        outs = ctx.align_level(pairs, task=0, gappy=gappy, cache_threshold=cache_threshold)
        for k, ((a, b, parent), o, r) in enumerate(zip(level, outs, recs)):
            tag = f"pair ({a},{b})"
            assert o.status == r.error == 0, tag
            if check_stages:
                for s in (0, 1):
                    assert np.array_equal(ctx.level_fetch(k, api.F_PROFILE_RAW[s]), r.profile_raw[s]), f"{tag} raw profile side {s}"
                    assert ctx.level_fetch(k, api.F_CONSENSUS[s]).tobytes() == r.consensus[s], f"{tag} consensus side {s}"
                    assert np.array_equal(ctx.level_fetch(k, api.F_RUNS[s]), np.asarray(r.runs[s], np.int32).reshape(-1, 2)), f"{tag} gappy runs side {s}"
                    dp = ctx.level_fetch(k, api.F_DP_PROFILE[s])
                    assert np.array_equal(dp[:, :cfg.P], r.profile[s]), f"{tag} compacted profile side {s}"
                    assert np.array_equal(dp[:, cfg.P], r.gap_op[s]) and np.array_equal(dp[:, cfg.P + 1], r.gap_ex[s]), f"{tag} PSGP side {s}"
                assert np.array_equal(ctx.level_fetch(k, api.F_PATH_WO), r.aln_wo), f"{tag} DP path"
            assert o.cells == r.cells and o.tiles == r.tiles, tag
            assert np.array_equal(o.path, r.aln_w), f"{tag} final path"
            ids = gpu[a]["ids"] + gpu[b]["ids"]
            freq = None
            if o.merged_freq:
                freq = ctx.level_fetch(k, api.F_FREQ_MERGED)
                assert r.merged.msa_freq is not None and np.array_equal(freq, r.merged.msa_freq), f"{tag} merged msaFreq"
            else:
                assert r.merged.msa_freq is None, tag
            gpu[parent] = dict(ids=ids, aln_len=len(o.path), aln_num=gpu[a]["aln_num"] + gpu[b]["aln_num"],
                               aln_weight=float(np.float32(gpu[a]["aln_weight"]) + np.float32(gpu[b]["aln_weight"])), freq=freq)
            cpu[parent] = r.merged
            rows = ctx.rows_download(ids)
            assert rows == r.merged.rows, f"{tag} rewritten rows"
            n_pairs += 1
    return n_pairs


@pytest.mark.parametrize("n,L,seed,marker,gappy,cache", [
    (8, 300, 0, 1024, 0.95, 1000),
    (24, 600, 1, 256, 0.6, 1000),     # low threshold: gappy-column runs on both sides, consensus re-alignment
    (16, 900, 2, 128, 0.95, 4),       # tiny cache threshold: msaFreq caching, cached-profile branch, frequency merge
    (12, 1500, 3, 1024, 1.0, 1000),   # gappy removal disabled
])
def test_level_pipeline_matches_oracle(n, L, seed, marker, gappy, cache):
    import twilight_b200
    tree = synth.random_tree(n, seed=seed, mean_blen=0.06)
    seqs = synth.evolve(tree, L, seed=seed, indel_rate=0.08)
    w = np.random.default_rng(seed + 1).uniform(0.5, 1.5, n).astype(np.float32)
    cfg = ol.TalcoCfg(marker=marker)
    ctx = twilight_b200.Context(marker=marker)
    done = run_tree_on_gpu(ctx, tree, seqs, w, cfg, gappy, cache)
    ctx.close()
    assert done == n - 1


def test_rows_roundtrip_and_errors():
    import twilight_b200
    ctx = twilight_b200.Context()
    rows = [b"ACGT-ACGTNN", b"A", b"", b"acgu" * 1000]
    ctx.rows_upload([3, 0, 7, 2], rows, [1.0, 2.0, 0.5, 1.5])
    assert ctx.rows_download([2, 7, 0, 3]) == [rows[3], rows[2], rows[1], rows[0]]
    with pytest.raises(twilight_b200.TwilightError):
        ctx.rows_download([5])
    side = twilight_b200.NodeSideIn([3], 999, 1, 1.0)   # wrong length
    with pytest.raises(twilight_b200.TwilightError):
        ctx.align_level([twilight_b200.LevelPairIn(side, twilight_b200.NodeSideIn([0], 1, 1, 2.0))])
    assert ctx.align_level([]) == []
    ctx.close()


def test_progressive_driver_end_to_end():
    """twilight_b200.msa.progressive_align (rows resident across levels) equals the oracle's MSA and passes the
    reference's --check criteria (sequencedb.cpp:87-120)."""
    import twilight_b200
    from twilight_b200 import msa
    n, L = 40, 500
    tree = synth.random_tree(n, seed=12, mean_blen=0.05)
    seqs = synth.evolve(tree, L, seed=12)
    w = np.random.default_rng(13).uniform(0.5, 1.5, n).astype(np.float32)
    ctx = twilight_b200.Context()
    rows, st = msa.progressive_align(ctx, tree, seqs, w)
    ctx.close()
    root, recs = ref_msa.progressive(tree, seqs, w, cfg=ol.TalcoCfg(), keep_records=True)
    assert st.pairs == n - 1 and st.cells == sum(r.cells for r in recs)
    assert len(set(len(r) for r in rows)) == 1 and len(rows[0]) == root.aln_len
    assert [r.replace(b"-", b"") for r in rows] == list(seqs)
    # same rows as the oracle (the oracle lists ref members first at every merge; compare as a multiset keyed by content)
    assert sorted(rows) == sorted(root.rows)


def test_level_is_chunked_transparently(monkeypatch):
    """A level larger than the scratch budget is processed in several chunks; results and rows must not depend on it."""
    import twilight_b200
    from twilight_b200 import msa
    n, L = 64, 400
    tree = synth.random_tree(n, seed=21, mean_blen=0.05, shape="balanced")
    seqs = synth.evolve(tree, L, seed=21)
    w = np.ones(n, np.float32)
    ctx = twilight_b200.Context()
    rows_a, st_a = msa.progressive_align(ctx, tree, seqs, w)
    monkeypatch.setenv("TWL_LEVEL_BUDGET_MB", "1")      # ~6 pairs per chunk at the first level
    rows_b, st_b = msa.progressive_align(ctx, tree, seqs, w)
    ctx.close()
    assert rows_a == rows_b and st_a.cells == st_b.cells and st_a.aln_len == st_b.aln_len
    assert st_b.launches > st_a.launches


@pytest.mark.parametrize("jobs", [None, "0", "1"])
@pytest.mark.parametrize("ins_len,expect_large", [(25, 0), (150, 1), (1300, 1)])
def test_coinciding_gappy_runs_are_realigned(ins_len, expect_large, jobs, monkeypatch):
    """Removed runs of both nodes that start at the same path position are aligned against each other (pairwiseGlobal,
    alignment-helper.cpp:243-322): in the restore kernel's shared memory when small, by a second pass of the same kernel
    with global scratch when the matrix exceeds it (150: by cells; 1300: also by row length). Same final path and rows as
    the oracle either way."""
    import twilight_b200
    # jobs: default = consensus alignments put off to the parallel job kernel + compaction; "0" = aligned in line by the walk;
    # "1" = a job list of one entry (the rest in line: the list-full path)
    if jobs is not None:
        monkeypatch.setenv("TWL_RESTORE_JOBS", jobs)
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
            mid = ins if m == 0 else np.full(ins_len, ord("-"), np.uint8)     # one member carries an insertion at column 250
            rows.append(np.concatenate([row[:250], mid, row[250:]]).tobytes())
        return rows

    ra, rb = family(1, 24), family(2, 24)
    w = np.ones(48, np.float32)
    cfg = ol.TalcoCfg()
    ctx = twilight_b200.Context()
    ctx.rows_upload(list(range(48)), ra + rb, w)
    L = len(ra[0])
    pair = twilight_b200.LevelPairIn(twilight_b200.NodeSideIn(list(range(24)), L, 24, 24.0), twilight_b200.NodeSideIn(list(range(24, 48)), L, 24, 24.0))
    out = ctx.align_level([pair], task=0, gappy=0.9)[0]
    sa = ref_msa.NodeState(ra, w[:24], L, 24, 24.0)
    sb = ref_msa.NodeState(rb, w[24:], L, 24, 24.0)
    rec = ref_msa.align_pair("n", cfg, sa, sb, 0.9, 0, None, 1000)
    assert len(rec.runs[0]) >= 1 and len(rec.runs[1]) >= 1          # the case is what it claims to be
    assert out.status == rec.error == 0
    assert np.array_equal(out.path, rec.aln_w)
    assert ctx.rows_download(list(range(48))) == rec.merged.rows
    assert ctx.large_restores() == expect_large
    ctx.close()


@pytest.mark.parametrize("seed", list(range(10)))
def test_level_pipeline_fuzz(seed):
    """Random small trees with random marker / gappy threshold / indel rate / cache threshold: every stage of every pair
    against the oracle. Low thresholds make many removed runs (also at column 0 and at the end) and many coinciding run
    pairs for the consensus re-alignment of the restore kernel."""
    import twilight_b200
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(5, 14))
    L = int(rng.integers(60, 500))
    marker = int(rng.choice([32, 64, 128, 1024]))
    gappy = float(rng.choice([0.3, 0.5, 0.7, 0.9, 0.95]))
    cache = int(rng.choice([2, 4, 1000]))
    tree = synth.random_tree(n, seed=seed, mean_blen=float(rng.uniform(0.03, 0.15)))
    seqs = synth.evolve(tree, L, seed=seed, indel_rate=float(rng.uniform(0.02, 0.25)))
    w = rng.uniform(0.5, 1.5, n).astype(np.float32)
    cfg = ol.TalcoCfg(marker=marker)
    ctx = twilight_b200.Context(marker=marker)
    done = run_tree_on_gpu(ctx, tree, seqs, w, cfg, gappy, cache)
    assert ctx.large_restores() == 0
    ctx.close()
    assert done == n - 1


def test_rows_export_import_device_roundtrip():
    """twl_rows_export / twl_rows_import: rows packed into one device buffer by one context and adopted by another
    (the GPU-to-GPU node migration of the sharded MSA) come back byte-identical."""
    import torch
    import twilight_b200
    rng = np.random.default_rng(3)
    rows = [bytes(rng.choice(np.frombuffer(b"ACGU-", np.uint8), int(n)).tobytes()) for n in (1, 15, 16, 17, 1000, 4097)]
    ids = [5, 0, 9, 2, 7, 3]
    a, b = twilight_b200.Context(), twilight_b200.Context()
    a.rows_upload(ids, rows, [1.0 + k for k in range(len(ids))])
    total = sum((len(r) + 15) & ~15 for r in rows)
    buf = torch.zeros(total, dtype=torch.uint8, device="cuda")
    lens, offs = a.rows_export(ids, buf.data_ptr(), buf.numel())
    assert list(lens) == [len(r) for r in rows]
    b.rows_import([10 + i for i in ids], lens, [2.0] * len(ids), buf.data_ptr(), offs)
    assert b.rows_download([10 + i for i in ids]) == rows
    with pytest.raises(twilight_b200.TwilightError):
        a.rows_export([99], buf.data_ptr(), buf.numel())
    with pytest.raises(twilight_b200.TwilightError):
        a.rows_export(ids, buf.data_ptr(), 16)
    a.close(); b.close()
