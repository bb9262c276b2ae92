"""Protein alphabet (P = 22, 21x21 matrix): the CPU port against the reference build, and the device against the port — the
similarity-matrix path (talco_sim.cu + register-resident wavefront kernel, the default), the generic kernel alone
(protein_sim=0, also what a batch whose matrices exceed the budget falls back to) and the level pipeline."""
import numpy as np
import pytest

from tests import oracle_lib as ol, ref_msa
from tests.helpers import records_to_pairs, synthetic_records
from twilight_b200 import synth

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def protein_cfg(marker=1024):
    return ol.TalcoCfg(score=ol.protein_matrix(), marker=marker)


@needs_ref
@pytest.mark.parametrize("n,L,seed,marker", [(8, 300, 0, 1024), (10, 700, 1, 128)])
def test_port_protein_equals_reference(n, L, seed, marker):
    cfg, _, seqs, root, recs = synthetic_records(n, L, seed, marker, kind="protein", cfg=protein_cfg(marker), mean_blen=0.15)
    for r in recs:
        a, e = ol.ref_talco(cfg, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        assert e == r.error and np.array_equal(a, r.aln_wo)
    assert sorted(x.replace(b"-", b"") for x in root.rows) == sorted(seqs)


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,seed,marker", [(8, 300, 0, 1024), (10, 700, 1, 128), (6, 1300, 2, 256)])
def test_device_protein_dp_matches_port(n, L, seed, marker):
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(n, L, seed, marker, kind="protein", cfg=protein_cfg(marker), mean_blen=0.15)
    ctx = twilight_b200.Context(score=cfg.score, marker=marker)
    outs = ctx.align_profiles(records_to_pairs(recs, cfg))
    ctx.close()
    for k, (o, r) in enumerate(zip(outs, recs)):
        assert o.status == r.error == 0
        assert o.cells == r.cells and o.tiles == r.tiles
        assert np.array_equal(o.path, r.aln_wo), f"pair {k}"


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [{"protein_sim": 0}, {"sim_budget_mb": 1}, {"latency_mode": 1}, {"latency_mode": 0}])
def test_device_protein_dp_variants_match_port(opts):
    """generic kernel only; matrices over budget (falls back to the generic kernel); one CTA per SM (512 x 2 window first); 128 x 4 first."""
    import twilight_b200
    marker = 256
    cfg, _, _, _, recs = synthetic_records(6, 1300, 2, marker, kind="protein", cfg=protein_cfg(marker), mean_blen=0.15)
    ctx = twilight_b200.Context(score=cfg.score, marker=marker)
    for k, v in opts.items():
        ctx.set_option(k, v)
    outs = ctx.align_profiles(records_to_pairs(recs, cfg))
    launches = ctx.launch_count()
    ctx.close()
    for k, (o, r) in enumerate(zip(outs, recs)):
        assert o.status == r.error == 0
        assert o.cells == r.cells and o.tiles == r.tiles
        assert np.array_equal(o.path, r.aln_wo), f"pair {k}"
    if "latency_mode" in opts:
        assert launches >= 2          # similarity kernel + at least one wavefront stage


@pytest.mark.gpu
def test_device_protein_level_chunks(monkeypatch):
    """a level cut into several chunks by the scratch budget (which counts the similarity matrices) gives the same MSA"""
    import twilight_b200
    from twilight_b200 import msa
    n, L = 16, 300
    tree = synth.random_tree(n, seed=9, mean_blen=0.12)
    seqs = synth.evolve(tree, L, seed=9, kind="protein", indel_rate=0.05)
    w = np.ones(n, np.float32)
    cfg = protein_cfg()
    monkeypatch.setenv("TWL_LEVEL_BUDGET_MB", "1")
    ctx = twilight_b200.Context(score=cfg.score)
    rows, st = msa.progressive_align(ctx, tree, seqs, w, gappy=0.7)
    ctx.close()
    root, recs = ref_msa.progressive(tree, seqs, w, type_="p", cfg=cfg, gappy=0.7)
    assert st.cells == sum(r.cells for r in recs)
    assert sorted(rows) == sorted(root.rows)


@pytest.mark.gpu
def test_device_protein_level_pipeline_matches_port():
    import twilight_b200
    from twilight_b200 import msa
    n, L = 24, 400
    tree = synth.random_tree(n, seed=5, mean_blen=0.12)
    seqs = synth.evolve(tree, L, seed=5, kind="protein", indel_rate=0.05)
    w = np.random.default_rng(6).uniform(0.5, 1.5, n).astype(np.float32)
    cfg = protein_cfg()
    ctx = twilight_b200.Context(score=cfg.score)
    rows, st = msa.progressive_align(ctx, tree, seqs, w, gappy=0.7)
    ctx.close()
    root, recs = ref_msa.progressive(tree, seqs, w, type_="p", cfg=cfg, gappy=0.7)
    assert st.cells == sum(r.cells for r in recs)
    assert sorted(rows) == sorted(root.rows)
    assert [r.replace(b"-", b"") for r in rows] == list(seqs)
