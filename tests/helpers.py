"""Shared helpers for the parity tests (TEST INFRASTRUCTURE)."""
import numpy as np

from tests import oracle_lib as ol, ref_msa
from twilight_b200 import synth


def synthetic_records(n_leaves, root_len, seed, marker=1024, kind="rna", mean_blen=0.04, shape="yule", cfg=None,
                      gappy=0.95, indel_rate=0.03):
    """Progressive alignment of a seeded synthetic set with the CPU port; returns (cfg, tree, seqs, root, records)."""
    tree = synth.random_tree(n_leaves, seed=seed, mean_blen=mean_blen, shape=shape)
    seqs = synth.evolve(tree, root_len, seed=seed, kind=kind, indel_rate=indel_rate)
    w = np.random.default_rng(seed + 1).uniform(0.5, 1.5, n_leaves).astype(np.float32)
    cfg = cfg or ol.TalcoCfg(marker=marker)
    root, recs = ref_msa.progressive(tree, seqs, w, type_="n" if kind != "protein" else "p", cfg=cfg, gappy=gappy)
    return cfg, tree, seqs, root, recs


def records_to_pairs(recs, cfg):
    from twilight_b200 import ProfilePairIn
    out = []
    for r in recs:
        out.append(ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1],
                                 r.ref.aln_num, r.qry.aln_num, gap_char_score=cfg.gap_char, xdrop=cfg.xdrop, flen=cfg.flen))
    return out
