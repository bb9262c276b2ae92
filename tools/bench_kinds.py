"""DP throughput by pair kind: profile x profile, leaf x profile, profile x leaf, leaf x leaf (one level of 4096 pairs each)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import twilight_b200
from twilight_b200 import LevelPairIn, NodeSideIn, synth

def batch(n_pairs, ref_members, qry_members, seed):
    rng = np.random.default_rng(seed)
    ids, rows, pairs = [], [], []
    for _ in range(n_pairs):
        root = rng.choice(synth.RNA, size=int(1500 * rng.uniform(0.97, 1.03)))
        a = synth._mutate(root, 0.075, rng, synth.RNA, 0.03)
        b = synth._mutate(root, 0.075, rng, synth.RNA, 0.03)
        sides = []
        for anc, mem in ((a, ref_members), (b, qry_members)):
            fr = synth.family_rows(anc, int(rng.choice(mem)), rng, "rna")
            mine = list(range(len(ids), len(ids) + len(fr)))
            ids += mine; rows += fr
            sides.append(NodeSideIn(mine, len(fr[0]), len(fr), float(len(fr))))
        pairs.append(LevelPairIn(sides[0], sides[1]))
    return ids, rows, [1.0] * len(ids), pairs

ctx = twilight_b200.Context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for name, rm, qm in (("profile x profile", (2, 4, 8), (2, 4, 8)), ("leaf x profile", (1,), (2, 4, 8)), ("profile x leaf", (2, 4, 8), (1,)), ("leaf x leaf", (1,), (1,))):
    ids, rows, w, pairs = batch(n, rm, qm, 5)
    for _ in range(3):
        ctx.rows_upload(ids, rows, w)
        outs = ctx.align_level(pairs)
        ph = ctx.level_phase_ms()
    cells = sum(o.cells for o in outs)
    print("%-18s dp %.2f ms  %.1f GCUPS  (failed %d)" % (name, ph[2], cells / ph[2] / 1e6, sum(o.status != 0 for o in outs)))
