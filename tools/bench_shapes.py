"""Device throughput of one level on the other BASELINE.json shapes: SARS-CoV-2-like 30 kb genomes and 400-aa proteins."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import twilight_b200
from twilight_b200 import LevelPairIn, NodeSideIn, synth
from tests import oracle_lib as ol   # only for the BLOSUM62 table (an input, not a checker here)

def batch(n_pairs, length, kind, members, divergence, indel, seed):
    rng = np.random.default_rng(seed)
    alphabet = {"rna": synth.RNA, "dna": synth.NT, "protein": synth.AA}[kind]
    probs = synth.AA_FREQ / synth.AA_FREQ.sum() if kind == "protein" else None
    ids, rows, pairs = [], [], []
    for _ in range(n_pairs):
        root = rng.choice(alphabet, size=int(length * rng.uniform(0.98, 1.02)), p=probs)
        sides = []
        for _s in range(2):
            anc = synth._mutate(root, divergence / 2, rng, alphabet, indel, probs)
            fr = synth.family_rows(anc, int(rng.choice(members)), rng, kind, sub_rate=divergence / 3, del_rate=0.01)
            mine = list(range(len(ids), len(ids) + len(fr)))
            ids += mine; rows += fr
            sides.append(NodeSideIn(mine, len(fr[0]), len(fr), float(len(fr))))
        pairs.append(LevelPairIn(sides[0], sides[1]))
    return ids, rows, [1.0] * len(ids), pairs

for name, kind, n, L, members, div, indel, score in (
        ("sars-like 30 kb (dna)", "dna", 296, 29700, (1, 2, 4), 0.002, 0.001, None),
        ("protein 400 aa", "protein", 4096, 400, (1, 2, 4, 8), 0.5, 0.02, ol.protein_matrix())):
    if os.environ.get("SHAPES") and os.environ["SHAPES"] not in name:
        continue
    ctx = twilight_b200.Context(score=score)
    ids, rows, w, pairs = batch(n, L, kind, members, div, indel, 3)
    for _ in range(2):
        ctx.rows_upload(ids, rows, w)
        t0 = time.perf_counter(); outs = ctx.align_level(pairs); wall = time.perf_counter() - t0
        ph = ctx.level_phase_ms()
    cells = sum(o.cells for o in outs)
    print("%-24s %5d pairs: dp %.2f ms  %.2f GCUPS | all phases %.2f ms | tiles/pair %.1f band %.0f failed %d" % (
        name, n, ph[2], cells / ph[2] / 1e6, sum(ph), np.mean([o.tiles for o in outs]), cells / max(1, sum(len(o.path) for o in outs)), sum(o.status != 0 for o in outs)),
        "| TWL_OPTIONS=" + os.environ.get("TWL_OPTIONS", ""))
    ctx.close()
