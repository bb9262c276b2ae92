"""Per-level device phases of a synthetic progressive MSA: pairs, longest critical path (diagonals), DP time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twilight_b200
from twilight_b200 import api, synth, msa

leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tree = synth.random_tree(leaves, seed=7, mean_blen=0.05)
seqs = synth.evolve(tree, 1500, seed=7, kind="rna")
w = np.ones(leaves, np.float32)
ctx = twilight_b200.Context()
for name, val in [a.split("=") for a in sys.argv[2:]]:
    ctx.set_option(name, int(val))
msa.progressive_align(ctx, tree, seqs, w)          # warm-up
ctx.rows_clear(); ctx.rows_upload(list(range(leaves)), seqs, w)
book = {i: msa.NodeBook([i], len(seqs[i]), 1, 1.0) for i in range(leaves)}
tot = 0.0
for lv, level in enumerate(synth.levels_bottom_up(tree)):
    pairs = [api.LevelPairIn(api.NodeSideIn(book[a].ids, book[a].aln_len, book[a].aln_num, book[a].aln_weight, book[a].msa_freq),
                             api.NodeSideIn(book[b].ids, book[b].aln_len, book[b].aln_num, book[b].aln_weight, book[b].msa_freq)) for a, b, _ in level]
    outs = ctx.align_level(pairs, task=0, gappy=0.95, cache_threshold=1000)
    ph = ctx.level_phase_ms()
    diag = max(o.ref_len_dp + o.qry_len_dp for o in outs)
    cells = sum(o.cells for o in outs)
    maxnum = max(max(book[a].aln_num, book[b].aln_num) for a, b, _ in level)
    tot += ph[2]
    print("level %2d pairs %4d maxDiag %6d tiles %2d maxNum %5d cells %.2e | prof %.2f pack %.2f dp %.2f upd %.2f ms | %.2f us/diag | %.1f GCUPS | launches %d"
          % (lv, len(level), diag, max(o.tiles for o in outs), maxnum, cells, ph[0], ph[1], ph[2], ph[3], ph[2] * 1e3 / max(diag, 1), cells / ph[2] / 1e6, ctx.launch_count()))
    for k, ((a, b, parent), o) in enumerate(zip(level, outs)):
        x, y = book.pop(a), book.pop(b)
        freq = ctx.level_fetch(k, api.F_FREQ_MERGED) if o.merged_freq else None
        book[parent] = msa.NodeBook(x.ids + y.ids, len(o.path), x.aln_num + y.aln_num, x.aln_weight + y.aln_weight, freq)
print("dp total %.1f ms" % tot)
