"""Host wall time of every step of a synthetic progressive MSA (upload, each level, download)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twilight_b200
from twilight_b200 import api, synth, msa

leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tree = synth.random_tree(leaves, seed=7, mean_blen=0.05)
seqs = synth.evolve(tree, 1500, seed=7, kind="rna")
w = np.ones(leaves, np.float32)
ctx = twilight_b200.Context()
msa.progressive_align(ctx, tree, seqs, w)
for rep in range(2):
    t0 = time.perf_counter()
    ctx.rows_clear(); t1 = time.perf_counter()
    ctx.rows_upload(list(range(leaves)), seqs, w); t2 = time.perf_counter()
    print("rep %d: clear %.2f ms, upload %.2f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    book = {i: msa.NodeBook([i], len(seqs[i]), 1, 1.0) for i in range(leaves)}
    for lv, level in enumerate(synth.levels_bottom_up(tree)):
        ta = time.perf_counter()
        pairs = [api.LevelPairIn(api.NodeSideIn(book[a].ids, book[a].aln_len, book[a].aln_num, book[a].aln_weight, book[a].msa_freq),
                                 api.NodeSideIn(book[b].ids, book[b].aln_len, book[b].aln_num, book[b].aln_weight, book[b].msa_freq)) for a, b, _ in level]
        tb = time.perf_counter()
        outs = ctx.align_level(pairs, task=0, gappy=0.95, cache_threshold=1000)
        tc = time.perf_counter()
        for k, ((a, b, parent), o) in enumerate(zip(level, outs)):
            x, y = book.pop(a), book.pop(b)
            freq = ctx.level_fetch(k, api.F_FREQ_MERGED) if o.merged_freq else None
            book[parent] = msa.NodeBook(x.ids + y.ids, len(o.path), x.aln_num + y.aln_num, x.aln_weight + y.aln_weight, freq)
        td = time.perf_counter()
        print("  level %2d pairs %4d: build %.2f ms, align_level %.2f ms (device %.2f), bookkeeping %.2f ms" % (lv, len(level), (tb - ta) * 1e3, (tc - tb) * 1e3, sum(ctx.level_phase_ms()), (td - tc) * 1e3))
    t3 = time.perf_counter()
    rows = ctx.rows_download(book[tree.root].ids)
    t4 = time.perf_counter()
    print("  download %.2f ms; total %.1f ms" % ((t4 - t3) * 1e3, (t4 - t0) * 1e3))
