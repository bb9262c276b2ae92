"""A/B of the wavefront kernel's score/barrier overlap on a throughput batch and on a latency-bound MSA."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import twilight_b200
from twilight_b200 import msa, synth

ids, rows, weights, pairs = bench.build_level_batch(4096, 1500, seed=1000)
tree = synth.random_tree(2048, seed=77, mean_blen=0.05)
seqs = synth.evolve(tree, 1500, seed=77)
w = np.ones(2048, np.float32)
ctx = twilight_b200.Context()
for mode in (0, 1, -1):
    ctx.set_option("overlap", mode)
    for _ in range(3):
        ctx.rows_upload(ids, rows, weights)
        outs = ctx.align_level(pairs)
        ph = ctx.level_phase_ms()
    cells = sum(o.cells for o in outs)
    _, st = msa.progressive_align(ctx, tree, seqs, w)
    _, st = msa.progressive_align(ctx, tree, seqs, w)
    print("overlap=%2d  level batch: dp %.2f ms %.1f GCUPS | msa 2048: dp %.1f ms wall %.3f s %.0f seqs/s" % (mode, ph[2], cells / ph[2] / 1e6, st.phase_ms[2], st.wall_s, 2048 / st.wall_s))
