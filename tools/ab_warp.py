"""A/B: CTA-per-pair vs warp-per-pair DP kernel on the bench level, plus a parity check of the warp kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, twilight_b200
ids, rows, weights, pairs = bench.build_level_batch(4096, 1500, seed=1000)
ctx = twilight_b200.Context()
ref = None
for name, opts in (("cta", {"dp_kernel": 1}),):
    for k, v in opts.items(): ctx.set_option(k, v)
    for _ in range(3):
        ctx.rows_upload(ids, rows, weights)
        outs = ctx.align_level(pairs)
        ph = ctx.level_phase_ms()
    cells = sum(o.cells for o in outs)
    sig = [(o.status, o.cells, o.tiles, o.path.tobytes()) for o in outs]
    if ref is None: ref = sig
    print("%-9s dp %.2f ms %.1f GCUPS launches %d identical-to-cta %s failed %d" % (name, ph[2], cells / ph[2] / 1e6, ctx.launch_count(), sig == ref, sum(o.status != 0 for o in outs)))
