"""A/B: the DP chain inside twl_align_level vs the DP-only batch API on the very same profiles."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import twilight_b200
from twilight_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
ids, rows, weights, pairs = bench.build_level_batch(n, 1500, seed=1000)
ctx = twilight_b200.Context()
for _ in range(3):
    ctx.rows_upload(ids, rows, weights)
    outs = ctx.align_level(pairs)
    ph = ctx.level_phase_ms()
cells = sum(o.cells for o in outs)
print("level pipeline: dp phase %.2f ms, %.1f GCUPS, launches %d" % (ph[2], cells / ph[2] / 1e6, ctx.launch_count()))
pp = []
for k in range(n):
    fr, fq = ctx.level_fetch(k, api.F_DP_PROFILE[0]), ctx.level_fetch(k, api.F_DP_PROFILE[1])
    pp.append(twilight_b200.ProfilePairIn(fr[:, :6].copy(), fq[:, :6].copy(), fr[:, 6].copy(), fr[:, 7].copy(), fq[:, 6].copy(), fq[:, 7].copy(),
                                          pairs[k].ref.aln_num, pairs[k].qry.aln_num))
ctx.stage(pp)
for _ in range(4):
    ctx.run(); ms = ctx.kernel_ms()
o2 = ctx.fetch(want_paths=False)
print("dp-only batch on the same profiles: %.2f ms, %.1f GCUPS, launches %d, cells equal %s" % (ms, cells / ms / 1e6, ctx.launch_count(), sum(o.cells for o in o2) == cells))
widths = [o.cells / max(o.diagonals, 1) for o in o2]
print("mean band %.1f max mean band %.1f" % (np.mean(widths), np.max(widths)))
raw, pp3 = bench.make_batch(n, 1500, 1000) if hasattr(bench, "make_batch") else (None, None)
import json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import bench_dp
raw, pp3 = bench_dp.make_batch(n, 1500, 1000)
ctx.stage(pp3)
for _ in range(4):
    ctx.run(); ms3 = ctx.kernel_ms()
o3 = ctx.fetch(want_paths=False)
c3 = sum(o.cells for o in o3)
print("dp-only batch on bench_dp profiles: %.2f ms, %.1f GCUPS" % (ms3, c3 / ms3 / 1e6))
for name, oo in (("level-data", o2), ("bench_dp-data", o3)):
    print(name, "tiles/pair %.2f diag/pair %.0f cells/pair %.0f band %.1f" % (np.mean([o.tiles for o in oo]), np.mean([o.diagonals for o in oo]), np.mean([o.cells for o in oo]), np.mean([o.cells / o.diagonals for o in oo])))
gq = np.mean([np.mean(p.freq_qry[:, 5] != 0) for p in pp]); gq3 = np.mean([np.mean(p.freq_qry[:, 5] != 0) for p in pp3])
print("fraction of columns with gaps: level-data %.3f bench_dp-data %.3f" % (gq, gq3))
print("denominators level-data", sorted(set(p.ref_num * p.qry_num for p in pp))[:8], "bench_dp", sorted(set(p.ref_num * p.qry_num for p in pp3))[:8])
