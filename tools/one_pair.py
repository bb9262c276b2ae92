"""One profile pair alone on the device (the top of a guide tree): kernel time, diagonals, time per diagonal."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twilight_b200
from twilight_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
members = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = twilight_b200.Context()
for name, val in [a.split("=") for a in sys.argv[3:]]:
    ctx.set_option(name, int(val))
recs = synth.profile_pair_batch(n, 1600, seed=3, members=(members,))
pairs = [twilight_b200.ProfilePairIn(r["freq_ref"], r["freq_qry"], r["gap_open_ref"], r["gap_ext_ref"], r["gap_open_qry"], r["gap_ext_qry"],
                                     r["ref_num"], r["qry_num"]) for r in recs]
ctx.stage(pairs)
for _ in range(3):
    ctx.run()
ms = ctx.kernel_ms()
outs = ctx.fetch()
d = max(o.diagonals for o in outs)
print("pairs %d members %d: %.3f ms, %d diagonals (max), %d tiles, %.2f us/diagonal, cells %.2e, band %.0f" %
      (n, members, ms, d, outs[0].tiles, ms * 1e3 / d, sum(o.cells for o in outs), outs[0].cells / outs[0].diagonals))
