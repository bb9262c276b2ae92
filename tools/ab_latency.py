"""A/B of the low-latency wavefront variant (256 threads x 2 rows) on a full synthetic MSA and on small batches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, twilight_b200

leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for mode, shape in ((-1, 2), (-1, 4), (1, 4)):
    ctx = twilight_b200.Context()
    ctx.set_option("latency_mode", mode)
    ctx.set_option("latency_shape", shape)
    r = bench.run_msa(ctx, leaves, 1500, seed=7, repeats=3)
    print(json.dumps({"latency_mode": mode, "shape": shape, "wall_s": r["wall_s"], "device_ms": r["device_ms"], "dp_chain_ms": r["phase_ms"]["dp_chain"],
                      "seqs_per_s_e2e": r["seqs_per_s_e2e"]}))
    ctx.close()
