"""bench.py's C4 block (a level of 592 pairs of 30 kb genomes through the level API) on its own: A/B of library builds via TWL_LIB."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
orig = bench.cpu_block
bench.cpu_block = lambda *a, **k: None
out = bench.run_level_config("C4", "dna", 592, 29700, 3, 2, flush, 16, dict(divergence=0.004, indel_rate=0.002, members=(1, 2, 4)), "C4")
print(json.dumps({k: out[k] for k in ("gcups_device", "gcups_dp_phase", "gcups_e2e", "ms_per_step_device", "failed_pairs", "cells_per_step")}))
