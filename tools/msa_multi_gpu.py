"""Sharded progressive MSA on N GPUs (torchrun, one rank per GPU): checks that the result equals the single-GPU MSA and
prints sequences/s. python -m torch.distributed.run --nproc-per-node N tools/msa_multi_gpu.py [leaves] [length]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import twilight_b200
from twilight_b200 import msa, synth

leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
length = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
tree = synth.random_tree(leaves, seed=31, mean_blen=0.05)
seqs = synth.evolve(tree, length, seed=31)
w = np.ones(leaves, np.float32)
ctx = twilight_b200.Context(device=local)
best = None
for _ in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows, st, root_owner = msa.progressive_align_sharded(ctx, tree, seqs, w, dist)
    dist.barrier(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
dev = torch.tensor([st.device_ms, float(st.cells), float(st.pairs)], dtype=torch.float64, device="cuda")
mx = dev.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
sm = dev.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
if rank == root_owner:
    single_rows, single = msa.progressive_align(ctx, tree, seqs, w)
    single_rows, single = msa.progressive_align(ctx, tree, seqs, w)
    same = rows == single_rows
    print(f"world {world}: {leaves} x {length}: sharded wall {best:.3f} s = {leaves / best:.0f} seqs/s (device max {float(mx[0]):.1f} ms, pairs {int(sm[2])}, "
          f"cells {int(sm[1])}) | single GPU wall {single.wall_s:.3f} s = {leaves / single.wall_s:.0f} seqs/s | identical MSA: {same}")
    assert same
dist.barrier()
ctx.close()
dist.destroy_process_group()
