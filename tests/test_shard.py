"""Host-side sharding logic (CPU): LPT partition, subtree affinity, and a world_size-2 gloo run in which every rank
aligns its own shard of a level (with the CPU oracle standing in for the device) and rank 0 gathers the results."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_balances_and_covers():
    from twilight_b200 import shard
    rng = np.random.default_rng(0)
    costs = rng.integers(1000, 9000, 257).astype(float)
    parts = shard.lpt_partition(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(257))
    loads = [costs[p].sum() for p in parts]
    assert max(loads) - min(loads) <= costs.max()


def test_subtree_affinity_keeps_subtrees_together():
    from twilight_b200 import shard, synth
    tree = synth.random_tree(200, seed=3)
    levels = synth.levels_bottom_up(tree)
    owner = shard.subtree_affinity(levels, 8, tree.n_nodes)
    assert owner.min() >= 0 and owner.max() <= 7
    counts = np.bincount(owner[:200], minlength=8)
    assert counts.min() > 0 and counts.max() <= 3 * 200 // 8
    moves = sum(1 for lv in levels for a, b, p in lv if owner[a] != owner[b])
    assert moves <= 7          # rows cross ranks only at the (world-1) joins at the top of the tree


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from tests import oracle_lib as ol
    from twilight_b200 import shard, synth
    raw = synth.profile_pair_batch(10, 300, seed=5)
    costs = [p["freq_ref"].shape[0] + p["freq_qry"].shape[0] for p in raw]
    mine = shard.lpt_partition(costs, world)[rank]
    cfg = ol.TalcoCfg()
    res = {}
    for i in mine:
        p = raw[i]
        path, err, cells, tiles, _ = ol.port_talco(cfg, p["freq_ref"], p["freq_qry"], p["gap_open_ref"], p["gap_ext_ref"], p["gap_open_qry"],
                                                   p["gap_ext_qry"], p["ref_num"], p["qry_num"])
        res[i] = (path.tobytes(), err, cells)
    slowest = shard.allreduce_max(float(len(mine)), dist)
    got = shard.gather_objects(res, dist, dst=0)
    if rank == 0:
        merged = {}
        for part in got:
            merged.update(part)
        np.save(out_path, np.array([len(merged), int(slowest), sum(v[2] for v in merged.values())]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_level_shard(tmp_path):
    import torch.multiprocessing as mp
    from tests import oracle_lib as ol
    from twilight_b200 import synth
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    n, slowest, cells = np.load(out)
    raw = synth.profile_pair_batch(10, 300, seed=5)
    cfg = ol.TalcoCfg()
    want = sum(ol.port_talco(cfg, p["freq_ref"], p["freq_qry"], p["gap_open_ref"], p["gap_ext_ref"], p["gap_open_qry"], p["gap_ext_qry"],
                             p["ref_num"], p["qry_num"])[2] for p in raw)
    assert n == 10 and slowest == 5 and cells == want


def _msa_worker(rank, world, port, out_path):
    import pickle
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from tests.fake_ctx import FakeContext
    from twilight_b200 import msa, synth
    tree = synth.random_tree(24, seed=9, mean_blen=0.05)
    seqs = synth.evolve(tree, 200, seed=9)
    w = np.random.default_rng(2).uniform(0.5, 1.5, 24).astype(np.float32)
    rows, st, root_owner = msa.progressive_align_sharded(FakeContext(), tree, seqs, w, dist)
    if rank == root_owner:
        pickle.dump((rows, st.aln_len), open(out_path, "wb"))
    counts = [None] * world
    dist.all_gather_object(counts, st.pairs)
    assert sum(counts) == 23 and min(counts) > 0          # every rank aligned something, nothing was aligned twice
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_msa_equals_single_process(tmp_path):
    """Subtree sharding + node migration over gloo (world_size 2) gives the same MSA as one process."""
    import pickle
    import torch.multiprocessing as mp
    from tests.fake_ctx import FakeContext
    from twilight_b200 import msa, synth
    out = str(tmp_path / "rows.pkl")
    port = 29500 + ((os.getpid() + 7) % 2000)
    mp.spawn(_msa_worker, args=(2, port, out), nprocs=2, join=True)
    rows2, aln_len2 = pickle.load(open(out, "rb"))
    tree = synth.random_tree(24, seed=9, mean_blen=0.05)
    seqs = synth.evolve(tree, 200, seed=9)
    w = np.random.default_rng(2).uniform(0.5, 1.5, 24).astype(np.float32)
    rows1, st1 = msa.progressive_align(FakeContext(), tree, seqs, w)
    assert aln_len2 == st1.aln_len
    assert rows2 == rows1


def test_subtree_affinity_on_a_deep_unbalanced_tree():
    """A caterpillar guide tree of 5000 leaves is 4999 levels deep: the partition must not recurse."""
    from twilight_b200 import shard, synth
    tree = synth.random_tree(5000, seed=3, shape="caterpillar")
    levels = synth.levels_bottom_up(tree)
    owner = shard.subtree_affinity(levels, 4, tree.n_nodes)
    assert (owner >= 0).all() and set(owner[:5000].tolist()) <= {0, 1, 2, 3}
    assert owner[tree.root] >= 0
