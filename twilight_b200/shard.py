"""Sharding of a guide-tree level (or of divide-and-conquer subtrees) across the GPUs of one box.

The path has no data-path collective: same-level node pairs are independent (src/progressive.cpp:52-68)
and divide-and-conquer subtrees are independent until the final merge (src/twilight-main.cpp:139-176), so every rank
runs the level kernel on its own shard. What is exchanged is small and happens once per level / once per subtree:
per-pair results (path lengths, status) and, for the cross-subtree merge, the subtree root profiles (msaFreq,
alnLen x P floats) gathered onto rank 0. Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
from typing import List, Sequence

import numpy as np


def lpt_partition(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to `world` ranks; returns the item indices per rank.
    The cost of a pair is its anti-diagonal count times the band (~ ref_len + qry_len)."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += costs[i]
    for lst in out:
        lst.sort()
    return out


def subtree_affinity(levels, world: int, n_nodes: int) -> np.ndarray:
    """Owner rank per tree node such that a node is aligned on the rank that holds its rows: leaves of one subtree stay
    together. `levels` is synth.levels_bottom_up(tree). The top log2(world) levels collapse onto fewer ranks."""
    owner = np.full(n_nodes, -1, np.int64)
    size = np.zeros(n_nodes, np.int64)
    children = {}
    for level in levels:
        for a, b, parent in level:
            children[parent] = (a, b)
    root = levels[-1][0][2]

    # leaf counts bottom-up: levels are already in dependency order (children before parents), so no recursion is needed
    # (a caterpillar-shaped guide tree is thousands of nodes deep)
    for level in levels:
        for a, b, parent in level:
            for c in (a, b):
                if size[c] == 0:
                    size[c] = 1
            size[parent] = size[a] + size[b]
    # split the tree top-down into `world` subtrees of roughly equal leaf count
    parts = [root]
    while len(parts) < world:
        big = max((p for p in parts if p in children), key=lambda p: size[p], default=None)
        if big is None:
            break
        parts.remove(big)
        parts.extend(children[big])
    bins = lpt_partition([float(size[p]) for p in parts], world)

    for r, items in enumerate(bins):
        stack = [parts[k] for k in items]
        while stack:                                   # explicit stack: depth is unbounded for unbalanced trees
            v = stack.pop()
            owner[v] = r
            if v in children:
                stack.extend(children[v])
    # ancestors of the split points: owned by the rank of their first child (rows migrate once, at the top of the tree)
    for level in levels:
        for a, b, parent in level:
            if owner[parent] < 0:
                owner[parent] = owner[a]
    return owner


def gather_objects(obj, dist, dst: int = 0):
    """Gathers one picklable object per rank onto `dst` (subtree root profiles, per-pair results)."""
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def allreduce_max(value: float, dist, device=None) -> float:
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
