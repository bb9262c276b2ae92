"""Python mirror of the reference's progressive driver for the level kernel: schedules a guide tree bottom-up
(getProgressivePairs mode 0, src/progressive.cpp:52-68), keeps the per-node bookkeeping the reference keeps in
phylogeny::Node (seqsIncluded, alnLen, alnNum, alnWeight, msaFreq) and calls `Context.align_level` once per level
(src/progressive.cpp:177-180). Rows stay resident on the device between levels. Used by bench.py and the tests; the
drop-in C++ binding for the real CLI is twilight_b200/host/alignment_b200.cpp."""
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import api, synth


@dataclass
class NodeBook:
    ids: List[int]
    aln_len: int
    aln_num: int
    aln_weight: float
    msa_freq: Optional[np.ndarray] = None


@dataclass
class MsaStats:
    levels: int = 0
    pairs: int = 0
    cells: int = 0
    failed: int = 0
    device_ms: float = 0.0
    phase_ms: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])
    wall_s: float = 0.0
    launches: int = 0
    profile_bytes: int = 0      # algorithmic bytes of the profile-build kernel (rows read + raw profile written)
    update_bytes: int = 0       # algorithmic bytes of the row-update kernel (rows read + written + path)
    aln_len: int = 0


def progressive_align(ctx: api.Context, tree: synth.Tree, seqs: Sequence[bytes], weights: Sequence[float], gappy: float = 0.95,
                      cache_threshold: int = 1000, download: bool = True):
    """Returns (aligned rows in leaf order or None, MsaStats)."""
    st = MsaStats()
    t0 = time.perf_counter()
    n = tree.n_leaves
    ctx.rows_clear()
    ctx.rows_upload(list(range(n)), seqs, weights)
    book: Dict[int, NodeBook] = {i: NodeBook([i], len(seqs[i]), 1, float(np.float32(weights[i]))) for i in range(n)}
    P = ctx.P
    for level in synth.levels_bottom_up(tree):
        pairs = []
        for a, b, _ in level:
            x, y = book[a], book[b]
            pairs.append(api.LevelPairIn(api.NodeSideIn(x.ids, x.aln_len, x.aln_num, x.aln_weight, x.msa_freq),
                                         api.NodeSideIn(y.ids, y.aln_len, y.aln_num, y.aln_weight, y.msa_freq)))
            for nd in (x, y):
                if nd.msa_freq is None:
                    st.profile_bytes += nd.aln_num * nd.aln_len + nd.aln_len * P * 4
        outs = ctx.align_level(pairs, task=0, gappy=gappy, cache_threshold=cache_threshold)
        ph = ctx.level_phase_ms()
        st.phase_ms = [p + q for p, q in zip(st.phase_ms, ph)]
        st.device_ms += sum(ph)
        st.launches += ctx.launch_count()
        st.levels += 1
        for k, ((a, b, parent), o) in enumerate(zip(level, outs)):
            x, y = book.pop(a), book.pop(b)
            st.pairs += 1
            st.cells += o.cells
            if o.status != 0:
                st.failed += 1
                raise api.TwilightError(f"pair ({a},{b}) failed with status {o.status}; deferred re-alignment is the caller's job")
            freq = ctx.level_fetch(k, api.F_FREQ_MERGED) if o.merged_freq else None
            new_len = len(o.path)
            st.update_bytes += (x.aln_num * (x.aln_len + new_len) + y.aln_num * (y.aln_len + new_len)) + new_len
            book[parent] = NodeBook(x.ids + y.ids, new_len, x.aln_num + y.aln_num,
                                    float(np.float32(x.aln_weight) + np.float32(y.aln_weight)), freq)
    root = book[tree.root]
    st.aln_len = root.aln_len
    rows = None
    if download:
        got = ctx.rows_download(root.ids)
        rows = [None] * n
        for i, r in zip(root.ids, got):
            rows[i] = r
    st.wall_s = time.perf_counter() - t0
    return rows, st


def progressive_align_sharded(ctx, tree: synth.Tree, seqs: Sequence[bytes], weights: Sequence[float], dist, gappy: float = 0.95,
                              cache_threshold: int = 1000, download: bool = True):
    """Multi-GPU progressive alignment, one process per GPU. The guide tree is cut into `world` subtrees of
    similar leaf count (shard.subtree_affinity); every rank aligns its subtrees with the level kernel, no collective on the
    data path. Where a join crosses ranks (at most world-1 joins, at the top of the tree) the finished child node — its
    rows and its profile bookkeeping — moves to the parent's rank with one all_gather_object per such level (NCCL over
    NVLink on GPUs, gloo in the CPU tests). Returns (rows in leaf order on the root's owner else None, MsaStats, owner_of_root)."""
    from . import shard
    rank, world = dist.get_rank(), dist.get_world_size()
    st = MsaStats()
    t0 = time.perf_counter()
    n = tree.n_leaves
    levels = synth.levels_bottom_up(tree)
    owner = shard.subtree_affinity(levels, world, tree.n_nodes)
    mine = [i for i in range(n) if owner[i] == rank]
    ctx.rows_clear()
    if mine:
        ctx.rows_upload(mine, [seqs[i] for i in mine], [weights[i] for i in mine])
    book: Dict[int, NodeBook] = {i: NodeBook([i], len(seqs[i]), 1, float(np.float32(weights[i]))) for i in mine}
    for level in levels:
        # nodes that finish on one rank and are consumed on another
        moving = [(c, int(owner[c]), int(owner[p])) for a, b, p in level for c in (a, b) if owner[c] != owner[p]]
        if moving:
            device_path = hasattr(ctx, "rows_export") and dist.get_backend() == "nccl"
            if device_path:
                # rows travel GPU -> GPU (NCCL send/recv over NVLink) in one packed device buffer per node; only the node's
                # bookkeeping (ids, lengths, weights, cached msaFreq) goes through the object collective
                import torch
                out = {}
                for c, src, dst in moving:
                    if src == rank:
                        nb = book[c]
                        out[c] = (dst, nb, [float(w) for w in np.asarray(weights)[nb.ids]])
                gathered = [None] * world
                dist.all_gather_object(gathered, out)
                meta = {c: v for part in gathered for c, v in part.items()}
                for c, src, dst in moving:
                    _, nb, w = meta[c]
                    total = len(nb.ids) * ((nb.aln_len + 15) & ~15)
                    if src == rank:
                        buf = torch.empty(max(total, 16), dtype=torch.uint8, device="cuda")
                        ctx.rows_export(nb.ids, buf.data_ptr(), buf.numel())
                        dist.send(buf, dst)
                        book.pop(c)
                    elif dst == rank:
                        buf = torch.empty(max(total, 16), dtype=torch.uint8, device="cuda")
                        dist.recv(buf, src)
                        torch.cuda.current_stream().synchronize()    # the import runs on the context's own stream
                        lens = np.full(len(nb.ids), nb.aln_len, np.int32)
                        offs = np.arange(len(nb.ids), dtype=np.int64) * ((nb.aln_len + 15) & ~15)
                        ctx.rows_import(nb.ids, lens, w, buf.data_ptr(), offs)
                        book[c] = nb
            else:
                out = {}
                for c, src, dst in moving:
                    if src == rank:
                        nb = book.pop(c)
                        out[c] = (dst, nb, ctx.rows_download(nb.ids), [float(w) for w in np.asarray(weights)[nb.ids]])
                gathered = [None] * world
                dist.all_gather_object(gathered, out)
                for part in gathered:
                    for c, (dst, nb, rows, w) in part.items():
                        if dst == rank:
                            ctx.rows_upload(nb.ids, rows, w)
                            book[c] = nb
        todo = [(a, b, p) for a, b, p in level if owner[p] == rank]
        if not todo:
            continue
        pairs = []
        for a, b, _ in todo:
            x, y = book[a], book[b]
            pairs.append(api.LevelPairIn(api.NodeSideIn(x.ids, x.aln_len, x.aln_num, x.aln_weight, x.msa_freq),
                                         api.NodeSideIn(y.ids, y.aln_len, y.aln_num, y.aln_weight, y.msa_freq)))
        outs = ctx.align_level(pairs, task=0, gappy=gappy, cache_threshold=cache_threshold)
        ph = ctx.level_phase_ms()
        st.phase_ms = [p + q for p, q in zip(st.phase_ms, ph)]
        st.device_ms += sum(ph)
        st.launches += ctx.launch_count()
        st.levels += 1
        for k, ((a, b, parent), o) in enumerate(zip(todo, outs)):
            x, y = book.pop(a), book.pop(b)
            st.pairs += 1
            st.cells += o.cells
            if o.status != 0:
                raise api.TwilightError(f"pair ({a},{b}) failed with status {o.status}")
            freq = ctx.level_fetch(k, api.F_FREQ_MERGED) if o.merged_freq else None
            book[parent] = NodeBook(x.ids + y.ids, len(o.path), x.aln_num + y.aln_num,
                                    float(np.float32(x.aln_weight) + np.float32(y.aln_weight)), freq)
    root_owner = int(owner[tree.root])
    rows = None
    if rank == root_owner:
        root = book[tree.root]
        st.aln_len = root.aln_len
        if download:
            got = ctx.rows_download(root.ids)
            rows = [None] * n
            for i, r in zip(root.ids, got):
                rows[i] = r
    st.wall_s = time.perf_counter() - t0
    return rows, st, root_owner
