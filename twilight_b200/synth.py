"""Seeded synthetic inputs of the shapes BASELINE.json names: sequences evolved along a random binary
guide tree with substitutions and short indels. Pure numpy; used by tests and by bench.py (data = "synthetic").

Tree representation: `Tree(parent, children, blen, names)` with integer node ids, leaves are 0..n_leaves-1.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

NT = np.frombuffer(b"ACGT", dtype=np.uint8)
RNA = np.frombuffer(b"ACGU", dtype=np.uint8)
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
# Robinson-Robinson background frequencies, order ACDEFGHIKLMNPQRSTVWY
AA_FREQ = np.array([0.0780, 0.0192, 0.0536, 0.0630, 0.0386, 0.0738, 0.0220, 0.0514, 0.0574, 0.0902, 0.0224, 0.0449,
                    0.0520, 0.0426, 0.0513, 0.0712, 0.0584, 0.0644, 0.0133, 0.0321])


@dataclass
class Tree:
    n_leaves: int
    parent: np.ndarray                      # [n_nodes] int32, -1 at the root
    children: List[List[int]]
    blen: np.ndarray                        # [n_nodes] float32 branch length to the parent
    root: int
    names: List[str] = field(default_factory=list)

    @property
    def n_nodes(self):
        return len(self.parent)

    def newick(self) -> str:
        out = {}
        order = self.postorder()
        for v in order:
            if not self.children[v]:
                out[v] = f"{self.names[v]}:{self.blen[v]:.6f}"
            else:
                inner = ",".join(out.pop(c) for c in self.children[v])
                out[v] = f"({inner}):{self.blen[v]:.6f}" if v != self.root else f"({inner});"
        return out[self.root]

    def postorder(self) -> List[int]:
        order, stack = [], [(self.root, False)]
        while stack:
            v, done = stack.pop()
            if done:
                order.append(v)
            else:
                stack.append((v, True))
                for c in reversed(self.children[v]):
                    stack.append((c, False))
        return order


def random_tree(n_leaves: int, seed: int = 0, shape: str = "yule", mean_blen: float = 0.05) -> Tree:
    """Random rooted binary tree. shape: 'yule' (random joins, RNASim-like depth), 'balanced', 'caterpillar'."""
    rng = np.random.default_rng(seed)
    n_nodes = 2 * n_leaves - 1
    parent = np.full(n_nodes, -1, np.int32)
    children: List[List[int]] = [[] for _ in range(n_nodes)]
    nxt = n_leaves
    if shape == "balanced":
        layer = list(range(n_leaves))
        while len(layer) > 1:
            up = []
            for a in range(0, len(layer) - 1, 2):
                children[nxt] = [layer[a], layer[a + 1]]
                parent[layer[a]] = parent[layer[a + 1]] = nxt
                up.append(nxt)
                nxt += 1
            if len(layer) % 2:
                up.append(layer[-1])
            layer = up
    elif shape == "caterpillar":
        cur = 0
        for leaf in range(1, n_leaves):
            children[nxt] = [cur, leaf]
            parent[cur] = parent[leaf] = nxt
            cur = nxt
            nxt += 1
    else:
        live = list(range(n_leaves))
        while len(live) > 1:
            a, b = rng.choice(len(live), size=2, replace=False)
            x, y = live[a], live[b]
            children[nxt] = [x, y]
            parent[x] = parent[y] = nxt
            for idx in sorted((a, b), reverse=True):
                live.pop(idx)
            live.append(nxt)
            nxt += 1
    root = n_nodes - 1
    blen = rng.exponential(mean_blen, n_nodes).astype(np.float32) + np.float32(1e-4)
    blen[root] = 0
    return Tree(n_leaves, parent, children, blen, root, [f"L{i}" for i in range(n_leaves)])


def _mutate(seq: np.ndarray, b: float, rng, alphabet: np.ndarray, indel_rate: float, probs=None) -> np.ndarray:
    L = len(seq)
    p_sub = 1.0 - np.exp(-b)
    hit = rng.random(L) < p_sub
    if hit.any():
        seq = seq.copy()
        seq[hit] = rng.choice(alphabet, size=int(hit.sum()), p=probs)
    n_ev = rng.poisson(indel_rate * b * L)
    if n_ev == 0:
        return seq
    pos = np.sort(rng.integers(0, max(L, 1), n_ev))
    parts, last = [], 0
    for p in pos:
        if p < last:
            continue
        parts.append(seq[last:p])
        ln = int(rng.geometric(0.4))
        if rng.random() < 0.5:
            parts.append(rng.choice(alphabet, size=ln, p=probs))
            last = p
        else:
            last = min(L, p + ln)
    parts.append(seq[last:])
    return np.concatenate(parts) if parts else seq


def evolve(tree: Tree, root_len: int, seed: int = 0, kind: str = "rna", indel_rate: float = 0.03,
           n_frac: float = 0.0) -> List[bytes]:
    """Evolve a random root sequence down the tree; returns the leaf sequences (bytes) in leaf-id order."""
    rng = np.random.default_rng(seed + 7919)
    alphabet = {"rna": RNA, "dna": NT, "protein": AA}[kind]
    probs = AA_FREQ / AA_FREQ.sum() if kind == "protein" else None
    seqs: List[Optional[np.ndarray]] = [None] * tree.n_nodes
    seqs[tree.root] = rng.choice(alphabet, size=root_len, p=probs)
    stack = [tree.root]
    leaves: List[Optional[bytes]] = [None] * tree.n_leaves
    while stack:
        v = stack.pop()
        s = seqs[v]
        for c in tree.children[v]:
            seqs[c] = _mutate(s, float(tree.blen[c]), rng, alphabet, indel_rate, probs)
            stack.append(c)
        if not tree.children[v]:
            if n_frac > 0:
                s = s.copy()
                s[rng.random(len(s)) < n_frac] = ord("N")
            leaves[v] = s.tobytes()
        seqs[v] = None
    return leaves  # type: ignore


def levels_bottom_up(tree: Tree):
    """Sibling pairs grouped by level, the schedule of getProgressivePairs mode 0 (progressive.cpp:52-68):
    level(node) = 1 + max(level(children)) with leaves at 0; returns [[(first_child, second_child, parent), ...], ...]."""
    lvl = np.zeros(tree.n_nodes, np.int32)
    out = {}
    for v in tree.postorder():
        ch = tree.children[v]
        if ch:
            lvl[v] = 1 + max(lvl[c] for c in ch)
            out.setdefault(int(lvl[v]) - 1, []).append((ch[0], ch[1], v))
    return [out[k] for k in sorted(out)]


# ------------------------------------------------------------------------------------------------------------------
# Level-shaped batches of profile pairs (inputs at the Align_freq boundary) without running an aligner: each side is a
# small star-shaped family whose members differ from the family ancestor by substitutions and masked-out deletions, so
# the family is trivially aligned; the two family ancestors are diverged with indels like siblings of a guide tree.
# ------------------------------------------------------------------------------------------------------------------
def _letters_to_index(seq: np.ndarray, kind: str) -> np.ndarray:
    lut = np.full(256, 4 if kind != "protein" else 20, np.int64)
    if kind == "protein":
        for n, ch in enumerate(AA):
            lut[ch] = n
        lut[ord("-")] = 21
    else:
        for ch, n in ((ord("A"), 0), (ord("C"), 1), (ord("G"), 2), (ord("T"), 3), (ord("U"), 3)):
            lut[ch] = n
        lut[ord("-")] = 5
    return lut[seq]


def family_profile(anc: np.ndarray, members: int, rng, kind: str, sub_rate=0.08, del_rate=0.02, gap_open=-50.0,
                   gap_extend=-5.0):
    """Weighted column counts [len][P] (columns sum to `members`), and position-specific gap penalties computed with the
    ClustalW-style rule of calculatePSGP (alignment-helper.cpp:168-219)."""
    alphabet = {"rna": RNA, "dna": NT, "protein": AA}[kind]
    P = 22 if kind == "protein" else 6
    L = len(anc)
    w = rng.uniform(0.5, 1.5, members).astype(np.float32)
    w = (w / w.sum() * members).astype(np.float32)
    prof = np.zeros((L, P), np.float32)
    for m in range(members):
        s = anc.copy()
        hit = rng.random(L) < sub_rate
        s[hit] = rng.choice(alphabet, size=int(hit.sum()))
        if members > 1 and del_rate > 0:
            starts = np.flatnonzero(rng.random(L) < del_rate / 3)
            for st in starts:
                s[st:st + int(rng.geometric(0.4))] = ord("-")
        idx = _letters_to_index(s, kind)
        np.add.at(prof, (np.arange(L), idx), w[m])
    g = prof[:, P - 1].astype(np.float64)
    scale = 1.0 if kind == "protein" else 0.5
    keep = (members - g) / members
    gop = np.where(g > 0, np.minimum(np.float32(gap_open * 0.1), (np.float32(gap_open * scale) * keep).astype(np.float32)), np.float32(gap_open))
    gex = np.where(g > 0, np.minimum(np.float32(gap_extend * 0.2), (gap_extend * keep).astype(np.float32)), np.float32(gap_extend))
    return prof, gop.astype(np.float32), gex.astype(np.float32), float(members)


def profile_pair_batch(n_pairs: int, length: int, seed: int = 0, kind: str = "rna", members=(1, 2, 4, 8),
                       divergence: float = 0.15, indel_rate: float = 0.03):
    """n_pairs sibling profile pairs of ~`length` columns. Returns a list of dicts with the twl_profile_pair fields."""
    rng = np.random.default_rng(seed)
    alphabet = {"rna": RNA, "dna": NT, "protein": AA}[kind]
    probs = AA_FREQ / AA_FREQ.sum() if kind == "protein" else None
    out = []
    for _ in range(n_pairs):
        root = rng.choice(alphabet, size=int(length * rng.uniform(0.97, 1.03)), p=probs)
        a = _mutate(root, divergence / 2, rng, alphabet, indel_rate, probs)
        b = _mutate(root, divergence / 2, rng, alphabet, indel_rate, probs)
        ma, mb = int(rng.choice(members)), int(rng.choice(members))
        fr, gor, ger, nr = family_profile(a, ma, rng, kind)
        fq, goq, geq, nq = family_profile(b, mb, rng, kind)
        out.append(dict(freq_ref=fr, freq_qry=fq, gap_open_ref=gor, gap_ext_ref=ger, gap_open_qry=goq, gap_ext_qry=geq,
                        ref_num=nr, qry_num=nq))
    return out


def family_rows(anc: np.ndarray, members: int, rng, kind: str, sub_rate=0.08, del_rate=0.02):
    """`members` aligned rows (bytes, '-' for deletions) derived from one ancestor; all rows have len(anc) columns."""
    alphabet = {"rna": RNA, "dna": NT, "protein": AA}[kind]
    L = len(anc)
    rows = []
    for _ in range(members):
        s = anc.copy()
        hit = rng.random(L) < sub_rate
        s[hit] = rng.choice(alphabet, size=int(hit.sum()))
        if members > 1 and del_rate > 0:
            for st in np.flatnonzero(rng.random(L) < del_rate / 3):
                s[st:st + int(rng.geometric(0.4))] = ord("-")
        rows.append(s.tobytes())
    return rows


def level_rows_batch(n_pairs: int, length: int, seed: int = 0, kind: str = "rna", members=(1, 2, 4, 8), divergence: float = 0.15,
                     indel_rate: float = 0.03):
    """A guide-tree-level-shaped batch at the row level: n_pairs sibling nodes, each node a small aligned family of
    1-8 rows of ~`length` columns. Returns [(ref_rows, qry_rows), ...]."""
    rng = np.random.default_rng(seed)
    alphabet = {"rna": RNA, "dna": NT, "protein": AA}[kind]
    probs = AA_FREQ / AA_FREQ.sum() if kind == "protein" else None
    out = []
    for _ in range(n_pairs):
        root = rng.choice(alphabet, size=int(length * rng.uniform(0.97, 1.03)), p=probs)
        a = _mutate(root, divergence / 2, rng, alphabet, indel_rate, probs)
        b = _mutate(root, divergence / 2, rng, alphabet, indel_rate, probs)
        out.append((family_rows(a, int(rng.choice(members)), rng, kind), family_rows(b, int(rng.choice(members)), rng, kind)))
    return out


# ------------------------------------------------------------------------------------------------------------------
# Whole data sets on disk (FASTA + Newick) for the drop-in CLI: the named shapes of BASELINE.json at reduced N.
# ------------------------------------------------------------------------------------------------------------------
def tree_depth(tree: Tree) -> int:
    depth = np.zeros(tree.n_nodes, np.int32)
    for v in reversed(tree.postorder()):
        for c in tree.children[v]:
            depth[c] = depth[v] + 1
    return int(depth.max())


def write_dataset(prefix: str, tree: Tree, seqs: List[bytes], width: int = 0):
    """<prefix>.fa (one record per leaf, names L<i>) and <prefix>.nwk."""
    with open(prefix + ".fa", "wb") as f:
        for name, s in zip(tree.names, seqs):
            f.write(b">" + name.encode() + b"\n")
            if width:
                for k in range(0, len(s), width):
                    f.write(s[k:k + width] + b"\n")
            else:
                f.write(s + b"\n")
    with open(prefix + ".nwk", "w") as f:
        f.write(tree.newick() + "\n")


DATASETS = {
    # name: (leaves, root length, kind, root-to-tip divergence, indel rate, N fraction, seed)
    "rna_1k": (1000, 1500, "rna", 0.30, 0.03, 0.0, 11),        # C3 rung 10^3 (forces nothing special: < 1000 per node until the top)
    "rna_3k": (3000, 1500, "rna", 0.30, 0.03, 0.0, 12),        # msaFreq caching + parking (>= 1000 sequences per node)
    "rna_10k": (10000, 1500, "rna", 0.30, 0.03, 0.0, 13),      # C3 rung 10^4
    "rna_100k": (100000, 1500, "rna", 0.30, 0.03, 0.0, 18),    # C3 rung 10^5 (bench only: the CPU reference needs ~20 minutes)
    "rna_1m": (1000000, 1500, "rna", 0.30, 0.03, 0.0, 19),     # C3 itself: 10^6 leaves (no golden md5: the CPU reference would need hours; run with --check)
    "sars_64": (64, 29700, "dna", 0.002, 0.002, 0.0001, 14),   # C4 shape: 30 kb, near-identical, rare short indels
    "sars_50k": (50000, 29700, "dna", 0.002, 0.002, 0.0001, 22),  # C4 itself: 5*10^4 genomes of 30 kb (no golden md5; run with --check)
    "prot_2k": (2000, 400, "protein", 0.45, 0.02, 0.0, 15),    # C5 shape: 400 aa, BLOSUM62
    "prot_200k": (200000, 400, "protein", 0.45, 0.02, 0.0, 21),  # C5 itself: 2*10^5 proteins (no golden md5; run with --check)
    # 300 leaves of which 3 carry 15 % N (low quality, io.cpp:131-163: excluded, or deferred with --no-filtering), 3 are
    # unrelated random sequences and 2 are half-length fragments (pairs that fail the x-drop rule are deferred and
    # re-aligned with the widening ladder, alignment-cpu.cpp:108-129, progressive.cpp:275-298)
    "rna_outliers": (300, 1500, "rna", 0.30, 0.03, 0.0, 16),
}


def make_dataset(name: str, out_dir: str) -> str:
    """Writes the named synthetic data set and returns the path prefix (.fa / .nwk)."""
    import os
    n, length, kind, div, indel, nfrac, seed = DATASETS[name]
    tree = random_tree(n, seed=seed, mean_blen=1.0)
    # scale branch lengths so that the mean root-to-tip distance is `div`
    dist = np.zeros(tree.n_nodes, np.float64)
    for v in reversed(tree.postorder()):
        for c in tree.children[v]:
            dist[c] = dist[v] + float(tree.blen[c])
    scale = div / float(dist[:n].mean())
    tree.blen = (tree.blen.astype(np.float64) * scale).astype(np.float32)
    seqs = evolve(tree, length, seed=seed, kind=kind, indel_rate=indel, n_frac=nfrac)
    if name == "rna_outliers":
        rng = np.random.default_rng(seed + 1)
        pick = rng.choice(n, size=8, replace=False)
        for k, leaf in enumerate(pick):
            s = np.frombuffer(seqs[leaf], np.uint8).copy()
            if k < 3:
                s[rng.random(len(s)) < 0.15] = ord("N")
            elif k < 6:
                s = rng.choice(RNA, size=len(s))
            else:
                s = s[: len(s) // 2]
            seqs[leaf] = s.tobytes()
    os.makedirs(out_dir, exist_ok=True)
    prefix = os.path.join(out_dir, name)
    write_dataset(prefix, tree, seqs)
    return prefix
