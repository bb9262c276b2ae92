"""Divergence report between two alignments of the same sequences (the north star's tolerance clause: where a
floating-point tie flips, the difference must stay below 0.1 % in SP and TC score and the affected columns are counted).

SP = residue pairs aligned in the reference alignment that are also aligned in the test alignment / pairs of the reference;
TC = reference columns (with >= 2 residues) reproduced exactly as one column of the test alignment / such columns.
Both are computed from per-residue column labels in O(total residues)."""
import numpy as np


def read_fasta(path):
    rows, name, parts = {}, None, []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    rows[name] = b"".join(parts)
                name, parts = line[1:].split()[0], []
            else:
                parts.append(line)
    if name is not None:
        rows[name] = b"".join(parts)
    return rows


def _residue_columns(rows, names):
    cols, owner = [], []
    for k, n in enumerate(names):
        a = np.frombuffer(rows[n], np.uint8)
        c = np.flatnonzero((a != ord("-")) & (a != ord(".")))
        cols.append(c)
        owner.append(np.full(len(c), k, np.int64))
    return np.concatenate(cols), np.concatenate(owner)


def _pairs(labels):
    _, cnt = np.unique(labels, return_counts=True)
    return int((cnt.astype(np.int64) * (cnt - 1) // 2).sum())


def compare(ref_path, test_path):
    ref, test = read_fasta(ref_path), read_fasta(test_path)
    if set(ref) != set(test):
        return {"same_rows": False, "missing": len(set(ref) - set(test)), "extra": len(set(test) - set(ref))}
    names = sorted(ref)
    cr, _ = _residue_columns(ref, names)
    ct, _ = _residue_columns(test, names)
    if len(cr) != len(ct):
        return {"same_rows": True, "same_residues": False}
    width_t = max(len(r) for r in test.values()) + 1
    both = cr.astype(np.int64) * width_t + ct
    sp_ref, sp_both = _pairs(cr), _pairs(both)
    # a reference column is reproduced when all its residues carry one test-column label and that test column holds nothing else
    order = np.argsort(cr, kind="stable")
    crs, cts = cr[order], ct[order]
    starts = np.flatnonzero(np.r_[True, crs[1:] != crs[:-1]])
    sizes = np.diff(np.r_[starts, len(crs)])
    one_label = np.minimum.reduceat(cts, starts) == np.maximum.reduceat(cts, starts)
    t_sizes = np.bincount(ct, minlength=width_t)
    exact = one_label & (t_sizes[cts[starts]] == sizes)
    multi = sizes >= 2
    n_cols = int(multi.sum())
    tc_hit = int((exact & multi).sum())
    return {"same_rows": True, "same_residues": True, "identical": all(ref[n] == test[n] for n in names),
            "sp": sp_both / max(sp_ref, 1), "tc": tc_hit / max(n_cols, 1), "ref_columns": int(len(starts)),
            "affected_columns": int((~exact).sum()), "ref_len": max(len(r) for r in ref.values()), "test_len": width_t - 1}
