"""The SP/TC divergence reporter on hand-made alignments (CPU)."""
from tests.msa_compare import compare


def _write(path, rows):
    with open(path, "wb") as f:
        for n, r in rows.items():
            f.write(b">" + n + b"\n" + r + b"\n")


def test_identical_and_shifted(tmp_path):
    a = {b"s1": b"ACGT-A", b"s2": b"AC-TTA", b"s3": b"ACGTTA"}
    _write(tmp_path / "a.fa", a)
    _write(tmp_path / "b.fa", a)
    r = compare(str(tmp_path / "a.fa"), str(tmp_path / "b.fa"))
    assert r["identical"] and r["sp"] == 1.0 and r["tc"] == 1.0 and r["affected_columns"] == 0
    # an all-gap column inserted in the test alignment changes bytes but no aligned pair
    b = {k: v[:2] + b"-" + v[2:] for k, v in a.items()}
    _write(tmp_path / "b.fa", b)
    r = compare(str(tmp_path / "a.fa"), str(tmp_path / "b.fa"))
    assert not r["identical"] and r["sp"] == 1.0 and r["tc"] == 1.0 and r["affected_columns"] == 0
    # s1's last residue moves one column to the left: the pairs (s1:A, s2:A) and (s1:A, s3:A) are lost, two columns are affected
    c = dict(a)
    c[b"s1"] = b"ACGTA-"
    _write(tmp_path / "c.fa", c)
    r = compare(str(tmp_path / "a.fa"), str(tmp_path / "c.fa"))
    assert r["affected_columns"] == 2 and 0 < r["sp"] < 1 and r["tc"] < 1
    total_pairs = 3 + 3 + 1 + 3 + 1 + 3
    assert abs(r["sp"] - (total_pairs - 2) / total_pairs) < 1e-12
