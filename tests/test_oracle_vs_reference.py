"""Pins the CPU restatement (oracle/twl_oracle.cpp) to the unmodified reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile) and the reference build to its golden outputs. CPU only."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from tests import oracle_lib as ol, ref_msa
from tests.cli_scenarios import DATA, ROOT, run_cli
from tests.helpers import synthetic_records

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("n,L,seed,marker", [(8, 300, 0, 1024), (10, 1500, 1, 1024), (12, 700, 2, 128), (6, 2500, 3, 256), (9, 500, 4, 32)])
def test_port_talco_equals_reference(n, L, seed, marker):
    cfg, _, seqs, root, recs = synthetic_records(n, L, seed, marker)
    for r in recs:
        a, e = ol.ref_talco(cfg, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        assert e == r.error
        assert np.array_equal(a, r.aln_wo)
    # structural validity, the reference's --check (sequencedb.cpp:87-120)
    assert sorted(x.replace(b"-", b"") for x in root.rows) == sorted(seqs)
    assert all(len(x) == root.aln_len for x in root.rows)


@needs_ref
def test_port_error_codes_equal_reference():
    cfg, _, _, _, recs = synthetic_records(4, 600, 7, 1024)
    r = recs[-1]
    for xdrop, flen in ((5, 4096), (5000, 8), (40, 4096), (5000, 40), (100, 64)):
        c = ol.TalcoCfg(xdrop=xdrop, flen=flen)
        args = (c, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        a, e = ol.ref_talco(*args)
        b, f = ol.port_talco(*args)[:2]
        assert e == f and np.array_equal(a, b), (xdrop, flen, e, f)


@needs_ref
def test_reference_cli_reproduces_golden(tmp_path):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))
    for name in ("rnasim_merge_msas", "rnasim_sub_prune"):
        out, _ = run_cli(ol.REF_CLI, name, str(tmp_path))
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == gold[name]["md5"]
