"""Full-size checks (the bench level: 4096 node pairs, ~36 k rows) through size-independent properties, where the oracle
would take minutes: residue preservation, path/length consistency, determinism across repeated runs and across the two
schedules of the DP chain (narrow and wide kernels co-running vs one after the other)."""
import hashlib

import numpy as np
import pytest

import bench

pytestmark = pytest.mark.gpu

N_PAIRS = 4096


def _run(ctx, prows, plevel):
    ctx.upload_prepared(prows)
    ctx.align_level_prepared(plevel)
    ctx.download_prepared(prows)
    paths = [plevel.path_buf[int(plevel.path_offs[k]):int(plevel.path_offs[k]) + int(plevel.res[k].path_len)].copy() for k in range(plevel.n)]
    rows = [prows.buf[int(prows.offs[i]):int(prows.offs[i]) + int(prows.out_lens[i])].tobytes() for i in range(prows.n)]
    return paths, rows


def _digest(paths, rows):
    h = hashlib.sha256()
    for p in paths:
        h.update(p.tobytes()); h.update(b"|")
    for r in rows:
        h.update(r); h.update(b"|")
    return h.hexdigest()


def test_bench_level_properties_and_determinism():
    import twilight_b200
    ids, rows_in, weights, pairs = bench.build_level_batch(N_PAIRS, 1500, seed=1000)
    caps = {}
    for p in pairs:
        for sd in (p.ref, p.qry):
            for i in sd.seq_ids:
                caps[i] = p.ref.aln_len + p.qry.aln_len + 16
    digests = []
    for workers in (8, 8, 0):                                   # co-run twice (scheduling differs run to run), then serial chain
        ctx = twilight_b200.Context()
        ctx.set_option("wide_workers", workers)
        prows = ctx.prepare_rows(ids, rows_in, weights, [caps[i] for i in ids])
        plevel = ctx.prepare_level(pairs)
        paths, rows_out = _run(ctx, prows, plevel)
        assert ctx.large_restores() == 0
        ctx.close()
        digests.append(_digest(paths, rows_out))
        if len(digests) > 1:
            continue
        cells = 0
        for k, p in enumerate(pairs):
            res = plevel.res[k]
            assert res.status == 0, k
            path = paths[k]
            assert len(path) > 0 and path.min() >= 0 and path.max() <= 2
            assert int(np.count_nonzero(path != 1)) == p.ref.aln_len, k      # ops 0 and 2 consume one ref column each
            assert int(np.count_nonzero(path != 2)) == p.qry.aln_len, k
            cells += int(res.cells)
            for sd in (p.ref, p.qry):
                for i in sd.seq_ids:
                    assert len(rows_out[i]) == len(path)
                    assert rows_out[i].replace(b"-", b"") == rows_in[i].replace(b"-", b"")
            # column consistency: where the path gives the ref side a gap, every ref member has '-' (and likewise for qry)
            i0 = p.ref.seq_ids[0]
            got = np.frombuffer(rows_out[i0], np.uint8)
            assert np.all(got[path == 1] == ord("-"))
        assert cells > 5e9
    assert digests[0] == digests[1] == digests[2]
