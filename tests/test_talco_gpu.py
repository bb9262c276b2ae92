"""GPU parity of the TALCO-XDrop DP + traceback (C ABI twl_align_profiles) against the CPU oracle.

Bar: bit-exact alignment paths, identical errorType, identical cell / tile / diagonal counts."""
import numpy as np
import pytest

from tests import oracle_lib as ol
from tests.helpers import records_to_pairs, synthetic_records

pytestmark = pytest.mark.gpu

CASES = [
    # n_leaves, root_len, seed, marker
    (8, 300, 0, 1024),     # single tile per pair
    (16, 1500, 1, 1024),   # RNASim-shaped, 2-4 tiles per pair
    (12, 700, 2, 128),     # small marker: many tiles, convergence logic on every tile
    (6, 4000, 3, 256),
    (10, 900, 4, 64),
    (5, 2600, 5, 1024),
]


@pytest.fixture(scope="module")
def ctxs():
    import twilight_b200
    cache = {}

    def get(marker):
        if marker not in cache:
            cache[marker] = twilight_b200.Context(marker=marker)
        return cache[marker]
    yield get
    for c in cache.values():
        c.close()


@pytest.mark.parametrize("n,L,seed,marker", CASES)
def test_paths_match_port(ctxs, n, L, seed, marker):
    cfg, _, _, _, recs = synthetic_records(n, L, seed, marker)
    ctx = ctxs(marker)
    outs = ctx.align_profiles(records_to_pairs(recs, cfg))
    assert len(outs) == len(recs)
    for k, (o, r) in enumerate(zip(outs, recs)):
        assert o.status == r.error == 0, f"pair {k}: status {o.status} vs {r.error}"
        assert o.tiles == r.tiles, f"pair {k}: tiles {o.tiles} vs {r.tiles}"
        assert o.cells == r.cells, f"pair {k}: cells {o.cells} vs {r.cells}"
        assert len(o.path) == len(r.aln_wo), f"pair {k}: path length {len(o.path)} vs {len(r.aln_wo)}"
        assert np.array_equal(o.path, r.aln_wo), f"pair {k}: first diff at {int(np.argmax(o.path != r.aln_wo))}"


def test_paths_match_reference_build(ctxs):
    if not ol.have_ref():
        pytest.skip("oracle/_ref/libtalco_ref.so not built")
    cfg, _, _, _, recs = synthetic_records(10, 1200, 11, 256)
    outs = ctxs(256).align_profiles(records_to_pairs(recs, cfg))
    for o, r in zip(outs, recs):
        a, e = ol.ref_talco(cfg, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        assert o.status == e
        assert np.array_equal(o.path, a)


def test_error_codes(ctxs):
    """errorType 1 (band dies) and 2 (band wider than fLen) must come back exactly as the reference reports them."""
    import twilight_b200
    rng = np.random.default_rng(5)
    cfg, _, _, _, recs = synthetic_records(4, 600, 7, 1024)
    r = recs[-1]
    ctx = ctxs(1024)
    for xdrop, flen in ((5, 4096), (5000, 8), (40, 4096), (5000, 40)):
        c = ol.TalcoCfg(xdrop=xdrop, flen=flen)
        want, err, cells, tiles, _ = ol.port_talco(c, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], r.ref.aln_num, r.qry.aln_num)
        pair = twilight_b200.ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1],
                                           r.ref.aln_num, r.qry.aln_num, xdrop=xdrop, flen=flen)
        out = ctx.align_profiles([pair])[0]
        assert out.status == err, (xdrop, flen, out.status, err)
        assert np.array_equal(out.path, want)


def test_wide_band_goes_through_global_state(ctxs):
    """A huge x-drop keeps the whole matrix alive: bands wider than the shared-memory capacity are re-run by the wide
    variant and must still match."""
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(2, 1800, 9, 1024)
    r = recs[0]
    c = ol.TalcoCfg(xdrop=2000000)
    want, err, cells, tiles, _ = ol.port_talco(c, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1)
    pair = twilight_b200.ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1, xdrop=2000000)
    out = ctxs(1024).align_profiles([pair])[0]
    assert out.status == err == 0
    assert out.cells == cells
    assert np.array_equal(out.path, want)


def test_empty_batch_and_tiny_pairs(ctxs):
    import twilight_b200
    ctx = ctxs(1024)
    assert ctx.align_profiles([]) == []
    cfg = ol.TalcoCfg()
    rng = np.random.default_rng(3)
    pairs, wants = [], []
    for (R, Q) in ((1, 1), (1, 5), (5, 1), (2, 2), (3, 40), (40, 3)):
        fr = np.zeros((R, 6), np.float32); fr[np.arange(R), rng.integers(0, 4, R)] = 1
        fq = np.zeros((Q, 6), np.float32); fq[np.arange(Q), rng.integers(0, 4, Q)] = 1
        go_r = np.full(R, -50, np.float32); ge_r = np.full(R, -5, np.float32)
        go_q = np.full(Q, -50, np.float32); ge_q = np.full(Q, -5, np.float32)
        wants.append(ol.port_talco(cfg, fr, fq, go_r, ge_r, go_q, ge_q, 1, 1))
        pairs.append(twilight_b200.ProfilePairIn(fr, fq, go_r, ge_r, go_q, ge_q, 1, 1))
    outs = ctx.align_profiles(pairs)
    for o, w in zip(outs, wants):
        assert o.status == w[1]
        assert np.array_equal(o.path, w[0])


@pytest.mark.parametrize("n,L,seed,marker", [(12, 700, 2, 128), (16, 1500, 1, 1024)])
def test_generic_kernel_matches_port(n, L, seed, marker):
    """The wide-band generic kernel (the fallback of the register-resident wavefront kernel) stays parity-green."""
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(n, L, seed, marker)
    ctx = twilight_b200.Context(marker=marker)
    ctx.set_option("force_generic", 1)
    outs = ctx.align_profiles(records_to_pairs(recs, cfg))
    ctx.close()
    for k, (o, r) in enumerate(zip(outs, recs)):
        assert o.status == r.error == 0
        assert o.cells == r.cells and o.tiles == r.tiles
        assert np.array_equal(o.path, r.aln_wo), f"pair {k}"


def test_band_overflow_chain(ctxs):
    """x-drop values that push the band over 512 and over 1024 cells walk the kernel chain wavefront<128> ->
    wavefront<256> -> generic and must still equal the oracle."""
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(2, 2400, 19, 1024)
    r = recs[0]
    for xdrop in (9000, 14000, 40000):
        c = ol.TalcoCfg(xdrop=xdrop)
        want, err, cells, tiles, _ = ol.port_talco(c, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1)
        pair = twilight_b200.ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1, xdrop=xdrop)
        out = ctxs(1024).align_profiles([pair])[0]
        assert out.status == err == 0, xdrop
        assert out.cells == cells, xdrop
        assert np.array_equal(out.path, want), xdrop


def test_exact_division_selftest(ctxs):
    """The reciprocal-based division inside the DP kernels must be bit-identical to the IEEE divide."""
    rng = np.random.default_rng(0)
    n = 4_000_000
    den = rng.integers(1, 20000, n).astype(np.float32) * rng.integers(1, 20000, n).astype(np.float32)
    den[: n // 4] = rng.integers(1, 5000, n // 4).astype(np.float32)
    num = (rng.standard_normal(n) * np.exp(rng.uniform(-12, 25, n))).astype(np.float32)
    num[::97] = 0.0
    num[1::97] = rng.integers(-4000, 4000, len(num[1::97])).astype(np.float32) * 18
    assert ctxs(1024).selftest_division(num, den) == 0


@pytest.mark.parametrize("which", ["wildcard", "random"])
def test_unstructured_matrix_path(which):
    """Matrices without the built-in nucleotide shape (e.g. --wildcard, --matrix) take the generic 5x5 score path of the
    wavefront kernel."""
    import twilight_b200
    if which == "wildcard":
        score = ol.nt_matrix(wildcard=True)
    else:
        score = np.random.default_rng(4).integers(-9, 12, (5, 5)).astype(np.float32)
        score = ((score + score.T) / 2).astype(np.float32)
    cfg = ol.TalcoCfg(score=score, marker=256)
    _, _, _, _, recs = synthetic_records(10, 900, 21, 256, cfg=cfg)
    ctx = twilight_b200.Context(score=score, marker=256)
    outs = ctx.align_profiles(records_to_pairs(recs, cfg))
    ctx.close()
    for k, (o, r) in enumerate(zip(outs, recs)):
        assert o.status == r.error == 0
        assert o.cells == r.cells
        assert np.array_equal(o.path, r.aln_wo), f"pair {k}"


@pytest.mark.parametrize("shape", [2, 3])
@pytest.mark.parametrize("n,L,seed,marker", [(12, 900, 29, 128), (6, 2500, 31, 1024)])
def test_low_latency_variant_matches_port(n, L, seed, marker, shape):
    """The instantiations used for levels with few pairs (twl_set_option latency_mode=1): 512 threads x 2 rows (shape 2) and
    512 x 1 followed by 512 x 2 (shape 3)."""
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(n, L, seed, marker)
    ctx = twilight_b200.Context(marker=marker)
    ctx.set_option("latency_mode", 1)
    ctx.set_option("latency_shape", shape)
    outs = ctx.align_profiles(records_to_pairs(recs, cfg))
    ctx.close()
    for k, (o, r) in enumerate(zip(outs, recs)):
        assert o.status == r.error == 0
        assert o.cells == r.cells and o.tiles == r.tiles
        assert np.array_equal(o.path, r.aln_wo), f"pair {k}"


@pytest.mark.parametrize("shape", [2, 3])
def test_low_latency_variant_overflows_to_wide(shape):
    """Bands wider than the 512-row window of the low-latency variant continue in the 1024-row kernel (and beyond),
    same bits as the oracle."""
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(2, 2400, 19, 1024)
    r = recs[0]
    ctx = twilight_b200.Context(marker=1024)
    ctx.set_option("latency_mode", 1)
    ctx.set_option("latency_shape", shape)
    for xdrop in (5000, 9000, 14000, 40000):
        c = ol.TalcoCfg(xdrop=xdrop)
        want, err, cells, tiles, _ = ol.port_talco(c, r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1)
        pair = twilight_b200.ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1, xdrop=xdrop)
        out = ctx.align_profiles([pair])[0]
        assert out.status == err == 0, xdrop
        assert out.cells == cells, xdrop
        assert np.array_equal(out.path, want), xdrop
    ctx.close()


@pytest.mark.parametrize("workers,every", [(0, 5), (1, 5), (8, 5), (8, 1)])
def test_co_running_wide_workers_match_port(workers, every):
    """More pairs than SMs: the narrow kernel and the wide workers run at the same time (twl_set_option wide_workers);
    pairs whose band outgrows 512 rows are handed over mid-flight. Same bits as the oracle, with and without co-run."""
    import twilight_b200
    cfg, _, _, _, recs = synthetic_records(2, 2400, 19, 1024)
    r = recs[0]
    pairs, want = [], []
    for k in range(320):
        xdrop = (9000, 14000, 40000, 9000)[k % 4] if k % every == 0 else 5000     # every == 1: every pair outgrows the narrow window
        key = xdrop
        pairs.append(twilight_b200.ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1, xdrop=xdrop))
        want.append(key)
    oracle = {}
    for xdrop in set(want):
        oracle[xdrop] = ol.port_talco(ol.TalcoCfg(xdrop=xdrop), r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1)
    ctx = twilight_b200.Context(marker=1024)
    ctx.set_option("wide_workers", workers)
    for _ in range(2):
        outs = ctx.align_profiles(pairs)
        for o, key in zip(outs, want):
            path, err, cells, tiles, _ = oracle[key]
            assert o.status == err == 0
            assert o.cells == cells and o.tiles == tiles
            assert np.array_equal(o.path, path)
    ctx.close()


def test_co_run_survives_serialised_kernels():
    """CUDA_LAUNCH_BLOCKING=1 (or a profiler's kernel replay) runs the wide workers to completion before the narrow kernel
    starts: they must give up waiting, and the clean-up launch must finish the pairs handed over afterwards."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent('''
        import numpy as np, twilight_b200
        from tests import oracle_lib as ol
        from tests.helpers import synthetic_records
        cfg, _, _, _, recs = synthetic_records(2, 2400, 19, 1024)
        r = recs[0]
        pairs, keys = [], []
        for k in range(200):
            xdrop = 9000 if k % 7 == 0 else 5000
            pairs.append(twilight_b200.ProfilePairIn(r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1, xdrop=xdrop))
            keys.append(xdrop)
        want = {x: ol.port_talco(ol.TalcoCfg(xdrop=x), r.profile[0], r.profile[1], r.gap_op[0], r.gap_ex[0], r.gap_op[1], r.gap_ex[1], 1, 1) for x in set(keys)}
        ctx = twilight_b200.Context(marker=1024)
        outs = ctx.align_profiles(pairs)
        for o, x in zip(outs, keys):
            path, err, cells, tiles, _ = want[x]
            assert o.status == err == 0 and o.cells == cells and np.array_equal(o.path, path)
        print("serialised-ok")
    ''')
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert "serialised-ok" in out.stdout, out.stdout + out.stderr
