"""Scale check: the first guide-tree level of a C3-shaped job (leaf x leaf pairs of ~1.5 kb RNA), n_pairs pairs in one
twl_align_level call (chunked internally). Prints wall, device phases, GCUPS."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import twilight_b200
from twilight_b200 import LevelPairIn, NodeSideIn, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(1)
t0 = time.perf_counter()
# cheap generator: a pool of 512 related sequences, pairs drawn from it (the kernels do not care that rows repeat)
root = rng.choice(synth.RNA, size=1500)
pool = [synth._mutate(root, 0.15, rng, synth.RNA, 0.03).tobytes() for _ in range(512)]
ids, rows, pairs = [], [], []
for p in range(n_pairs):
    a, b = pool[rng.integers(512)], pool[rng.integers(512)]
    ids += [2 * p, 2 * p + 1]; rows += [a, b]
    pairs.append(LevelPairIn(NodeSideIn([2 * p], len(a), 1, 1.0), NodeSideIn([2 * p + 1], len(b), 1, 1.0)))
print("generated %d pairs in %.1f s" % (n_pairs, time.perf_counter() - t0))
ctx = twilight_b200.Context()
prows = ctx.prepare_rows(ids, rows, [1.0] * len(ids), [3200] * len(ids)); plevel = ctx.prepare_level(pairs)
for it in range(2):
    t0 = time.perf_counter(); ctx.upload_prepared(prows); t1 = time.perf_counter(); ctx.align_level_prepared(plevel); t2 = time.perf_counter(); ctx.download_prepared(prows); t3 = time.perf_counter()
    ph = ctx.level_phase_ms()
    cells = sum(int(plevel.res[k].cells) for k in range(n_pairs)); bad = sum(1 for k in range(n_pairs) if plevel.res[k].status != 0)
    print("run %d: upload %.2f s | align_level %.2f s (device %.1f ms: %s) | download %.2f s | %.1f GCUPS device, %.1f GCUPS e2e, %.0f seqs/s e2e, failed %d" % (
        it, t1 - t0, t2 - t1, sum(ph), [round(x, 1) for x in ph], t3 - t2, cells / sum(ph) / 1e6, cells / (t3 - t0) / 1e9, 2 * n_pairs / (t3 - t0), bad))
# spot check: rows still de-gap to the inputs
k = n_pairs // 2
off = int(prows.offs[2 * k]); ln = int(prows.out_lens[2 * k])
assert bytes(prows.buf[off:off + ln]).replace(b"-", b"") == rows[2 * k]
print("spot check ok, aligned length", ln)
