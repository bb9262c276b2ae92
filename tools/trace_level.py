import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, twilight_b200
ids, rows, weights, pairs = bench.build_level_batch(4096, 1500, seed=1000)
ctx = twilight_b200.Context()
caps = {}
for p in pairs:
    for sd in (p.ref, p.qry):
        for i in sd.seq_ids: caps[i] = p.ref.aln_len + p.qry.aln_len + 16
prows = ctx.prepare_rows(ids, rows, weights, [caps[i] for i in ids]); plevel = ctx.prepare_level(pairs)
for it in range(3):
    os.environ.pop("TWL_TRACE", None)
    if it == 2: os.environ["TWL_TRACE"] = "1"
    t0 = time.perf_counter(); ctx.upload_prepared(prows); t1 = time.perf_counter(); ctx.align_level_prepared(plevel); t2 = time.perf_counter(); ctx.download_prepared(prows); t3 = time.perf_counter()
print("upload %.2f ms | align_level %.2f ms | download %.2f ms | device phases %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, ctx.level_phase_ms()))
