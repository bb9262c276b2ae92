#!/usr/bin/env python
"""tools/dp_ab.py — DP-only A/B timing of the TALCO-XDrop kernel chain (twl_batch_stage / run / fetch) on level-shaped
batches of synthetic profile pairs, several seeds, with a checksum of every path so that two builds / option sets can be
compared for bit-identical results without the oracle.

    python tools/dp_ab.py [--pairs 4096] [--length 1500] [--seeds 3] [--reps 5] [--kind rna] [--opts name=value,...] [--tag T]
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--length", type=int, default=1500)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kind", default="rna")
    ap.add_argument("--opts", default="")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import twilight_b200
    from twilight_b200 import api, synth
    score = api.protein_matrix() if args.kind == "protein" else None
    ctx = twilight_b200.Context(score=score)
    for kv in filter(None, args.opts.split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    out = {"tag": args.tag, "opts": args.opts, "pairs": args.pairs, "length": args.length, "kind": args.kind, "seeds": []}
    for s in range(args.seeds):
        batch = synth.profile_pair_batch(args.pairs, args.length, seed=1000 + s, kind=args.kind)
        pairs = [twilight_b200.ProfilePairIn(**b) for b in batch]
        ctx.stage(pairs)
        ms = []
        for _ in range(args.reps + 1):
            ctx.run()
            ms.append(ctx.kernel_ms())
        res = ctx.fetch()
        h = hashlib.md5()
        for r in res:
            h.update(np.int32(r.status).tobytes()); h.update(np.int64(r.cells).tobytes()); h.update(r.path.tobytes())
        cells = sum(r.cells for r in res)
        best, med = min(ms[1:]), float(np.median(ms[1:]))
        out["seeds"].append({"seed": 1000 + s, "cells": int(cells), "failed": sum(1 for r in res if r.status), "ms_best": best, "ms_median": med,
                             "gcups_median": cells / med / 1e6, "gcups_best": cells / best / 1e6, "launches": ctx.launch_count(), "md5": h.hexdigest()})
    out["gcups_median_over_seeds"] = float(np.median([s["gcups_median"] for s in out["seeds"]]))
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
