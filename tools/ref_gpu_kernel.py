#!/usr/bin/env python
"""tools/ref_gpu_kernel.py — the reference's own CUDA kernel (device_function::parallelProfileAlignment_Fast,
src/cuda/device-function.cu:753, unmodified, built for sm_100 into oracle/_ref/librefgpu.so by `make -C oracle refgpu`) against
the B200-native DP kernel chain on the SAME host-fed profile pairs of one guide-tree level, kernels only, inputs resident.

The two are not result-equivalent (the reference GPU build scores in int16 with tile marker 200, wavefront cap 1350 and
x-drop 600*|gapExtend| and gives up on pairs that exceed them; the B200 kernels reproduce the reference CPU path bit for bit),
so what is compared is time per level, pairs given up, and how many paths happen to coincide.

    python tools/ref_gpu_kernel.py [--pairs 4096] [--length 1500] [--seed 1000] [--kind rna]
Prints one JSON object (also used by bench.py's `reference_gpu_kernel` block)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "librefgpu.so")


def run(n_pairs=4096, length=1500, seed=1000, kind="rna", repeats=3, device=0):
    if not os.path.exists(LIB):
        return {"unavailable": "oracle/_ref/librefgpu.so missing (make -C oracle refgpu where /root/reference is mounted)"}
    import twilight_b200
    from twilight_b200 import api, synth
    P = 22 if kind == "protein" else 6
    score = api.protein_matrix() if kind == "protein" else api.nucleotide_matrix()
    batch = synth.profile_pair_batch(n_pairs, length, seed=seed, kind=kind)
    seq_len = max(max(len(b["freq_ref"]), len(b["freq_qry"])) for b in batch)
    freq = np.zeros((n_pairs, 2, seq_len, P), np.float32)
    gop = np.zeros((n_pairs, 2, seq_len), np.float32)
    gex = np.zeros((n_pairs, 2, seq_len), np.float32)
    lens = np.zeros(2 * n_pairs, np.int32)
    nums = np.zeros(2 * n_pairs, np.int32)
    for k, b in enumerate(batch):
        r, q = len(b["freq_ref"]), len(b["freq_qry"])
        freq[k, 0, :r], freq[k, 1, :q] = b["freq_ref"], b["freq_qry"]
        gop[k, 0, :r], gop[k, 1, :q] = b["gap_open_ref"], b["gap_open_qry"]
        gex[k, 0, :r], gex[k, 1, :q] = b["gap_ext_ref"], b["gap_ext_qry"]
        lens[2 * k], lens[2 * k + 1] = r, q
        nums[2 * k], nums[2 * k + 1] = int(b["ref_num"]), int(b["qry_num"])
    gap_open, gap_extend = -50.0, -5.0
    # hostParam (alignment-gpu.cu:75-81): matrix, gapOpen, gapExtend, gapBoundary, xdrop = 600 * |gapExtend| (scoring-matrix.cpp:95)
    param = np.concatenate([score.reshape(-1), [gap_open, gap_extend, gap_extend, 600.0 * -gap_extend]]).astype(np.float32)
    aln = np.zeros((n_pairs, 2 * seq_len), np.int8)
    aln_len = np.zeros(n_pairs, np.int32)
    ms = C.c_float(0)
    lib = C.CDLL(LIB)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
    lib.refgpu_level.restype = C.c_int
    lib.refgpu_level.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, i32p, i32p, f32p, i8p, i32p, C.POINTER(C.c_float), C.c_int]
    rc = lib.refgpu_level(device, P, n_pairs, seq_len, freq, gop, gex, lens, nums, param, aln, aln_len, C.byref(ms), repeats)
    if rc != 0:
        return {"unavailable": "refgpu_level failed (see stderr)"}
    ctx = twilight_b200.Context(device=device, score=None if kind != "protein" else score)
    ctx.stage([twilight_b200.ProfilePairIn(**b) for b in batch])
    mine = []
    for _ in range(repeats + 1):
        ctx.run()
        mine.append(ctx.kernel_ms())
    res = ctx.fetch()
    ctx.close()
    cells = int(sum(r.cells for r in res))
    gave_up = int((aln_len < 0).sum())
    same = 0
    for k, r in enumerate(res):
        if aln_len[k] > 0 and r.status == 0 and aln_len[k] == len(r.path) and np.array_equal(aln[k, :aln_len[k]], r.path):
            same += 1
    mine_ms = float(min(mine[1:]))
    return {"pairs": n_pairs, "columns": length, "kind": kind, "seed": seed, "cells_b200_definition": cells,
            "reference_cuda_kernel": {"kernel": "device_function::parallelProfileAlignment_Fast<<<2048,256>>> per round of 2048 pairs (src/cuda/device-function.cu:753), sm_100 build",
                                      "ms": float(ms.value), "pairs_per_s": n_pairs / (ms.value * 1e-3), "pairs_given_up": gave_up,
                                      "equivalent_gcups": cells / (ms.value * 1e-3) / 1e9},
            "b200_kernels": {"kernel": "talcoWavefrontKernel chain (twl_batch_run)", "ms": mine_ms, "pairs_per_s": n_pairs / (mine_ms * 1e-3),
                             "pairs_failed": int(sum(1 for r in res if r.status)), "gcups": cells / (mine_ms * 1e-3) / 1e9},
            "speedup_kernel_time": float(ms.value) / mine_ms, "paths_identical": same,
            "note": "not result-equivalent: the reference GPU kernel scores in int16 with marker 200 / wavefront cap 1350 / x-drop 3000 and no gappy-column "
                    "removal; 'equivalent_gcups' divides the B200 kernel's cell count (reference CPU definition) by the reference GPU kernel's time"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--length", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--kind", default="rna")
    a = ap.parse_args()
    print(json.dumps(run(a.pairs, a.length, a.seed, a.kind)))
