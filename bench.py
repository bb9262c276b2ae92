#!/usr/bin/env python
"""bench.py — throughput of the per-level TALCO-XDrop alignment path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--length L] [--impl reference]

A "step" is one pass of the hot path over one synthetic guide-tree level: 4096 node pairs per GPU, every node a small
aligned family of 1-8 RNASim-shaped rows (~1.5 kb; BASELINE.json configs[1] shape, synthetic because a single
579-sequence tree cannot fill a B200). The step runs the whole per-pair pipeline of the reference's level kernel on
the device: profile build, gappy-column removal, PSGP, TALCO-XDrop DP + traceback, gappy restore, row update.
Metric = DP giga cell-updates per second (GCUPS), cells counted by the kernel with the reference's definition (sum over
anti-diagonals of the live band width).

  value  : device time of all pipeline phases with the rows resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e    : rows in host memory -> twl_rows_upload -> twl_align_level -> twl_rows_download -> host, wall clock per step
  msa    : a full progressive alignment of 2048 synthetic sequences through the same API (sequences/s)
  roofline: the DP kernel against the FP32 pipe peak (117 FP32 op per nucleotide cell-update)
  cpu_baseline: the reference's own Talco_xdrop::Align_freq (oracle/_ref/libtalco_ref.so when present, else the
           oracle port) on a bounded sample of the same batch, all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CELL_NT = 117.0
N_SM, LANES = 148, 128


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    sm_mhz, hbm, src = 1965.0, 6650.0, "fallback"
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        sm_mhz, hbm, src = float(p.get("sm_max_mhz", sm_mhz)), float(p.get("hbm_gbs", hbm)), "measured"
    fp32_tflops = N_SM * LANES * 2 * sm_mhz * 1e6 / 1e12
    return dict(fp32_tflops=fp32_tflops, hbm_gbs=hbm, source=src, sm_mhz=sm_mhz)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CELLS_CACHE = {}


def cpu_reference_level(ids, rows, weights, pairs, sample_pairs, threads):
    """The reference's own CPU implementation of the whole per-pair path (calculateProfile, getConsensus,
    removeGappyColumns, calculatePSGP, Talco_xdrop::Align_freq, addGappyColumnsBack, updateFrequency, updateAlignment —
    the body of parallelAlignmentCPU, src/alignment-cpu.cpp:46-176) on the first `sample_pairs` pairs of the level,
    `threads` host threads, one pair per task like the reference's tbb::parallel_for. Uses oracle/_ref/libtalco_ref.so
    (the unmodified reference sources); falls back to the CPU port when that library has not been built.
    Returns (GCUPS, cells, seconds, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from tests import oracle_lib as ol, ref_msa
    cfg = ol.TalcoCfg()
    use_ref = ol.have_ref()
    row_of = dict(zip(ids, rows))
    w_of = dict(zip(ids, weights))

    def state(side):
        return ref_msa.NodeState([row_of[i] for i in side.seq_ids], np.array([w_of[i] for i in side.seq_ids], np.float32),
                                 side.aln_len, side.aln_num, side.aln_weight, None)
    sample = [(state(p.ref), state(p.qry)) for p in pairs[:sample_pairs]]
    # cell counts (same definition as the kernel's counter) come from the port, outside the timed region
    key = (id(pairs), sample_pairs)
    if key not in _CELLS_CACHE:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            _CELLS_CACHE[key] = sum(ex.map(lambda p: ref_msa.align_pair("n", cfg, state(p.ref), state(p.qry)).cells, pairs[:sample_pairs]))
    cells = _CELLS_CACHE[key]

    def one(ab):
        if use_ref:
            ol.ref_pipeline("n", cfg, ab[0], ab[1])
        else:
            ref_msa.align_pair("n", cfg, ab[0], ab[1])
        return 0
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, sample))
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, cells, dt, ("reference" if use_ref else "port")


def run_msa(ctx, n_leaves, length, seed, repeats=3):
    """Full progressive MSA of a synthetic RNASim-shaped set through the device-resident level pipeline
    (twl_rows_upload -> twl_align_level per guide-tree level -> twl_rows_download): sequences/s end to end and the
    per-phase device times with the HBM roofline of the two byte-moving kernels."""
    from twilight_b200 import msa, synth
    tree = synth.random_tree(n_leaves, seed=seed, mean_blen=0.05)
    seqs = synth.evolve(tree, length, seed=seed, kind="rna")
    w = np.ones(n_leaves, np.float32)
    best = None
    for _ in range(repeats):
        rows, st = msa.progressive_align(ctx, tree, seqs, w)
        if best is None or st.wall_s < best.wall_s:
            best = st
    hbm = peaks()["hbm_gbs"]
    prof_gbs = best.profile_bytes / max(best.phase_ms[0], 1e-6) / 1e6
    upd_gbs = best.update_bytes / max(best.phase_ms[3], 1e-6) / 1e6
    return {"leaves": n_leaves, "root_len": length, "aln_len": best.aln_len, "levels": best.levels, "pairs": best.pairs,
            "cells": best.cells, "seqs_per_s_e2e": n_leaves / best.wall_s, "wall_s": best.wall_s,
            "seqs_per_s_device": n_leaves / (best.device_ms * 1e-3), "device_ms": best.device_ms,
            "phase_ms": {"profile_build": best.phase_ms[0], "gappy_psgp_pack": best.phase_ms[1], "dp_chain": best.phase_ms[2],
                         "row_update_freq_merge": best.phase_ms[3]},
            "gcups_dp_phase": best.cells / max(best.phase_ms[2], 1e-6) / 1e6, "launches": best.launches,
            "hbm_kernels": {"profile_build": {"bytes": best.profile_bytes, "GB/s": prof_gbs, "frac_of_measured_copy": prof_gbs / hbm},
                            "row_update": {"bytes": best.update_bytes, "GB/s": upd_gbs, "frac_of_measured_copy": upd_gbs / hbm}},
            "note": "upper guide-tree levels hold 1-8 pairs and are latency bound (inherent to progressive alignment)"}


def run_msa_sharded(ctx, dist, world, n_leaves, length, seed, repeats=3):
    """N > 1: the same kind of job with world x as many leaves, sharded by subtree over the ranks (no data-path collective;
    finished child nodes move to the parent's rank where a join crosses ranks). Every rank calls this; wall clock is the max
    over ranks between two barriers."""
    import torch
    from twilight_b200 import msa, synth
    tree = synth.random_tree(n_leaves, seed=seed, mean_blen=0.05)
    seqs = synth.evolve(tree, length, seed=seed, kind="rna")
    w = np.ones(n_leaves, np.float32)
    best, keep = None, None
    for _ in range(repeats):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows, st, root_owner = msa.progressive_align_sharded(ctx, tree, seqs, w, dist)
        dist.barrier(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, keep = dt, st
    v = torch.tensor([best, keep.device_ms, float(keep.cells), float(keep.pairs)], dtype=torch.float64, device="cuda")
    mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = v.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    wall = float(mx[0])
    return {"leaves": n_leaves, "root_len": length, "ranks": world, "pairs": int(sm[3]), "cells": int(sm[2]), "wall_s": wall,
            "seqs_per_s_e2e": n_leaves / wall, "device_ms_max_rank": float(mx[1]),
            "note": "sharded by subtree (twilight_b200/shard.py); rows of a finished child move between ranks only at the top joins"}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same workload generator,
    metric and unit as the B200 arm; each step is a bounded sample of the level (4 pairs per host thread)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = max(threads, min(args.pairs, 16 * threads))
    ids, rows, weights, pairs = build_level_batch(n_sample, args.length, seed=1000)
    vals, ms = [], []
    kind = "port"
    for s in range(args.warmup + args.steps):
        g, cells, dt, kind = cpu_reference_level(ids, rows, weights, pairs, n_sample, threads)
        if s >= args.warmup:
            vals.append(g)
            ms.append(dt * 1e3)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "dp_gcups", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"one guide-tree level, RNASim-shaped node pairs (~{args.length} columns, 1-8 member sequences per node): "
                                   "the reference's whole per-pair path (profile build ... row update) on the CPU",
                       "pairs_per_step": n_sample, "l2": "inputs are host-resident (CPU run)"},
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": kind,
                             "sample": f"{n_sample} pairs of the B200 arm's level generator per step"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def build_level_batch(n_pairs, length, seed):
    """Row-level batch: ids, rows, weights and the LevelPairIn list of one synthetic guide-tree level."""
    from twilight_b200 import LevelPairIn, NodeSideIn, synth
    fam = synth.level_rows_batch(n_pairs, length, seed=seed, kind="rna")
    ids, rows, pairs = [], [], []
    for ref_rows, qry_rows in fam:
        sides = []
        for fr in (ref_rows, qry_rows):
            mine = list(range(len(ids), len(ids) + len(fr)))
            ids += mine
            rows += fr
            sides.append(NodeSideIn(mine, len(fr[0]), len(fr), float(len(fr))))
        pairs.append(LevelPairIn(sides[0], sides[1]))
    weights = [1.0] * len(ids)
    return ids, rows, weights, pairs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=4096, help="node pairs per GPU per step")
    ap.add_argument("--length", type=int, default=1500)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--msa-leaves", type=int, default=2048, help="leaves of the synthetic MSA job (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import twilight_b200
    from twilight_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)

    # weak scaling: every rank aligns its own shard of same-level node pairs (no data-path collective)
    ids, rows, weights, pairs = build_level_batch(args.pairs, args.length, seed=1000 + rank)
    row_bytes = sum(len(r) for r in rows)
    ctx = twilight_b200.Context(device=local)
    # L2 hygiene: a flush buffer larger than L2 (126 MB) is written between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # the ctypes argument blocks (host buffers: rows in, rewritten rows + paths out) are built once; a step is then
    # exactly three C-ABI calls
    caps = {}
    for p in pairs:
        for sd in (p.ref, p.qry):
            for i in sd.seq_ids:
                caps[i] = p.ref.aln_len + p.qry.aln_len + 16
    prows = ctx.prepare_rows(ids, rows, weights, [caps[i] for i in ids])
    plevel = ctx.prepare_level(pairs)

    def one_step():
        """rows -> HBM, one level through the device pipeline, rewritten rows -> host. Returns (phase_ms, wall_ms)."""
        t0 = time.perf_counter()
        ctx.upload_prepared(prows)
        ctx.align_level_prepared(plevel)
        ctx.download_prepared(prows)
        t1 = time.perf_counter()
        return ctx.level_phase_ms(), (t1 - t0) * 1e3

    for _ in range(args.warmup):
        one_step()
    cells = sum(int(plevel.res[k].cells) for k in range(plevel.n))
    bad = sum(1 for k in range(plevel.n) if plevel.res[k].status != 0)

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms, phases, launches, e2e_ms = [], [0.0] * 4, 0, []
    d2h = 0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ph, wall = one_step()
        dev_ms.append(sum(ph))
        phases = [a + b for a, b in zip(phases, ph)]
        launches += ctx.launch_count()
        e2e_ms.append(wall)
        d2h = int(sum(prows.out_lens)) + sum(int(plevel.res[k].path_len) + 40 for k in range(plevel.n))
    barrier()
    clocks = sampler.stop()
    dev_total, e2e_total = float(np.sum(dev_ms)), float(np.sum(e2e_ms))

    tot = torch.tensor([dev_total, e2e_total, float(cells)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_all, e2e_all, cells_all = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        dev_all, e2e_all, cells_all = dev_total, e2e_total, float(cells)

    msa_sharded = None
    if dist is not None and args.msa_leaves > 0:
        msa_sharded = run_msa_sharded(ctx, dist, world, args.msa_leaves * world, args.length, seed=77)

    if rank == 0:
        pk = peaks()
        gcups = cells_all * args.steps / (dev_all * 1e-3) / 1e9
        e2e_gcups = cells_all * args.steps / (e2e_all * 1e-3) / 1e9
        dp_gcups_gpu = cells * args.steps / (phases[2] * 1e-3) / 1e9            # dominant kernel, this rank
        achieved_tflops = dp_gcups_gpu * 1e9 * FLOP_PER_CELL_NT / 1e12
        n_seqs = len(ids) * world
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):   # DRAM bytes per cell of the dominant kernel from the committed ncu --set full capture
            traffic = json.load(open(tpath))["dram_bytes_per_cell"] * cells
        line = {"metric": "dp_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"one guide-tree level of {args.pairs} node pairs per GPU, RNASim-shaped (~{args.length} columns, 1-8 member "
                                       "sequences per node): profile build + gappy-column removal + PSGP + TALCO-XDrop DP/traceback + "
                                       "gappy restore + row update", "pairs_per_gpu": args.pairs, "sequences_per_gpu": len(ids),
                           "cells_per_step": cells_all, "failed_pairs": bad, "l2": "256 MiB flush buffer written between timed steps",
                           "timed_region": "value: device time of the four pipeline phases with the rows resident in HBM; e2e: rows from host "
                                           "memory -> twl_rows_upload -> twl_align_level -> twl_rows_download -> host"},
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(row_bytes), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_all / args.steps, "seqs_per_s": n_seqs * args.steps / (e2e_all * 1e-3)},
                "gpu_launches": launches,
                "phase_ms_per_step": {"profile_build": phases[0] / args.steps, "gappy_psgp_pack": phases[1] / args.steps,
                                      "dp_chain": phases[2] / args.steps, "row_update_freq_merge": phases[3] / args.steps},
                "roofline": {"bound": "fp32-pipe", "achieved": achieved_tflops, "peak": pk["fp32_tflops"], "unit": "TFLOP/s",
                             "frac": achieved_tflops / pk["fp32_tflops"], "traffic": traffic, "kernel": "talcoWavefrontKernel<128,1>",
                             "note": f"dominant kernel (DP) is CUDA-core bound (SURVEY.md §8d): 117 FP32 op per cell x {dp_gcups_gpu:.1f} GCUPS "
                                     f"in the DP phase; peak = 148 SM x 128 lanes x 2 x {pk['sm_mhz']:.0f} MHz ({pk['source']} sm_max_mhz); per GPU"},
                "seqs_per_s": n_seqs * args.steps / (dev_all * 1e-3),
                "clocks": clocks}
        # the HBM-bound kernels of the step against the measured copy bandwidth (algorithmic bytes)
        P = 6
        prof_bytes = row_bytes + sum((p.ref.aln_len + p.qry.aln_len) * P * 4 for p in pairs)
        new_len = {k: int(plevel.res[k].path_len) for k in range(plevel.n)}
        upd_bytes = sum(p.ref.aln_num * (p.ref.aln_len + new_len[k]) + p.qry.aln_num * (p.qry.aln_len + new_len[k]) + new_len[k] for k, p in enumerate(pairs))
        pack_bytes = sum((p.ref.aln_len + p.qry.aln_len) * (P * 4 + (P + 2) * 4) for p in pairs)
        line["hbm_kernels"] = {
            name: {"bytes_per_step": int(b), "GB/s": b / (ms / args.steps) / 1e6, "frac_of_measured_copy": b / (ms / args.steps) / 1e6 / pk["hbm_gbs"]}
            for name, b, ms in (("profile_build", prof_bytes, phases[0]), ("gappy_psgp_pack", pack_bytes, phases[1]), ("row_update", upd_bytes, phases[3]))}
        if args.msa_leaves > 0:
            line["msa"] = run_msa(ctx, args.msa_leaves, args.length, seed=77)
        if msa_sharded is not None:
            line["msa_sharded"] = msa_sharded
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_sample = max(threads, min(len(pairs), 16 * threads))
            g, c, dt, kind = cpu_reference_level(ids, rows, weights, pairs, n_sample, threads)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": kind,
                                    "sample": f"the first {n_sample} pairs of the step's level ({c} cells, {dt:.1f} s): the reference's whole per-pair "
                                              "path from oracle/_ref/libtalco_ref.so (unmodified alignment-helper.cpp + TALCO-XDrop.cpp)"}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
