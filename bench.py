#!/usr/bin/env python
"""bench.py — throughput of the per-level TALCO-XDrop alignment path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--length L] [--seeds S] [--configs all|none] [--impl reference]

Main line. A "step" is one pass of the hot path over one synthetic guide-tree level: 4096 node pairs per GPU, every node a
small aligned family of 1-8 RNASim-shaped rows (~1.5 kb; BASELINE.json configs[1]/[2] shape, synthetic because the bundled
579-sequence tree cannot fill a B200). The step runs the whole per-pair pipeline of the reference's level kernel on the
device: profile build, gappy-column removal, PSGP, TALCO-XDrop DP + traceback, gappy restore, row update. Every rank holds
S (default 3) levels drawn with different seeds and cycles through them, so the number is not one seed's luck; per-rank,
per-seed times are in `per_rank`.
Metric = DP giga cell-updates per second (GCUPS), cells counted by the kernel with the reference's definition (sum over
anti-diagonals of the live band width).

  value  : device time of all pipeline phases with the rows resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e    : rows in host memory -> twl_rows_upload -> twl_align_level -> twl_rows_download -> host, wall clock per step
  roofline: the DP kernel against the FP32 pipe peak (117 FP32 op per nucleotide cell-update)
  cpu_baseline: the reference's own level entry point (cpu::alignmentKernel_CPU -> parallelAlignmentCPU from
           oracle/_ref/libtalco_ref.so, unmodified sources; the oracle port when that library is absent) on a bounded
           sample of the same level, all host cores.
  configs: (N = 1) the configurations BASELINE.json names, each with its own roofline / cpu_baseline:
           C1 sars_20 and C2 RNASim through the drop-in CLI (byte-identity checked), C3 a 10^4-leaf RNA rung through the
           CLI, C4 a level of 30 kb pairs and C5 a level of 400-aa protein pairs through the level API.
  msa    : a full progressive alignment of 2048 synthetic sequences through the Python mirror of the level API
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CELL = {"n": 117.0, "p": 1461.0}
N_SM, LANES = 148, 128
CLI = os.path.join(ROOT, "build", "twilight_b200")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "twilight_ref")
REF_DATA = os.path.join(ROOT, "oracle", "_ref", "dataset")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    sm_mhz, hbm, src = 1965.0, 6650.0, "fallback"
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        sm_mhz, hbm, src = float(p.get("sm_max_mhz", sm_mhz)), float(p.get("hbm_gbs", hbm)), "measured"
    fp32_tflops = N_SM * LANES * 2 * sm_mhz * 1e6 / 1e12
    return dict(fp32_tflops=fp32_tflops, hbm_gbs=hbm, source=src, sm_mhz=sm_mhz)


def dp_roofline(gcups, kind="n", kernel="talcoWavefrontKernel<128,1,4>", traffic=None, traffic_source=None):
    pk = peaks()
    achieved = gcups * 1e9 * FLOP_PER_CELL[kind] / 1e12
    out = {"bound": "fp32-pipe", "achieved": achieved, "peak": pk["fp32_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["fp32_tflops"],
           "traffic": traffic, "kernel": kernel,
           "note": f"DP kernel is CUDA-core bound (SURVEY.md §8d): {FLOP_PER_CELL[kind]:.0f} FP32 op per cell x {gcups:.1f} GCUPS in the DP phase; "
                   f"peak = 148 SM x 128 lanes x 2 x {pk['sm_mhz']:.0f} MHz ({pk['source']} sm_max_mhz); per GPU"}
    if traffic_source:
        out["traffic_source"] = traffic_source
    return out


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


# ----------------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------------
def level_config(args):
    """The workload both arms (B200 and --impl reference) are measured on; a function of the arguments only."""
    return {"workload": f"one guide-tree level of {args.pairs} node pairs per GPU, RNASim-shaped (~{args.length} columns, 1-8 member sequences per "
                        "node): profile build + gappy-column removal + PSGP + TALCO-XDrop DP/traceback + gappy restore + row update "
                        "(the body of parallelAlignmentCPU, src/alignment-cpu.cpp:46-176)",
            "pairs_per_gpu": args.pairs, "columns": args.length, "generator": "twilight_b200.synth.level_rows_batch(kind='rna')",
            "seeds": f"{args.seeds} levels per rank (seed 1000 + {args.seeds}*rank + s), steps cycle through them",
            "l2": "256 MiB flush buffer written between timed steps",
            "timed_region": "value: device time of the four pipeline phases with the rows resident in HBM; e2e: rows from host memory -> "
                            "twl_rows_upload -> twl_align_level -> twl_rows_download -> host"}


def build_level_batch(n_pairs, length, seed, kind="rna", **kw):
    """Row-level batch: ids, rows, weights and the LevelPairIn list of one synthetic guide-tree level."""
    from twilight_b200 import LevelPairIn, NodeSideIn, synth
    fam = synth.level_rows_batch(n_pairs, length, seed=seed, kind=kind, **kw)
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


def cpu_reference_level(type_, cfg, ids, rows, weights, pairs, sample_pairs, threads):
    """The reference's own level entry point (cpu::alignmentKernel_CPU -> parallelAlignmentCPU, src/alignment-cpu.cpp:32-183,
    one tbb::parallel_for over the pairs on `threads` workers) on the first `sample_pairs` pairs of the level, from
    oracle/_ref/libtalco_ref.so (the unmodified reference sources). When that library has not been built the CPU port
    (oracle/twl_oracle.cpp) is timed instead, pair by pair on a thread pool. Returns (GCUPS, cells, seconds, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from tests import oracle_lib as ol, ref_msa
    row_of = dict(zip(ids, rows))
    w_of = dict(zip(ids, weights))

    def state(side):
        return ref_msa.NodeState([row_of[i] for i in side.seq_ids], np.array([w_of[i] for i in side.seq_ids], np.float32),
                                 side.aln_len, side.aln_num, side.aln_weight, None)
    sample = [(state(p.ref), state(p.qry)) for p in pairs[:sample_pairs]]
    # cell counts (the kernel's definition) come from the port, outside the timed region
    with ThreadPoolExecutor(max_workers=threads) as ex:
        cells = sum(ex.map(lambda ab: ref_msa.align_pair(type_, cfg, state_copy(ab[0]), state_copy(ab[1])).cells, sample))
    if ol.have_ref():
        _, dt, _ = ol.ref_level(type_, cfg, sample, threads)
        kind = "reference"
    else:
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(lambda ab: ref_msa.align_pair(type_, cfg, ab[0], ab[1]) and 0, sample))
        dt = time.perf_counter() - t0
        kind = "port"
    return cells / dt / 1e9, cells, dt, kind


def state_copy(st):
    from tests import ref_msa
    return ref_msa.NodeState(st.rows, st.weights, st.aln_len, st.aln_num, st.aln_weight, None)


def cpu_block(type_, cfg, ids, rows, weights, pairs, n_sample, what):
    threads = os.cpu_count() or 1
    g, c, dt, kind = cpu_reference_level(type_, cfg, ids, rows, weights, pairs, n_sample, threads)
    return {"value": g, "unit": "GCUPS", "cores": threads, "kind": kind, "mcups_per_core": g * 1e3 / threads,
            "sample": f"the first {n_sample} pairs of {what} ({c} cells, {dt:.1f} s in the level call): the reference's level entry point "
                      "cpu::alignmentKernel_CPU (parallelAlignmentCPU, tbb::parallel_for over pairs) from oracle/_ref/libtalco_ref.so"}


# ----------------------------------------------------------------------------------------------------------------------
# one level workload on the device
# ----------------------------------------------------------------------------------------------------------------------
class LevelJob:
    def __init__(self, ctx, ids, rows, weights, pairs):
        self.ctx, self.ids, self.rows, self.weights, self.pairs = ctx, ids, rows, weights, pairs
        caps = {}
        for p in pairs:
            for sd in (p.ref, p.qry):
                for i in sd.seq_ids:
                    caps[i] = p.ref.aln_len + p.qry.aln_len + 16
        self.prows = ctx.prepare_rows(ids, rows, weights, [caps[i] for i in ids])
        self.plevel = ctx.prepare_level(pairs)
        self.row_bytes = sum(len(r) for r in rows)

    def step(self):
        """rows -> HBM, one level through the device pipeline, rewritten rows -> host. Returns (phase_ms, wall_ms)."""
        t0 = time.perf_counter()
        self.ctx.upload_prepared(self.prows)
        self.ctx.align_level_prepared(self.plevel)
        self.ctx.download_prepared(self.prows)
        t1 = time.perf_counter()
        return self.ctx.level_phase_ms(), (t1 - t0) * 1e3

    def cells(self):
        return sum(int(self.plevel.res[k].cells) for k in range(self.plevel.n))

    def failed(self):
        return sum(1 for k in range(self.plevel.n) if self.plevel.res[k].status != 0)

    def d2h_bytes(self):
        return int(sum(self.prows.out_lens)) + sum(int(self.plevel.res[k].path_len) + 40 for k in range(self.plevel.n))


def run_level_config(name, kind, n_pairs, length, steps, warmup, flush, sample_pairs, gen_kw, note):
    """A BASELINE.json config that is one level of pairs (C4: 30 kb genomes, C5: 400-aa proteins) through the level API."""
    import torch
    import twilight_b200
    from twilight_b200 import api
    from tests import oracle_lib as ol
    type_ = "p" if kind == "protein" else "n"
    score = api.protein_matrix() if type_ == "p" else None
    ctx = twilight_b200.Context(score=score)
    ids, rows, weights, pairs = build_level_batch(n_pairs, length, seed=4242, kind=kind, **gen_kw)
    job = LevelJob(ctx, ids, rows, weights, pairs)
    for _ in range(warmup):
        job.step()
    dev, dp, e2e = [], [], []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ph, wall = job.step()
        dev.append(sum(ph)); dp.append(ph[2]); e2e.append(wall)
    cells = job.cells()
    g_dev, g_dp, g_e2e = (cells / (float(np.median(x)) * 1e-3) / 1e9 for x in (dev, dp, e2e))
    cfg = ol.TalcoCfg(score=score) if type_ == "p" else ol.TalcoCfg()
    out = {"workload": note, "pairs": n_pairs, "sequences": len(ids), "cells_per_step": cells, "failed_pairs": job.failed(),
           "gcups_device": g_dev, "gcups_dp_phase": g_dp, "gcups_e2e": g_e2e, "ms_per_step_device": float(np.median(dev)),
           "ms_per_step_e2e": float(np.median(e2e)), "seqs_per_s_device": len(ids) / (float(np.median(dev)) * 1e-3),
           "roofline": dp_roofline(g_dp, type_, "simMatrixAaKernel + talcoWavefrontKernel<128,0,4,1>" if type_ == "p" else "talcoWavefrontKernel<512,1,2> + talcoGenericKernel<6>"),
           "cpu_baseline": cpu_block(type_, cfg, ids, rows, weights, pairs, sample_pairs, f"this level ({name})")}
    if type_ == "p":
        out["roofline"]["note"] += ("; the algorithmic count charges all 21 x 21 terms of the reference's sum, the similarity kernel skips reference letters "
                                    "whose count is zero (exact zeros), so the fraction of the as-written roofline can exceed what the FP32 pipe executes")
    ctx.close()
    return out


def parse_stats(stderr_text):
    m = re.search(r"\[twl-stats\] (\{.*\})", stderr_text)
    return json.loads(m.group(1)) if m else None


def run_cli_config(name, argv, rows_expected, golden_md5, cpu_argv=None, cpu_note=None, repeats=2):
    """A BASELINE.json config that is a whole data set through the drop-in CLI (unchanged TWILIGHT host + B200 level kernel):
    wall clock read-FASTA -> write-FASTA, the device's share (TWL_STATS), byte-identity against the reference's md5."""
    if not os.path.exists(CLI):
        return {"unavailable": "build/twilight_b200 missing (built by __graft_entry__.build() where /root/reference is mounted)"}
    best = None
    with tempfile.TemporaryDirectory() as tmp:
        for r in range(repeats):
            out = os.path.join(tmp, f"{name}.{r}.aln")
            env = dict(os.environ, TWL_STATS="1")
            t0 = time.perf_counter()
            res = subprocess.run([CLI] + argv + ["-o", out, "-d", os.path.join(tmp, f"tmp{r}")], cwd=tmp, env=env, capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if res.returncode != 0:
                return {"unavailable": f"CLI failed: {res.stderr[-300:]}"}
            st = parse_stats(res.stderr) or {}
            h = hashlib.md5()
            with open(out, "rb") as fh:
                for blk in iter(lambda: fh.read(1 << 24), b""):
                    h.update(blk)
            md5 = h.hexdigest()
            os.remove(out)
            if best is None or wall < best["wall_s"]:
                best = {"wall_s": wall, "stats": st, "md5": md5}
        cpu = None
        if cpu_argv is not None and os.path.exists(REF_CLI):
            threads = os.cpu_count() or 1
            t0 = time.perf_counter()
            res = subprocess.run([REF_CLI] + cpu_argv + ["-o", os.path.join(tmp, "cpu.aln"), "-d", os.path.join(tmp, "cputmp"), "-C", str(threads)],
                                 cwd=tmp, capture_output=True, text=True)
            cpu_wall = time.perf_counter() - t0
            n_cpu = sum(blk.count(b">") for blk in iter(lambda f=open(os.path.join(tmp, "cpu.aln"), "rb"): f.read(1 << 24), b"")) if res.returncode == 0 else 0
            cpu = {"value": n_cpu / cpu_wall, "unit": "sequences/s", "cores": threads, "kind": "reference", "wall_s": cpu_wall,
                   "sample": cpu_note or "the same input through oracle/_ref/twilight_ref (unmodified reference CLI, CPU path)"}
    st = best["stats"]
    dp_ms = st.get("phase_ms", {}).get("dp_chain", 0.0)
    g_dp = st.get("cells", 0) / max(dp_ms, 1e-9) / 1e6
    return {"sequences": rows_expected, "wall_s": best["wall_s"], "seqs_per_s_e2e": rows_expected / best["wall_s"],
            "device_ms": st.get("device_ms"), "seqs_per_s_device": rows_expected / max(st.get("device_ms", 0.0) * 1e-3, 1e-9),
            "level_calls_wall_ms": st.get("level_calls_wall_ms"), "levels": st.get("levels"), "pairs": st.get("pairs"), "cells": st.get("cells"),
            "phase_ms": st.get("phase_ms"), "gcups_dp_phase": g_dp, "h2d_row_bytes": st.get("h2d_row_bytes"), "d2h_row_bytes": st.get("d2h_row_bytes"),
            "byte_identical_to_reference": (best["md5"] == golden_md5) if golden_md5 else None,
            "roofline": dp_roofline(g_dp, "n", "talcoWavefrontKernel (latency shape 512x2 where a level has fewer pairs than SMs)"),
            "cpu_baseline": cpu,
            "note": "wall clock includes process start, CUDA context creation (~1 s), FASTA read and write"}


def run_named_configs(flush):
    """BASELINE.json configs C1..C5 at sizes that fit the default run (rank 0, one GPU)."""
    from twilight_b200 import synth
    out = {}
    gold_cli = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))
    gold_syn = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_synth_md5.json")))
    if os.path.isdir(REF_DATA):
        a1 = ["-t", f"{REF_DATA}/sars_20.nwk", "-i", f"{REF_DATA}/sars_20.fa"]
        if os.path.exists(CLI):   # one untimed run: on a fresh box the first start of the program pages in the CUDA libraries (seconds)
            with tempfile.TemporaryDirectory() as warm:
                subprocess.run([CLI] + a1 + ["-o", os.path.join(warm, "w.aln"), "-d", os.path.join(warm, "t")], cwd=warm, capture_output=True)
        out["C1_sars_20_cli"] = run_cli_config("sars_20", a1, 20, gold_cli["sars_20_default"]["md5"], cpu_argv=a1)
        a2 = ["-t", f"{REF_DATA}/RNASim.nwk", "-i", f"{REF_DATA}/RNASim.fa"]
        out["C2_rnasim_cli"] = run_cli_config("rnasim", a2, 579, gold_cli["rnasim_default"]["md5"], cpu_argv=a2)
    else:
        out["C1_sars_20_cli"] = out["C2_rnasim_cli"] = {"unavailable": "oracle/_ref/dataset missing"}
    with tempfile.TemporaryDirectory() as tmp:
        big = synth.make_dataset("rna_10k", tmp)
        small = synth.make_dataset("rna_1k", tmp)
        out["C3_rna_10k_cli"] = run_cli_config(
            "rna_10k", ["-t", big + ".nwk", "-i", big + ".fa"], 10000, gold_syn["rna_10k_default"]["md5"],
            cpu_argv=["-t", small + ".nwk", "-i", small + ".fa"],
            cpu_note="the 10^3-leaf rung of the same generator (rna_1k) through oracle/_ref/twilight_ref: the 10^4-leaf run takes the CPU "
                     f"{gold_syn['rna_10k_default']['ref_seconds_8_threads']} s on 8 threads (tests/golden/cli_synth_md5.json)", repeats=1)
        out["C3_rna_10k_cli"]["workload"] = "synthetic RNA, 10^4 leaves x 1.5 kb, random tree (C3 ladder rung), default mode through the drop-in CLI"
        if "rna_100k_default" in gold_syn:    # the 10^5 rung: ~20 minutes on the CPU reference (golden md5 recorded once), seconds here
            huge = synth.make_dataset("rna_100k", tmp)
            out["C3_rna_100k_cli"] = run_cli_config("rna_100k", ["-t", huge + ".nwk", "-i", huge + ".fa"], 100000, gold_syn["rna_100k_default"]["md5"], repeats=1)
            out["C3_rna_100k_cli"]["workload"] = "synthetic RNA, 10^5 leaves x 1.5 kb, random tree (C3 ladder rung), default mode through the drop-in CLI"
            out["C3_rna_100k_cli"]["cpu_baseline"] = {
                "value": 100000 / gold_syn["rna_100k_default"]["ref_seconds_8_threads"], "unit": "sequences/s", "cores": 8, "kind": "reference",
                "sample": "NOT timed in this run: oracle/_ref/twilight_ref on the same input, 8 threads of the build container "
                          f"({gold_syn['rna_100k_default']['ref_seconds_8_threads']} s, tests/golden/cli_synth_md5.json)"}
    out["C4_level_30kb"] = run_level_config(
        "C4", "dna", 592, 29700, 3, 2, flush, 16, dict(divergence=0.004, indel_rate=0.002, members=(1, 2, 4)),
        "one guide-tree level of 592 node pairs (4 per SM) of SARS-CoV-2-length genomes (~29.7 kb, 1-4 members per node, tip identity ~99.6 %): "
        "long pairs, ~58 TALCO tiles each, bands ~850 wide (the 1024-row wavefront instantiation)")
    try:   # the reference's own CUDA kernel on the same B200 and level shape ("the existing GPU kernel to beat")
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_gpu_kernel", os.path.join(ROOT, "tools", "ref_gpu_kernel.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out["reference_gpu_kernel_same_level"] = mod.run(4096, 1500, 1000, "rna")
    except Exception as e:
        out["reference_gpu_kernel_same_level"] = {"error": repr(e)}
    out["C5_level_protein"] = run_level_config(
        "C5", "protein", 4096, 400, 3, 2, flush, 16 * (os.cpu_count() or 1), dict(divergence=0.4, indel_rate=0.02, members=(1, 2, 4, 8)),
        "one guide-tree level of 4096 node pairs of ~400-aa protein families (1-8 members per node), 5 x BLOSUM62, --type p semantics")
    return out


def run_msa(ctx, n_leaves, length, seed, repeats=3):
    """Full progressive MSA of a synthetic RNASim-shaped set through the Python mirror of the level pipeline."""
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
    """--impl reference: the reference's own CPU implementation of the path (its level entry point, all host cores) on the
    B200 arm's config, metric and unit; each step is a bounded sample of the level (16 pairs per host thread, at most the level)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import oracle_lib as ol
    threads = os.cpu_count() or 1
    n_sample = max(threads, min(args.pairs, 16 * threads))
    ids, rows, weights, pairs = build_level_batch(n_sample, args.length, seed=1000)   # the first pairs of rank 0's first level
    cfg = ol.TalcoCfg()
    vals, ms = [], []
    kind, cells = "port", 0
    for s in range(args.warmup + args.steps):
        g, cells, dt, kind = cpu_reference_level("n", cfg, ids, rows, weights, pairs, n_sample, threads)
        if s >= args.warmup:
            vals.append(g)
            ms.append(dt * 1e3)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "dp_gcups", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": level_config(args),
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": kind, "mcups_per_core": v * 1e3 / threads,
                             "sample": f"the first {n_sample} pairs of the level of seed 1000 per step ({cells} cells): cpu::alignmentKernel_CPU "
                                       "(parallelAlignmentCPU, src/alignment-cpu.cpp:32-183) from the unmodified reference sources, "
                                       "tbb::parallel_for over the pairs on all host threads"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=4096, help="node pairs per GPU per step")
    ap.add_argument("--length", type=int, default=1500)
    ap.add_argument("--seeds", type=int, default=3, help="different levels per rank the steps cycle through")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="auto", choices=["auto", "all", "none"], help="the BASELINE.json config blocks (auto: at 1 GPU)")
    ap.add_argument("--msa-leaves", type=int, default=2048, help="leaves of the synthetic MSA job (0 = skip)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3, args.seeds)

    import torch
    import twilight_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        # the process group prints an "NCCL version ..." banner on stdout when the communicator is created; stdout carries the ONE
        # JSON line, so file descriptor 1 points at stderr while the group comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    else:
        torch.cuda.set_device(local)

    # weak scaling: every rank aligns its own shards of same-level node pairs (no data-path collective)
    ctx = twilight_b200.Context(device=local)
    seeds = [1000 + args.seeds * rank + s for s in range(args.seeds)]
    jobs = [LevelJob(ctx, *build_level_batch(args.pairs, args.length, seed=sd)) for sd in seeds]
    # L2 hygiene: a flush buffer larger than L2 (126 MB) is written between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cells_of = [0] * args.seeds
    for w in range(args.warmup):
        jobs[w % args.seeds].step()
        cells_of[w % args.seeds] = jobs[w % args.seeds].cells()
    bad = sum(j.failed() for j in jobs)

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms, phases, launches, e2e_ms, cells_timed = [], [0.0] * 4, 0, [], 0
    by_seed = [[] for _ in seeds]
    d2h = h2d = 0
    for s in range(args.steps):
        job = jobs[s % args.seeds]
        flush.fill_(1)
        torch.cuda.synchronize()
        ph, wall = job.step()
        dev_ms.append(sum(ph))
        by_seed[s % args.seeds].append((sum(ph), ph[2], wall))
        phases = [a + b for a, b in zip(phases, ph)]
        launches += ctx.launch_count()
        e2e_ms.append(wall)
        cells_timed += cells_of[s % args.seeds]
        d2h += job.d2h_bytes()
        h2d += job.row_bytes
    barrier()
    clocks = sampler.stop()
    dev_total, e2e_total = float(np.sum(dev_ms)), float(np.sum(e2e_ms))

    # per rank: totals + per-seed medians (step device ms, DP phase ms, e2e wall ms, cells)
    mine = [dev_total, e2e_total, float(cells_timed), phases[2]]
    for k in range(args.seeds):
        t = by_seed[k] or [(0.0, 0.0, 0.0)]
        mine += [float(np.median([x[0] for x in t])), float(np.median([x[1] for x in t])), float(np.median([x[2] for x in t])), float(cells_of[k])]
    tot = torch.tensor(mine, dtype=torch.float64, device="cuda")
    if dist is not None:
        allr = [torch.zeros_like(tot) for _ in range(world)]
        dist.all_gather(allr, tot)
        allr = [t.cpu().numpy() for t in allr]
    else:
        allr = [tot.cpu().numpy()]
    dev_all = max(float(t[0]) for t in allr)
    e2e_all = max(float(t[1]) for t in allr)
    cells_all = sum(float(t[2]) for t in allr)

    msa_sharded = None
    if dist is not None and args.msa_leaves > 0:
        msa_sharded = run_msa_sharded(ctx, dist, world, args.msa_leaves * world, args.length, seed=77)

    if rank == 0:
        pk = peaks()
        gcups = cells_all / (dev_all * 1e-3) / 1e9
        e2e_gcups = cells_all / (e2e_all * 1e-3) / 1e9
        dp_gcups_gpu = cells_timed / (phases[2] * 1e-3) / 1e9            # dominant kernel, this rank
        n_seqs = len(jobs[0].ids) * world
        traffic, tsrc = None, None
        for tname in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):   # DRAM bytes per cell of the dominant kernel from the committed ncu --set full capture
                traffic = json.load(open(tpath))["dram_bytes_per_cell"] * cells_timed / args.steps
                tsrc = f"not measured in this run: profiles/{tname} (dram__bytes_read+write of one ncu --set full launch of this kernel, per cell) x cells per step"
                break
        per_rank = []
        for r, t in enumerate(allr):
            seeds_r = [{"seed": 1000 + args.seeds * r + k, "cells": int(t[4 + 4 * k + 3]), "step_device_ms_median": float(t[4 + 4 * k]),
                        "dp_phase_ms_median": float(t[4 + 4 * k + 1]), "e2e_ms_median": float(t[4 + 4 * k + 2]),
                        "gcups_dp_phase": float(t[4 + 4 * k + 3]) / max(float(t[4 + 4 * k + 1]), 1e-9) / 1e6} for k in range(args.seeds)]
            per_rank.append({"rank": r, "device_ms_total": float(t[0]), "e2e_ms_total": float(t[1]), "cells_total": int(t[2]),
                             "dp_phase_ms_total": float(t[3]), "gcups_device": float(t[2]) / max(float(t[0]), 1e-9) / 1e6, "levels": seeds_r})
        slowest = max(per_rank, key=lambda x: x["device_ms_total"])
        line = {"metric": "dp_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": level_config(args),
                "workload_stats": {"sequences_per_gpu": len(jobs[0].ids), "cells_timed_all_ranks": cells_all, "failed_pairs": bad,
                                   "limiting_rank": slowest["rank"], "limiting_rank_device_ms": slowest["device_ms_total"]},
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps),
                        "ms_per_step": e2e_all / args.steps, "seqs_per_s": n_seqs * args.steps / (e2e_all * 1e-3)},
                "gpu_launches": launches,
                "phase_ms_per_step": {"profile_build": phases[0] / args.steps, "gappy_psgp_pack": phases[1] / args.steps,
                                      "dp_chain": phases[2] / args.steps, "row_update_freq_merge": phases[3] / args.steps},
                "roofline": dp_roofline(dp_gcups_gpu, "n", "talcoWavefrontKernel<128,1,4>", traffic, tsrc),
                "seqs_per_s": n_seqs * args.steps / (dev_all * 1e-3),
                "per_rank": per_rank,
                "clocks": clocks}
        # the HBM-bound kernels of the step against the measured copy bandwidth (algorithmic bytes), first level of this rank
        P = 6
        j0 = jobs[0]
        n0 = max(1, len(by_seed[0]))
        ph0 = [0.0] * 4
        ph_l, _ = j0.step()
        ph0 = list(ph_l)
        split = ctx.level_update_split_ms()
        ph0[3] = split[1]                                   # the row rewrite alone; the gappy-column restore is latency bound and reported beside it
        prof_bytes = j0.row_bytes + sum((p.ref.aln_len + p.qry.aln_len) * P * 4 for p in j0.pairs)
        new_len = {k: int(j0.plevel.res[k].path_len) for k in range(j0.plevel.n)}
        upd_bytes = sum(p.ref.aln_num * (p.ref.aln_len + new_len[k]) + p.qry.aln_num * (p.qry.aln_len + new_len[k]) + new_len[k] for k, p in enumerate(j0.pairs))
        pack_bytes = sum((p.ref.aln_len + p.qry.aln_len) * (P * 4 + (P + 2) * 4) for p in j0.pairs)
        line["hbm_kernels"] = {
            name: {"bytes_per_step": int(b), "GB/s": b / max(ms, 1e-9) / 1e6, "frac_of_measured_copy": b / max(ms, 1e-9) / 1e6 / pk["hbm_gbs"]}
            for name, b, ms in (("profile_build", prof_bytes, ph0[0]), ("gappy_psgp_pack", pack_bytes, ph0[1]), ("row_update", upd_bytes, ph0[3]))}
        line["hbm_kernels"]["gappy_restore_ms"] = split[0]
        if args.msa_leaves > 0:
            line["msa"] = run_msa(ctx, args.msa_leaves, args.length, seed=77)
        if msa_sharded is not None:
            line["msa_sharded"] = msa_sharded
        if not args.no_cpu_baseline:
            from tests import oracle_lib as ol
            threads = os.cpu_count() or 1
            n_sample = max(threads, min(len(j0.pairs), 16 * threads))
            line["cpu_baseline"] = cpu_block("n", ol.TalcoCfg(), j0.ids, j0.rows, j0.weights, j0.pairs, n_sample, "the step's first level (seed 1000)")
        if args.configs == "all" or (args.configs == "auto" and world == 1):
            ctx.close()
            ctx = None
            try:
                line["configs"] = run_named_configs(flush)
            except Exception as e:   # a failing side block must not take the contract line with it
                line["configs"] = {"error": repr(e)}
        print(json.dumps(line))
    if ctx is not None:
        ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
