"""Synthetic end-to-end scenarios for the drop-in CLI: the named shapes of BASELINE.json (C3 RNA ladder rungs, C4 30 kb
genomes, C5 protein) at sizes the CPU reference finishes in minutes, each in default and divide-and-conquer (-m) mode, plus
a set with low-quality / unrelated / fragment sequences (deferred pairs, re-alignment, --filter). Data sets are
regenerated from seeds by twilight_b200.synth.make_dataset; tests/golden/cli_synth_md5.json holds the md5 of the input
files and of the FASTA the UNMODIFIED reference CLI wrote for them (tests/golden/make_cli_synth_golden.py)."""
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name: (data set, extra CLI arguments)
SCENARIOS = {
    "rna_1k_default": ("rna_1k", []),
    "rna_1k_m300": ("rna_1k", ["-m", "300"]),
    "rna_3k_default": ("rna_3k", []),                 # nodes >= 1000 sequences: msaFreq caching + parking (helper.cpp:14,35-40,479-500)
    "rna_3k_m1000": ("rna_3k", ["-m", "1000"]),
    "rna_10k_default": ("rna_10k", []),
    "rna_100k_default": ("rna_100k", []),             # bench.py's C3 rung; not part of the pytest parametrisation (see test_cli_synth_gpu.py)
    "sars_64_default": ("sars_64", []),
    "sars_64_m20": ("sars_64", ["-m", "20"]),
    "prot_2k_default": ("prot_2k", ["--type", "p"]),
    "prot_2k_m500": ("prot_2k", ["--type", "p", "-m", "500"]),
    "rna_outliers_default": ("rna_outliers", []),     # low-quality sequences are deferred and re-aligned (task 1)
    "rna_outliers_filter": ("rna_outliers", ["--filter"]),   # ... or excluded: their nodes have length 0
}


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run_cli(binary, name, data_dir, out_dir, threads=None, timeout=3600):
    ds, extra = SCENARIOS[name]
    out = os.path.join(out_dir, name + ".aln")
    cmd = [binary, "-t", os.path.join(data_dir, ds + ".nwk"), "-i", os.path.join(data_dir, ds + ".fa"), "-o", out,
           "-d", os.path.join(out_dir, name + "_tmp")] + list(extra)
    if threads:
        cmd += ["-C", str(threads)]
    res = subprocess.run(cmd, cwd=out_dir, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0 or not os.path.exists(out):
        raise RuntimeError(f"{' '.join(cmd)} failed ({res.returncode}):\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    return out, res.stdout + res.stderr
