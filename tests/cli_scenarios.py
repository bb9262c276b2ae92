"""The reference's documented end-to-end scenarios (README.md:219-263) as command lines over the bundled datasets."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "oracle", "_ref", "dataset")

SCENARIOS = {
    "sars_20_default": ["-t", "{D}/sars_20.nwk", "-i", "{D}/sars_20.fa"],
    "rnasim_default": ["-t", "{D}/RNASim.nwk", "-i", "{D}/RNASim.fa"],
    "rnasim_divide_m200": ["-t", "{D}/RNASim.nwk", "-i", "{D}/RNASim.fa", "-m", "200"],
    "rnasim_merge_msas": ["-f", "{D}/RNASim_subalignments/"],
    "rnasim_add_with_tree": ["-a", "{D}/RNASim_backbone.aln", "-i", "{D}/RNASim_sub.fa", "-t", "{D}/RNASim.nwk"],
    "rnasim_add_without_tree": ["-a", "{D}/RNASim_backbone.aln", "-i", "{D}/RNASim_sub.fa"],
    "rnasim_sub_prune": ["-t", "{D}/RNASim.nwk", "-i", "{D}/RNASim_sub.fa", "--prune"],
}


def run_cli(binary, name, tmpdir, threads=None, extra=(), timeout=900):
    out = os.path.join(tmpdir, name + ".aln")
    args = [a.replace("{D}", DATA) for a in SCENARIOS[name]]
    cmd = [binary] + args + ["-o", out, "-d", os.path.join(tmpdir, name + "_tmp")] + list(extra)
    if threads:
        cmd += ["-C", str(threads)]
    res = subprocess.run(cmd, cwd=tmpdir, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0 or not os.path.exists(out):
        raise RuntimeError(f"{' '.join(cmd)} failed ({res.returncode}):\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    return out, res.stdout + res.stderr
