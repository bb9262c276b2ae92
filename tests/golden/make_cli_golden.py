"""Generates tests/golden/cli_md5.json: md5 of the FASTA the UNMODIFIED reference CLI (oracle/_ref/twilight_ref, built by
oracle/Makefile from /root/reference/src) writes for the bundled scenarios of README.md:219-263 / SURVEY.md §4.
Run in the build container (needs oracle/_ref):  python tests/golden/make_cli_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.cli_scenarios import SCENARIOS, run_cli  # noqa: E402


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "twilight_ref")
    out = {}
    for name in SCENARIOS:
        with tempfile.TemporaryDirectory() as tmp:
            path, log = run_cli(ref, name, tmp, threads=8)
            data = open(path, "rb").read()
            out[name] = {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data), "rows": data.count(b">")}
            print(name, out[name])
    with open(os.path.join(ROOT, "tests", "golden", "cli_md5.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
