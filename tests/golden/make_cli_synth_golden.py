"""Generates tests/golden/cli_synth_md5.json: for every scenario of tests/synth_scenarios.py the md5 of the seeded
synthetic input (FASTA + Newick) and of the alignment the UNMODIFIED reference CLI (oracle/_ref/twilight_ref, CPU path)
writes for it. Run where /root/reference is mounted (after `make -C oracle ref`):

    python tests/golden/make_cli_synth_golden.py [scenario ...]
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.synth_scenarios import SCENARIOS, md5_file, run_cli   # noqa: E402
from twilight_b200 import synth   # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "twilight_ref")
OUT = os.path.join(ROOT, "tests", "golden", "cli_synth_md5.json")


def main():
    names = sys.argv[1:] or list(SCENARIOS)
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    with tempfile.TemporaryDirectory() as tmp:
        made = {}
        for name in names:
            ds = SCENARIOS[name][0]
            if ds not in made:
                made[ds] = synth.make_dataset(ds, os.path.join(tmp, "data"))
            t0 = time.time()
            out, log = run_cli(REF, name, os.path.join(tmp, "data"), tmp)
            data = open(out, "rb").read()
            gold[name] = {"input_fa_md5": md5_file(made[ds] + ".fa"), "input_nwk_md5": md5_file(made[ds] + ".nwk"),
                          "md5": md5_file(out), "bytes": len(data), "rows": data.count(b">"),
                          "deferred": "Realign profiles that have been deferred" in log,
                          "ref_seconds_8_threads": round(time.time() - t0, 1)}
            print(name, gold[name], flush=True)
            os.remove(out)
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
