"""End-to-end drop-in check: the unchanged TWILIGHT host + the B200 level kernel (build/twilight_b200) must write a
FASTA byte-identical to the reference CPU path (golden md5 from oracle/_ref/twilight_ref, tests/golden/cli_md5.json)
on every bundled scenario: default, divide-and-conquer, merge, add-sequences (with/without tree) and prune."""
import hashlib
import json
import os

import pytest

from tests.cli_scenarios import DATA, ROOT, SCENARIOS, run_cli

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "build", "twilight_b200")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_fasta_byte_identical(name, tmp_path):
    if not os.path.exists(CLI) or not os.path.isdir(DATA):
        pytest.skip("build/twilight_b200 or oracle/_ref/dataset missing (built by __graft_entry__.build() where /root/reference is mounted)")
    out, log = run_cli(CLI, name, str(tmp_path), extra=["--check"] if "default" in name else [])
    data = open(out, "rb").read()
    assert data.count(b">") == GOLD[name]["rows"]
    assert len(data) == GOLD[name]["bytes"], log[-1500:]
    assert hashlib.md5(data).hexdigest() == GOLD[name]["md5"], log[-1500:]
    assert "did not match" not in log


def test_out_of_memory_spills_and_retries(tmp_path):
    """A level that fails with TWL_E_NOMEM (injected at the 3rd and the 9th level call) makes the adapter send every dirty row
    back to the host, empty the device row store and retry the level with its rows re-sent: same bytes as an undisturbed run."""
    import subprocess
    if not os.path.exists(CLI) or not os.path.isdir(DATA):
        pytest.skip("build/twilight_b200 or oracle/_ref/dataset missing")
    from tests.cli_scenarios import SCENARIOS
    name = "rnasim_default"
    args = [a.replace("{D}", DATA) for a in SCENARIOS[name]]
    for inject in ("3", "9"):
        out = str(tmp_path / f"o{inject}.aln")
        env = dict(os.environ, TWL_OPTIONS=f"inject_nomem={inject}")
        res = subprocess.run([CLI] + args + ["-o", out, "-d", str(tmp_path / f"tmp{inject}")], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stderr[-1500:]
        assert "spilling the row store" in res.stderr
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == GOLD[name]["md5"]
