"""Several device contexts in ONE process (the reference GPU build's host-thread-per-GPU scheme, src/cuda/alignment-gpu.cu:
226-253, with device-resident rows): the drop-in CLI with TWL_DEVICES must write the same bytes as with one device. On a
one-GPU box the contexts share the GPU (TWL_DEVICES=0,0,0), which exercises the whole multi-device path of the adapter —
pair placement by row affinity, row migration between contexts (twl_rows_migrate), one host thread per context — with real
GPUs only changing where the copies go. With >= 2 GPUs visible the same scenarios also run on distinct devices."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests.cli_scenarios import DATA, ROOT, SCENARIOS
from tests import synth_scenarios as syn

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "build", "twilight_b200")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))
GOLD_SYN = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_synth_md5.json")))


def _run(cmd, cwd, devices):
    env = dict(os.environ, TWL_DEVICES=devices, TWL_STATS="1")
    res = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=1800)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stderr.splitlines() if l.startswith("[twl-stats]")]
    return json.loads(line[-1][len("[twl-stats] "):]) if line else {}


def _device_sets():
    import torch
    sets = ["0,0,0"]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append(",".join(str(i) for i in range(min(n, 8))))
    return sets


@pytest.mark.parametrize("name", ["rnasim_default", "rnasim_divide_m200", "sars_20_default"])
def test_bundled_scenarios_on_several_contexts(name, tmp_path):
    if not os.path.exists(CLI) or not os.path.isdir(DATA):
        pytest.skip("build/twilight_b200 or oracle/_ref/dataset missing")
    for devs in _device_sets():
        out = str(tmp_path / f"{name}.{devs.count(',')}.aln")
        args = [a.replace("{D}", DATA) for a in SCENARIOS[name]]
        st = _run([CLI] + args + ["-o", out, "-d", str(tmp_path / f"tmp{devs.count(',')}")], str(tmp_path), devs)
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == GOLD[name]["md5"], devs
        assert st["devices"] == devs.count(",") + 1
        assert sum(1 for p in st["pairs_per_device"] if p > 0) >= 2, st       # the work really was spread
        assert st["rows_migrated_between_devices"] > 0, st                     # and joins across contexts moved rows


def test_parking_and_caching_on_several_contexts(tmp_path):
    """rna_3k: nodes >= 1000 sequences (msaFreq caching, parking) with rows spread over contexts."""
    if not os.path.exists(CLI):
        pytest.skip("build/twilight_b200 missing")
    from twilight_b200 import synth
    prefix = synth.make_dataset("rna_3k", str(tmp_path / "data"))
    for devs in _device_sets():
        out = str(tmp_path / f"rna_3k.{devs.count(',')}.aln")
        st = _run([CLI, "-t", prefix + ".nwk", "-i", prefix + ".fa", "-o", out, "-d", str(tmp_path / f"tmp{devs.count(',')}")], str(tmp_path), devs)
        assert syn.md5_file(out) == GOLD_SYN["rna_3k_default"]["md5"], devs
        assert sum(1 for p in st["pairs_per_device"] if p > 0) >= 2


def test_rows_migrate_between_contexts():
    """twl_rows_migrate: rows leave one context and arrive byte-identical on the other (same GPU here; cudaMemcpyPeer between
    GPUs when the contexts sit on different devices); twl_rows_drop forgets rows and their buffers are reused."""
    import twilight_b200
    rng = np.random.default_rng(4)
    rows = [bytes(rng.choice(np.frombuffer(b"ACGU-", np.uint8), int(n)).tobytes()) for n in (1, 15, 16, 17, 1000, 4097, 0)]
    ids = [5, 0, 9, 2, 7, 3, 11]
    a, b = twilight_b200.Context(), twilight_b200.Context()
    a.rows_upload(ids, rows, [1.0 + k for k in range(len(ids))])
    a.rows_migrate_to(b, ids[:4])
    assert b.rows_download(ids[:4]) == rows[:4]
    assert a.rows_download(ids[4:]) == rows[4:]
    with pytest.raises(twilight_b200.TwilightError):
        a.rows_download(ids[:1])                      # gone from the source
    b.rows_migrate_to(a, ids[:4])                     # and back
    assert a.rows_download(ids) == rows
    a.rows_drop(ids[:2])
    with pytest.raises(twilight_b200.TwilightError):
        a.rows_download(ids[:1])
    a.rows_upload([20, 21], rows[:2], [1.0, 1.0])     # recycled buffers
    assert a.rows_download([20, 21] + ids[2:]) == rows[:2] + rows[2:]
    a.close(); b.close()
