"""bench.py --impl reference on the CPU (no GPU needed): the JSON line of the reference arm -- the unmodified reference on
one rank and, when the host has the cores, on several ranks of the multi-process MPI stand-in."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmgpicola_ref_lcdm.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "lcdm", "--ref-nmesh", "32",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                    # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-updates/sec per COLA PM step"
    assert d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if cb["cores"] > 1:                                       # several ranks: the one-rank figure rides along
        assert cb["one_core"]["value"] > 0


def test_usable_cpus_is_sane():
    sys.path.insert(0, ROOT)
    import bench
    n = bench.usable_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)
