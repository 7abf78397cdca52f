"""bench.py --impl reference on the CPU (no GPU needed): the JSON line of the reference arm -- the unmodified reference as
the multi-rank program it is (multi-process MPI stand-in), timed per iteration from its own output."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + extra,
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                    # exactly one JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MG_PICOLA_lcdm_mp")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import bench
    d = _run(["--model", "lcdm", "--nmesh", "32", "--steps", "2", "--warmup", "1"])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC
    assert d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the workload is named exactly as the GPU arm names it, and the sample ran that very mesh
    args = type("A", (), dict(model="lcdm", grid_bytes=8, scale_dependent=-1))()
    assert d["config"] == bench.config_of(args, 32)
    assert d["same_config"] is True and d["sample_nmesh"] == 32
    # steps / warmup are what was timed, not what was asked for
    assert d["steps"] == 2 and d["warmup"] == 1 and d["steps_requested"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert abs(cb["ms_per_step"] * 1e-3 * cb["value"] - 32 ** 3) < 1e-3 * 32 ** 3


def test_reference_arm_samples_a_bounded_mesh_and_says_so():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MG_PICOLA_lcdm_mp")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    d = _run(["--model", "lcdm", "--nmesh", "64", "--ref-nmesh", "32", "--steps", "1", "--warmup", "1"])
    assert d["config"]["nmesh"] == 64 and d["sample_nmesh"] == 32 and d["same_config"] is False
    assert "32^3" in d["cpu_baseline"]["sample"]


def test_rank_count_of_the_reference_keeps_slabs_thick():
    """16 ranks at 128^3 made the reference abort on the 32-CPU box of round 1 (more than Buffer - 1 of a thin slab's
    particles leave it in one long step, auxPM.c:178-199): at least 16 planes per rank, at most 16 ranks."""
    import bench
    for n in (32, 64, 128, 256, 512):
        k = bench.ref_ranks(n)
        assert k >= 1 and (k & (k - 1)) == 0 and k <= 16 and (k == 1 or n // k >= 16)
    assert 1 <= bench.usable_cpus() <= (os.cpu_count() or 1)
