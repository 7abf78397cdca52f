"""bench.py host logic that needs no GPU: the workload description shared by both arms, the P(k) comparison behind the
`parity` object, the roofline arithmetic."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(model="fofr", grid_bytes=8, scale_dependent=-1, sd_mode="merged", deposit_mode=0, sort_interval=4)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_config_names_the_workload_once_for_both_arms():
    a = _args()
    c = bench.config_of(a, 512)
    assert c["nmesh"] == 512 and c["npart"] == 512 ** 3 and c["scale_dependent"] == 1
    assert "512^3" in c["workload"] and "SCALEDEPENDENT" in c["workload"] and "model" not in c
    assert bench.config_of(_args(model="lcdm"), 128)["scale_dependent"] == 0
    assert bench.box_for(512) == 400.0 and bench.box_for(128) == 200.0


def test_pofk_compare_looks_below_half_nyquist_only():
    N, box = 64, 100.0
    knyq = np.pi * N / box
    k = np.linspace(0.05, 1.2 * knyq, 40)
    n = np.full(40, 10.0)
    p = 1000.0 / k
    q = p.copy()
    q[k > 0.5 * knyq] *= 1.5                       # garbage above the cut must not count
    q[3] *= 1.0 + 3e-5
    r = bench.pofk_compare((q, k, n), (p, k, n), N, box, "unit test")
    assert abs(r["max_rel_err"] - 3e-5) < 1e-9 and r["bins"] == int((k < 0.5 * knyq).sum()) and r["modes_equal"]
    assert r["k_max"] < 0.5 * knyq
    n2 = n.copy(); n2[1] = 0                       # empty bins are skipped
    assert bench.pofk_compare((q, k, n2), (p, k, n2), N, box, "x")["bins"] == r["bins"] - 1


def test_step_roofline_adds_hbm_and_nvlink_time():
    a = _args()
    rec = {"nmesh": 1024, "ms_per_step": 61.6, "phases_ms": {"PtoMesh": 4.5, "MtoParticles": 3.8}}
    r = bench.roofline_of(a, rec, 8)
    st = r["step"]
    assert st["algorithmic_bytes_per_particle"] == 120 + 51 * 8 + 52 * 8 == 944 and st["ffts_per_step"] == 12
    n_loc = 1024 ** 3 / 8
    assert abs(st["nvlink_bytes_per_gpu"] - 12 * 8 * n_loc * 7 / 8) < 1
    assert abs(st["frac"] - (st["hbm_ms"] + st["nvlink_ms"]) / 61.6) < 1e-12
    assert 0.5 < st["frac"] < 0.62                 # the round-2 target record
    assert r["bound"] == "hbm" and r["kernel"].startswith("k_deposit_atomic") and 0 < r["frac"] < 1
    one = bench.roofline_of(a, {"nmesh": 512, "ms_per_step": 37.3, "phases_ms": {"PtoMesh": 3.4, "MtoParticles": 3.2}}, 1)
    assert one["step"]["nvlink_ms"] == 0 and (one["traffic"] is None or one["traffic"] > 1e9)
