"""Drift_Lightcone (lightcone.c:265-474).

Chain of evidence:
  CPU  the numpy restatement (oracle/pm_oracle.py::drift_lightcone) reproduces the UNMODIFIED reference
       (oracle/_ref/libmgpicola_ref_lcdm_lc.so, -DLIGHTCONE -DUNFORMATTED) bit for bit: every row of every replicate
       file and every particle position;
  CPU  the per-particle function the CUDA kernels consist of (csrc/lightcone.cuh), run particle by particle on the host
       (tests/host/lightcone_emul.cu), gives the same rows and positions;
  GPU  mgp_lightcone_count / mgp_drift_lightcone against the reference's files (rows as sets: the device appends them in
       no particular order), positions bit-exact; and the reference's own driver bound to the library against the
       unmodified reference on a whole lightcone run (tests/test_dropin_driver.py::test_lightcone_driver_...)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import lightcone_case as lcc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    c = lcc.reference_case(str(tmp_path_factory.mktemp("lc")))
    if c is None:
        pytest.skip("oracle/_ref lightcone build missing (make -C oracle; needs /root/reference)")
    assert len(c["reps"]) > 8 and sum(r.shape[0] for r in c["ref_rows"]) > 1000     # the case exercises the loop
    return c


def _oracle(c):
    from oracle import pm_oracle as po
    i, s = c["inputs"], c["scalars"]
    return po.drift_lightcone(i["pos"], i["vel"], i["D"], i["D2"], i["sumxyz"], i["box"], i["use_cola"], s["A"], s["AFF"],
                              s["dyyy"], s["da1"], s["da2"], s["dv1"], s["dv2"], s["rcomov_old"], s["rcomov_new"], s["origin"],
                              c["reps"], s["al_tab"], s["da1_tab"], s["da2_tab"], s["dyyy_tab"], s["lengthfac"],
                              s["velfac_times_fac"], s["boundary"])


def test_oracle_matches_reference_bit_for_bit(case):
    newpos, rows, exceeded = _oracle(case)
    assert not exceeded
    assert np.array_equal(newpos.view(np.uint32), case["ref_pos"].view(np.uint32))
    for got, ref in zip(rows, case["ref_rows"]):
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))            # same particle order: row by row


def test_oracle_flags_displacement_beyond_boundary(case):
    from oracle import pm_oracle as po
    i, s = case["inputs"], case["scalars"]
    vel = i["vel"].copy()
    vel[5, 1] = 1e6
    _, _, exceeded = po.drift_lightcone(i["pos"], vel, i["D"], i["D2"], i["sumxyz"], i["box"], i["use_cola"], s["A"], s["AFF"],
                                        s["dyyy"], s["da1"], s["da2"], s["dv1"], s["dv2"], s["rcomov_old"], s["rcomov_new"],
                                        s["origin"], case["reps"][:1], s["al_tab"], s["da1_tab"], s["da2_tab"], s["dyyy_tab"],
                                        s["lengthfac"], s["velfac_times_fac"], s["boundary"])
    assert exceeded


def _scal(c):
    i, s = c["inputs"], c["scalars"]
    return np.array([s["A"], s["AFF"], s["dyyy"], s["da1"], s["da2"], s["dv1"], s["dv2"], *i["sumxyz"], s["rcomov_old"],
                     s["rcomov_new"], *s["origin"], i["box"], s["boundary"], s["lengthfac"], s["velfac_times_fac"],
                     float(i["use_cola"])], np.float64)


def test_kernel_functions_on_the_host_match_reference(case, tmp_path):
    """csrc/lightcone.cuh run on the CPU: counting pass, offsets, drift pass."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = str(tmp_path / "liblc_emul.so")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-gencode",
                    "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "mg-picola-public_b200", "csrc"), "-o", so,
                    os.path.join(ROOT, "tests", "host", "lightcone_emul.cu")], check=True)
    L = C.CDLL(so)
    i, s = case["inputs"], case["scalars"]
    n = i["pos"].shape[0]
    reps = np.ascontiguousarray(case["reps"], np.int32)
    nrep = reps.shape[0]
    total = sum(r.shape[0] for r in case["ref_rows"])
    cnt = np.zeros(nrep, np.uint64)
    rows = np.zeros((total, 6), np.float32)
    newpos = np.zeros((n, 3), np.float32)
    arrs = [np.ascontiguousarray(i[k], np.float32) for k in ("pos", "vel", "D", "D2")]
    tabs = [np.ascontiguousarray(s[k], np.float64) for k in ("al_tab", "da1_tab", "da2_tab", "dyyy_tab")]
    scal = _scal(case)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.lc_emul.argtypes = [C.c_long] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4
    rc = L.lc_emul(n, *[p(a) for a in arrs], p(scal), tabs[0].size, *[p(t) for t in tabs], nrep, p(reps), p(cnt), p(rows), p(newpos))
    assert rc == 0
    assert np.array_equal(cnt, np.array([r.shape[0] for r in case["ref_rows"]], np.uint64))
    assert np.array_equal(newpos.view(np.uint32), case["ref_pos"].view(np.uint32))
    o = 0
    for ref in case["ref_rows"]:
        got = rows[o:o + ref.shape[0]]
        o += ref.shape[0]
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.gpu
def test_cuda_drift_lightcone_matches_reference(mgp, require_gpu, case):
    i, s = case["inputs"], case["scalars"]
    N = i["nmesh"]
    pm = mgp.PM(N, N, i["box"], omega=case["omega"], grid_bytes=8, use_cola=i["use_cola"], sort_particles=0)
    pm.upload_particles(i["pos"], i["vel"], i["D"], i["D2"], i["ids"])
    want = np.array([r.shape[0] for r in case["ref_rows"]], np.uint64)
    cnt = pm.lightcone_count(s, case["reps"], i["sumxyz"])
    assert np.array_equal(cnt, want)
    before = pm.download_particles(("pos", "id"))
    assert np.array_equal(before["pos"], i["pos"])                                 # the count changes nothing
    # a block that is too small: refused with MGP_ERR_BUFFER, nothing moved
    with pytest.raises(mgp.MgpError) as e:
        pm.Drift_Lightcone(s, case["reps"], i["sumxyz"], cap=int(want.max()) - 1)
    assert e.value.code == -3
    assert np.array_equal(pm.download_particles(("pos",))["pos"], i["pos"])
    rows = pm.Drift_Lightcone(s, case["reps"], i["sumxyz"])
    for got, ref in zip(rows, case["ref_rows"]):
        assert got.shape == ref.shape
        assert np.array_equal(lcc.sort_rows(got).view(np.uint32), lcc.sort_rows(ref).view(np.uint32))
    after = pm.download_particles(("pos", "id"))
    assert np.array_equal(after["id"], i["ids"])
    assert np.array_equal(after["pos"].view(np.uint32), case["ref_pos"].view(np.uint32))
    pm.close()


@pytest.mark.gpu
def test_cuda_lightcone_edge_cases(mgp, require_gpu, case):
    i, s = case["inputs"], case["scalars"]
    N = i["nmesh"]
    # no replicate listed: a plain drift with Drift_Lightcone's arithmetic (one rounding instead of Drift's two)
    pm = mgp.PM(N, N, i["box"], omega=case["omega"], grid_bytes=8, use_cola=i["use_cola"], sort_particles=0)
    pm.upload_particles(i["pos"], i["vel"], i["D"], i["D2"], i["ids"])
    assert pm.Drift_Lightcone(s, np.zeros((0, 3), np.int32), i["sumxyz"]) == []
    assert np.array_equal(pm.download_particles(("pos",))["pos"].view(np.uint32), case["ref_pos"].view(np.uint32))
    # a displacement beyond the boundary is the reference's FatalError (lightcone.c:403-407)
    vel = i["vel"].copy()
    vel[11, 2] = 1e6
    pm.upload_particles(i["pos"], vel, i["D"], i["D2"], i["ids"])
    with pytest.raises(mgp.MgpError) as e:
        pm.Drift_Lightcone(s, case["reps"], i["sumxyz"])
    assert e.value.code == -1 and "boundary" in str(e.value)
    pm.close()
    # not with scale-dependent growth (the reference refuses that build)
    pm = mgp.PM(N, N, i["box"], omega=case["omega"], grid_bytes=8, scale_dependent=1, sort_particles=0)
    with pytest.raises(mgp.MgpError) as e:
        pm.lightcone_count(s, case["reps"], i["sumxyz"])
    assert e.value.code == -4
    pm.close()
