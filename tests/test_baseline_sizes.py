"""Parity at the sizes BASELINE.json names, driver against driver: the reference's own main.c bound to the CUDA library
(adapter/_build/MG_PICOLA_CUDA_<v>) and the unmodified CPU reference (oracle/_ref, on several ranks of the host where the
multi-process stand-in is built: bit-identical to one rank, tests/test_ref_multirank.py) on the same parameter file.

  config 1   LCDM, Npart = Nmesh = 128^3, Box = 200, 10 steps z = 49 -> 0: every in-step P(k) file to 1e-8, final positions
             to 3e-5 cells, IDs exact
  config 2'  f(R) + screening, SCALEDEPENDENT, merged two-order fields (what bench.py times), 128^3, 30 steps z = 9 -> 0:
             every in-step P(k) within the north-star bound 1e-4 (runtest_fofr.sh:57-63 is the reference's own such run)
  FOFRNU     massive neutrinos through the driver (auxPM.c:383-427) on the reference's bundled CAMB tables
             (tests/golden/camb_nu0.2.tgz = camb_data/example_data_nu0.2 of the reference, transfer files only)
"""
import os
import subprocess
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from test_dropin_driver import read_gadget, read_pofk  # noqa: E402

pytestmark = pytest.mark.gpu


def _gpu_exe(variant):
    p = os.path.join(ROOT, "adapter", "_build", "MG_PICOLA_CUDA_%s" % variant)
    if not os.path.exists(p):
        pytest.skip("%s not built (needs /root/reference at build time)" % p)
    return p


def _run_cpu(variant, pf, wd, nmesh):
    """The unmodified reference: on ranks when the multi-rank build exists (same numbers, a fraction of the time)."""
    import bench
    from oracle import mprun
    K = bench.ref_ranks(nmesh) if mprun.available(variant) else 1
    if K > 1:
        rc, so, errs = mprun.run([mprun.exe_path(variant), pf], K, scratch_mb=mprun.scratch_mb_for(nmesh), timeout=1500, cwd=wd)
        assert rc == 0, (rc, so[-1500:], errs)
    else:
        exe = os.path.join(ROOT, "oracle", "_ref", "MG_PICOLA_%s" % variant)
        if not os.path.exists(exe):
            pytest.skip("%s not built (needs /root/reference at build time)" % exe)
        r = subprocess.run([exe, pf], capture_output=True, text=True, cwd=wd, timeout=3000)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return K


def _run_gpu(variant, pf, wd, env=None):
    r = subprocess.run([_gpu_exe(variant), pf], capture_output=True, text=True, cwd=wd, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def _snapshot(outdir, base):
    files = sorted((f for f in os.listdir(outdir) if f.startswith(base + ".")), key=lambda s: int(s.rsplit(".", 1)[1]))
    assert files
    parts = [read_gadget(os.path.join(outdir, f)) for f in files]
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


def _pofk_files(outdir, suffix="_CDM.txt"):
    return sorted(f for f in os.listdir(outdir) if f.startswith("pofk_") and f.endswith(suffix) and "RSD" not in f)


def test_config1_lcdm_128_ten_steps_from_z49(require_gpu, tmp_path):
    import bench
    N, box, nsteps = 128, 200.0, 10
    out = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, "lcdm", nsteps, z_init=49.0)
        if kind == "cpu":
            _run_cpu("lcdm", pf, wd, N)
        else:
            _run_gpu("lcdm", pf, wd)
        out[kind] = os.path.join(wd, "output")
    fc, fg = _pofk_files(out["cpu"]), _pofk_files(out["gpu"])
    assert fc == fg and len(fc) >= nsteps
    shot = (box / N) ** 3
    for f in fc:
        a, b = read_pofk(os.path.join(out["cpu"], f)), read_pofk(os.path.join(out["gpu"], f))
        assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
        # the files carry %10.5f: print resolution + 1e-8 relative
        assert np.all(np.abs(a[:, 1] - b[:, 1]) <= 2e-5 + 1e-8 * (np.abs(a[:, 1]) + shot)), f
    pc, vc, ic = _snapshot(out["cpu"], "bench_z0p000")
    pg, vg, ig = _snapshot(out["gpu"], "bench_z0p000")
    oc, og = np.argsort(ic), np.argsort(ig)
    assert np.array_equal(ic[oc], np.arange(N ** 3, dtype=np.uint64)) and np.array_equal(ig[og], ic[oc])
    dp = np.abs(pc[oc].astype(np.float64) - pg[og])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < 3e-5 * box / N
    assert np.abs(vc[oc] - vg[og]).max() < 3e-4 * np.abs(vc).max()


def test_scale_dependent_merged_fofr_128_thirty_steps(require_gpu, tmp_path):
    import bench
    N, box, nsteps = 128, 100.0, 30
    out = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, "fofr", nsteps, lcdm_growth=0)
        if kind == "cpu":
            _run_cpu("fofr", pf, wd, N)
        else:
            _run_gpu("fofr", pf, wd, env={"MGP_SD_MERGED": "1"})
        out[kind] = os.path.join(wd, "output")
    fc, fg = _pofk_files(out["cpu"]), _pofk_files(out["gpu"])
    assert fc == fg and len(fc) >= nsteps
    knyq = np.pi * N / box
    worst = 0.0
    for f in fc:
        a, b = read_pofk(os.path.join(out["cpu"], f)), read_pofk(os.path.join(out["gpu"], f))
        assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
        sel = (a[:, 0] < 0.5 * knyq) & (np.abs(a[:, 1]) > 1e-2)
        worst = max(worst, float(np.max(np.abs(a[sel, 1] - b[sel, 1]) / np.abs(a[sel, 1]))))
    assert worst < 1e-4, worst                                  # BASELINE.json north_star: 1e-4 at k < k_Nyquist / 2
    pc, vc, ic = _snapshot(out["cpu"], "bench_z0p000")
    pg, vg, ig = _snapshot(out["gpu"], "bench_z0p000")
    assert np.array_equal(np.sort(ic), np.sort(ig))
    oc, og = np.argsort(ic), np.argsort(ig)
    dp = np.abs(pc[oc].astype(np.float64) - pg[og])
    dp = np.minimum(dp, box - dp)
    assert np.sqrt((dp ** 2).mean()) < 1e-3 * box / N          # rms: merged fields round differently, one ulp per step


def test_fofrnu_driver_matches_cpu_reference(require_gpu, tmp_path):
    import bench
    N, box, nsteps = 64, 256.0, 8
    camb = str(tmp_path / "camb")
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "camb_nu0.2.tgz")) as t:
        t.extractall(camb)
    lines = open(os.path.join(camb, "picola_transfer_info_nu0.2.txt")).read().split("\n")
    lines[0] = "%s 50" % camb                                   # the file ships with its author's absolute path
    open(os.path.join(camb, "info.txt"), "w").write("\n".join(lines))
    extra = "nu_FilenameTransferInfofile %s/info.txt\nnu_include_massive_neutrinos 1\nnu_SumMassNuEV 0.2\n" % camb
    out = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, "fofr", nsteps, lcdm_growth=0, extra=extra)
        txt = open(pf).read().replace("Omega 0.267", "Omega 0.3175").replace("HubbleParam 0.71", "HubbleParam 0.671")
        open(pf, "w").write(txt)
        if kind == "cpu":
            _run_cpu("fofrnu", pf, wd, N)
        else:
            _run_gpu("fofrnu", pf, wd, env={"MGP_SD_MERGED": "0"})
        out[kind] = os.path.join(wd, "output")
    shot = (box / N) ** 3
    nfiles = 0
    for suffix in ("_CDM.txt", "_total.txt"):
        fc = sorted(f for f in os.listdir(out["cpu"]) if f.startswith("pofk_") and f.endswith(suffix))
        fg = sorted(f for f in os.listdir(out["gpu"]) if f.startswith("pofk_") and f.endswith(suffix))
        assert fc == fg
        for f in fc:
            a, b = read_pofk(os.path.join(out["cpu"], f)), read_pofk(os.path.join(out["gpu"], f))
            assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
            assert np.all(np.abs(a[:, 1] - b[:, 1]) <= 2e-5 + 1e-6 * (np.abs(a[:, 1]) + shot)), f
            nfiles += 1
    assert nfiles >= 2 * (nsteps - 1)                          # both spectra of (almost) every step
    pc, vc, ic = _snapshot(out["cpu"], "bench_z0p000")
    pg, vg, ig = _snapshot(out["gpu"], "bench_z0p000")
    oc, og = np.argsort(ic), np.argsort(ig)
    assert np.array_equal(ic[oc], ig[og])
    dp = np.abs(pc[oc].astype(np.float64) - pg[og])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < 3e-5 * box / N
