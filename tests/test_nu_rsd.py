"""Massive-neutrino density add (auxPM.c:383-420) and redshift-space multipoles (PtoMesh_RSD,
compute_RSD_powerspectrum, bin_up_RSD_power_spectrum; compute_pofk.c:280-753).

Fixtures tests/golden/step_nu.npz (reference build MODEL=FOFRNU on its bundled CAMB data) and
tests/golden/rsd_lcdm.npz were written by tools/make_golden.py --nu --rsd from oracle/_ref.  The reference only
prints these spectra with %10.5f, so the fixtures hold the per-bin sums it hands to MPI_Allreduce.
CPU: oracle vs fixture.  GPU: CUDA path through the C ABI vs fixture."""
import os

import numpy as np
import pytest

from oracle import pm_oracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


def pk_cfg(g):
    c = g["pofk_cfg"]
    return dict(nbins=int(c[0]), bintype=int(c[1]), subtract_shotnoise=int(c[2]), kmin_hmpc=float(c[3]), kmax_hmpc=float(c[4]))


def pofk_from(s, box, nsample):
    n = s[1]
    good = n > 0
    p = np.zeros_like(n)
    p[good] = s[0][good] / n[good] * box ** 3 - (box / nsample) ** 3
    return p, n


# ----------------------------------------------------------------------------- CPU

def test_oracle_nu_add():
    g = load("step_nu.npz")
    N, box = int(g["N"]), float(g["box"])
    dens = po.ptomesh_deposit(g["pos"], N, N, box)
    P3D = po.r2c(dens, N)
    p, k, n = po.compute_power_spectrum(P3D, N, N, box, **pk_cfg(g))
    pr, nr = pofk_from(g["pofk_cdm"], box, N)
    assert np.array_equal(n, nr) and np.allclose(p, pr, rtol=1e-11, atol=1e-11 * (box / N) ** 3)
    tot = po.nu_add(P3D, g["cdelta_cdm"], g["nu_by_k2"], float(g["cdmfac"]), N)
    assert np.abs(tot - g["density_k"]).max() <= 1e-13 * np.abs(g["density_k"]).max()
    p, k, n = po.compute_power_spectrum(tot, N, N, box, **pk_cfg(g))
    pr, nr = pofk_from(g["pofk_total"], box, N)
    assert np.array_equal(n, nr) and np.allclose(p, pr, rtol=1e-11, atol=1e-11 * (box / N) ** 3)
    assert np.abs(tot - P3D).max() > 1e-3 * np.abs(P3D).max()          # the neutrinos do something


def test_oracle_rsd_multipoles():
    g = load("rsd_lcdm.npz")
    N, box = int(g["N"]), float(g["box"])
    Vy = po.rsd_velocity(g["vel"], g["D"], g["D2"], 1, 1, float(g["dDdy"]), float(g["dD2dy"]))
    Vz = po.rsd_velocity(g["vel"], g["D"], g["D2"], 2, 1, float(g["dDdy"]), float(g["dD2dy"]))
    out = po.compute_rsd_powerspectrum(g["pos"], Vy, Vz, float(g["vnorm"]), N, N, box, pk_cfg(g))
    for ax in ("y", "z"):
        ref = g["sums_" + ax]                       # P0, P2, P4, n, k sums
        got = out[ax + "_sums"]
        assert np.array_equal(got[3], ref[3])       # mode counts exact
        for q in (0, 1, 2, 4):
            assert np.abs(got[q] - ref[q]).max() <= 1e-12 * np.abs(ref[q]).max()


# ----------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("gb", [8, 4])
def test_cuda_nu_add(mgp, require_gpu, gb):
    """PtoMesh with mgp_step_scalars.nu_by_k2: density_k after the add and both in-step spectra."""
    g = load("step_nu.npz")
    N, box = int(g["N"]), float(g["box"])
    pm = mgp.PM(N, N, box, omega=float(g["omega"]), model=mgp.MODEL_FOFR, include_screening=1, grid_bytes=gb, scale_dependent=1)
    c = g["pofk_cfg"]
    pm.set_pofk(int(c[0]), int(c[1]), int(c[2]), float(c[3]), float(c[4]))
    pm.upload_grid_k(mgp.GRID_SD_DELTA1, g["cdelta_cdm"])
    pm.upload_particles(g["pos"])
    pm.MoveParticles()
    pm.PtoMesh(pm.scalars(a=float(g["a"]), compute_pofk=1, nu_by_k2=g["nu_by_k2"], nu_cdmfac=float(g["cdmfac"])))
    dk = pm.download_grid_k(mgp.GRID_DENSITY)
    tol = 1e-12 if gb == 8 else 2e-5
    assert np.abs(dk - g["density_k"]).max() <= tol * np.abs(g["density_k"]).max()
    for fn, key in ((pm.step_power_spectrum, "pofk_cdm"), (pm.step_power_spectrum_total, "pofk_total")):
        p, k, n = fn()
        pr, nr = pofk_from(g[key], box, N)
        assert np.array_equal(n, nr)
        good = nr > 0
        rel = np.abs(p[good] - pr[good]) / (np.abs(pr[good]) + (box / N) ** 3)
        assert rel.max() < (1e-11 if gb == 8 else 1e-4)
    pm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gb", [8, 4])
def test_cuda_rsd_multipoles(mgp, require_gpu, gb):
    """mgp_compute_rsd_power_spectrum == the reference's compute_RSD_powerspectrum: mode counts exact, P0 / P2 / P4
    within 1e-11 (f64 grids) / 1e-4 (f32 grids) of the value range."""
    g = load("rsd_lcdm.npz")
    N, box = int(g["N"]), float(g["box"])
    pm = mgp.PM(N, N, box, omega=float(g["omega"]), grid_bytes=gb)
    c = g["pofk_cfg"]
    pm.set_pofk(int(c[0]), int(c[1]), int(c[2]), float(c[3]), float(c[4]))
    pm.upload_particles(g["pos"], g["vel"], g["D"], g["D2"], g["id"])
    pm.MoveParticles()
    out = pm.compute_RSD_powerspectrum(float(g["vnorm"]), float(g["dDdy"]), float(g["dD2dy"]))
    for ax in ("y", "z"):
        sums = tuple(g["sums_" + ax])
        cfg = po.adjust_pofk_parameters(N, box, int(c[0]), int(c[1]), int(c[2]), float(c[3]), float(c[4]))
        n, k, P0, P2, P4 = po.rsd_multipoles(sums, cfg, box, N)
        got = out[ax]
        assert np.array_equal(got[0], n)
        assert np.allclose(got[1], k, rtol=1e-13)
        tol = 1e-11 if gb == 8 else 1e-4
        for have, want in ((got[2], P0), (got[3], P2), (got[4], P4)):
            assert np.abs(have - want).max() <= tol * (np.abs(want).max() + (box / N) ** 3)
    # the combination written to file (compute_pofk.c:465-476)
    assert np.allclose(out["P0"], (out["y"][2] + out["z"][2]) / 2) and np.all(out["err0"] >= 0)
    pm.close()


@pytest.mark.gpu
def test_cuda_rsd_scale_dependent_velocity(mgp, require_gpu):
    """SCALEDEPENDENT branch of PtoMesh_RSD (V = Vel + (dDdy + dD2dy), float sum) against the oracle."""
    g = load("rsd_lcdm.npz")
    N, box = int(g["N"]), float(g["box"])
    rng = np.random.default_rng(4)
    n = g["pos"].shape[0]
    f1 = (rng.standard_normal((n, 3)) * 0.8).astype(np.float32)
    f2 = (rng.standard_normal((n, 3)) * 0.1).astype(np.float32)
    pm = mgp.PM(N, N, box, omega=float(g["omega"]), scale_dependent=1, sort_particles=0)
    c = g["pofk_cfg"]
    pm.set_pofk(int(c[0]), int(c[1]), int(c[2]), float(c[3]), float(c[4]))
    pm.upload_particles(g["pos"], g["vel"], None, None, g["id"])
    pm.upload_sd_fields(f1, f2)
    out = pm.compute_RSD_powerspectrum(float(g["vnorm"]))
    Vy = po.rsd_velocity(g["vel"], f1, f2, 1, 1, 0, 0, scale_dependent=True)
    Vz = po.rsd_velocity(g["vel"], f1, f2, 2, 1, 0, 0, scale_dependent=True)
    ref = po.compute_rsd_powerspectrum(g["pos"], Vy, Vz, float(g["vnorm"]), N, N, box, pk_cfg(g))
    for ax in ("y", "z"):
        for q in range(5):
            assert np.abs(out[ax][q] - ref[ax][q]).max() <= 1e-10 * (np.abs(ref[ax][q]).max() + 1.0)
    pm.close()
