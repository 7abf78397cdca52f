"""Pins the numpy oracle (oracle/pm_oracle.py) against the UNMODIFIED reference sources compiled into
oracle/_ref (oracle/Makefile), function by function, on the same seeded inputs.  CPU only.

The reference publishes no golden vectors, so the pin is the reference's own
code executed here: PtoMesh, Forces, MtoParticles, ComputeFifthForce, Kick, Drift,
compute_power_spectrum, called through ctypes on globals set exactly as main() would.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import pm_oracle as po
from oracle import ref_lib

OMEGA = 0.267


def _need(variant):
    if not ref_lib.available(variant):
        pytest.skip("oracle/_ref/%s not built (build needs /root/reference)" % variant)


def particles(n_side, box, seed, clustered=True):
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    nc = n // 2 if clustered else 0
    centres = rng.random((6, 3)) * box
    c = centres[rng.integers(0, 6, nc)] + rng.standard_normal((nc, 3)) * box * 0.02
    pos = np.concatenate([rng.random((n - nc, 3)) * box, np.mod(c, box)]).astype(np.float32)
    pos[pos >= np.float32(box)] = 0
    top = np.nextafter(np.float32(box), np.float32(0))
    h = np.float32(box / n_side)
    pos[:4] = [[0, 0, 0], [top, top, top], [h, 2 * h, 3 * h], [top, 0, h / 2]]
    D = (rng.standard_normal((n, 3)) * 0.5).astype(np.float32)
    D2 = (rng.standard_normal((n, 3)) * 0.1).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 0.2).astype(np.float32)
    return pos, vel, D, D2


def ref_setup(variant, N, box, pos, vel, D, D2, mg=False, **globals_):
    r = ref_lib.RefLib(variant)
    with ref_lib._silenced(True):
        r.setup_grid(N, N, box, omega=OMEGA)
        r.set(TotNumPart=pos.shape[0], modified_gravity_active=int(mg), allocate_mg_arrays=int(mg),
              pofk_compute_every_step=0, pofk_compute_rsd_pofk=0, timeStep_global=1, StdDA=0, fullT=1, nLPT=-2.5,
              **globals_)
        r.set_particles(pos, vel, D, D2)
        r.alloc_step_grids(mg=mg)
    return r


@pytest.mark.parametrize("variant,gdt", [("lcdm", np.float64), ("lcdm_sp", np.float32)])
def test_lcdm_step_functions(variant, gdt):
    """PtoMesh -> Forces -> MtoParticles of the reference == oracle (auxPM.c:280-644)."""
    _need(variant)
    N, box = 16, 50.0
    pos, vel, D, D2 = particles(N, box, 3)
    r = ref_setup(variant, N, box, pos, vel, D, D2)
    with ref_lib._silenced(True):
        r.lib.PtoMesh()
    dk = r.grid_k("density").astype(np.complex128)
    dens = po.ptomesh_deposit(pos, N, N, box, grid_dtype=gdt)
    ok = po.r2c(dens, N)
    tol = 1e-13 if gdt == np.float64 else 2e-6
    assert np.abs(dk - ok).max() / np.abs(ok).max() < tol
    with ref_lib._silenced(True):
        r.lib.Forces()
    Fo = po.forces(dk, N, box)            # from the reference's own density so that only Forces is compared
    for nm, f in zip(("N11", "N12", "N13"), Fo):
        g = r.grid(nm)
        assert np.abs(g[:N, :, :N] - f).max() / np.abs(f).max() < tol
        assert np.array_equal(g[N], g[0])                      # ghost slice filled from slice 0 (auxPM.c:546-551)
    with ref_lib._silenced(True):
        r.alloc_disp()
        r.lib.MtoParticles()
    F = [r.grid(nm)[:N, :, :N].astype(np.float64) for nm in ("N11", "N12", "N13")]
    do, sD = po.mtoparticles(pos, *F, N, box)
    d = r.disp()
    assert np.array_equal(d.view(np.uint32), do.view(np.uint32))      # same grids in -> bit-identical floats out
    assert np.allclose(r.get3("sumDxyz"), sD, rtol=0, atol=1e-15)


def test_kick_drift_bit_exact():
    """Kick (main.c:688-742) and Drift (main.c:747-787) of the reference == oracle, bit for bit."""
    _need("lcdm")
    N, box = 12, 40.0
    pos, vel, D, D2 = particles(N, box, 5)
    n = pos.shape[0]
    r = ref_setup("lcdm", N, box, pos, vel, D, D2)
    with ref_lib._silenced(True):
        pf = _paramfile(N, box)
        r.init_from_paramfile(pf)               # growth-factor splines for growth_ddDddy etc.
        r.set(TotNumPart=n)
        r.set_particles(pos, vel, D, D2)
    L = r.lib
    rng = np.random.default_rng(0)
    disp = (rng.standard_normal((n, 3)) * 0.3).astype(np.float32)
    bufs = r.alloc_disp()
    for a in range(3):
        bufs[a][:n] = disp[:, a]
    sumD = np.array([1e-3, -2e-3, 5e-4])
    r.set3("sumDxyz", sumD)
    AI, AF, A, AFF = 0.31, 0.33, 0.32, 0.34
    Di, Di2 = L.growth_D(A), L.growth_D2(A)
    L.Kick(AI, AF, A, Di)
    P = r.particles()
    vref, dref, svref = po.kick(vel, disp, D, D2, sumD, OMEGA, 1, A, L.Sphi(AI, AF, A), L.growth_ddDddy(A), L.growth_ddD2ddy(A))
    assert np.array_equal(P["Vel"].view(np.uint32), vref.view(np.uint32))
    assert np.array_equal(r.disp().view(np.uint32), dref.view(np.uint32))
    sv = r.get3("sumxyz")
    assert np.allclose(sv, svref, rtol=0, atol=1e-14)
    L.Drift(A, AFF, AF, Di, Di2)
    pref = po.drift(pos, vref, D, D2, sv, box, 1, L.Sq(A, AFF, AF), L.growth_D(AFF) - Di, L.growth_D2(AFF) - Di2)
    assert np.array_equal(r.particles()["Pos"].view(np.uint32), pref.view(np.uint32))


def _paramfile(N, box, model="fofr"):
    import sys
    import tempfile
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    wd = tempfile.mkdtemp(prefix="mgp_test_")
    return bench.write_paramfile(wd, N, box, model, 10)


@pytest.mark.parametrize("screening", [1, 0])
def test_fofr_fifth_force(screening):
    """ComputeFifthForce_PotentialScreening (mg.h:147-189) + the scalars of user_defined_functions.h."""
    _need("lcdm")
    N, box, a = 16, 50.0, 0.7
    pos, vel, D, D2 = particles(N, box, 7)
    r = ref_setup("lcdm", N, box, pos, vel, D, D2, mg=True, include_screening=screening, fofr0=1e-5, nfofr=1.0, aexp_global=a)
    with ref_lib._silenced(True):
        r.lib.PtoMesh()
    dk = r.grid_k("density").copy()
    dens_real = r.grid("mgarray_two")[:N, :, :N].copy()           # CopyDensityArray (mg.h:381-386)
    with ref_lib._silenced(True):
        r.lib.ComputeFifthForce()
    phik = r.grid_k("mgarray_two")
    phicrit, coupling, massterm2 = po.fofr_scalars(a, OMEGA, box, 1e-5, 1.0)
    L = r.lib
    assert abs(L.coupling_function(a) - coupling) < 1e-15
    assert abs(a * a * L.mass2_of_a(a) / ((2 * po.PI) * po.INVERSE_H0_MPCH / box) ** 2 / massterm2 - 1) < 1e-14
    if screening:     # how the C adapter obtains Phi_crit without touching user_defined_functions.h
        big = -1e30                                                # Phi_crit = factor * |Phi| for deeply screened cells
        assert abs(L.screening_factor_potential(a, big) * abs(big) / phicrit - 1) < 1e-13
    ok = po.fifth_force_potential_screening(dk, dens_real, N, box, OMEGA, a, phicrit, coupling, massterm2, bool(screening))
    assert np.abs(phik - ok).max() / np.abs(ok).max() < 1e-12


@pytest.mark.parametrize("screening", [1, 0])
def test_dgp_fifth_force(screening):
    """ComputeFifthForce_DensityScreening with the Gaussian filter (mg.h:197-309)."""
    _need("dgp")
    N, box, a = 16, 50.0, 0.8
    pos, vel, D, D2 = particles(N, box, 9)
    r = ref_setup("dgp", N, box, pos, vel, D, D2, mg=True, include_screening=screening, rcH0_DGP=1.2, Rsmooth_global=1.0,
                  aexp_global=a)
    with ref_lib._silenced(True):
        r.lib.PtoMesh()
    dk = r.grid_k("density").copy()
    dens_real = r.grid("mgarray_two")[:N, :, :N].copy()
    with ref_lib._silenced(True):
        r.lib.ComputeFifthForce()
    phik = r.grid_k("mgarray_two")
    coupling, fac0 = po.dgp_scalars(a, OMEGA, 1.2)
    assert abs(r.lib.coupling_function(a) - coupling) < 1e-15
    ok = po.fifth_force_density_screening(dk, dens_real, N, box, coupling, fac0, 1.0, bool(screening))
    assert np.abs(phik - ok).max() / np.abs(ok).max() < 1e-12


@pytest.mark.parametrize("nbins,bintype,kmin,kmax", [(64, 1, 0.03, 2.0), (0, 0, 0.0, 0.0), (10, 0, 0.2, 1.5)])
def test_power_spectrum_bins(tmp_path, nbins, bintype, kmin, kmax):
    """compute_power_spectrum (compute_pofk.c:71-236): the per-bin sums the reference all-reduces
    (read through the MPI stand-in's tap) == oracle."""
    _need("lcdm")
    N, box = 16, 120.0
    pos, vel, D, D2 = particles(N, box, 11)
    r = ref_setup("lcdm", N, box, pos, vel, D, D2, pofk_nbins=nbins, pofk_bintype=bintype, pofk_subtract_shotnoise=1,
                  pofk_kmin=kmin, pofk_kmax=kmax)
    r.set_str("OutputDir", str(tmp_path))
    r.set_str("FileBase", "t")
    with ref_lib._silenced(True):
        r.lib.PtoMesh()
    dk = r.grid_k("density").copy()
    ref_lib.tap_reset(r)
    with ref_lib._silenced(True):
        r.lib.compute_power_spectrum(r._keep["density"].ctypes.data, 0.5, b"CDM")
    sums = ref_lib.tap_arrays(r)[-3:]                              # pofk_bin_all, n_bin_all, k_bin_all (225-227)
    p, k, n = po.compute_power_spectrum(dk, N, N, box, nbins, bintype, 1, kmin, kmax)
    nb = len(p)
    assert all(len(s) == nb for s in sums)
    assert np.array_equal(sums[1], n)
    good = n > 0
    pref = np.zeros(nb)
    pref[good] = sums[0][good] / n[good] * box ** 3 - (box / N) ** 3
    kref = np.zeros(nb)
    kref[good] = sums[2][good] / n[good] * 2 * np.pi / box
    assert np.allclose(p, pref, rtol=1e-12, atol=1e-12 * (box / N) ** 3)
    assert np.allclose(k, kref, rtol=1e-14)


def test_full_run_matches_oracle_stepping():
    """Three COLA steps of the reference on its own ICs (GetDisplacements/Kick/Drift as compiled,
    MEMORY_MODE allocation churn included) == the oracle stepped with the same scalars."""
    _need("lcdm")
    N, box = 16, 60.0
    pf = _paramfile(N, box, "fofr")
    run = ref_lib.RefRun("lcdm", pf)
    L = run.r.lib
    P0 = run.particles().copy()
    pos, vel, D, D2 = P0["Pos"].copy(), P0["Vel"].copy(), P0["D"].copy(), P0["D2"].copy()
    for it in range(3):
        A, AI, da, Di, Di2 = run.A, run.AI, run.da, run.Di, run.Di2
        AF, AFF = A + 0.5 * da, A + da
        phicrit, coupling, massterm2 = po.fofr_scalars(A, OMEGA, box, 1e-5, 1.0)
        ref = po.get_displacements(pos, N, N, box, model="fofr",
                                   mg=dict(omega=OMEGA, a=A, phi_crit=phicrit, coupling=coupling, massterm2=massterm2))
        vel, _, sv = po.kick(vel, ref["disp"], D, D2, ref["sumDxyz"], OMEGA, 1, A, L.Sphi(AI, AF, A), L.growth_ddDddy(A),
                             L.growth_ddD2ddy(A))
        pos = po.drift(pos, vel, D, D2, sv, box, 1, L.Sq(A, AFF, AF), L.growth_D(AFF) - Di, L.growth_D2(AFF) - Di2)
        run.step()
        P = run.particles()
        assert np.array_equal(P["ID"], P0["ID"])
        dp = np.abs(P["Pos"].astype(np.float64) - pos)
        dp = np.minimum(dp, box - dp)
        # float32 positions; the two sides differ only by FFT rounding (1e-16) amplified through float stores
        assert dp.max() < 2e-5 * box / N
        assert np.abs(P["Vel"] - vel).max() < 1e-5 * np.abs(vel).max()
