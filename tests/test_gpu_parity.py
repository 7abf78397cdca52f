"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs.  Tolerances are stated per test; integer / ownership results must be exact."""
import numpy as np
import pytest

from oracle import pm_oracle as po

pytestmark = pytest.mark.gpu

OMEGA = 0.267


def make_particles(n_side, box, seed, clustered=False):
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    if clustered:
        # half uniform, half in a few tight clumps (many particles per cell, incl. box edges)
        nc = n // 2
        centres = rng.random((8, 3)) * box
        centres[0] = [0.01, box - 0.01, box / 2]
        c = centres[rng.integers(0, 8, nc)] + rng.standard_normal((nc, 3)) * box * 0.01
        pos = np.concatenate([rng.random((n - nc, 3)) * box, np.mod(c, box)])
    else:
        pos = rng.random((n, 3)) * box
    pos = pos.astype(np.float32)
    pos[pos >= np.float32(box)] = 0
    pos[pos < 0] = 0
    D = (rng.standard_normal((n, 3)) * 0.5).astype(np.float32)
    D2 = (rng.standard_normal((n, 3)) * 0.1).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 0.2).astype(np.float32)
    return pos, vel, D, D2


def adversarial_positions(N, box):
    """Particles exactly on cell boundaries, at 0, and at the largest float below Box."""
    h = box / N
    top = np.nextafter(np.float32(box), np.float32(0))
    pts = [[0, 0, 0], [top, top, top], [h, 2 * h, 3 * h], [box - h, box - h, box - h], [top, 0, h / 2],
           [h / 2, top, 0], [0.5 * h, 0.5 * h, top], [box / 2, box / 2, box / 2]]
    return np.array(pts, dtype=np.float32)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("gb", [8, 4])
@pytest.mark.parametrize("clustered", [False, True])
def test_deposit_matches_oracle(mgp, require_gpu, mode, gb, clustered):
    N, box = 32, 100.0
    pos, vel, D, D2 = make_particles(N, box, 3, clustered)
    pos[:8] = adversarial_positions(N, box)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=gb, deposit_mode=mode)
    pm.upload_particles(pos, vel, D, D2)
    pm.MoveParticles()
    # deposit only: run PtoMesh then undo the FFT by comparing in k-space against the oracle
    pm.PtoMesh()
    dk = pm.download_grid_k(mgp.GRID_DENSITY)
    dens = po.ptomesh_deposit(pos, N, N, box)
    ref = po.r2c(dens, N)
    scale = np.abs(ref).max()
    tol = 1e-12 if gb == 8 else 2e-5
    assert np.abs(dk - ref).max() / scale < tol
    # mass conservation: the k = 0 mode is sum(delta) = Np*W - Ncells = 0 here
    assert abs(dk[0, 0, 0]) / N ** 3 < (1e-12 if gb == 8 else 1e-5)
    pm.close()


def test_deposit_nmesh_ne_nsample(mgp, require_gpu):
    """W = (Nmesh/Nsample)^3 weighting (auxPM.c:289, 316-317) and an odd particle count."""
    N, Ns, box = 32, 20, 75.0
    pos, vel, D, D2 = make_particles(Ns, box, 5)
    for mode in (0, 1, 2):
        pm = mgp.PM(N, Ns, box, omega=OMEGA, grid_bytes=8, deposit_mode=mode)
        pm.upload_particles(pos, vel, D, D2)
        pm.PtoMesh()
        dk = pm.download_grid_k(mgp.GRID_DENSITY)
        ref = po.r2c(po.ptomesh_deposit(pos, N, Ns, box), N)
        assert np.abs(dk - ref).max() / np.abs(ref).max() < 1e-12
        pm.close()


def test_deposit_deterministic_bitwise(mgp, require_gpu):
    """MGP_DEPOSIT_DETERMINISTIC: two runs give bit-identical grids (fp32 grids, clustered)."""
    N, box = 32, 100.0
    pos, vel, D, D2 = make_particles(N, box, 9, clustered=True)
    grids = []
    for _ in range(2):
        pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=4, deposit_mode=2)
        pm.upload_particles(pos, vel, D, D2)
        pm.PtoMesh()
        grids.append(pm.download_grid(mgp.GRID_DENSITY).copy())
        pm.close()
    assert np.array_equal(grids[0].view(np.uint32), grids[1].view(np.uint32))


def test_empty_particle_set(mgp, require_gpu):
    N, box = 16, 50.0
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8)
    pm.upload_particles(np.zeros((0, 3), np.float32))
    pm.PtoMesh()
    dk = pm.download_grid_k(mgp.GRID_DENSITY)
    assert abs(dk[0, 0, 0] + N ** 3) < 1e-9          # grid is -1 everywhere
    assert np.abs(dk.reshape(-1)[1:]).max() < 1e-9
    pm.close()


@pytest.mark.parametrize("gb", [8, 4])
def test_fft_roundtrip_and_nonhermitian_c2r(mgp, require_gpu, gb):
    N, box = 16, 50.0
    pm = mgp.PM(N, N, box, grid_bytes=gb)
    rng = np.random.default_rng(0)
    nzp = 2 * (N // 2 + 1)
    g = np.zeros((N + 1, N, nzp), pm.gdtype)
    g[:N, :, :N] = rng.standard_normal((N, N, N))
    pm.upload_grid(mgp.GRID_DENSITY, g)
    pm.fft_r2c(mgp.GRID_DENSITY)
    k = pm.download_grid_k(mgp.GRID_DENSITY)
    ref = po.r2c(g[:N].astype(np.float64), N)
    tol = 1e-12 if gb == 8 else 1e-5
    assert np.abs(k - ref).max() / np.abs(ref).max() < tol
    # c2r of a spectrum that is NOT Hermitian on the kz = 0 / Nyquist planes (as Forces produces)
    ck = (rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))).astype(pm.cdtype)
    pm.upload_grid_k(mgp.GRID_DENSITY, ck)
    pm.fft_c2r(mgp.GRID_DENSITY)
    r = pm.download_grid(mgp.GRID_DENSITY)[:N, :, :N]
    ref = po.c2r(ck.astype(np.complex128), N)
    assert np.abs(r - ref).max() / np.abs(ref).max() < tol
    pm.close()


@pytest.mark.parametrize("gb", [8, 4])
def test_get_displacements_lcdm(mgp, require_gpu, gb):
    """Full GetDisplacements (auxPM.c:37-103) against the oracle: force grids, Disp, sumDxyz."""
    N, box = 32, 100.0
    pos, vel, D, D2 = make_particles(N, box, 21, clustered=True)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=gb)
    pm.upload_particles(pos, vel, D, D2)
    sumD = pm.GetDisplacements()
    ref = po.get_displacements(pos, N, N, box)
    tol = 1e-11 if gb == 8 else 3e-5
    for a in range(3):
        F = pm.download_grid(mgp.GRID_FORCE_X + a)
        Fr = ref["force_grids"][a]
        assert np.abs(F[:N, :, :N] - Fr).max() / np.abs(Fr).max() < tol
        # ghost plane == plane 0 (auxPM.c:546-551 with a single task)
        assert np.array_equal(F[N, :, :N], F[0, :, :N])
    got = pm.download_particles()
    order = np.argsort(got["id"])
    disp = pm.download_disp()[order]
    assert np.array_equal(got["id"][order], np.arange(N ** 3, dtype=np.uint64))
    dtol = 2e-7 if gb == 8 else 5e-5      # float32 storage of Disp: 1 ulp ~ 6e-8 relative
    assert np.abs(disp - ref["disp"]).max() / np.abs(ref["disp"]).max() < dtol
    assert np.abs(sumD - ref["sumDxyz"]).max() < 1e-7 * np.abs(ref["disp"]).max()
    pm.close()


def test_kick_drift_bit_exact(mgp, require_gpu):
    """Kick / Drift are streaming kernels written to round exactly like the reference's C
    (double arithmetic without FMA contraction, float storage): bit-identical floats."""
    N, box = 16, 64.0
    pos, vel, D, D2 = make_particles(N, box, 33)
    n = N ** 3
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8, sort_particles=0)
    pm.upload_particles(pos, vel, D, D2)
    sumD = pm.GetDisplacements()
    disp = pm.download_disp()
    A, dda, ddD, ddD2 = 0.3, 0.0173, 1.234, -0.567
    sv = pm.Kick(A, dda, ddD, ddD2)
    got = pm.download_particles()
    vref, dref, svref = po.kick(vel, disp, D, D2, sumD, OMEGA, 1, A, dda, ddD, ddD2)
    assert np.array_equal(got["vel"].view(np.uint32), vref.view(np.uint32))
    assert np.allclose(sv, svref, rtol=0, atol=1e-12)
    # second kick of an output step: sumDxyz = 0, Disp already mean-subtracted (main.c:566-569)
    sv2 = pm.Kick(A, 0.5 * dda, ddD, ddD2, sumDxyz=np.zeros(3))
    v2, _, sv2ref = po.kick(vref, dref, D, D2, np.zeros(3), OMEGA, 1, A, 0.5 * dda, ddD, ddD2)
    got = pm.download_particles()
    assert np.array_equal(got["vel"].view(np.uint32), v2.view(np.uint32))
    dyyy, dD, dD2 = 0.731, 0.0421, -0.0113
    pm.Drift(dyyy, dD, dD2)
    got = pm.download_particles()
    pref = po.drift(pos, v2, D, D2, sv2, box, 1, dyyy, dD, dD2)
    assert np.array_equal(got["pos"].view(np.uint32), pref.view(np.uint32))
    assert (got["pos"] >= 0).all() and (got["pos"] < np.float32(box)).all()
    pm.close()


def test_drift_wraps_far_particles(mgp, require_gpu):
    N, box = 8, 10.0
    n = N ** 3
    rng = np.random.default_rng(1)
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 50).astype(np.float32)     # several box lengths per step
    D = np.zeros((n, 3), np.float32)
    pm = mgp.PM(N, N, box, omega=OMEGA, sort_particles=0)
    pm.upload_particles(pos, vel, D, D)
    pm.Drift(1.0, 0.0, 0.0, sumxyz=np.zeros(3))
    got = pm.download_particles()
    pref = po.drift(pos, vel, D, D, np.zeros(3), box, 1, 1.0, 0.0, 0.0)
    assert np.array_equal(got["pos"].view(np.uint32), pref.view(np.uint32))
    pm.close()


@pytest.mark.parametrize("bintype,nbins,kmin,kmax", [(1, 64, 0.03, 2.0), (0, 0, 0.0, 0.0), (0, 16, 0.1, 0.9)])
@pytest.mark.parametrize("gb", [8, 4])
def test_power_spectrum(mgp, require_gpu, bintype, nbins, kmin, kmax, gb):
    N, box = 32, 200.0
    pos, vel, D, D2 = make_particles(N, box, 17, clustered=True)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=gb)
    pm.set_pofk(nbins, bintype, 1, kmin, kmax)
    pm.upload_particles(pos, vel, D, D2)
    pm.PtoMesh(pm.scalars(compute_pofk=1))
    p, k, n = pm.step_power_spectrum()
    P3D = po.r2c(po.ptomesh_deposit(pos, N, N, box), N)
    pr, kr, nr = po.compute_power_spectrum(P3D, N, N, box, nbins, bintype, 1, kmin, kmax)
    assert p.shape == pr.shape
    assert np.array_equal(n, nr)                       # mode counts per bin: exact
    assert np.allclose(k, kr, rtol=1e-13, atol=0)
    good = nr > 0
    shot = (box / N) ** 3
    rel = np.abs(p[good] - pr[good]) / (np.abs(pr[good]) + shot)
    assert rel.max() < (1e-11 if gb == 8 else 1e-4)
    p2, k2, n2 = pm.compute_power_spectrum()           # stand-alone call on the same grid
    assert np.allclose(p2, p, rtol=1e-12, atol=1e-12 * shot)
    pm.close()


@pytest.mark.parametrize("screening", [1, 0])
@pytest.mark.parametrize("gb", [8, 4])
def test_fofr_fifth_force(mgp, require_gpu, screening, gb):
    N, box = 32, 100.0
    pos, vel, D, D2 = make_particles(N, box, 41, clustered=True)
    a = 0.7
    phicrit, coupling, massterm2 = po.fofr_scalars(a, OMEGA, box, 1e-5, 1.0)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=gb, model=mgp.MODEL_FOFR, include_screening=screening)
    pm.upload_particles(pos, vel, D, D2)
    s = pm.scalars(a=a, phi_crit=phicrit, coupling=coupling, massterm2=massterm2)
    sumD = pm.GetDisplacements(s)
    mgk = pm.download_grid_k(mgp.GRID_MG_TWO)
    ref = po.get_displacements(pos, N, N, box, model="fofr",
                               mg=dict(omega=OMEGA, a=a, phi_crit=phicrit, coupling=coupling, massterm2=massterm2,
                                       screening=bool(screening)))
    dens_real = ref["density"][:, :, :N].astype(np.float64)
    phik = po.fifth_force_potential_screening(po.r2c(ref["density"], N), dens_real, N, box, OMEGA, a, phicrit,
                                              coupling, massterm2, bool(screening))
    tol = 1e-10 if gb == 8 else 2e-4
    assert np.abs(mgk - phik).max() / np.abs(phik).max() < tol
    got = pm.download_particles()
    disp = pm.download_disp()[np.argsort(got["id"])]
    assert np.abs(disp - ref["disp"]).max() / np.abs(ref["disp"]).max() < (3e-7 if gb == 8 else 2e-4)
    # the fifth force must actually matter in this configuration
    lcdm = po.get_displacements(pos, N, N, box)
    assert np.abs(lcdm["disp"] - ref["disp"]).max() / np.abs(ref["disp"]).max() > 1e-3
    pm.close()


@pytest.mark.parametrize("screening", [1, 0])
def test_dgp_fifth_force(mgp, require_gpu, screening):
    N, box = 32, 100.0
    pos, vel, D, D2 = make_particles(N, box, 43, clustered=True)
    a = 0.8
    coupling, fac0 = po.dgp_scalars(a, OMEGA, 1.2)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8, model=mgp.MODEL_DGP, include_screening=screening)
    pm.upload_particles(pos, vel, D, D2)
    s = pm.scalars(a=a, coupling=coupling, dgp_fac0=fac0, rsmooth=1.0)
    pm.GetDisplacements(s)
    ref = po.get_displacements(pos, N, N, box, model="dgp",
                               mg=dict(coupling=coupling, dgp_fac0=fac0, rsmooth=1.0, screening=bool(screening)))
    got = pm.download_particles()
    disp = pm.download_disp()[np.argsort(got["id"])]
    assert np.abs(disp - ref["disp"]).max() / np.abs(ref["disp"]).max() < 3e-7
    pm.close()


def test_geff_model(mgp, require_gpu):
    N, box = 16, 50.0
    pos, vel, D, D2 = make_particles(N, box, 47)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8, model=mgp.MODEL_GEFF)
    pm.upload_particles(pos, vel, D, D2)
    pm.GetDisplacements(pm.scalars(a=0.5, geff=1.25))
    ref = po.get_displacements(pos, N, N, box, model="geff", mg=dict(geff=1.25))
    got = pm.download_particles()
    disp = pm.download_disp()[np.argsort(got["id"])]
    assert np.abs(disp - ref["disp"]).max() / np.abs(ref["disp"]).max() < 3e-7
    pm.close()


def test_errors_are_loud(mgp, require_gpu):
    with pytest.raises(mgp.MgpError):
        mgp.PM(15, 15, 10.0)                     # odd Nmesh
    pm = mgp.PM(8, 8, 10.0)
    with pytest.raises(mgp.MgpError):
        pm.Kick(0.5, 0.1, 1.0, 1.0)              # Kick before GetDisplacements
    with pytest.raises(mgp.MgpError):
        pm.upload_particles(np.zeros((8 ** 3 + 1000, 3), np.float32))   # over capacity
    with pytest.raises(mgp.MgpError):
        pm.PtoMesh(pm.scalars(compute_pofk=1))   # P(k) without binning configuration
    pm.close()


def test_launch_counter(mgp, require_gpu):
    N, box = 16, 50.0
    pos, vel, D, D2 = make_particles(N, box, 2)
    pm = mgp.PM(N, N, box)
    pm.upload_particles(pos, vel, D, D2)
    pm.launch_count(reset=True)
    pm.GetDisplacements()
    assert pm.launch_count() >= 10
    pm.close()
