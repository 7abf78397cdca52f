"""Scale-dependent growth (-DSCALEDEPENDENT, reference build MODEL=FOFR): from_cdisp_store_to_ZA +
assign_displacment_field_to_particles (2LPT.c:1539-2005), the SCALEDEPENDENT branches of the particle
initialisation, Kick and Drift (main.c:231-304, 705-717, 760-770).

CPU: the oracle against the compiled reference (oracle/_ref, when built) and against the committed fixture
tests/golden/sd_fofr.npz (written by tools/make_golden.py --sd from the reference run with
use_lcdm_growth_factors = 0, seed 5001).  GPU: the CUDA path through the C ABI against the same fixture.
"""
import os

import numpy as np
import pytest

from oracle import pm_oracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SLOT = {po.FIELD_D: ("D", "D2"), po.FIELD_ddDddy: ("D", "D2"), po.FIELD_dDdy: ("dDdy", "dD2dy"), po.FIELD_deltaD: ("dDdy", "dD2dy")}


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(G, "sd_fofr.npz")))


def pk_cfg(g):
    c = g["pofk_cfg"]
    return dict(nbins=int(c[0]), bintype=int(c[1]), subtract_shotnoise=int(c[2]), kmin_hmpc=float(c[3]), kmax_hmpc=float(c[4]))


# ----------------------------------------------------------------------------- CPU: oracle vs fixture

def test_oracle_sd_fields_at_init(g):
    """The four initial fields (main.c:246-251) from the stored delta_k and the growth tables: the oracle
    reproduces the reference's floats exactly (same FFT back end)."""
    N, box = int(g["N"]), float(g["box"])
    want = [g["D0"], g["D20"], g["dDdy0"], g["dD2dy0"]]
    for idx, (ft, order) in enumerate([(0, 1), (0, 2), (1, 1), (1, 2)]):
        za = po.sd_displacement_field(g["delta1"] if order == 1 else g["delta2"], g["G_init"][idx], order, N, N, box)
        assert np.array_equal(za.astype(np.float32).view(np.uint32), want[idx].view(np.uint32))
    pos, vel, ids = po.init_particles_sd(g["D0"], g["D20"], g["dDdy0"], g["dD2dy0"], N, box, 1)
    assert np.array_equal(ids, g["id0"])
    assert np.array_equal(pos.view(np.uint32), g["pos0"].view(np.uint32))
    assert np.array_equal(vel, g["vel0"])


def test_oracle_sd_first_step(g):
    """Step 1 taken apart: the four per-step fields, Kick and Drift (bit-exact floats)."""
    N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
    A, AI, AF, AFF, dda, dyyy = g["steps"][0]
    names = ["dDdy_s0", "dD2dy_s0", "D_s0", "D2_s0"]                  # G_steps order: deltaD 1/2, ddDddy 1/2
    for idx, order in enumerate([1, 2, 1, 2]):
        za = po.sd_displacement_field(g["delta1"] if order == 1 else g["delta2"], g["G_steps"][0][idx], order, N, N, box)
        assert np.array_equal(za.astype(np.float32).view(np.uint32), g[names[idx]].view(np.uint32))
    v, d, sv = po.kick_sd(g["vel_s0"], g["disp_s0"], g["D_s0"], g["D2_s0"], g["sumDxyz_s0"], om, 1, A, dda)
    assert np.array_equal(v.view(np.uint32), g["vel_k0"].view(np.uint32))
    assert np.allclose(sv, g["sumxyz_k0"], rtol=0, atol=1e-14)
    p = po.drift_sd(g["pos_s0"], v, g["dDdy_s0"], g["dD2dy_s0"], g["sumxyz_k0"], box, 1, dyyy)
    assert np.array_equal(p.view(np.uint32), g["pos_d0"].view(np.uint32))


def test_oracle_sd_run(g):
    """Three full scale-dependent f(R) steps from the reference's ICs."""
    N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    d = [g["delta1"], g["delta2"]]
    for it, (A, AI, AF, AFF, dda, dyyy) in enumerate(g["steps"]):
        pc, c, m2 = po.fofr_scalars(A, om, box, float(g["fofr0"]), float(g["nfofr"]))
        out = po.get_displacements(pos, N, N, box, model="fofr", mg=dict(omega=om, a=A, phi_crit=pc, coupling=c, massterm2=m2), pofk=pk_cfg(g))
        f = [po.sd_displacement_field(d[i % 2], g["G_steps"][it][i], 1 + i % 2, N, N, box).astype(np.float32) for i in range(4)]
        vel, _, sv = po.kick_sd(vel, out["disp"], f[2], f[3], out["sumDxyz"], om, 1, A, dda)
        pos = po.drift_sd(pos, vel, f[0], f[1], sv, box, 1, dyyy)
        s = g["pofk_sums"][it]
        assert np.array_equal(out["pofk"][2], s[1])
    dp = np.abs(pos.astype(np.float64) - g["pos1"])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < 2e-5 * box / N
    assert np.abs(vel - g["vel1"]).max() < 1e-5 * np.abs(g["vel1"]).max()


def test_oracle_sd_vs_compiled_reference():
    """Same checks against oracle/_ref executed here (other mesh / sample sizes than the fixture, incl.
    Nsample != Nmesh so that the trilinear read-out weights matter)."""
    from oracle import ref_lib
    if not ref_lib.available("fofr"):
        pytest.skip("oracle/_ref/fofr not built (build needs /root/reference)")
    import tempfile
    import bench
    for (N, Ns) in ((12, 12), (12, 8)):
        wd = tempfile.mkdtemp(prefix="mgp_sd_")
        pf = bench.write_paramfile(wd, N, 40.0, "fofr", 8, lcdm_growth=0)
        txt = open(pf).read().replace("Nsample %d" % N, "Nsample %d" % Ns)
        open(pf, "w").write(txt)
        run = ref_lib.RefRun("fofr", pf)
        r = run.r
        P = run.particles()
        d = [r.sd_delta(1), r.sd_delta(2)]
        for ft, nm in ((0, ("D", "D2")), (1, ("dDdy", "dD2dy"))):
            for order in (1, 2):
                tab = r.sd_growth_table(ft, order, run.A)
                za = po.sd_displacement_field(d[order - 1], tab, order, N, Ns, 40.0).astype(np.float32)
                ref = P[nm[order - 1]]
                assert np.abs(za - ref).max() <= 2e-7 * np.abs(ref).max()     # table at k(m) vs per-mode kmag: last-bit input change
        pos, vel, ids = po.init_particles_sd(P["D"], P["D2"], P["dDdy"], P["dD2dy"], Ns, 40.0, 1)
        assert np.array_equal(pos.view(np.uint32), P["Pos"].view(np.uint32))


def test_host_sd_growth_tables_match_reference(g):
    """mgpicola_b200.cosmology.ScaleDependentGrowth (what bench.py uses in place of cosmo.c) against the tables the
    reference's growth_X_scaledependent produced for the fixture: 1e-7 relative (the reference's own ODE tolerance)."""
    from mgpicola_b200 import cosmology
    N, box = int(g["N"]), float(g["box"])
    sd = cosmology.ScaleDependentGrowth(cosmology.LCDM(float(g["omega"]), 9.0), box, N, "fofr", fofr0=float(g["fofr0"]), nfofr=float(g["nfofr"]))
    A0 = float(g["A0"])
    for idx, (ft, o) in enumerate([(0, 1), (0, 2), (1, 1), (1, 2)]):
        assert np.abs(sd.table(ft, o, A0)[1:] / g["G_init"][idx][1:] - 1).max() < 1e-7
    for it, (A, AI, AF, AFF, dda, dyyy) in enumerate(g["steps"]):
        for idx, (ft, o) in enumerate([(3, 1), (3, 2), (2, 1), (2, 2)]):
            assert np.abs(sd.table(ft, o, A, AFF)[1:] / g["G_steps"][it][idx][1:] - 1).max() < 1e-7
    # mg_pofk_ratio(k, 1) as folded into the fixture's IC power table
    t = np.load(os.path.join(G, "input_power_spectrum.npz"))
    assert np.all(sd.pofk_ratio_by_k2()[1:] >= 1.0) and sd.pofk_ratio_by_k2()[1:].max() < 2.0


# ----------------------------------------------------------------------------- GPU: CUDA path vs fixture

def _pm(mgp, g, gb=8, **kw):
    N, box = int(g["N"]), float(g["box"])
    pm = mgp.PM(N, N, box, omega=float(g["omega"]), model=mgp.MODEL_FOFR, include_screening=1, grid_bytes=gb,
                scale_dependent=1, **kw)
    c = g["pofk_cfg"]
    pm.set_pofk(int(c[0]), int(c[1]), int(c[2]), float(c[3]), float(c[4]))
    return pm


@pytest.mark.gpu
@pytest.mark.parametrize("gb", [8, 4])
def test_cuda_sd_init(mgp, require_gpu, g, gb):
    """mgp_assign_displacement_field x 4 + mgp_init_particles from uploaded delta_k == the reference's
    initial particles: IDs exact, fields within float32 rounding (f64 grids) / 2e-4 (f32 grids)."""
    N, box = int(g["N"]), float(g["box"])
    pm = _pm(mgp, g, gb)
    pm.upload_grid_k(mgp.GRID_SD_DELTA1, g["delta1"])
    pm.upload_grid_k(mgp.GRID_SD_DELTA2, g["delta2"])
    for idx, (ft, order) in enumerate([(0, 1), (0, 2), (1, 1), (1, 2)]):
        pm.assign_displacment_field_to_particles(ft, order, g["G_init"][idx])
    pm.init_particles(0.0, 0.0)
    got = pm.download_particles()
    dd, dd2 = pm.download_sd_fields()
    assert np.array_equal(got["id"], g["id0"])                      # Lagrangian order, exact IDs
    tol = 3e-7 if gb == 8 else 2e-4
    for have, want in ((got["D"], g["D0"]), (got["D2"], g["D20"]), (dd, g["dDdy0"]), (dd2, g["dD2dy0"])):
        assert np.abs(have - want).max() <= tol * np.abs(want).max()
    dp = np.abs(got["pos"].astype(np.float64) - g["pos0"])
    dp = np.minimum(dp, box - dp)
    assert dp.max() <= (8e-6 if gb == 8 else 5e-4)                   # ulp(60) = 3.8e-6
    assert np.array_equal(got["vel"], g["vel0"])
    pm.close()


@pytest.mark.gpu
def test_cuda_sd_kick_drift_bit_exact(mgp, require_gpu, g):
    """Kick / Drift SCALEDEPENDENT branches on the reference's own inputs: bit-identical floats."""
    N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
    A, AI, AF, AFF, dda, dyyy = g["steps"][0]
    pm = _pm(mgp, g, sort_particles=0)
    pm.upload_particles(g["pos_s0"], g["vel_s0"], g["D_s0"], g["D2_s0"], g["id_s0"])
    pm.upload_sd_fields(g["dDdy_s0"], g["dD2dy_s0"])
    pm.upload_disp(g["disp_s0"])
    sv = pm.Kick(A, dda, 0.0, 0.0, sumDxyz=g["sumDxyz_s0"])
    got = pm.download_particles()
    assert np.array_equal(got["vel"].view(np.uint32), g["vel_k0"].view(np.uint32))
    assert np.allclose(sv, g["sumxyz_k0"], rtol=0, atol=1e-13)
    pm.Drift(dyyy, 0.0, 0.0, sumxyz=g["sumxyz_k0"])
    got = pm.download_particles()
    assert np.array_equal(got["pos"].view(np.uint32), g["pos_d0"].view(np.uint32))
    pm.close()


def _sd_run(mgp, g, gb, merged):
    N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
    pm = _pm(mgp, g, gb)
    pm.upload_grid_k(mgp.GRID_SD_DELTA1, g["delta1"])
    pm.upload_grid_k(mgp.GRID_SD_DELTA2, g["delta2"])
    pm.upload_particles(g["pos0"], g["vel0"], None, None, g["id0"])
    for it, (A, AI, AF, AFF, dda, dyyy) in enumerate(g["steps"]):
        pc, c, m2 = po.fofr_scalars(A, om, box, float(g["fofr0"]), float(g["nfofr"]))
        pm.GetDisplacements(pm.scalars(a=A, phi_crit=pc, coupling=c, massterm2=m2, compute_pofk=1))
        p, k, n = pm.step_power_spectrum()
        s = g["pofk_sums"][it]
        assert np.array_equal(n, s[1])
        good = n > 0
        pr = s[0][good] / n[good] * box ** 3 - (box / N) ** 3
        assert np.allclose(p[good], pr, rtol=1e-4, atol=1e-4 * (box / N) ** 3)
        T = g["G_steps"][it]
        if merged:
            pm.assign_displacement_fields_merged(mgp.FIELD_deltaD, T[0], T[1])
            pm.assign_displacement_fields_merged(mgp.FIELD_ddDddy, T[2], T[3])
        else:
            for idx, (ft, order) in enumerate([(3, 1), (3, 2), (2, 1), (2, 2)]):       # main.c:496-501
                pm.assign_displacment_field_to_particles(ft, order, T[idx])
        if it == 0 and not merged:
            got = pm.download_particles()
            dd, dd2 = pm.download_sd_fields()
            o = np.argsort(got["id"])
            ro = np.argsort(g["id_s0"])
            tol = 3e-7 if gb == 8 else 2e-4
            for have, want in ((got["D"], g["D_s0"]), (got["D2"], g["D2_s0"]), (dd, g["dDdy_s0"]), (dd2, g["dD2dy_s0"])):
                assert np.abs(have[o] - want[ro]).max() <= tol * np.abs(want).max()
        pm.Kick(A, dda, 0.0, 0.0)
        pm.Drift(dyyy, 0.0, 0.0)
    got = pm.download_particles()
    pm.close()
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("gb,merged", [(8, False), (4, False), (8, True)])
def test_cuda_sd_run(mgp, require_gpu, g, gb, merged):
    """Three full scale-dependent f(R) steps from the reference's ICs; `merged` builds D + D2 and
    dDdy + dD2dy in one pass each (6 instead of 12 inverse FFTs)."""
    N, box = int(g["N"]), float(g["box"])
    got = _sd_run(mgp, g, gb, merged)
    o = np.argsort(got["id"])
    ro = np.argsort(g["id1"])
    assert np.array_equal(got["id"][o], g["id1"][ro])
    dp = np.abs(got["pos"][o].astype(np.float64) - g["pos1"][ro])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < (2e-5 if gb == 8 else 2e-3) * box / N
    assert np.abs(got["vel"][o] - g["vel1"][ro]).max() < (1e-5 if gb == 8 else 1e-3) * np.abs(g["vel1"]).max()


@pytest.mark.gpu
def test_cuda_sd_ic_stores_reference_delta(mgp, require_gpu, g):
    """mgp_ic_generate with scale_dependent = 1 keeps delta1_k and delta2_k = -S_k exactly as the reference's
    displacement_fields() does (2LPT.c:424-426, 1336-1344); same seed, same P(k) table."""
    N = int(g["N"])
    pm = _pm(mgp, g)
    pm.ic_generate(g["power_by_k2"], seed=int(g["seed"]))
    d1 = pm.download_grid_k(mgp.GRID_SD_DELTA1)
    d2 = pm.download_grid_k(mgp.GRID_SD_DELTA2)
    assert np.abs(d1 - g["delta1"]).max() <= 1e-13 * np.abs(g["delta1"]).max()
    assert np.abs(d2 - g["delta2"]).max() <= 1e-11 * np.abs(g["delta2"]).max()
    pm.close()


@pytest.mark.gpu
def test_cuda_sd_nsample_ne_nmesh(mgp, require_gpu):
    """Nsample != Nmesh: the trilinear read-out weights and the lattice mean matter.  Checked against the oracle
    on random delta_k and a k-dependent table."""
    N, Ns, box = 16, 12, 30.0
    rng = np.random.default_rng(2)
    d = (rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1)))
    x = rng.standard_normal((N, N, N))
    d = np.fft.rfftn(x) / N ** 1.5                                   # a Hermitian-consistent spectrum
    d[0, 0, 0] = 0
    tab = 1.0 + 0.3 * np.sin(np.arange(3 * (N // 2) ** 2 + 1) * 0.1)
    pm = mgp.PM(N, Ns, box, model=mgp.MODEL_NONE, scale_dependent=1)
    pm.upload_grid_k(mgp.GRID_SD_DELTA1, d)
    pm.upload_grid_k(mgp.GRID_SD_DELTA2, d * 3.0)
    for order, slot in ((1, "D"), (2, "D2")):
        pm.assign_displacment_field_to_particles(mgp.FIELD_D, order, tab)
    got = pm.download_particles(want=("D", "D2", "id"))
    assert np.array_equal(got["id"], np.arange(Ns ** 3, dtype=np.uint64))
    for order, slot, src in ((1, "D", d), (2, "D2", d * 3.0)):
        za = po.sd_displacement_field(src, tab, order, N, Ns, box).astype(np.float32)
        assert np.abs(got[slot] - za).max() <= 3e-7 * np.abs(za).max()
    pm.close()
