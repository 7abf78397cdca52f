"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref.

Run in the build container (needs oracle/_ref, i.e. /root/reference at build time):
    python tools/make_golden.py
The fixtures are what lets the oracle and the CUDA path be checked against the reference's own
results on machines where the reference cannot be built.  Sizes are small on purpose (16^3).

  step_<model>.npz   one force evaluation on a seeded clustered particle set: inputs (pos, vel, D, D2,
                     step scalars) and the reference's density_k, phi_k (MG), Disp, sumDxyz, P(k) sums
  kickdrift.npz      Kick + Drift on seeded inputs (bit-exact floats)
  run_fofr.npz       reference ICs (seed 5001) + 3 full COLA steps with f(R) screening: particle
                     state before / after and the per-step in-code P(k) sums
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_lib                    # noqa: E402
from oracle import pm_oracle as po            # noqa: E402
import test_oracle_vs_ref as T                # noqa: E402
import bench                                  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
OMEGA = T.OMEGA
POFK = dict(pofk_nbins=24, pofk_bintype=1, pofk_subtract_shotnoise=1, pofk_kmin=0.05, pofk_kmax=2.0)


def one_step(model, variant, N=16, box=50.0, a=0.7, seed=101, extra=None):
    pos, vel, D, D2 = T.particles(N, box, seed)
    g = dict(POFK)
    g.update(extra or {})
    mg = model != "lcdm"
    r = T.ref_setup(variant, N, box, pos, vel, D, D2, mg=mg, aexp_global=a, **g)
    r.set_str("OutputDir", "/tmp")
    r.set_str("FileBase", "golden")
    with ref_lib._silenced(True):
        r.lib.PtoMesh()
        dk = r.grid_k("density").copy()
        ref_lib.tap_reset(r)
        r.lib.compute_power_spectrum(r._keep["density"].ctypes.data, a, b"CDM")
        sums = ref_lib.tap_arrays(r)[-3:]
        phik = None
        if mg:
            r.lib.ComputeFifthForce()
            phik = r.grid_k("mgarray_two").copy()
        r.lib.Forces()
        F = np.stack([r.grid(nm)[:N, :, :N].copy() for nm in ("N11", "N12", "N13")])
        r.alloc_disp()
        r.lib.MtoParticles()
    out = dict(N=N, box=box, a=a, omega=OMEGA, pos=pos, vel=vel, D=D, D2=D2, density_k=dk, force=F, disp=r.disp(),
               sumDxyz=r.get3("sumDxyz"), pofk_sum=sums[0], pofk_n=sums[1], pofk_ksum=sums[2],
               pofk_cfg=np.array([g["pofk_nbins"], g["pofk_bintype"], g["pofk_subtract_shotnoise"], g["pofk_kmin"], g["pofk_kmax"]]))
    if phik is not None:
        out["phik"] = phik
    if extra:
        for k, v in extra.items():
            out["par_" + k] = v
    np.savez_compressed(os.path.join(OUT, "step_%s.npz" % model), **out)
    print("step_%s" % model, {k: getattr(v, "shape", v) for k, v in out.items() if k in ("density_k", "disp")})


def kickdrift(N=12, box=40.0):
    pos, vel, D, D2 = T.particles(N, box, 5)
    n = pos.shape[0]
    r = T.ref_setup("lcdm", N, box, pos, vel, D, D2)
    with ref_lib._silenced(True):
        r.init_from_paramfile(T._paramfile(N, box))
        r.set(TotNumPart=n)
        r.set_particles(pos, vel, D, D2)
    L = r.lib
    disp = (np.random.default_rng(0).standard_normal((n, 3)) * 0.3).astype(np.float32)
    bufs = r.alloc_disp()
    for a in range(3):
        bufs[a][:n] = disp[:, a]
    sumD = np.array([1e-3, -2e-3, 5e-4])
    r.set3("sumDxyz", sumD)
    AI, AF, A, AFF = 0.31, 0.33, 0.32, 0.34
    Di, Di2 = L.growth_D(A), L.growth_D2(A)
    sc = dict(A=A, dda=L.Sphi(AI, AF, A), ddDddy=L.growth_ddDddy(A), ddD2ddy=L.growth_ddD2ddy(A), dyyy=L.Sq(A, AFF, AF),
              deltaD=L.growth_D(AFF) - Di, deltaD2=L.growth_D2(AFF) - Di2)
    L.Kick(AI, AF, A, Di)
    vel1 = r.particles()["Vel"].copy()
    sumxyz = r.get3("sumxyz")
    L.Drift(A, AFF, AF, Di, Di2)
    pos1 = r.particles()["Pos"].copy()
    np.savez_compressed(os.path.join(OUT, "kickdrift.npz"), N=N, box=box, omega=OMEGA, pos=pos, vel=vel, D=D, D2=D2, disp=disp,
                        sumDxyz=sumD, vel_after=vel1, disp_after=r.disp(), sumxyz=sumxyz, pos_after=pos1, **sc)
    print("kickdrift", n)


def run_fofr(N=16, box=60.0, nsteps=3):
    pf = bench.write_paramfile("/tmp/mgp_golden_run", N, box, "fofr", 10)
    run = ref_lib.RefRun("lcdm", pf)
    run.r.set(**POFK)
    L = run.r.lib
    P0 = run.particles().copy()
    steps = []
    pk = []
    for it in range(nsteps):
        A, AI, da, Di, Di2 = run.A, run.AI, run.da, run.Di, run.Di2
        AF, AFF = A + 0.5 * da, A + da
        steps.append([A, L.Sphi(AI, AF, A), L.growth_ddDddy(A), L.growth_ddD2ddy(A), L.Sq(A, AFF, AF), L.growth_D(AFF) - Di,
                      L.growth_D2(AFF) - Di2])
        ref_lib.tap_reset(run.r)
        run.step()
        taps = [t for t in ref_lib.tap_arrays(run.r) if len(t) == POFK["pofk_nbins"]]
        pk.append(np.stack(taps[:3]))
    P1 = run.particles().copy()
    np.savez_compressed(os.path.join(OUT, "run_fofr.npz"), N=N, box=box, omega=OMEGA, fofr0=1e-5, nfofr=1.0,
                        id0=P0["ID"], pos0=P0["Pos"], vel0=P0["Vel"], D=P0["D"], D2=P0["D2"],
                        id1=P1["ID"], pos1=P1["Pos"], vel1=P1["Vel"], steps=np.array(steps), pofk_sums=np.stack(pk),
                        pofk_cfg=np.array([POFK["pofk_nbins"], POFK["pofk_bintype"], 1, POFK["pofk_kmin"], POFK["pofk_kmax"]]))
    print("run_fofr", P0.shape, np.stack(pk).shape)


if __name__ == "__main__" and len(sys.argv) == 1:
    one_step("lcdm", "lcdm")
    one_step("fofr", "lcdm", extra=dict(include_screening=1, fofr0=1e-5, nfofr=1.0))
    one_step("dgp", "dgp", a=0.8, extra=dict(include_screening=1, rcH0_DGP=1.2, Rsmooth_global=1.0))
    kickdrift()
    run_fofr()


def ic_fixture(N=16, box=60.0):
    """Reference displacement_fields() + particle init on seed 5001 (LCDM: plain PowerSpec amplitudes)."""
    pf = bench.write_paramfile("/tmp/mgp_golden_ic", N, box, "lcdm", 10)
    r = ref_lib.RefLib("lcdm")
    with ref_lib._silenced(True):
        r.init_from_paramfile(pf)
        h = N // 2
        m = np.arange(3 * h * h + 1)
        power = np.array([r.lib.PowerSpec(2 * np.pi / box * np.sqrt(float(v))) for v in m])
        ic = r.make_ic()
    P = r.particles().copy()
    np.savez_compressed(os.path.join(OUT, "ic_lcdm.npz"), N=N, box=box, seed=5001, power_by_k2=power, ZA=ic["ZA"], LPT=ic["LPT"],
                        Di=ic["Di"], Di2=ic["Di2"], pos=P["Pos"], id=P["ID"])
    print("ic_lcdm", ic["ZA"].std(0), ic["LPT"].std(0))


if __name__ == "__main__" and "--ic" in sys.argv:
    ic_fixture()


def sd_fixture(N=16, box=60.0, nsteps=3):
    """SCALEDEPENDENT f(R) build (MODEL=FOFR), use_lcdm_growth_factors = 0: the stored delta1_k / delta2_k of the
    reference's IC generator, the growth tables the adapter would pass, the four per-particle fields at
    initialisation and in the first step, and the particle state along three full steps."""
    pf = bench.write_paramfile("/tmp/mgp_golden_sd", N, box, "fofr", 10, lcdm_growth=0)
    run = ref_lib.RefRun("fofr", pf)
    r, L = run.r, run.r.lib
    r.set(**POFK)
    h = N // 2
    mm = np.arange(3 * h * h + 1)
    with ref_lib._silenced(True):
        power = np.array([L.PowerSpec(2 * np.pi / box * np.sqrt(float(v))) * (L_mg_ratio(L, 2 * np.pi / box * np.sqrt(float(v))) if v else 0.0)
                          for v in mm])
    out = dict(N=N, box=box, omega=OMEGA, fofr0=1e-5, nfofr=1.0, seed=5001, power_by_k2=power,
               delta1=r.sd_delta(1), delta2=r.sd_delta(2), A0=run.A)
    P0 = run.particles().copy()
    out.update(id0=P0["ID"], pos0=P0["Pos"], vel0=P0["Vel"], D0=P0["D"], D20=P0["D2"], dDdy0=P0["dDdy"], dD2dy0=P0["dD2dy"])
    out["G_init"] = np.stack([r.sd_growth_table(ft, o, run.A) for ft in (0, 1) for o in (1, 2)])   # [D1, D2, dD1, dD2]
    steps, tabs, pk = [], [], []
    for it in range(nsteps):
        A, AI, da = run.A, run.AI, run.da
        AF, AFF = A + 0.5 * da, A + da
        steps.append([A, AI, AF, AFF, L.Sphi(AI, AF, A), L.Sq(A, AFF, AF)])
        tabs.append(np.stack([r.sd_growth_table(ft, o, A, AFF) for ft in (3, 2) for o in (1, 2)]))   # [dD1, dD2, ddD1, ddD2]
        if it == 0:
            # first step taken apart: GetDisplacements, the four assigns, Kick, Drift
            with ref_lib._silenced(True):
                r.set(timeStep_global=0, NoutputStart_global=0, aexp_global=A)
                ref_lib.tap_reset(r)
                L.GetDisplacements()
                taps = [t for t in ref_lib.tap_arrays(r) if len(t) == POFK["pofk_nbins"]]
                out["disp_s0"] = r.ref_disp()
                out["sumDxyz_s0"] = r.get3("sumDxyz")
                for ft in (3, 2):
                    for o in (1, 2):
                        L.assign_displacment_field_to_particles(A, AF, AFF, ft, o)
                Pm = run.particles().copy()
                out.update(id_s0=Pm["ID"], pos_s0=Pm["Pos"], vel_s0=Pm["Vel"], D_s0=Pm["D"], D2_s0=Pm["D2"], dDdy_s0=Pm["dDdy"], dD2dy_s0=Pm["dD2dy"])
                L.Kick(AI, AF, A, run.Di)
                out["vel_k0"] = run.particles()["Vel"].copy()
                out["sumxyz_k0"] = r.get3("sumxyz")
                r.free_disp()
                L.Drift(A, AFF, AF, run.Di, run.Di2)
                out["pos_d0"] = run.particles()["Pos"].copy()
            run.A, run.AI = AFF, AF
            run.Di, run.Di2 = L.growth_D(run.A), L.growth_D2(run.A)
            run.istep += 1
        else:
            ref_lib.tap_reset(r)
            run.step()
            taps = [t for t in ref_lib.tap_arrays(r) if len(t) == POFK["pofk_nbins"]]
        pk.append(np.stack(taps[:3]))
    P1 = run.particles().copy()
    out.update(steps=np.array(steps), G_steps=np.stack(tabs), pofk_sums=np.stack(pk), id1=P1["ID"], pos1=P1["Pos"], vel1=P1["Vel"],
               pofk_cfg=np.array([POFK["pofk_nbins"], POFK["pofk_bintype"], 1, POFK["pofk_kmin"], POFK["pofk_kmax"]]))
    np.savez_compressed(os.path.join(OUT, "sd_fofr.npz"), **out)
    print("sd_fofr", {k: getattr(v, "shape", v) for k, v in out.items() if k in ("delta1", "G_steps", "pos1")},
          "G spread", out["G_init"][0][1:].min(), out["G_init"][0][1:].max())


def L_mg_ratio(L, k):
    import ctypes as C
    L.mg_pofk_ratio.restype = C.c_double
    L.mg_pofk_ratio.argtypes = [C.c_double, C.c_double]
    return L.mg_pofk_ratio(k, 1.0)


if __name__ == "__main__" and "--sd" in sys.argv:
    sd_fixture()


def nu_paramfile(wd, N, box):
    """fofrnu (MODEL=FOFRNU) parameter file on the reference's bundled CAMB data (camb_data/example_data_nu0.2)."""
    src = "/root/reference/camb_data/example_data_nu0.2"
    lines = open(os.path.join(src, "picola_transfer_info_nu0.2.txt")).read().split("\n")
    lines[0] = "%s 50" % src                       # the file ships with the author's absolute path
    os.makedirs(wd, exist_ok=True)
    open(os.path.join(wd, "info.txt"), "w").write("\n".join(lines))
    extra = "nu_FilenameTransferInfofile %s/info.txt\nnu_include_massive_neutrinos 1\nnu_SumMassNuEV 0.2\n" % wd
    pf = bench.write_paramfile(wd, N, box, "fofr", 10, lcdm_growth=0, extra=extra)
    txt = open(pf).read().replace("Omega 0.267", "Omega 0.3175").replace("HubbleParam 0.71", "HubbleParam 0.671")
    open(pf, "w").write(txt)
    return pf


def nu_fixture(N=16, box=200.0):
    """MASSIVE_NEUTRINOS: the reference's PtoMesh on its own ICs after two steps, with the neutrino add
    (auxPM.c:383-420) and both in-step spectra ("CDM" before, "total" after the add)."""
    import ctypes as C
    pf = nu_paramfile("/tmp/mgp_golden_nu", N, box)
    run = ref_lib.RefRun("fofrnu", pf)
    r, L = run.r, run.r.lib
    r.set(**POFK)
    run.step()
    run.step()
    A = run.A
    for nm in ("get_nu_transfer_function", "get_cdm_baryon_transfer_function"):
        f = getattr(L, nm)
        f.restype = C.c_double
        f.argtypes = [C.c_double, C.c_double]
    omega, omega_nu = C.c_double.in_dll(L, "Omega").value, C.c_double.in_dll(L, "OmegaNu").value
    kk = po.sd_k_of_m(N, box)
    tab = np.zeros(kk.size)
    for m in range(1, kk.size):
        tab[m] = omega_nu / omega * float(N * N * N) * L.get_nu_transfer_function(float(kk[m]), A) / L.get_cdm_baryon_transfer_function(float(kk[m]), 1.0)
    P = run.particles().copy()
    with ref_lib._silenced(True):
        r.set(aexp_global=A, timeStep_global=1, NoutputStart_global=0, allocate_mg_arrays=1)
        r.alloc_step_grids(mg=True)
        ref_lib.tap_reset(r)
        L.PtoMesh()
    taps = [t for t in ref_lib.tap_arrays(r) if len(t) == POFK["pofk_nbins"]]
    assert len(taps) == 6
    np.savez_compressed(os.path.join(OUT, "step_nu.npz"), N=N, box=box, a=A, omega=omega, omega_nu=omega_nu, pos=P["Pos"],
                        cdelta_cdm=r.sd_delta(1), nu_by_k2=tab, cdmfac=(omega - omega_nu) / omega, density_k=r.grid_k("density").copy(),
                        pofk_cdm=np.stack(taps[:3]), pofk_total=np.stack(taps[3:]),
                        pofk_cfg=np.array([POFK["pofk_nbins"], POFK["pofk_bintype"], 1, POFK["pofk_kmin"], POFK["pofk_kmax"]]))
    print("step_nu", omega_nu, tab[1:4], np.abs(taps[3] / np.maximum(taps[0], 1e-300) - 1).max())


def rsd_fixture(N=16, box=60.0):
    """compute_RSD_powerspectrum (compute_pofk.c:403) on the reference's own particles after three f(R) steps:
    the ten per-bin sums (P0, P2, P4, n, k for the y and the z line of sight)."""
    import ctypes as C
    pf = bench.write_paramfile("/tmp/mgp_golden_rsd", N, box, "fofr", 10)
    run = ref_lib.RefRun("lcdm", pf)
    r, L = run.r, run.r.lib
    r.set(**POFK)
    for _ in range(3):
        run.step()
    A = run.A
    P = run.particles().copy()
    L.compute_RSD_powerspectrum.argtypes = [C.c_double, C.c_int]
    ref_lib.tap_reset(r)
    with ref_lib._silenced(True):
        L.compute_RSD_powerspectrum(A, 1)
    taps = ref_lib.tap_arrays(r)
    assert len(taps) == 10
    hub = C.c_double.in_dll(L, "Hubble").value
    vnorm = (hub / A) / (100.0 * A * L.hubble(A)) * N / box                 # compute_pofk.c:291-292
    np.savez_compressed(os.path.join(OUT, "rsd_lcdm.npz"), N=N, box=box, a=A, omega=OMEGA, pos=P["Pos"], vel=P["Vel"], D=P["D"], D2=P["D2"],
                        id=P["ID"], vnorm=vnorm, dDdy=L.growth_dDdy(A), dD2dy=L.growth_dD2dy(A), sums_y=np.stack(taps[:5]),
                        sums_z=np.stack(taps[5:]),
                        pofk_cfg=np.array([POFK["pofk_nbins"], POFK["pofk_bintype"], 1, POFK["pofk_kmin"], POFK["pofk_kmax"]]))
    print("rsd_lcdm", vnorm, taps[0][:4])


if __name__ == "__main__" and "--nu" in sys.argv:
    nu_fixture()
if __name__ == "__main__" and "--rsd" in sys.argv:
    rsd_fixture()
