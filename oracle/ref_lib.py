"""ctypes driver for oracle/_ref/libmgpicola_ref_<variant>.so -- the UNMODIFIED reference sources
compiled against the stand-ins of standins (see oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  Used to (a) pin oracle/pm_oracle.py against the reference's own code,
(b) generate the committed fixtures in tests/golden/ (oracle/make_golden.py) and (c) time the
reference's CPU path for bench.py's cpu_baseline / --impl reference.

Two ways of driving the reference:
  * function level: set the globals of src/vars.h, hand it a particle array and call PtoMesh /
    Forces / MtoParticles / ComputeFifthForce / compute_power_spectrum exactly as GetDisplacements
    (auxPM.c:37-103) sequences them;
  * run level: call the reference's own set-up functions in main()'s order (main.c:68-210) on a
    parameter file, then step with its GetDisplacements / Kick / Drift; the A/AF/AFF bookkeeping of
    main.c:394-611 is inline code in main() and is restated here.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

VARIANT_FLAGS = {
    "lcdm": dict(scaledependent=False, single=False),
    "lcdm_sp": dict(scaledependent=False, single=True),
    "fofr": dict(scaledependent=True, single=False),
    "dgp": dict(scaledependent=False, single=False),
    "dgp_sd": dict(scaledependent=True, single=False),
    "fofrnu": dict(scaledependent=True, single=False),
    "lcdm_lc": dict(scaledependent=False, single=False),     # -DLIGHTCONE -DUNFORMATTED, no GADGET_STYLE
    "lcdm_mm": dict(scaledependent=False, single=False),     # -DMATCHMAKER_HALOFINDER
    "fofr_ric": dict(scaledependent=True, single=False),     # -DREADICFROMFILE (+ SCALEDEPENDENT: delta(k) is kept)
    "lcdm_ric": dict(scaledependent=False, single=False),    # -DREADICFROMFILE
    "jbd": dict(scaledependent=False, single=False),         # -DBRANSDICKE
    "mbeta": dict(scaledependent=True, single=False),        # -DMBETAMODEL -DSCALEDEPENDENT (the symmetron)
}


def lib_path(variant):
    return os.path.join(REF_DIR, "libmgpicola_ref_%s.so" % variant)


def exe_path(variant):
    return os.path.join(REF_DIR, "MG_PICOLA_%s" % variant)


def available(variant="lcdm"):
    return os.path.exists(lib_path(variant))


def part_dtype(scaledependent):
    """struct part_data with -DMEMORY_MODE -DPARTICLE_ID (vars.h:293-321)."""
    f = [("ID", np.uint64), ("Pos", np.float32, 3), ("Vel", np.float32, 3), ("D", np.float32, 3), ("D2", np.float32, 3)]
    if scaledependent:
        f += [("dDdy", np.float32, 3), ("dD2dy", np.float32, 3), ("coord_q", np.uint32), ("init_cpu_id", np.uint32)]
    return np.dtype(f, align=True)


class RefLib:
    def __init__(self, variant="lcdm"):
        if not available(variant):
            raise FileNotFoundError("%s not built (run `make -C oracle`; needs /root/reference)" % lib_path(variant))
        self.variant = variant
        self.flags = VARIANT_FLAGS[variant]
        # RTLD_LOCAL + a private copy per variant: the variants export the same global names
        self.lib = C.CDLL(lib_path(variant), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        self.fk = np.float32 if self.flags["single"] else np.float64
        self.pdt = part_dtype(self.flags["scaledependent"])
        self._keep = {}
        L = self.lib
        for name in ("growth_D", "growth_D2", "growth_dDdy", "growth_dD2dy", "growth_ddDddy", "growth_ddD2ddy",
                     "coupling_function", "mass2_of_a", "hubble", "PowerSpec"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double]
        for name in ("Sphi", "Sq"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_double, C.c_double]
        L.screening_factor_potential.restype = C.c_double
        L.screening_factor_potential.argtypes = [C.c_double, C.c_double]
        L.screening_factor_density.restype = C.c_double
        L.screening_factor_density.argtypes = [C.c_double, C.c_double]
        L.Kick.argtypes = [C.c_double] * 4
        L.Drift.argtypes = [C.c_double] * 5
        L.compute_power_spectrum.argtypes = [C.c_void_p, C.c_double, C.c_char_p]
        L.read_parameterfile.argtypes = [C.c_char_p]
        L.my_malloc.restype = C.c_void_p
        L.my_malloc.argtypes = [C.c_size_t]
        L.my_free.argtypes = [C.c_void_p]
        for name in ("my_fftw_mpi_plan_dft_r2c_3d", "my_fftw_mpi_plan_dft_c2r_3d"):
            f = getattr(L, name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
        L.my_fftw_execute.argtypes = [C.c_void_p]
        L.my_fftw_destroy_plan.argtypes = [C.c_void_p]
        if self.flags["scaledependent"]:
            for name in ("growth_D_scaledependent", "growth_dDdy_scaledependent", "growth_ddDddy_scaledependent",
                         "growth_D2_scaledependent", "growth_dD2dy_scaledependent", "growth_ddD2ddy_scaledependent"):
                f = getattr(L, name)
                f.restype = C.c_double
                f.argtypes = [C.c_double, C.c_double]
            L.assign_displacment_field_to_particles.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]

    # ---- globals ----
    def _g(self, name, ctype):
        return ctype.in_dll(self.lib, name)

    def set(self, **kw):
        types = dict(Nmesh=C.c_int, Nsample=C.c_int, UseCOLA=C.c_int, NTask=C.c_int, ThisTask=C.c_int,
                     modified_gravity_active=C.c_int, include_screening=C.c_int, allocate_mg_arrays=C.c_int,
                     use_lcdm_growth_factors=C.c_int, pofk_nbins=C.c_int, pofk_bintype=C.c_int,
                     pofk_subtract_shotnoise=C.c_int, pofk_compute_every_step=C.c_int, pofk_compute_rsd_pofk=C.c_int,
                     StdDA=C.c_int, fullT=C.c_int, timeStep_global=C.c_int, NoutputStart_global=C.c_int,
                     Box=C.c_double, Buffer=C.c_double, Omega=C.c_double, aexp_global=C.c_double, fofr0=C.c_double,
                     nfofr=C.c_double, Rsmooth_global=C.c_double, rcH0_DGP=C.c_double, pofk_kmin=C.c_double,
                     pofk_kmax=C.c_double, nLPT=C.c_double, NumPart=C.c_uint, TotNumPart=C.c_ulonglong,
                     Origin_x=C.c_double, Origin_y=C.c_double, Origin_z=C.c_double, Nrep_neg_x=C.c_int, Nrep_neg_y=C.c_int,
                     Nrep_neg_z=C.c_int, Nrep_pos_x=C.c_int, Nrep_pos_y=C.c_int, Nrep_pos_z=C.c_int)
        for k, v in kw.items():
            self._g(k, types[k]).value = v

    def get(self, name, ctype):
        return self._g(name, ctype).value

    def get3(self, name):
        a = (C.c_double * 3).in_dll(self.lib, name)
        return np.array([a[0], a[1], a[2]])

    def set3(self, name, v):
        a = (C.c_double * 3).in_dll(self.lib, name)
        for i in range(3):
            a[i] = float(v[i])

    def set_str(self, name, s, size=500):
        buf = (C.c_char * size).in_dll(self.lib, name)
        buf.value = s.encode()

    # ---- function-level set-up (no parameter file) ----
    def setup_grid(self, nmesh, nsample, box, omega=0.267, use_cola=1, buffer=1.5):
        self.set(NTask=1, ThisTask=0, Nmesh=nmesh, Nsample=nsample, Box=box, Omega=omega, UseCOLA=use_cola, Buffer=buffer)
        self.lib.initialize_ffts()      # 2LPT.c:47-113
        self.lib.initialize_parts()     # 2LPT.c:118-176
        self.N = nmesh
        self.total_size = C.c_ssize_t.in_dll(self.lib, "Total_size").value
        self.create_particle_type()

    def create_particle_type(self):
        self.lib.create_MPI_type_for_Particles(C.byref(C.c_int.in_dll(self.lib, "PartDataMPIType")))

    def set_particles(self, pos, vel=None, D=None, D2=None, ids=None, capacity=None):
        n = pos.shape[0]
        cap = capacity or int(np.ceil(n * 1.5)) + 16
        P = np.zeros(cap, dtype=self.pdt)
        P["Pos"][:n] = pos
        if vel is not None:
            P["Vel"][:n] = vel
        if D is not None:
            P["D"][:n] = D
        if D2 is not None:
            P["D2"][:n] = D2
        P["ID"][:n] = np.arange(n, dtype=np.uint64) if ids is None else ids
        self.P = P
        C.c_void_p.in_dll(self.lib, "P").value = P.ctypes.data
        self.set(NumPart=n)

    def particles(self):
        n = self.get("NumPart", C.c_uint)
        return self.P[:n]

    def _grid_alloc(self, name, cname):
        """Allocate a float_kind grid of 2*Total_size values owned by numpy and bind the reference
        globals `name` (real view) and `cname` (complex view) to it."""
        arr = np.zeros(2 * self.total_size, dtype=self.fk)
        self._keep[name] = arr
        C.c_void_p.in_dll(self.lib, name).value = arr.ctypes.data
        if cname:
            C.c_void_p.in_dll(self.lib, cname).value = arr.ctypes.data
        return arr

    def alloc_step_grids(self, mg):
        """What GetDisplacements allocates under MEMORY_MODE (auxPM.c:46-51, 61-71) -- done here so
        the grids stay inspectable between PtoMesh / ComputeFifthForce / Forces / MtoParticles."""
        N = self.N
        L = self.lib
        d = self._grid_alloc("density", "P3D")
        self._plans = {}
        self._plans["plan"] = L.my_fftw_mpi_plan_dft_r2c_3d(N, N, N, d.ctypes.data, d.ctypes.data, 0, 64)
        C.c_void_p.in_dll(L, "plan").value = self._plans["plan"]
        for nm, cn, pn in (("N11", "FN11", "p11"), ("N12", "FN12", "p12"), ("N13", "FN13", "p13")):
            g = self._grid_alloc(nm, cn)
            p = L.my_fftw_mpi_plan_dft_c2r_3d(N, N, N, g.ctypes.data, g.ctypes.data, 0, 64)
            C.c_void_p.in_dll(L, pn).value = p
        if mg:
            one = self._grid_alloc("mgarray_one", "P3D_mgarray_one")
            two = self._grid_alloc("mgarray_two", "P3D_mgarray_two")
            C.c_void_p.in_dll(L, "plan_mg_phinewton").value = L.my_fftw_mpi_plan_dft_c2r_3d(N, N, N, one.ctypes.data, one.ctypes.data, 0, 64)
            C.c_void_p.in_dll(L, "plan_mg_phik").value = L.my_fftw_mpi_plan_dft_r2c_3d(N, N, N, two.ctypes.data, two.ctypes.data, 0, 64)

    def grid(self, name):
        """Real view [Local_nx+1][N][2*(N/2+1)] of a bound grid."""
        N = self.N
        return self._keep[name].reshape(N + 1, N, 2 * (N // 2 + 1))

    def grid_k(self, name):
        N = self.N
        ck = np.complex64 if self.fk == np.float32 else np.complex128
        return self._keep[name].view(ck).reshape(N + 1, N, N // 2 + 1)[:N]

    def alloc_disp(self):
        n = self.get("NumPart", C.c_uint)
        arr = (C.c_void_p * 3).in_dll(self.lib, "Disp")
        self._disp = []
        for a in range(3):
            d = np.zeros(max(n, 1), dtype=np.float32)
            self._disp.append(d)
            arr[a] = d.ctypes.data
        return self._disp

    def disp(self):
        n = self.get("NumPart", C.c_uint)
        return np.stack([d[:n] for d in self._disp], axis=1)

    # ---- run-level driver ----
    def init_from_paramfile(self, path):
        """main.c:68-162 in order (everything before displacement_fields)."""
        L = self.lib
        self.set(NTask=1, ThisTask=0)
        L.my_fftw_mpi_init()
        L.msg_init()
        L.read_parameterfile(path.encode())
        L.read_outputs()
        L.set_units()
        use_cola = self.get("UseCOLA", C.c_int)
        if use_cola:                      # main.c:75-85
            self.set(StdDA=0)
        else:
            self.set(StdDA=2)
        if self.get("StdDA", C.c_int) == 0:
            self.set(fullT=1, nLPT=-2.5)
        L.initialize_transferfunction()
        L.initialize_powerspectrum()
        L.initialize_ffts()
        L.initialize_parts()
        self.create_particle_type()
        L.init_modified_version()
        self.N = self.get("Nmesh", C.c_int)
        self.Ns = self.get("Nsample", C.c_int)
        self.box = self.get("Box", C.c_double)
        self.total_size = C.c_ssize_t.in_dll(L, "Total_size").value

    def make_ic(self):
        """displacement_fields() + the particle initialisation loop of main.c:231-309
        (non-SCALEDEPENDENT branch; the loop is inline in main() and restated here)."""
        assert not self.flags["scaledependent"]
        L = self.lib
        Ns, box = self.Ns, self.box
        init_z = self.get("Init_Redshift", C.c_double)
        A = 1.0 / (1.0 + init_z)
        Di, Di2 = L.growth_D(A), L.growth_D2(A)
        dDdy, dD2dy = L.growth_dDdy(A), L.growth_dD2dy(A)
        L.displacement_fields()
        n = Ns ** 3
        za = (C.c_void_p * 3).in_dll(L, "ZA")
        lpt = (C.c_void_p * 3).in_dll(L, "LPT")
        ZA = np.stack([np.ctypeslib.as_array(C.cast(za[a], C.POINTER(C.c_float)), shape=(n,)).copy() for a in range(3)], 1)
        LPT = np.stack([np.ctypeslib.as_array(C.cast(lpt[a], C.POINTER(C.c_float)), shape=(n,)).copy() for a in range(3)], 1)
        for a in range(3):
            L.my_free(za[a])
            L.my_free(lpt[a])
        use_cola = self.get("UseCOLA", C.c_int)
        q = np.stack(np.meshgrid(np.arange(Ns), np.arange(Ns), np.arange(Ns), indexing="ij"), -1).reshape(-1, 3)
        if use_cola == 0:
            vel = (ZA.astype(np.float64) * dDdy + LPT.astype(np.float64) * dD2dy).astype(np.float32)   # main.c:284
        else:
            vel = np.zeros((n, 3), np.float32)
        arg = q.astype(np.float64) * (box / float(Ns)) + ZA.astype(np.float64) * Di + LPT.astype(np.float64) * Di2  # main.c:302-304
        from . import pm_oracle
        pos = pm_oracle.periodic_wrap(arg.astype(np.float32), box)
        ids = (q[:, 0].astype(np.uint64) * Ns + q[:, 1].astype(np.uint64)) * Ns + q[:, 2].astype(np.uint64)  # main.c:263
        self.set_particles(pos, vel, ZA, LPT, ids, capacity=int(np.ceil(n * self.get("Buffer", C.c_double))) + 16)
        self.set(TotNumPart=n)
        return dict(A=A, Di=Di, Di2=Di2, ZA=ZA, LPT=LPT)

    # ---- SCALEDEPENDENT builds ----
    SD_FUNCS = {1: ("growth_D_scaledependent", "growth_dDdy_scaledependent", "growth_ddDddy_scaledependent"),
                2: ("growth_D2_scaledependent", "growth_dD2dy_scaledependent", "growth_ddD2ddy_scaledependent")}

    def sd_growth_table(self, fieldtype, order, A, AFF=None):
        """What the C adapter hands to mgp_assign_displacement_field: the growth factor from_cdisp_store_to_ZA
        evaluates per mode (2LPT.c:1611-1614, without normfactor) at every integer m = |d|^2, k = 2 pi sqrt(m) / Box."""
        from . import pm_oracle
        kk = pm_oracle.sd_k_of_m(self.N, self.box)
        fD, fdD, fddD = (getattr(self.lib, n) for n in self.SD_FUNCS[order])
        out = np.zeros(kk.size)
        for m in range(1, kk.size):
            k = float(kk[m])
            if fieldtype == 0:
                out[m] = fD(k, A)
            elif fieldtype == 1:
                out[m] = fdD(k, A)
            elif fieldtype == 2:
                out[m] = fddD(k, A)
            else:
                out[m] = fD(k, AFF) - fD(k, A)
        return out

    def sd_delta(self, order):
        """cdelta_cdm / cdelta_cdm2 (vars.h:272-273) as complex [N][N][N/2+1]."""
        N = self.N
        ptr = C.c_void_p.in_dll(self.lib, "cdelta_cdm" if order == 1 else "cdelta_cdm2").value
        ck = np.complex64 if self.fk == np.float32 else np.complex128
        n = N * N * (N // 2 + 1)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double if self.fk == np.float64 else C.c_float)), shape=(2 * n,))
        return raw.view(ck).reshape(N, N, N // 2 + 1).copy()

    def sd_assign(self, A, AF, AFF, fieldtype, order):
        self.lib.assign_displacment_field_to_particles(A, AF, AFF, fieldtype, order)

    def make_ic_sd(self):
        """displacement_fields() + the particle initialisation of main.c:231-309, SCALEDEPENDENT branch (the
        loop is inline in main() and restated here; the four assign calls are the reference's own)."""
        assert self.flags["scaledependent"]
        L = self.lib
        Ns, box = self.Ns, self.box
        init_z = self.get("Init_Redshift", C.c_double)
        A = 1.0 / (1.0 + init_z)
        L.displacement_fields()
        n = Ns ** 3
        cap = int(np.ceil(n * self.get("Buffer", C.c_double))) + 16
        P = np.zeros(cap, dtype=self.pdt)
        P["coord_q"][:n] = np.arange(n, dtype=np.uint32)        # main.c:233-241
        P["init_cpu_id"][:n] = 0
        self.P = P
        C.c_void_p.in_dll(L, "P").value = P.ctypes.data
        self.set(NumPart=n, TotNumPart=n)
        for ft in (0, 1):                                       # main.c:246-251
            for order in (1, 2):
                L.assign_displacment_field_to_particles(A, A, A, ft, order)
        from . import pm_oracle
        use_cola = self.get("UseCOLA", C.c_int)
        pos, vel, ids = pm_oracle.init_particles_sd(P["D"][:n], P["D2"][:n], P["dDdy"][:n], P["dD2dy"][:n], Ns, box, use_cola)
        P["Pos"][:n] = pos
        P["Vel"][:n] = vel
        P["ID"][:n] = ids
        return dict(A=A, Di=L.growth_D(A), Di2=L.growth_D2(A))

    # ---- LIGHTCONE build ----
    def lightcone_scalars(self, A, AFF, AF, Di, Di2, ntab=1000):
        """The host part of Drift_Lightcone (lightcone.c:281-347) evaluated with the reference's own functions: what
        the C adapter hands to mgp_drift_lightcone.  StdDA = 0 (COLA)."""
        L = self.lib
        for name, nargs in (("SphiStd", 2), ("Sq", 3)):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double] * nargs
        light = self.get("Light", C.c_double)
        hub = self.get("Hubble", C.c_double)
        al = np.array([(i * (AFF - A)) / (ntab - 1.0) + A for i in range(ntab)])
        return dict(A=A, AFF=AFF, dyyy=L.Sq(A, AFF, AF), da1=L.growth_D(AFF) - Di, da2=L.growth_D2(AFF) - Di2,
                    dv1=L.growth_dDdy(AF), dv2=L.growth_dD2dy(AF),
                    rcomov_old=light / hub * L.SphiStd(A, 1.0), rcomov_new=light / hub * L.SphiStd(AFF, 1.0),
                    al_tab=al, da1_tab=np.array([L.growth_D(a) - Di for a in al]),
                    da2_tab=np.array([L.growth_D2(a) - Di2 for a in al]), dyyy_tab=np.array([L.Sq(A, a, AF) for a in al]),
                    lengthfac=self.get("UnitLength_in_cm", C.c_double) / 3.085678e24,
                    velfac_times_fac=self.get("UnitVelocity_in_cm_per_s", C.c_double) / 1.0e5 * (hub / AF), boundary=20.0)

    def lightcone_replicates(self):
        """Offsets (i, j, k) of the replicates with repflag == 0 in the order of the loops of lightcone.c:411-413, and
        their `coord` (file number = coord * NTask + ThisTask, lightcone.c:509); valid after flag_replicates."""
        g = lambda n: self.get(n, C.c_int)
        nmax = (C.c_int * 3).in_dll(self.lib, "Nrep_neg_max"), (C.c_int * 3).in_dll(self.lib, "Nrep_pos_max")
        dims = [nmax[0][a] + nmax[1][a] + 1 for a in range(3)]
        flags = C.POINTER(C.c_int).in_dll(self.lib, "repflag")
        reps, coords = [], []
        for i in range(-g("Nrep_neg_x"), g("Nrep_pos_x") + 1):
            for j in range(-g("Nrep_neg_y"), g("Nrep_pos_y") + 1):
                for k in range(-g("Nrep_neg_z"), g("Nrep_pos_z") + 1):
                    coord = ((i + nmax[0][0]) * dims[1] + (j + nmax[0][1])) * dims[2] + (k + nmax[0][2])
                    if flags[coord] == 0:
                        reps.append((i, j, k))
                        coords.append(coord)
        return np.array(reps, np.int32).reshape(-1, 3), coords

    def get_displacements(self):
        """GetDisplacements() as compiled (MEMORY_MODE: allocates and frees its own grids)."""
        self.lib.GetDisplacements()

    def free_disp(self):
        arr = (C.c_void_p * 3).in_dll(self.lib, "Disp")
        for a in range(3):
            self.lib.my_free(arr[a])

    def ref_disp(self):
        n = self.get("NumPart", C.c_uint)
        arr = (C.c_void_p * 3).in_dll(self.lib, "Disp")
        return np.stack([np.ctypeslib.as_array(C.cast(arr[a], C.POINTER(C.c_float)), shape=(n,)).copy() for a in range(3)], 1)


class RefRun:
    """Run-level driver: the reference's own set-up, IC generator and GetDisplacements / Kick / Drift
    stepped along main()'s schedule (main.c:394-611, stepDistr = 0; the bookkeeping is inline code in
    main() and is restated here).  Output steps are not taken: the
    schedule is the regular sequence of one output interval."""

    def __init__(self, variant, paramfile, quiet=True):
        self.r = RefLib(variant)
        self.quiet = quiet
        self.sd = self.r.flags["scaledependent"]
        with _silenced(quiet):
            self.r.init_from_paramfile(paramfile)
            ic = self.r.make_ic_sd() if self.sd else self.r.make_ic()
        L = self.r.lib
        self.A = ic["A"]
        self.AI = self.A
        self.Di, self.Di2 = ic["Di"], ic["Di2"]
        nout = C.c_int.in_dll(L, "Noutputs").value
        assert nout >= 1

        class _Out(C.Structure):
            _fields_ = [("Nsteps", C.c_int), ("Redshift", C.c_double)]
        ol = C.POINTER(_Out).in_dll(L, "OutputList")
        self.nsteps = ol[0].Nsteps
        ao = 1.0 / (1.0 + ol[0].Redshift)
        self.da = (ao - self.A) / float(self.nsteps)
        self.istep = 0

    def step(self):
        r, L = self.r, self.r.lib
        A, da = self.A, self.da
        AF, AFF = A + 0.5 * da, A + da
        with _silenced(self.quiet):
            r.set(timeStep_global=self.istep, NoutputStart_global=0, aexp_global=A)
            L.GetDisplacements()
            if self.sd:                                   # main.c:485-503
                for ft in (3, 2):
                    for order in (1, 2):
                        L.assign_displacment_field_to_particles(A, AF, AFF, ft, order)
            L.Kick(self.AI, AF, A, self.Di)
            r.free_disp()
            L.Drift(A, AFF, AF, self.Di, self.Di2)
        self.A, self.AI = AFF, AF
        self.Di, self.Di2 = L.growth_D(self.A), L.growth_D2(self.A)
        self.istep += 1

    def particles(self):
        return self.r.particles()


class _silenced:
    """Redirects the C library's stdout (the reference prints a lot) to /dev/null."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        if not self.on:
            return
        import sys
        sys.stdout.flush()
        C.CDLL(None).fflush(None)
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        if not self.on:
            return
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


def tap_reset(reflib):
    reflib.lib.mgp_shim_tap_reset()


def tap_arrays(reflib):
    """Payloads of the MPI_Allreduce calls (arrays of >= 4 doubles) since tap_reset, oldest first."""
    L = reflib.lib
    L.mgp_shim_tap_total.restype = C.c_long
    L.mgp_shim_tap_get.argtypes = [C.c_long, C.POINTER(C.c_double)]
    tot = L.mgp_shim_tap_total()
    out = []
    buf = (C.c_double * 8192)()
    for i in range(max(0, tot - 16), tot):
        n = L.mgp_shim_tap_get(i, buf)
        out.append(np.array(buf[:n]))
    return out
