"""mgpicola_b200 -- Python host side of the B200-native COLA particle-mesh library.

The product is the C-ABI shared library built from csrc/ (include/mgpicola.h).  This module is the
ctypes binding used by the tests and bench.py; its `PM` class mirrors the reference's operator
interface for the path (the argument-less C functions of src/proto.h:45-48, 211-215 that talk
through globals) with the same names and argument meaning:

    MoveParticles / PtoMesh / ComputeFifthForce / Forces / MtoParticles / GetDisplacements
    Kick / Drift / compute_power_spectrum

There is no CPU fallback: importing works anywhere (so that CPU-only tests can check the exported
symbols), but constructing a `PM` without the built library or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmgpicola_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mgpicola.h")

MODEL_NONE, MODEL_FOFR, MODEL_DGP, MODEL_GEFF = 0, 1, 2, 3
DEPOSIT_ATOMIC, DEPOSIT_TILE, DEPOSIT_DETERMINISTIC, DEPOSIT_ROWS = 0, 1, 2, 3
GRID_DENSITY, GRID_FORCE_X, GRID_FORCE_Y, GRID_FORCE_Z, GRID_MG_ONE, GRID_MG_TWO, GRID_SD_DELTA1, GRID_SD_DELTA2 = range(8)
FIELD_D, FIELD_dDdy, FIELD_ddDddy, FIELD_deltaD = range(4)     # proto.h:155-158


class MgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mgpicola error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("nmesh", C.c_int), ("nsample", C.c_int), ("box", C.c_double), ("buffer", C.c_double),
                ("omega", C.c_double), ("use_cola", C.c_int), ("model", C.c_int), ("include_screening", C.c_int),
                ("grid_bytes", C.c_int), ("deposit_mode", C.c_int), ("sort_particles", C.c_int),
                ("scale_dependent", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int), ("nccl_unique_id", C.c_void_p)]


class IcConfig(C.Structure):
    _fields_ = [("seed", C.c_uint), ("sphere_mode", C.c_int), ("amplitude_fixed", C.c_int), ("inverted", C.c_int),
                ("power_by_k2", C.c_void_p), ("n_power", C.c_size_t), ("seedtable", C.c_void_p)]


class PofkConfig(C.Structure):
    _fields_ = [("nbins", C.c_int), ("bintype", C.c_int), ("subtract_shotnoise", C.c_int),
                ("kmin", C.c_double), ("kmax", C.c_double)]


class LightconeStep(C.Structure):
    """mgp_lightcone_step (include/mgpicola.h): the host scalars and tables of Drift_Lightcone (lightcone.c:281-347)."""
    _fields_ = [("A", C.c_double), ("AFF", C.c_double), ("dyyy", C.c_double), ("da1", C.c_double), ("da2", C.c_double),
                ("dv1", C.c_double), ("dv2", C.c_double), ("sumxyz", C.c_double * 3), ("rcomov_old", C.c_double),
                ("rcomov_new", C.c_double), ("origin", C.c_double * 3), ("boundary", C.c_double), ("lengthfac", C.c_double),
                ("velfac_times_fac", C.c_double), ("ntab", C.c_int), ("al_tab", C.c_void_p), ("da1_tab", C.c_void_p),
                ("da2_tab", C.c_void_p), ("dyyy_tab", C.c_void_p), ("nrep", C.c_int), ("rep_ijk", C.c_void_p)]


class FofConfig(C.Structure):
    """mgp_fof_config (include/mgpicola.h): what main.c:834-868 hands to MatchMaker()."""
    _fields_ = [("norm_pos", C.c_double), ("norm_vel", C.c_double), ("boxsize", C.c_double), ("dx_extra", C.c_double),
                ("b_fof", C.c_double), ("np_min", C.c_int), ("mass_part", C.c_double), ("dDdy", C.c_double), ("dD2dy", C.c_double)]


# mgp_fof_halo == FoFHalo of mm_common.h:115-128
FOF_HALO_DTYPE = np.dtype([("np", np.int32), ("m_halo", np.float32), ("x_avg", np.float32, 3), ("x_rms", np.float32, 3),
                           ("v_avg", np.float32, 3), ("v_rms", np.float32, 3), ("lam", np.float32, 3), ("b", np.float32),
                           ("c", np.float32), ("ea", np.float32, 3), ("eb", np.float32, 3), ("ec", np.float32, 3)])


class StepScalars(C.Structure):
    _fields_ = [("a", C.c_double), ("phi_crit", C.c_double), ("coupling", C.c_double), ("massterm2", C.c_double),
                ("dgp_fac0", C.c_double), ("rsmooth", C.c_double), ("geff", C.c_double), ("compute_pofk", C.c_int),
                ("nu_by_k2", C.c_void_p), ("n_nu", C.c_size_t), ("nu_cdmfac", C.c_double)]


_lib = None


def load_library(path=None):
    """dlopen the C-ABI library.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError("libmgpicola_cuda.so not built at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    L = C.CDLL(p)
    L.mgp_last_error.restype = C.c_char_p
    L.mgp_device_count.restype = C.c_int
    L.mgp_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.mgp_destroy.argtypes = [C.c_void_p]
    L.mgp_nccl_unique_id.argtypes = [C.c_void_p]
    L.mgp_get_layout.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_uint64)]
    L.mgp_kspace_layout.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
    L.mgp_debug_time_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    fp, dp, up = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    L.mgp_upload_particles.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgp_download_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgp_download_disp.argtypes = [C.c_void_p, C.c_void_p]
    L.mgp_pack_snapshot.argtypes = [C.c_void_p, C.c_double, C.c_double, dp, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgp_alloc_host.argtypes = [C.c_size_t]
    L.mgp_alloc_host.restype = C.c_void_p
    L.mgp_free_host.argtypes = [C.c_void_p]
    L.mgp_upload_disp.argtypes = [C.c_void_p, C.c_void_p]
    L.mgp_ic_generate.argtypes = [C.c_void_p, C.POINTER(IcConfig)]
    L.mgp_ic_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgp_ic_particles_begin.argtypes = [C.c_void_p]
    L.mgp_ic_particles_add.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.mgp_ic_particles_finish.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]
    L.mgp_init_particles.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
    L.mgp_seedtable.argtypes = [C.c_uint, C.c_int, C.c_void_p]
    L.mgp_ranlxd1_draw.argtypes = [C.c_ulong, C.c_long]
    L.mgp_ranlxd1_draw.restype = C.c_double
    L.mgp_move_particles.argtypes = [C.c_void_p]
    L.mgp_assign_displacement_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.mgp_assign_displacement_fields_merged.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.mgp_download_sd_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgp_upload_sd_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgp_ptomesh.argtypes = [C.c_void_p, C.POINTER(StepScalars)]
    L.mgp_compute_fifth_force.argtypes = [C.c_void_p, C.POINTER(StepScalars)]
    L.mgp_forces.argtypes = [C.c_void_p]
    L.mgp_mtoparticles.argtypes = [C.c_void_p, dp]
    L.mgp_get_displacements.argtypes = [C.c_void_p, C.POINTER(StepScalars), dp]
    L.mgp_kick.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
    L.mgp_drift.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, dp]
    L.mgp_fof_find.argtypes = [C.c_void_p, C.POINTER(FofConfig), C.POINTER(C.c_uint64)]
    L.mgp_fof_get.argtypes = [C.c_void_p, C.c_void_p]
    L.mgp_lightcone_count.argtypes = [C.c_void_p, C.POINTER(LightconeStep), C.c_void_p]
    L.mgp_drift_lightcone.argtypes = [C.c_void_p, C.POINTER(LightconeStep), C.c_uint64, C.c_void_p, C.c_void_p]
    L.mgp_set_pofk_config.argtypes = [C.c_void_p, C.POINTER(PofkConfig)]
    L.mgp_pofk_nbins.argtypes = [C.c_void_p]
    L.mgp_compute_power_spectrum.argtypes = [C.c_void_p, dp, dp, dp]
    L.mgp_get_step_power_spectrum.argtypes = [C.c_void_p, dp, dp, dp]
    L.mgp_get_step_power_spectrum_total.argtypes = [C.c_void_p, dp, dp, dp]
    L.mgp_compute_rsd_power_spectrum.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, dp, dp]
    L.mgp_simple_pofk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp]
    L.mgp_grid_local_values.argtypes = [C.c_void_p]
    L.mgp_grid_local_values.restype = C.c_size_t
    L.mgp_download_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mgp_upload_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mgp_fft_r2c.argtypes = [C.c_void_p, C.c_int]
    L.mgp_fft_c2r.argtypes = [C.c_void_p, C.c_int]
    L.mgp_launch_count.argtypes = [C.c_void_p, C.c_int]
    L.mgp_launch_count.restype = C.c_uint64
    L.mgp_phase_name.argtypes = [C.c_int]
    L.mgp_phase_name.restype = C.c_char_p
    L.mgp_phase_times_ms.argtypes = [C.c_void_p, dp, up, C.c_int]
    L.mgp_set_phase_timing.argtypes = [C.c_void_p, C.c_int]
    L.mgp_stream.argtypes = [C.c_void_p]
    L.mgp_stream.restype = C.c_void_p
    if path is None:
        _lib = L
    return L


def declared_symbols(header=HEADER_PATH):
    """Names of all functions include/mgpicola.h declares (used by the CPU-side symbol test)."""
    import re
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgp_[a-z0-9_]+)\s*\(", txt)))


def seedtable(seed, nmesh):
    L = load_library()
    out = np.zeros((nmesh, nmesh), np.uint32)
    rc = L.mgp_seedtable(seed, nmesh, out.ctypes.data)
    if rc:
        raise MgpError(rc, L.mgp_last_error().decode())
    return out


def ranlxd1_draw(seed, n):
    return load_library().mgp_ranlxd1_draw(seed, n)


def nccl_unique_id():
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.mgp_nccl_unique_id(buf)
    if rc:
        raise MgpError(rc, L.mgp_last_error().decode())
    return buf.raw


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class PM:
    """One GPU-resident particle-mesh context (one per process per GPU)."""

    def __init__(self, nmesh, nsample, box, omega=0.267, use_cola=1, model=MODEL_NONE, include_screening=0,
                 grid_bytes=8, deposit_mode=DEPOSIT_DETERMINISTIC, sort_particles=1, buffer=1.5, rank=0, nranks=1,
                 device=0, nccl_id=None, scale_dependent=0):
        self.L = load_library()
        self._idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        self.cfg = Config(nmesh, nsample, box, buffer, omega, use_cola, model, include_screening, grid_bytes,
                          deposit_mode, sort_particles, scale_dependent, rank, nranks, device,
                          C.cast(self._idbuf, C.c_void_p) if self._idbuf is not None else None)
        self.ctx = C.c_void_p()
        self._ck(self.L.mgp_create(C.byref(self.cfg), C.byref(self.ctx)))
        self.N, self.Ns, self.box = nmesh, nsample, box
        self.gdtype = np.float32 if grid_bytes == 4 else np.float64
        self.cdtype = np.complex64 if grid_bytes == 4 else np.complex128
        lay = [C.c_int() for _ in range(4)]
        npart = C.c_uint64()
        self._ck(self.L.mgp_get_layout(self.ctx, *[C.byref(x) for x in lay], C.byref(npart)))
        self.local_nx, self.local_x_start, self.local_np, self.local_p_start = [x.value for x in lay]
        kl = [C.c_int() for _ in range(3)]
        self._ck(self.L.mgp_kspace_layout(self.ctx, *[C.byref(x) for x in kl]))
        self.k_transposed, self.ky_start, self.ky_local = [x.value for x in kl]
        self.sumDxyz = np.zeros(3)
        self.sumxyz = np.zeros(3)

    def _ck(self, rc):
        if rc != 0:
            raise MgpError(rc, self.L.mgp_last_error().decode())

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.L.mgp_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def numpart(self):
        n = C.c_uint64()
        self._ck(self.L.mgp_get_layout(self.ctx, None, None, None, None, C.byref(n)))
        return n.value

    # ---- particles ----
    def upload_particles(self, pos, vel=None, D=None, D2=None, ids=None):
        pos, vel, D, D2 = _f32(pos), _f32(vel), _f32(D), _f32(D2)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        self._ck(self.L.mgp_upload_particles(self.ctx, pos.shape[0], _ptr(pos), _ptr(vel), _ptr(D), _ptr(D2), _ptr(ids)))

    def download_particles(self, want=("pos", "vel", "D", "D2", "id")):
        n = self.numpart
        out = {}
        for k in ("pos", "vel", "D", "D2"):
            out[k] = np.empty((n, 3), np.float32) if k in want else None
        out["id"] = np.empty(n, np.uint64) if "id" in want else None
        self._ck(self.L.mgp_download_particles(self.ctx, _ptr(out["pos"]), _ptr(out["vel"]), _ptr(out["D"]),
                                               _ptr(out["D2"]), _ptr(out["id"])))
        return {k: v for k, v in out.items() if v is not None}

    def upload_raw(self, pos, vel, D, D2, ids, n):
        """Host pointers (ints) to [n][3] float32 arrays (+ [n] uint64), e.g. pinned staging buffers."""
        self._ck(self.L.mgp_upload_particles(self.ctx, n, pos, vel or None, D or None, D2 or None, ids or None))

    def download_raw(self, pos, vel, D, D2, ids):
        self._ck(self.L.mgp_download_particles(self.ctx, pos or None, vel or None, D or None, D2 or None, ids or None))

    def pack_snapshot(self, lengthfac, velfac_times_fac, sumxyz=None, dDdy=0.0, dD2dy=0.0):
        """The GADGET blocks of Output() (main.c:915-997): (pos [n][3] float32, vel [n][3] float32, id [n] uint64)."""
        n = self.numpart
        pos, vel, ids = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty(n, np.uint64)
        sv = (C.c_double * 3)(*(self.sumxyz if sumxyz is None else sumxyz))
        self._ck(self.L.mgp_pack_snapshot(self.ctx, lengthfac, velfac_times_fac, sv, dDdy, dD2dy, _ptr(pos), _ptr(vel), _ptr(ids)))
        return pos, vel, ids

    def download_disp(self):
        d = np.empty((self.numpart, 3), np.float32)
        self._ck(self.L.mgp_download_disp(self.ctx, _ptr(d)))
        return d

    # ---- initial conditions ----
    def ic_generate(self, power_by_k2, seed=5001, sphere_mode=0, amplitude_fixed=0, inverted=0, seedtable=None):
        """displacement_fields(): power_by_k2[m] = P(k = 2 pi sqrt(m) / Box), m = 0 .. 3 (Nmesh/2)^2."""
        pw = np.ascontiguousarray(power_by_k2, dtype=np.float64)
        st = None if seedtable is None else np.ascontiguousarray(seedtable, dtype=np.uint32)
        ic = IcConfig(seed, sphere_mode, amplitude_fixed, inverted, pw.ctypes.data, pw.size, None if st is None else st.ctypes.data)
        self._ck(self.L.mgp_ic_generate(self.ctx, C.byref(ic)))

    def ic_from_particles(self, pos01_files, normfac, rescale_by_k2):
        """READICFROMFILE (readICfromfile.c:533-778): pos01_files = one [n][3] float32 array in [0, 1) per particle file.
        Returns how many of the particles fell into this rank's slab."""
        self._ck(self.L.mgp_ic_particles_begin(self.ctx))
        taken = C.c_uint64(0)
        for p in pos01_files:
            p = np.ascontiguousarray(p, dtype=np.float32)
            self._ck(self.L.mgp_ic_particles_add(self.ctx, _ptr(p), p.shape[0], C.byref(taken)))
        r = np.ascontiguousarray(rescale_by_k2, dtype=np.float64)
        self._ck(self.L.mgp_ic_particles_finish(self.ctx, normfac, _ptr(r), r.size))
        return taken.value

    def ic_download(self):
        n = self.local_np * self.Ns * self.Ns
        za, lpt = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        self._ck(self.L.mgp_ic_download(self.ctx, _ptr(za), _ptr(lpt)))
        return za, lpt

    def init_particles(self, Di, Di2, dDdy=0.0, dD2dy=0.0):
        self._ck(self.L.mgp_init_particles(self.ctx, Di, Di2, dDdy, dD2dy))

    def upload_disp(self, disp):
        d = _f32(disp)
        assert d.shape == (self.numpart, 3)
        self._ck(self.L.mgp_upload_disp(self.ctx, _ptr(d)))

    # ---- scale-dependent growth (2LPT.c:1758) ----
    def assign_displacment_field_to_particles(self, fieldtype, lpt_order, growth_by_k2):
        """Same name (sic) and meaning as the reference's function; the (A, AF, AFF) dependence has been
        evaluated by the caller into growth_by_k2[m], m = |d|^2."""
        g = np.ascontiguousarray(growth_by_k2, dtype=np.float64)
        self._ck(self.L.mgp_assign_displacement_field(self.ctx, fieldtype, lpt_order, _ptr(g), g.size))

    def assign_displacement_fields_merged(self, fieldtype, growth1_by_k2, growth2_by_k2):
        g1 = np.ascontiguousarray(growth1_by_k2, dtype=np.float64)
        g2 = np.ascontiguousarray(growth2_by_k2, dtype=np.float64)
        assert g1.size == g2.size
        self._ck(self.L.mgp_assign_displacement_fields_merged(self.ctx, fieldtype, _ptr(g1), _ptr(g2), g1.size))

    def download_sd_fields(self):
        n = self.numpart
        a, b = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        self._ck(self.L.mgp_download_sd_fields(self.ctx, _ptr(a), _ptr(b)))
        return a, b

    def upload_sd_fields(self, dDdy=None, dD2dy=None):
        a, b = _f32(dDdy), _f32(dD2dy)
        self._ck(self.L.mgp_upload_sd_fields(self.ctx, _ptr(a), _ptr(b)))

    def upload_grid_k(self, gid, arr_k):
        """k-space array [N][N][N/2+1] complex (all kx, ky, kz) -> this rank's k-space layout: the padded slab
        [kx][ky][kz] on a single rank, the transposed rows [ky_local][kz][kx] with slab-decomposed transforms."""
        N = self.N
        full = np.zeros(((self.local_nx + 1) * N * (N // 2 + 1)), self.cdtype)
        if self.k_transposed:
            rows = np.asarray(arr_k)[:, self.ky_start:self.ky_start + self.ky_local, :].transpose(1, 2, 0)
            full[: rows.size] = rows.reshape(-1)
        else:
            full[: self.local_nx * N * (N // 2 + 1)] = np.asarray(arr_k).reshape(-1)
        self.upload_grid(gid, full.view(self.gdtype))

    # ---- the reference's per-step functions ----
    @staticmethod
    def scalars(a=1.0, phi_crit=0.0, coupling=0.0, massterm2=0.0, dgp_fac0=0.0, rsmooth=0.0, geff=1.0, compute_pofk=0,
                nu_by_k2=None, nu_cdmfac=1.0):
        s = StepScalars(a, phi_crit, coupling, massterm2, dgp_fac0, rsmooth, geff, compute_pofk, None, 0, nu_cdmfac)
        if nu_by_k2 is not None:
            s._nu = np.ascontiguousarray(nu_by_k2, dtype=np.float64)      # kept alive with the struct
            s.nu_by_k2 = s._nu.ctypes.data
            s.n_nu = s._nu.size
        return s

    def MoveParticles(self):
        self._ck(self.L.mgp_move_particles(self.ctx))

    def PtoMesh(self, s=None):
        self._ck(self.L.mgp_ptomesh(self.ctx, C.byref(s) if s is not None else None))

    def ComputeFifthForce(self, s):
        self._ck(self.L.mgp_compute_fifth_force(self.ctx, C.byref(s)))

    def Forces(self):
        self._ck(self.L.mgp_forces(self.ctx))

    def MtoParticles(self):
        out = (C.c_double * 3)()
        self._ck(self.L.mgp_mtoparticles(self.ctx, out))
        self.sumDxyz = np.array(out[:])
        return self.sumDxyz

    def GetDisplacements(self, s=None):
        out = (C.c_double * 3)()
        self._ck(self.L.mgp_get_displacements(self.ctx, C.byref(s) if s is not None else None, out))
        self.sumDxyz = np.array(out[:])
        return self.sumDxyz

    def Kick(self, A, dda, ddDddy, ddD2ddy, sumDxyz=None):
        sd = (C.c_double * 3)(*(self.sumDxyz if sumDxyz is None else sumDxyz))
        out = (C.c_double * 3)()
        self._ck(self.L.mgp_kick(self.ctx, A, dda, ddDddy, ddD2ddy, sd, out))
        self.sumxyz = np.array(out[:])
        return self.sumxyz

    def Drift(self, dyyy, deltaD, deltaD2, sumxyz=None):
        sv = (C.c_double * 3)(*(self.sumxyz if sumxyz is None else sumxyz))
        self._ck(self.L.mgp_drift(self.ctx, dyyy, deltaD, deltaD2, sv))

    # ---- FoF halo finder (MatchMaker, mm_main.c / mm_fof.c) ----
    def MatchMaker(self, norm_pos, norm_vel, boxsize, dx_extra, b_fof, np_min, mass_part, dDdy=0.0, dD2dy=0.0):
        """The halos of this rank's slab as FoFHalo records (FOF_HALO_DTYPE), by decreasing np."""
        cfg = FofConfig(norm_pos, norm_vel, boxsize, dx_extra, b_fof, np_min, mass_part, dDdy, dD2dy)
        n = C.c_uint64()
        self._ck(self.L.mgp_fof_find(self.ctx, C.byref(cfg), C.byref(n)))
        out = np.zeros(n.value, FOF_HALO_DTYPE)
        self._ck(self.L.mgp_fof_get(self.ctx, _ptr(out) if n.value else None))
        return out

    # ---- lightcone (lightcone.c:265-474) ----
    def _lightcone_step(self, sc, reps, sumxyz):
        keep = [np.ascontiguousarray(sc[k], dtype=np.float64) for k in ("al_tab", "da1_tab", "da2_tab", "dyyy_tab")]
        reps = np.ascontiguousarray(reps, dtype=np.int32).reshape(-1, 3)
        ls = LightconeStep(sc["A"], sc["AFF"], sc["dyyy"], sc["da1"], sc["da2"], sc["dv1"], sc["dv2"],
                           (C.c_double * 3)(*(self.sumxyz if sumxyz is None else sumxyz)), sc["rcomov_old"], sc["rcomov_new"],
                           (C.c_double * 3)(*sc["origin"]), sc.get("boundary", 20.0), sc["lengthfac"], sc["velfac_times_fac"],
                           keep[0].size, keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data,
                           reps.shape[0], reps.ctypes.data if reps.size else None)
        return ls, (keep, reps)

    def lightcone_count(self, sc, reps, sumxyz=None):
        """How many particle images leave the lightcone in the step described by `sc`, per replicate of `reps`."""
        ls, keep = self._lightcone_step(sc, reps, sumxyz)
        cnt = np.zeros(max(ls.nrep, 1), np.uint64)
        self._ck(self.L.mgp_lightcone_count(self.ctx, C.byref(ls), _ptr(cnt)))
        return cnt[:ls.nrep]

    def Drift_Lightcone(self, sc, reps, sumxyz=None, cap=None, pinned=False):
        """The particle loop of Drift_Lightcone: drifts the particles and returns [rows of replicate r, float32 [count][6]].
        pinned: receive the rows in a page-locked block (mgp_alloc_host), as the C adapter does."""
        ls, keep = self._lightcone_step(sc, reps, sumxyz)
        if cap is None:
            c0 = self.lightcone_count(sc, reps, sumxyz)
            cap = int(c0.max()) if c0.size else 0
        cnt = np.zeros(max(ls.nrep, 1), np.uint64)
        shape = (max(ls.nrep, 1), max(cap, 1), 6)
        host = self.L.mgp_alloc_host(4 * shape[0] * shape[1] * shape[2]) if pinned else None
        try:
            if host:
                block = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=shape)
            else:
                block = np.empty(shape, np.float32)
            self._ck(self.L.mgp_drift_lightcone(self.ctx, C.byref(ls), cap, _ptr(block), _ptr(cnt)))
            return [block[r, :int(cnt[r])].copy() for r in range(ls.nrep)]
        finally:
            if host:
                self.L.mgp_free_host(host)

    # ---- P(k) ----
    def set_pofk(self, nbins, bintype, subtract_shotnoise, kmin, kmax):
        pc = PofkConfig(nbins, bintype, subtract_shotnoise, kmin, kmax)
        self._ck(self.L.mgp_set_pofk_config(self.ctx, C.byref(pc)))

    def _pofk_out(self, fn):
        nb = self.L.mgp_pofk_nbins(self.ctx)
        if nb < 0:
            raise MgpError(nb, "P(k) binning not configured")
        p, k, n = np.zeros(nb), np.zeros(nb), np.zeros(nb)
        dp = C.POINTER(C.c_double)
        self._ck(fn(self.ctx, p.ctypes.data_as(dp), k.ctypes.data_as(dp), n.ctypes.data_as(dp)))
        return p, k, n

    def compute_power_spectrum(self):
        return self._pofk_out(self.L.mgp_compute_power_spectrum)

    def step_power_spectrum(self):
        return self._pofk_out(self.L.mgp_get_step_power_spectrum)

    def step_power_spectrum_total(self):
        return self._pofk_out(self.L.mgp_get_step_power_spectrum_total)

    def compute_RSD_powerspectrum(self, vnorm, dDdy=0.0, dD2dy=0.0):
        """compute_RSD_powerspectrum (compute_pofk.c:403): returns dict(k, P0, P2, P4, err0, err2, err4, n, y=..., z=...)
        with the combination the reference writes to file (465-476)."""
        nb = self.L.mgp_pofk_nbins(self.ctx)
        if nb < 0:
            raise MgpError(nb, "P(k) binning not configured")
        oy, oz = np.zeros((5, nb)), np.zeros((5, nb))
        dp = C.POINTER(C.c_double)
        self._ck(self.L.mgp_compute_rsd_power_spectrum(self.ctx, vnorm, dDdy, dD2dy, oy.ctypes.data_as(dp), oz.ctypes.data_as(dp)))
        out = dict(y=oy, z=oz, n=oy[0], k=oy[1])
        for i, nm in ((2, "0"), (3, "2"), (4, "4")):
            out["P" + nm] = (oy[i] + oz[i]) / 2.0
            out["err" + nm] = np.abs(oy[i] - oz[i]) / np.sqrt(2.0)
        return out

    def simple_pofk(self, scheme="CIC", subtract_shotnoise=False, tsc_as_published=True):
        """SimplePofk/main.cpp on the context's particles: (pofk[N], nmodes[N]); k_i = (2 i + 1) pi / Box, P_i = pofk[i] Box^3."""
        p, n = np.zeros(self.N), np.zeros(self.N)
        dp = C.POINTER(C.c_double)
        self._ck(self.L.mgp_simple_pofk(self.ctx, {"NGP": 1, "CIC": 2, "TSC": 3}[scheme], int(subtract_shotnoise), int(tsc_as_published),
                                        p.ctypes.data_as(dp), n.ctypes.data_as(dp)))
        return p, n

    # ---- grids ----
    def download_grid(self, gid):
        n = self.L.mgp_grid_local_values(self.ctx)
        a = np.empty(n, self.gdtype)
        self._ck(self.L.mgp_download_grid(self.ctx, gid, _ptr(a)))
        return a.reshape(self.local_nx + 1, self.N, 2 * (self.N // 2 + 1))

    def download_grid_k(self, gid):
        """k-space view [kx][ky][kz] on a single rank; [kx][ky_local][kz] (this rank's ky rows) when the transforms
        are slab-decomposed and the library keeps k-space transposed."""
        a = self.download_grid(gid).reshape(-1).view(self.cdtype)
        N, NZ = self.N, self.N // 2 + 1
        if self.k_transposed:
            return a[: self.ky_local * NZ * N].reshape(self.ky_local, NZ, N).transpose(2, 0, 1)
        return a.reshape(self.local_nx + 1, N, NZ)[: self.local_nx]

    def upload_grid(self, gid, arr):
        a = np.ascontiguousarray(arr, dtype=self.gdtype).reshape(-1)
        assert a.size == self.L.mgp_grid_local_values(self.ctx)
        self._ck(self.L.mgp_upload_grid(self.ctx, gid, _ptr(a)))

    def fft_r2c(self, gid):
        self._ck(self.L.mgp_fft_r2c(self.ctx, gid))

    def fft_c2r(self, gid):
        self._ck(self.L.mgp_fft_c2r(self.ctx, gid))

    # ---- instrumentation ----
    def debug_time_exchange(self, which, reps=5):
        ms = C.c_float()
        self._ck(self.L.mgp_debug_time_exchange(self.ctx, which, reps, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        return int(self.L.mgp_launch_count(self.ctx, int(reset)))

    def set_phase_timing(self, on):
        self._ck(self.L.mgp_set_phase_timing(self.ctx, int(on)))

    def phase_times(self, reset=False):
        n = self.L.mgp_phase_count()
        ms = (C.c_double * n)()
        calls = (C.c_uint64 * n)()
        self._ck(self.L.mgp_phase_times_ms(self.ctx, ms, calls, int(reset)))
        return {self.L.mgp_phase_name(i).decode(): (ms[i], int(calls[i])) for i in range(n)}

    @property
    def stream(self):
        return self.L.mgp_stream(self.ctx)
