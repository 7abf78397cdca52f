"""Seeded particle sets for the FoF halo finder (MatchMaker, mm_*.c) and the UNMODIFIED reference's answer on them
(oracle/_ref/libmgpicola_ref_lcdm_mm.so: MatchMaker() called with the reference's own struct, binary catalogue read back).
Shared by tests/test_fof.py (oracle, host emulation, CUDA path)."""
import ctypes as C
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_particles(n1d, box, seed, nblobs=48):
    """n1d^3 particles: Gaussian blobs of 12 .. 600 members (several of them across the periodic faces, edges and one
    corner of the box, one with exactly 20 members far from everything else) over a uniform background; D, D2 and Vel seeded."""
    rng = np.random.default_rng(seed)
    n = n1d ** 3
    ipd = box / n1d
    centres = rng.uniform(0, box, (nblobs, 3))
    centres[0] = [0.05 * ipd, 0.3 * box, 0.6 * box]            # across x = 0 (the slab / buffer boundary)
    centres[1] = [box - 0.05 * ipd, 0.7 * box, 0.2 * box]      # across x = L
    centres[2] = [0.4 * box, 0.02 * ipd, 0.5 * box]            # across y = 0
    centres[3] = [0.6 * box, 0.5 * box, box - 0.02 * ipd]      # across z = L
    centres[4] = [0.01 * ipd, box - 0.01 * ipd, 0.01 * ipd]    # a corner
    centres[5] = [0.5 * box + 0.01, 0.25 * box, 0.25 * box]    # across the middle (a rank boundary on 2 tasks)
    sizes = rng.integers(12, 600, nblobs)
    sizes[6] = 20
    parts = []
    for c, m in zip(centres, sizes):
        parts.append(c[None, :] + rng.standard_normal((m, 3)) * 0.12 * ipd * (m / 50.0) ** (1.0 / 3.0))
    blob = np.concatenate(parts)
    bg = rng.uniform(0, box, (n - blob.shape[0], 3))
    pos = np.mod(np.concatenate([blob, bg]), box).astype(np.float32)
    pos[pos >= np.float32(box)] = 0.0
    perm = rng.permutation(n)
    pos = pos[perm]
    vel = (rng.standard_normal((n, 3)) * 2.0).astype(np.float32)
    D = (rng.standard_normal((n, 3)) * 1.5).astype(np.float32)
    D2 = (rng.standard_normal((n, 3)) * 0.4).astype(np.float32)
    return pos, vel, D, D2


MM_TAGS = ("mm_run_matchmaker 1\nmm_output_pernode 0\nmm_output_format 2\nmm_min_npart_halo 20\nmm_linking_length 0.2\n"
           "mm_dx_extra_mpc 3.0\n")


class _MMData(C.Structure):
    """struct PicolaToMatchMakerData, vars.h:428-450."""
    _fields_ = [("output_format", C.c_int), ("output_pernode", C.c_int), ("np_min", C.c_int), ("n_part_1d", C.c_int),
                ("omega_m", C.c_double), ("omega_l", C.c_double), ("HubbleParam", C.c_double), ("boxsize", C.c_double),
                ("redshift", C.c_double), ("dx_extra", C.c_double), ("b_fof", C.c_double), ("norm_vel", C.c_double),
                ("norm_pos", C.c_double), ("mass_part", C.c_double), ("Local_p_start", C.c_int), ("P", C.c_void_p),
                ("NumPart", C.c_uint), ("FileBase", C.c_char * 500), ("OutputDir", C.c_char * 500),
                ("growth_dDdy", C.c_void_p), ("growth_dD2dy", C.c_void_p)]


def read_halo_file(path):
    """write_halos_binary, mm_snap_io.c:62-91: [256][FoFHeader][256] [bytes][FoFHalo x n][bytes]."""
    from oracle import pm_oracle as po
    with open(path, "rb") as f:
        assert struct.unpack("i", f.read(4))[0] == 256
        head = f.read(256)
        assert struct.unpack("i", f.read(4))[0] == 256
        nb = struct.unpack("i", f.read(4))[0]
        h = np.frombuffer(f.read(nb), po.FOF_HALO_DTYPE).copy()
        assert struct.unpack("i", f.read(4))[0] == nb
    n_total, n_here = struct.unpack("ll", head[:16])
    assert n_here == h.size
    return h


FOF_DEFAULTS = dict(norm_pos=1.0, norm_vel=0.37, dx_extra=3.0, b_fof=0.2, np_min=20, mass_part=7.25, dDdy=0.81, dD2dy=-0.23)


def reference_halos(workdir, pos, vel, D, D2, n1d, box, use_cola=1, **kw):
    """MatchMaker() of the unmodified reference on one task.  growth_dDdy / growth_dD2dy are the reference's own
    functions at the redshift handed over, so the caller gets the two values back to feed the other implementations."""
    from oracle import ref_lib
    if not ref_lib.available("lcdm_mm"):
        return None
    import bench
    cfg = dict(FOF_DEFAULTS, **kw)
    os.makedirs(os.path.join(workdir, "output"), exist_ok=True)
    pf = bench.write_paramfile(workdir, n1d, box, "lcdm", 4, extra=MM_TAGS)
    R = ref_lib.RefLib("lcdm_mm")
    with ref_lib._silenced(True):
        R.init_from_paramfile(pf)
    R.set(UseCOLA=use_cola)
    R.set_particles(pos, vel, D, D2)
    R.set(TotNumPart=pos.shape[0])
    z = 0.25
    A = 1.0 / (1.0 + z)
    cfg["dDdy"], cfg["dD2dy"] = R.lib.growth_dDdy(A), R.lib.growth_dD2dy(A)
    d = _MMData()
    d.output_format, d.output_pernode, d.np_min, d.n_part_1d = 2, 0, cfg["np_min"], n1d
    d.omega_m, d.omega_l, d.HubbleParam, d.boxsize, d.redshift = 0.267, 0.733, 0.71, box * cfg["norm_pos"], z
    d.dx_extra, d.b_fof, d.norm_vel, d.norm_pos, d.mass_part = cfg["dx_extra"], cfg["b_fof"], cfg["norm_vel"], cfg["norm_pos"], cfg["mass_part"]
    d.Local_p_start, d.P, d.NumPart = 0, R.P.ctypes.data, pos.shape[0]
    d.FileBase, d.OutputDir = b"fof", os.path.join(workdir, "output").encode()
    d.growth_dDdy = C.cast(R.lib.growth_dDdy, C.c_void_p).value
    d.growth_dD2dy = C.cast(R.lib.growth_dD2dy, C.c_void_p).value
    R.lib.MatchMaker.argtypes = [_MMData]
    R.lib.MatchMaker.restype = None
    with ref_lib._silenced(True):
        R.lib.MatchMaker(d)
    zint = int(z)
    path = os.path.join(workdir, "output", "matchmaker_fof_z%d.%03d.dat" % (zint, int((z - zint) * 1000)))
    return read_halo_file(path), cfg


def canonical(h):
    """Halo records in an order that does not depend on how ties in np were broken (qsort, mm_fof.c:459)."""
    return h[np.lexsort((h["x_avg"][:, 2], h["x_avg"][:, 1], h["x_avg"][:, 0], -h["np"].astype(np.int64)))]


# ---- the GPU path's own steps run on the CPU (tests/host/fof_emul.cu) ----

_EMUL = {}


def build_emulation(tmpdir):
    """nvcc-compiled shared library of tests/host/fof_emul.cu (host code only); None without nvcc."""
    import shutil
    import subprocess
    if "lib" in _EMUL:
        return _EMUL["lib"]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        _EMUL["lib"] = None
        return None
    so = os.path.join(tmpdir, "libfof_emul.so")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-gencode",
                    "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "mg-picola-public_b200", "csrc"), "-I",
                    os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "host", "fof_emul.cu")], check=True)
    _EMUL["lib"] = C.CDLL(so)
    return _EMUL["lib"]


def split_tasks(pos, vel, D, D2, ntask, nmesh, box):
    """Particles by owner slab, as MoveParticles leaves them (auxPM.c:151-153), with every task's Local_p_start."""
    ix = (pos[:, 0].astype(np.float64) * (np.float64(nmesh) / np.float64(box))).astype(np.uint32).astype(np.int64)
    ix[ix >= nmesh] = nmesh - 1
    nx = nmesh // ntask
    tasks = []
    for t in range(ntask):
        m = (ix >= t * nx) & (ix < (t + 1) * nx)
        tasks.append(dict(pos=np.ascontiguousarray(pos[m]), vel=np.ascontiguousarray(vel[m]), D=np.ascontiguousarray(D[m]),
                          D2=np.ascontiguousarray(D2[m]), local_p_start=t * nx))
    return tasks


def emulated_halos(lib, tasks, nsample, nmesh, cfg, boxsize, use_cola=1, scale_dependent=0):
    """find_halos of csrc/fof_impl.cuh on the CPU, one thread per task.  Returns the per-task FoFHalo arrays."""
    from oracle import pm_oracle as po

    class _Cfg(C.Structure):
        _fields_ = [("norm_pos", C.c_double), ("norm_vel", C.c_double), ("boxsize", C.c_double), ("dx_extra", C.c_double),
                    ("b_fof", C.c_double), ("np_min", C.c_int), ("mass_part", C.c_double), ("dDdy", C.c_double), ("dD2dy", C.c_double)]
    P = len(tasks)
    c = _Cfg(cfg["norm_pos"], cfg["norm_vel"], boxsize, cfg["dx_extra"], cfg["b_fof"], cfg["np_min"], cfg["mass_part"],
             cfg["dDdy"], cfg["dD2dy"])
    n = (C.c_long * P)(*[t["pos"].shape[0] for t in tasks])
    keep = []

    def ptrs(k):
        arrs = [np.ascontiguousarray(t[k], np.float32) for t in tasks]
        keep.extend(arrs)
        return (C.c_void_p * P)(*[a.ctypes.data for a in arrs])
    p_start = (C.c_int * P)(*[t["local_p_start"] for t in tasks])
    frac = (C.c_double * P)(*[1.0 / P] * P)
    nh = (C.c_long * P)()
    cap = 1 << 16
    out = np.zeros(cap, po.FOF_HALO_DTYPE)
    lib.fof_emul.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    rc = lib.fof_emul(P, n, ptrs("pos"), ptrs("vel"), ptrs("D"), ptrs("D2"), p_start, frac, nsample, use_cola, scale_dependent,
                      C.byref(c), nh, out.ctypes.data, cap)
    assert rc == 0
    res, at = [], 0
    for t in range(P):
        res.append(out[at:at + nh[t]].copy())
        at += nh[t]
    return res


def assert_same_halos(got, ref, exact_vectors=True):
    """Catalogues equal record by record (order inside equal np aside); eigenvectors up to sign when not exact."""
    a, b = canonical(got), canonical(ref)
    assert a.size == b.size and np.array_equal(a["np"], b["np"])
    for f in ("m_halo", "x_avg", "x_rms", "v_avg", "v_rms", "lam", "b", "c"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
    for f in ("ea", "eb", "ec"):
        if exact_vectors:
            assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
        else:
            s = np.sign((a[f].astype(np.float64) * b[f]).sum(axis=1))
            assert np.abs(a[f] - b[f] * s[:, None].astype(np.float32)).max() < 2e-5, f
