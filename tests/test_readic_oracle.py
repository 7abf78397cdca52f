"""READICFROMFILE (readICfromfile.c): the numpy restatement oracle/pm_oracle.py::readic_delta_k pinned to the UNMODIFIED
reference built with -DREADICFROMFILE -DSCALEDEPENDENT (oracle/_ref/libmgpicola_ref_fofr_ric.so).  The reference reads GADGET
files written here (a perturbed lattice in two files), runs ReadFilesMakeDisplacementField -> displacement_fields ->
AssignDisplacementField and keeps delta(k) in cdelta_cdm (readICfromfile.c:749-753); the restatement gets the same
particles in [0, 1) and the host scalars the adapter would hand to the library (normfac, the LCDM -> MG rescaling by
|d|^2).  The CUDA path is checked against the restatement in tests/test_readic.py."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from oracle import pm_oracle as po
from oracle import ref_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _glass(ns, seed, amp=0.3):
    rng = np.random.default_rng(seed)
    q = (np.stack(np.meshgrid(*[np.arange(ns)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / ns
    k = 2 * np.pi
    psi = amp / ns * np.stack([np.sin(k * q[:, 1]) + 0.5 * np.cos(2 * k * q[:, 2]), np.sin(k * q[:, 2] + 1.0), np.cos(k * q[:, 0])], -1)
    return np.mod(q + psi + rng.standard_normal(q.shape) * 0.02 / ns, 1.0)


def _write_gadget(path, pos_box, box):
    """The part of a GADGET-1 file readICfromfile.c:300-312, 486-527 reads: header (npart[1], BoxSize), position block."""
    n = pos_box.shape[0]
    head = struct.pack("6I6ddd2i6I2i4d", 0, n, 0, 0, 0, 0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0, 0, 0, n, 0, 0, 0, 0, 0, 2,
                       box, 0.267, 0.733, 0.71)
    head += b"\0" * (256 - len(head))
    blk = np.ascontiguousarray(pos_box, np.float32).tobytes()
    with open(path, "wb") as f:
        for b in (head, blk):
            f.write(struct.pack("i", len(b)))
            f.write(b)
            f.write(struct.pack("i", len(b)))


@pytest.mark.parametrize("N,ns,sigma8_lcdm", [(32, 16, 1), (16, 16, 0)])
def test_readic_restatement_matches_reference(tmp_path, N, ns, sigma8_lcdm):
    if not ref_lib.available("fofr_ric"):
        pytest.skip("oracle/_ref READICFROMFILE build missing (make -C oracle; needs /root/reference)")
    import bench
    box = 100.0
    wd = str(tmp_path)
    pos = _glass(ns, 3)
    pos[:5] = [[0.999999, 0.5, 0.5], [0.0, 0.0, 0.0], [0.5, 0.9999999, 0.25], [0.25, 0.5, 0.99999994], [1.0 - 1e-9, 1.0 - 1e-9, 0.3]]
    pbox = (pos * box).astype(np.float32)
    pbox[pbox >= np.float32(box)] = np.float32(box)            # up to and including the box edge: the reader wraps it
    half = len(pbox) // 3
    files = [pbox[:half], pbox[half:]]
    for i, f in enumerate(files):
        _write_gadget(os.path.join(wd, "part.%d" % i), f, box)
    tags = ("ReadParticlesFromFile 1\nNumInputParticleFiles 2\nInputParticleFileDir %s\nInputParticleFilePrefix part\n"
            "RamsesOutputNumber 1\nTypeInputParticleFiles 3\n" % wd)
    pf = bench.write_paramfile(wd, N, box, "fofr", 4, lcdm_growth=0, extra=tags)
    # WhichSpectrum 2 (Eisenstein-Hu): with ReadParticlesFromFile the reference does not read the tabulated spectrum
    # (power.c:240) and would then dereference the missing table while normalising sigma8 (power.c:246, 408)
    txt = (open(pf).read().replace("Nsample %d" % N, "Nsample %d" % ns).replace("WhichSpectrum 1", "WhichSpectrum 2")
           .replace("input_sigma8_is_for_lcdm 1", "input_sigma8_is_for_lcdm %d" % sigma8_lcdm))
    open(pf, "w").write(txt)
    R = ref_lib.RefLib("fofr_ric")
    L = R.lib
    with ref_lib._silenced(True):
        R.init_from_paramfile(pf)
        L.ReadFilesMakeDisplacementField()
    ref = R.sd_delta(1)
    # what the reader makes of the file (readICfromfile.c:507-511): float *= double, wrapped once
    nf = np.float64(1.0) / np.float64(box)
    files01 = []
    for f in files:
        u = (f.astype(np.float64) * nf).astype(np.float32)
        u = np.where(u >= np.float32(1.0), (u.astype(np.float64) - 1.0).astype(np.float32), u)
        files01.append(u)
    # host scalars of readICfromfile.c:641-643, 735-741 from the reference's own functions
    for name, args in (("growth_DLCDM", [C.c_double]), ("mg_pofk_ratio", [C.c_double, C.c_double]), ("mg_sigma8_enhancement", [C.c_double])):
        getattr(L, name).restype = C.c_double
        getattr(L, name).argtypes = args
    zi = R.get("Init_Redshift", C.c_double)
    normfac = 1.0 / float(N) ** 3 * (L.growth_DLCDM(1.0) / L.growth_DLCDM(1.0 / (1.0 + zi)))
    kk = po.sd_k_of_m(N, box)
    rescale = np.ones(kk.size)
    for m in range(1, kk.size):
        rescale[m] = np.sqrt(L.mg_pofk_ratio(float(kk[m]), 1.0))
    if not sigma8_lcdm:
        rescale /= L.mg_sigma8_enhancement(1.0)
    got = po.readic_delta_k(files01, N, ns, normfac, rescale)
    assert np.abs(ref).max() > 0 and rescale.max() > 1.0001            # the f(R) rescaling is in play
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
    if N > ns:
        assert np.count_nonzero(ref) < ref.size // 2                      # the sharp-k filter removed the modes beyond ns / 2
    else:
        _check_second_order_density(R, N, box, str(tmp_path))


def _check_second_order_density(R, N, box, tmpdir):
    """The stored second-order scalar cdelta_cdm2 = -S_k (2LPT.c:1336-1344) of the SCALEDEPENDENT -DREADICFROMFILE build from
    the reference's own cdelta_cdm through the kernels' gradient arithmetic (csrc/ic_modes.cuh on the CPU), to rounding --
    with AssignDisplacementField's convention on the Nyquist planes; with the other convention it is off by its own size."""
    lib = _build_ic_emulation(tmpdir)
    if lib is None:
        return
    NZ = N // 2 + 1
    d1, d2 = np.ascontiguousarray(R.sd_delta(1), np.complex128), R.sd_delta(2)
    lib.ic_mode_f64.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 3

    def second_order(ext):
        g = []
        for m in (1, 2):
            o = [np.zeros((N, N, NZ), np.complex128) for _ in range(3)]
            assert lib.ic_mode_f64(m, N, box, ext, d1.ctypes.data, None, 1.0, *[x.ctypes.data for x in o]) == 0
            for F in o:
                t = np.fft.ifft(np.fft.ifft(F, axis=0), axis=1) * N * N
                g.append(np.fft.irfft(t, n=N, axis=2) * N)
        S = g[0] * (g[1] + g[2]) + g[1] * g[2] - g[3] ** 2 - g[4] ** 2 - g[5] ** 2
        Sk = -np.fft.rfftn(S)
        Sk[0, 0, 0] = 0.0
        return Sk
    assert np.abs(second_order(1) - d2).max() < 1e-12 * np.abs(d2).max()
    assert np.abs(second_order(0) - d2).max() > 0.1 * np.abs(d2).max()


# ---- the displacements behind delta(k): the library's own k-space arithmetic run on the CPU against the reference's ZA / LPT

def _build_ic_emulation(tmpdir):
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return None
    so = os.path.join(tmpdir, "libic_emul.so")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", os.path.join(ROOT, "mg-picola-public_b200", "csrc"), "-o", so, os.path.join(ROOT, "tests", "host", "ic_emul.cu")],
                   check=True)
    return C.CDLL(so)


def reference_readic_displacements(wd, N, box):
    """ZA [n][3], LPT [n][3] of the unmodified reference (non-SCALEDEPENDENT -DREADICFROMFILE build) after
    ReadFilesMakeDisplacementField on the GADGET files of tests/test_zz_late_additions.py::readic_case, with the particles
    in [0, 1) and the normalisation the adapter hands to the library.  None without the build."""
    if not ref_lib.available("lcdm_ric"):
        return None
    import test_zz_late_additions as tz
    pf = tz.readic_case(wd, N, box, "lcdm_ric", "lcdm", 3)
    R = ref_lib.RefLib("lcdm_ric")
    L = R.lib
    with ref_lib._silenced(True):
        R.init_from_paramfile(pf)
        L.ReadFilesMakeDisplacementField()
    n = N ** 3
    out = []
    for name in ("ZA", "LPT"):
        p = (C.c_void_p * 3).in_dll(L, name)
        out.append(np.stack([np.ctypeslib.as_array(C.cast(p[a], C.POINTER(C.c_float)), shape=(n,)).copy() for a in range(3)], 1))
    pbox = (_glass(N, 3) * box).astype(np.float32)
    half = len(pbox) // 3
    files01 = []
    for f in (pbox[:half], pbox[half:]):
        u = (f.astype(np.float64) * (np.float64(1.0) / np.float64(box))).astype(np.float32)
        files01.append(np.where(u >= np.float32(1.0), (u.astype(np.float64) - 1.0).astype(np.float32), u))
    L.growth_DLCDM.restype = C.c_double
    L.growth_DLCDM.argtypes = [C.c_double]
    zi = R.get("Init_Redshift", C.c_double)
    normfac = 1.0 / float(N) ** 3 * (L.growth_DLCDM(1.0) / L.growth_DLCDM(1.0 / (1.0 + zi)))
    return dict(ZA=out[0], LPT=out[1], files01=files01, normfac=normfac)


def test_library_kspace_arithmetic_gives_the_reference_displacements(tmp_path):
    """ic_generate_t's sequence for external particles (csrc/ic.cu) with the kernels' arithmetic (csrc/ic_modes.cuh, run on the
    CPU) and numpy transforms: delta_k -> six gradients -> second-order source -> ZA and 2LPT displacements at the Lagrangian
    points, against the ZA / LPT arrays of the unmodified reference.  Nmesh = Nsample, so the Nyquist planes carry power and
    AssignDisplacementField's wave-vector convention there (+N/2 in psi, -N/2 in the gradients) decides the result: with the
    convention of the generated initial conditions everywhere the 2LPT field is off by several times its rms."""
    N, box = 16, 100.0
    ref = reference_readic_displacements(str(tmp_path), N, box)
    if ref is None:
        pytest.skip("oracle/_ref READICFROMFILE build missing")
    lib = _build_ic_emulation(str(tmp_path))
    if lib is None:
        pytest.skip("nvcc not available")
    NZ = N // 2 + 1
    lib.ic_mode_f64.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 3
    dk = np.ascontiguousarray(po.readic_delta_k(ref["files01"], N, N, ref["normfac"], np.ones(3 * (N // 2) ** 2 + 1)), np.complex128)

    def mode(m, src, ext):
        o = [np.zeros((N, N, NZ), np.complex128) for _ in range(3)]
        assert lib.ic_mode_f64(m, N, box, ext, src.ctypes.data, None, 1.0, *[x.ctypes.data for x in o]) == 0
        return o

    def c2r(F):                                                         # FFTW's c2r: complex over x and y, then z (unnormalised)
        t = np.fft.ifft(np.fft.ifft(F, axis=0), axis=1) * N * N
        return np.fft.irfft(t, n=N, axis=2) * N

    def displacements(ext):
        g = [c2r(x) for x in mode(1, dk, ext)] + [c2r(x) for x in mode(2, dk, ext)]      # 00 11 22, 01 02 12
        S = g[0] * (g[1] + g[2]) + g[1] * g[2] - g[3] ** 2 - g[4] ** 2 - g[5] ** 2
        Sk = np.ascontiguousarray(np.fft.rfftn(S), np.complex128)
        za = np.stack([c2r(x).reshape(-1) for x in mode(0, dk, ext)], 1)
        lpt = np.stack([c2r(x).reshape(-1) for x in mode(3, Sk, 0)], 1) * ((-3.0 / 7.0) / float(N) ** 3)
        return za - za.mean(0), lpt - lpt.mean(0)

    za, lpt = displacements(1)
    assert np.abs(za - ref["ZA"]).max() < 2e-6 * np.abs(ref["ZA"]).max()
    assert np.abs(lpt - ref["LPT"]).max() < 2e-6 * np.abs(ref["LPT"]).max()
    za0, lpt0 = displacements(0)                                        # the convention of the generated ICs is NOT the reference's here
    assert np.abs(lpt0 - ref["LPT"]).max() > 0.5 * ref["LPT"].std()
