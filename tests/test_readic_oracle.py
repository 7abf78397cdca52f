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
