"""One Drift_Lightcone call of the UNMODIFIED reference (oracle/_ref/libmgpicola_ref_lcdm_lc.so: -DLIGHTCONE -DUNFORMATTED)
on seeded particles, with everything a checker needs beside it: the inputs, the host scalars and tables of
lightcone.c:281-347 evaluated by the reference's own functions, the rows the reference wrote per replicate file and the
particle positions it left.  Shared by tests/test_lightcone.py (oracle, host emulation, CUDA path)."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LC_TAGS = ("Origin_x 50.0\nOrigin_y 30.0\nOrigin_z 60.0\nNrep_neg_x 2\nNrep_pos_x 2\nNrep_neg_y 2\nNrep_pos_y 2\n"
           "Nrep_neg_z 2\nNrep_pos_z 2\n")


def write_lightcone_paramfile(workdir, nmesh, box, z_start, steps_before, steps_lc, z_init=None):
    """Parameter file of a lightcone run: the cone starts at z_start (first entry of the output list, main.c:627-633)
    and is followed down to z = 0."""
    import bench
    pf = bench.write_paramfile(workdir, nmesh, box, "lcdm", steps_before, extra=LC_TAGS, **({} if z_init is None else dict(z_init=z_init)))
    with open(os.path.join(workdir, "out.dat"), "w") as f:
        f.write("%g, %d\n0.0, %d\n" % (z_start, steps_before, steps_lc))
    return pf


def read_lightcone_file(path):
    """All records of one <FileBase>_lightcone.<n> file written with -DUNFORMATTED (lightcone.c:528-538): per flush
    [4][count][4][nbytes][count x 6 floats][nbytes]."""
    rows = []
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            assert struct.unpack("i", h)[0] == 4
            cnt = struct.unpack("I", f.read(4))[0]
            assert struct.unpack("i", f.read(4))[0] == 4
            nb = struct.unpack("i", f.read(4))[0]
            assert nb == 24 * cnt
            rows.append(np.frombuffer(f.read(nb), np.float32).reshape(cnt, 6))
            assert struct.unpack("i", f.read(4))[0] == nb
    return np.concatenate(rows) if rows else np.zeros((0, 6), np.float32)


def reference_case(workdir, nmesh=16, box=100.0, seed=7, z_a=0.05, z_aff=0.02):
    """Runs the reference's Drift_Lightcone once (A = 1/(1+z_a) -> AFF = 1/(1+z_aff), velocity at the mid-point) on
    nmesh^3 particles: 2LPT initial conditions made at z_a by the reference itself, plus seeded residual velocities."""
    import ctypes as C
    from oracle import ref_lib
    if not ref_lib.available("lcdm_lc"):
        return None
    os.makedirs(workdir, exist_ok=True)
    pf = write_lightcone_paramfile(workdir, nmesh, box, z_a, 1, 1, z_init=z_a)
    R = ref_lib.RefLib("lcdm_lc")
    with ref_lib._silenced(True):
        R.init_from_paramfile(pf)
        R.lib.set_lightcone()                                   # main.c:72
        ic = R.make_ic()
    n = nmesh ** 3
    rng = np.random.default_rng(seed)
    P = R.particles()
    P["Vel"][:] = (rng.standard_normal((n, 3)) * 3.0).astype(np.float32)
    sumxyz = P["Vel"].astype(np.float64).mean(axis=0)
    R.set3("sumxyz", sumxyz)
    A, AFF = 1.0 / (1.0 + z_a), 1.0 / (1.0 + z_aff)
    AF = A + 0.5 * (AFF - A)
    inputs = dict(pos=P["Pos"].copy(), vel=P["Vel"].copy(), D=P["D"].copy(), D2=P["D2"].copy(), ids=P["ID"].copy(),
                  sumxyz=sumxyz, box=box, use_cola=R.get("UseCOLA", C.c_int), nmesh=nmesh)
    sc = R.lightcone_scalars(A, AFF, AF, ic["Di"], ic["Di2"])
    sc["origin"] = np.array([R.get("Origin_x", C.c_double), R.get("Origin_y", C.c_double), R.get("Origin_z", C.c_double)])
    R.lib.Drift_Lightcone.argtypes = [C.c_double] * 5
    with ref_lib._silenced(True):
        R.lib.Drift_Lightcone(A, AFF, AF, ic["Di"], ic["Di2"])
    reps, coords = R.lightcone_replicates()
    out = os.path.join(workdir, "output")
    ref_rows = []
    for coord in coords:
        f = os.path.join(out, "bench_lightcone.%d" % coord)        # coord * NTask + ThisTask with one task
        ref_rows.append(read_lightcone_file(f) if os.path.exists(f) else np.zeros((0, 6), np.float32))
    return dict(inputs=inputs, scalars=sc, reps=reps, coords=coords, ref_rows=ref_rows, ref_pos=R.particles()["Pos"].copy(),
                omega=R.get("Omega", C.c_double))


def sort_rows(a):
    a = np.asarray(a, np.float32).reshape(-1, 6)
    return a[np.lexsort(a.T[::-1])]
