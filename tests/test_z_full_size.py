"""The CUDA path at the size BASELINE.json quotes the metric on (Npart = Nmesh = 256^3, configs[1]) through
size-independent properties -- the oracle cannot follow to this size in seconds, these can:

  * mass conservation of the CIC deposit: the k = 0 mode of delta_k vanishes;
  * a checksum of checksums for the binned P(k): the oracle's binning (compute_pofk.c:71-236 restated) applied to the
    delta_k the GPU produced equals the GPU's own bins (mode counts exact);
  * momentum conservation of deposit -> Green's function -> gradient -> gather (same CIC kernel both ways): the mean
    displacement vanishes against its rms;
  * nobody lost, nobody duplicated: the IDs after the sort are a permutation of 0 .. N^3 - 1;
  * r2c -> c2r returns N^3 times the field.
Particles: half uniform, half in eight tight clumps (thousands per cell: the worst case for the deposit's reductions)."""
import numpy as np
import pytest

from oracle import pm_oracle as po

pytestmark = pytest.mark.gpu


def test_properties_at_baseline_size(mgp, require_gpu):
    import test_gpu_parity as T
    N, box = 256, 200.0
    pos, vel, D, D2 = T.make_particles(N, box, 2024, clustered=True)
    pm = mgp.PM(N, N, box, omega=0.267, grid_bytes=8, deposit_mode=0, sort_particles=4)
    pm.set_pofk(64, 1, 1, 0.03, 2.0)                       # paramfiles/additions_compute_pofk.txt
    pm.upload_particles(pos, vel, D, D2)
    del vel, D, D2

    # ---- deposit + r2c + in-step P(k)
    pm.MoveParticles()
    pm.PtoMesh(pm.scalars(compute_pofk=1))
    p, k, n = pm.step_power_spectrum()
    dk = np.ascontiguousarray(pm.download_grid_k(mgp.GRID_DENSITY))
    assert abs(dk[0, 0, 0]) < 1e-6                         # sum of delta over 1.7e7 cells, each O(1) .. O(1e3)
    pr, kr, nr = po.compute_power_spectrum(dk, N, N, box, 64, 1, 1, 0.03, 2.0)
    assert np.array_equal(n, nr)                           # modes per bin: exact
    good = nr > 0
    assert good.sum() > 30
    assert np.allclose(k[good], kr[good], rtol=1e-11, atol=0)
    shot = (box / N) ** 3
    rel = np.abs(p[good] - pr[good]) / (np.abs(pr[good]) + shot)
    assert rel.max() < 1e-10
    del dk

    # ---- the whole force evaluation
    sumD = pm.GetDisplacements()
    disp = pm.download_disp()
    rms = float(np.sqrt(np.mean(disp.astype(np.float64) ** 2)))
    assert rms > 0 and np.isfinite(rms)
    assert np.abs(sumD).max() < 1e-6 * rms                 # oracle at 32^3: 1e-10 of the rms
    mean = disp.astype(np.float64).mean(axis=0)
    assert np.abs(mean - sumD).max() < 1e-6 * rms          # the returned sumDxyz is the mean of Disp (auxPM.c:636-640)
    del disp
    ids = pm.download_particles(want=("id",))["id"]
    ids.sort()
    assert np.array_equal(ids, np.arange(N ** 3, dtype=np.uint64))
    del ids

    # ---- transforms: c2r(r2c(F)) = N^3 F on a force grid
    g = pm.download_grid(mgp.GRID_FORCE_X)[:N, :, :N].copy()
    pm.fft_r2c(mgp.GRID_FORCE_X)
    pm.fft_c2r(mgp.GRID_FORCE_X)
    g2 = pm.download_grid(mgp.GRID_FORCE_X)[:N, :, :N]
    assert np.abs(g2 / float(N) ** 3 - g).max() < 1e-12 * np.abs(g).max()
    pm.close()
