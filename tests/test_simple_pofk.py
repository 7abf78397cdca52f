"""mgp_simple_pofk: the reference's stand-alone estimator SimplePofk/main.cpp (NGP / CIC / TSC assignment, window
deconvolution, integer bins) on the GPU against the numpy restatement oracle/pm_oracle.py::simple_pofk."""
import numpy as np
import pytest

from oracle import pm_oracle as po
from test_gpu_parity import OMEGA, adversarial_positions, make_particles


def test_oracle_simple_pofk_properties():
    """CPU: every scheme conserves the counts (k = 0 excluded, mode counts of the full cube) and a Poisson sample shows
    the shot-noise level after deconvolution."""
    rng = np.random.default_rng(5)
    N, box = 16, 40.0
    n = N ** 3                       # the tool never divides by the mean count: its shot-noise term 1 / Npart fits Npart = N^3
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    for s in ("NGP", "CIC", "TSC"):
        p, nm = po.simple_pofk(pos, N, box, s)
        assert nm[0] == 0 and nm.sum() == sum(1 for i in range(-N // 2, N // 2) for j in range(-N // 2, N // 2) for k in range(-N // 2, N // 2)
                                              if 0 < int(np.sqrt(i * i + j * j + k * k) + 0.5) < N)
        low = p[1:4].mean()
        assert 0.5 / n < low < 2.0 / n                      # shot noise 1 / Npart at low k (the window is ~ 1 there)
    p0, _ = po.simple_pofk(pos, N, box, "CIC", subtract_shotnoise=False)
    p1, _ = po.simple_pofk(pos, N, box, "CIC", subtract_shotnoise=True)
    assert np.allclose(p0 - p1, 1.0 / n)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["NGP", "CIC", "TSC"])
@pytest.mark.parametrize("gb", [8, 4])
def test_simple_pofk_matches_oracle(mgp, require_gpu, scheme, gb):
    N, box = 32, 100.0
    pos, vel, D, D2 = make_particles(N, box, 13, clustered=True)
    pos[:8] = adversarial_positions(N, box)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=gb, sort_particles=0)
    pm.upload_particles(pos, vel, D, D2)
    p, n = pm.simple_pofk(scheme, subtract_shotnoise=True)
    pr, nr = po.simple_pofk(pos, N, box, scheme, subtract_shotnoise=True)
    assert np.array_equal(n, nr)                            # mode counts per bin: exact
    tol = 1e-10 if gb == 8 else 2e-4
    good = nr > 0
    assert np.abs(p[good] - pr[good]).max() < tol * np.abs(pr[good]).max()
    if scheme == "TSC":                                     # the textbook stencil differs from the published one
        p2, _ = pm.simple_pofk(scheme, subtract_shotnoise=True, tsc_as_published=False)
        assert np.abs(p2[good] - pr[good]).max() > 1e-4 * np.abs(pr[good]).max()
    pm.close()
