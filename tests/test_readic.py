"""READICFROMFILE on the GPU (mgp_ic_particles_begin / _add / _finish; readICfromfile.c:133-215, 533-778) against the numpy
restatement oracle/pm_oracle.py::readic_delta_k (pinned to the reference in tests/test_readic_oracle.py).  The displacements
behind delta_k are compared with the reference's own ZA / LPT arrays in tests/test_zz_late_additions.py (an earlier check
here compared them with the library's scale-dependent read-out of the same delta_k, which holds only without power on the
Nyquist planes: there the reference's two code paths use different wave-vector conventions, csrc/ic_modes.cuh)."""
import numpy as np
import pytest

from oracle import pm_oracle as po

pytestmark = pytest.mark.gpu


def _glass(ns, seed, amp=0.3):
    """A perturbed lattice in [0, 1), split into two "files"."""
    rng = np.random.default_rng(seed)
    q = (np.stack(np.meshgrid(*[np.arange(ns)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / ns
    k = 2 * np.pi
    psi = amp / ns * np.stack([np.sin(k * q[:, 1]) + 0.5 * np.cos(2 * k * q[:, 2]), np.sin(k * q[:, 2] + 1.0), np.cos(k * q[:, 0])], -1)
    pos = np.mod(q + psi + rng.standard_normal(q.shape) * 0.02 / ns, 1.0).astype(np.float32)
    pos[pos >= 1.0] = 0.0
    h = len(pos) // 3
    return [pos[:h], pos[h:]]


@pytest.mark.parametrize("N,ns,gb", [(32, 32, 8), (32, 16, 8), (32, 32, 4)])
def test_readic_delta_k_matches_oracle(mgp, require_gpu, N, ns, gb):
    files = _glass(ns, 3)
    box = 100.0
    mmax = 3 * (N // 2) ** 2 + 1
    rescale = 1.0 + 0.1 * np.sqrt(np.arange(mmax) / mmax)            # a k-dependent LCDM -> MG rescaling
    normfac = 1.0 / N ** 3 * 7.3                                     # 1/N^3 * D(1) / D(a_init)
    pm = mgp.PM(N, ns, box, grid_bytes=gb, scale_dependent=1, model=mgp.MODEL_FOFR, include_screening=1)
    taken = pm.ic_from_particles(files, normfac, rescale)
    assert taken == ns ** 3                                          # one rank: every particle is in the slab
    dk = pm.download_grid_k(mgp.GRID_SD_DELTA1)
    ref = po.readic_delta_k(files, N, ns, normfac, rescale, grid_dtype=np.float64 if gb == 8 else np.float32)
    tol = 1e-12 if gb == 8 else 2e-5
    assert np.abs(dk - ref).max() < tol * np.abs(ref).max()
    assert dk[0, 0, 0] == 0
    if N > ns:                                                       # sharp-k filter: nothing beyond the particle Nyquist
        k1 = np.arange(N)
        d0 = np.where(k1 > N // 2, k1 - N, k1)
        a, b, c = np.meshgrid(d0, d0, np.arange(N // 2 + 1), indexing="ij")
        assert (dk[np.sqrt(a * a + b * b + c * c) > ns // 2] == 0).all()
    pm.close()
