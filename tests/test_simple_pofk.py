"""mgp_simple_pofk: the reference's stand-alone estimator SimplePofk/main.cpp (NGP / CIC / TSC assignment, window
deconvolution, integer bins).
  CPU  the numpy restatement oracle/pm_oracle.py::simple_pofk against the tool itself: SimplePofk/main.cpp compiled
       UNMODIFIED (oracle/Makefile -> oracle/_ref/simplepofk_{NGP,CIC,TSC}) and run on a GADGET file;
  GPU  mgp_simple_pofk against the restatement."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import pm_oracle as po
from test_gpu_parity import OMEGA, adversarial_positions, make_particles


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_gadget_positions(path, pos, box):
    """The two blocks SimplePofk reads of a GADGET-1 file (io_gadget.h:57-75, 224-262): the 256-byte header (npart[1],
    BoxSize) and the position block."""
    n = pos.shape[0]
    head = struct.pack("6i6ddd2i6I2i4d", 0, n, 0, 0, 0, 0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0, 0, 0, n, 0, 0, 0, 0, 0, 1,
                       box, 0.267, 0.733, 0.71)
    head += b"\0" * (256 - len(head))
    blk = np.ascontiguousarray(pos, np.float32).tobytes()
    with open(path, "wb") as f:
        for b in (head, blk):
            f.write(struct.pack("i", len(b)))
            f.write(b)
            f.write(struct.pack("i", len(b)))


@pytest.mark.parametrize("scheme", ["NGP", "CIC", "TSC"])
@pytest.mark.parametrize("npart", [16 ** 3, 5000])
def test_oracle_matches_the_compiled_tool(tmp_path, scheme, npart):
    """The restatement against the tool's own output file (k, P(k) with the six digits `<<` prints), clustered particles,
    Npart equal to and different from Ngrid^3 (the tool normalises to the density contrast, main.cpp:513-533), TSC with
    the stencil as published."""
    exe = os.path.join(ROOT, "oracle", "_ref", "simplepofk_" + scheme)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/simplepofk_* missing (make -C oracle; needs /root/reference)")
    rng = np.random.default_rng(3)
    N, box = 16, 50.0
    pos = (rng.random((npart, 3)) * box).astype(np.float32)
    pos[:300] = (np.array([10.0, 20.0, 30.0]) + rng.standard_normal((300, 3)) * 1.5) % box
    pos[pos >= np.float32(box)] = 0.0
    write_gadget_positions(str(tmp_path / "snap"), pos, box)
    r = subprocess.run([exe, "snap", "pofk.txt", str(N), "1", "GADGET"], cwd=str(tmp_path), capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    tool = np.loadtxt(str(tmp_path / "pofk.txt"))
    p, nm = po.simple_pofk(pos, N, box, scheme, subtract_shotnoise=True)
    assert tool.shape == (N // 2, 2) and np.all(nm[1:N // 2 + 1] > 0)
    assert np.allclose(tool[:, 0], (2 * np.arange(1, N // 2 + 1) + 1) * np.pi / box, rtol=1e-5)       # main.cpp:441
    mine = p[1:N // 2 + 1] * box ** 3
    assert np.abs(mine - tool[:, 1]).max() < 6e-6 * np.abs(tool[:, 1]).max()
    assert np.all(np.abs(mine / tool[:, 1] - 1.0) < 2e-5)
    if scheme == "TSC":                                                 # the slip is visible at this precision
        assert np.abs(mine / tool[:, 1] - 1.0).max() < 0.1 * _textbook_tsc_deviation(pos, N, box, tool[:, 1])


def _textbook_tsc_deviation(pos, N, box, tool_p):
    """How far a textbook TSC (weights P, T, N on the planes iz-1, iz, iz+1 everywhere) lies from the tool's output."""
    X = pos.astype(np.float32).astype(np.float64) / box * N
    I = X.astype(np.int64)
    d = X - I
    I = np.where(I >= N, I - N, I)
    w = [0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2]
    grid = np.zeros((N, N, N))
    for a in range(3):
        for b in range(3):
            for c in range(3):
                np.add.at(grid, ((I[:, 0] + a - 1) % N, (I[:, 1] + b - 1) % N, (I[:, 2] + c - 1) % N), w[a][:, 0] * w[b][:, 1] * w[c][:, 2])
    grid = grid / grid.mean() - 1.0
    dk = np.fft.fftn(grid)
    kk = np.where(np.arange(N) < N // 2, np.arange(N), np.arange(N) - N)
    ii, jj, ll = np.meshgrid(kk, kk, kk, indexing="ij")
    kind = (np.sqrt((ii * ii + jj * jj + ll * ll).astype(np.float64)) + 0.5).astype(np.int64)
    sinc = lambda k: np.where(k != 0, np.sin(k * np.pi / N) / np.where(k != 0, k * np.pi / N, 1.0), 1.0)
    w3 = (sinc(ii) * sinc(jj) * sinc(ll)) ** 3
    val = np.abs(dk) ** 2 / float(N) ** 6 / (w3 * w3)
    sel = (kind < N) & (kind > 0)
    p = np.bincount(kind[sel], weights=val[sel], minlength=N) / np.maximum(np.bincount(kind[sel], minlength=N), 1) - 1.0 / pos.shape[0]
    return np.abs(p[1:N // 2 + 1] * box ** 3 / tool_p - 1.0).max()


def test_oracle_simple_pofk_properties():
    """CPU: every scheme conserves the counts (k = 0 excluded, mode counts of the full cube) and a Poisson sample shows
    the shot-noise level after deconvolution, whatever the number of particles per cell."""
    rng = np.random.default_rng(5)
    N, box = 16, 40.0
    n = 3 * N ** 3 // 2
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
