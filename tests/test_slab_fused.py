"""The slab-decomposed transform path (2-D cuFFT + x-transform fused with the exchange over peer memory,
csrc/xfft.cuh) on ONE GPU: MGP_FORCE_SLAB=1 makes a single rank run exactly what P ranks run (transposed k-space,
flag barriers, stores / loads through the peer table, which then holds only this rank).

  * the transforms themselves against numpy at the sizes and tile widths the fused kernels use;
  * the whole single-GPU parity suite (oracle, golden fixtures, scale-dependent runs, ICs, neutrinos, RSD) re-run
    under MGP_FORCE_SLAB=1, with the fused kernels and with the cuFFT 1-D plan + transpose kernels they replace.
Real multi-rank runs of the same kernels: tests/test_multi_gpu.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nonhermitian_c2r(mgp, N, gb, ck):
    pm = mgp.PM(N, N, 100.0, grid_bytes=gb)
    pm.upload_grid_k(mgp.GRID_DENSITY, ck.astype(pm.cdtype))
    pm.fft_c2r(mgp.GRID_DENSITY)
    r = pm.download_grid(mgp.GRID_DENSITY)[:N, :, :N].copy()
    pm.close()
    return r


@pytest.mark.parametrize("dma", ["1", "0"])
@pytest.mark.parametrize("N,gb", [(16, 8), (32, 8), (64, 8), (128, 8), (256, 8), (16, 4), (64, 4), (128, 4), (256, 4)])
def test_fused_transforms_match_numpy(mgp, require_gpu, monkeypatch, N, gb, dma):
    monkeypatch.setenv("MGP_FORCE_SLAB", "1")
    monkeypatch.setenv("MGP_XFFT", "1")
    monkeypatch.setenv("MGP_XFFT_DMA", dma)          # 1: copy-engine exchange around a local pack / unpack; 0: peer loads / stores
    pm = mgp.PM(N, N, 100.0, grid_bytes=gb)
    assert pm.k_transposed == 1 and pm.ky_local == N
    rng = np.random.default_rng(N + gb)
    nzp = 2 * (N // 2 + 1)
    tol = 2e-13 if gb == 8 else 2e-5
    x = rng.standard_normal((N, N, N))
    g = np.zeros((N + 1, N, nzp), pm.gdtype)
    g[:N, :, :N] = x
    pm.upload_grid(mgp.GRID_DENSITY, g)
    pm.fft_r2c(mgp.GRID_DENSITY)                      # 2-D r2c, barrier, pull + x-transform, barrier
    k = pm.download_grid_k(mgp.GRID_DENSITY)
    ref = np.fft.rfftn(x.astype(pm.gdtype).astype(np.float64))
    assert np.abs(k - ref).max() / np.abs(ref).max() < tol
    pm.fft_c2r(mgp.GRID_DENSITY)                      # barrier, x-transform + push, barrier, 2-D c2r
    r = pm.download_grid(mgp.GRID_DENSITY)[:N, :, :N]
    assert np.abs(r / N ** 3 - x).max() / np.abs(x).max() < tol
    pm.close()
    # a spectrum that is NOT Hermitian on the kz = 0 / Nyquist planes (what Forces produces)
    ck = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
    r = _nonhermitian_c2r(mgp, N, gb, ck)
    if gb == 8:
        # FFTW's definition: inverse c2c over x and y, then c2r over z (imaginary parts of the self-conjugate kz dropped)
        ref = np.fft.irfft(np.fft.ifft2(ck, axes=(0, 1)), n=N, axis=2) * N ** 3
        assert np.abs(r - ref).max() / np.abs(ref).max() < tol
    # whatever cuFFT's 2-D c2r makes of such input (single precision plans differ from FFTW's definition at some sizes),
    # the fused x-transform + exchange must deliver what the 1-D plan + transpose kernels deliver
    monkeypatch.setenv("MGP_XFFT", "0")
    r0 = _nonhermitian_c2r(mgp, N, gb, ck)
    assert np.abs(r - r0).max() / np.abs(r0).max() < tol


def test_forces_fused_pipeline_matches_single_rank_plans(mgp, require_gpu, monkeypatch):
    """GetDisplacements (batched inverse transforms = the two-stream pipeline of fused kernels) in the forced-slab
    mode against the 3-D cuFFT plans on the same particles: same Disp to float32 rounding."""
    import test_gpu_parity as T
    N, box = 64, 100.0
    pos, vel, D, D2 = T.make_particles(N, box, 5, clustered=True)
    out = []
    for slab in ("0", "1"):
        monkeypatch.setenv("MGP_FORCE_SLAB", slab)
        pm = mgp.PM(N, N, box, omega=0.267, grid_bytes=8)
        assert pm.k_transposed == int(slab)
        pm.upload_particles(pos, vel, D, D2)
        s = pm.GetDisplacements()
        got = pm.download_particles(want=("id",))
        out.append((pm.download_disp()[np.argsort(got["id"])], s))
        pm.close()
    scale = np.abs(out[0][0]).max()
    assert np.abs(out[0][0] - out[1][0]).max() < 2e-7 * scale
    assert np.abs(out[0][1] - out[1][1]).max() < 1e-9 * scale


SUITE = ["tests/test_gpu_parity.py", "tests/test_golden.py", "tests/test_sd.py", "tests/test_ic.py", "tests/test_nu_rsd.py"]


@pytest.mark.parametrize("xfft,dma,select", [("1", "1", ""), ("1", "0", "fft or displacements or sd_run or reference_run or fifth"),
                                             ("0", "0", "fft or displacements or sd_run or reference_run or fifth")])
def test_single_gpu_suite_on_the_slab_path(require_gpu, xfft, dma, select):
    env = dict(os.environ, MGP_FORCE_SLAB="1", MGP_XFFT=xfft, MGP_XFFT_DMA=dma)
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + SUITE
    if select:
        cmd += ["-k", select]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]


@pytest.mark.parametrize("N,gb,knob", [(320, 8, "MGP_XFFT_MIXED"), (400, 8, "MGP_XFFT_MIXED"), (320, 4, "MGP_XFFT_MIXED"),
                                       (128, 8, "MGP_XFFT_WIDE"), (256, 8, "MGP_XFFT_WIDE"), (256, 4, "MGP_XFFT_WIDE")])
def test_mixed_radix_and_wide_instances_match_numpy(mgp, require_gpu, monkeypatch, N, gb, knob):
    """The mixed-radix (Nmesh = 320, 400) and wide-tile instances of the fused kernel (host emulation:
    tests/test_xfft_host.py; first green GPU run: profiles/r02a_experimental.log)."""
    monkeypatch.setenv("MGP_FORCE_SLAB", "1")
    monkeypatch.setenv("MGP_XFFT", "1")
    monkeypatch.setenv("MGP_XFFT_DMA", "0")
    monkeypatch.setenv(knob, "1")
    pm = mgp.PM(N, N, 100.0, grid_bytes=gb)
    rng = np.random.default_rng(N + gb)
    tol = 2e-13 if gb == 8 else 2e-5
    x = rng.standard_normal((N, N, N))
    g = np.zeros((N + 1, N, 2 * (N // 2 + 1)), pm.gdtype)
    g[:N, :, :N] = x
    pm.upload_grid(mgp.GRID_DENSITY, g)
    pm.fft_r2c(mgp.GRID_DENSITY)
    k = pm.download_grid_k(mgp.GRID_DENSITY)
    ref = np.fft.rfftn(x.astype(pm.gdtype).astype(np.float64))
    assert np.abs(k - ref).max() / np.abs(ref).max() < tol
    pm.fft_c2r(mgp.GRID_DENSITY)
    r = pm.download_grid(mgp.GRID_DENSITY)[:N, :, :N]
    assert np.abs(r / float(N) ** 3 - x).max() / np.abs(x).max() < tol
    pm.close()
