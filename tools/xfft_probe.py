"""Drives the slab transforms of one forced-slab context (MGP_FORCE_SLAB=1) for profiling: N^3 grid, a few r2c / c2r."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MGP_FORCE_SLAB", "1")
import mgpicola_b200 as mgp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
pm = mgp.PM(N, N, 100.0 * N / 128, grid_bytes=gb)
g = np.zeros((N + 1, N, 2 * (N // 2 + 1)), pm.gdtype)
g[:N, :, :N] = np.random.default_rng(0).standard_normal((N, N, N)).astype(pm.gdtype)
pm.upload_grid(mgp.GRID_DENSITY, g)
for name, fn in (("r2c", pm.fft_r2c), ("c2r", pm.fft_c2r)):
    fn(mgp.GRID_DENSITY)                 # warm-up (plan work areas, clocks)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(mgp.GRID_DENSITY)
    dt = (time.perf_counter() - t0) / reps
    nb = (N ** 3) * gb
    print("%s N=%d gb=%d xfft=%s tk=%s: %.3f ms per transform (%.0f GB/s at 6 passes of %d MB)" % (
        name, N, gb, os.environ.get("MGP_XFFT", "1"), os.environ.get("MGP_XFFT_TK", "auto"), dt * 1e3, 6 * nb / dt / 1e9, nb / 1e6))
pm.close()
