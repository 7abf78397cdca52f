"""mgp_pack_snapshot: the GADGET blocks of Output() (main.c:915-997) formed on the device, bit for bit the floats the
reference's host loops produce.  (The drop-in driver tests compare whole snapshot files written this way with the
unmodified reference's: tests/test_dropin_driver.py, tests/test_baseline_sizes.py.)"""
import numpy as np
import pytest

from test_gpu_parity import OMEGA, make_particles

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("usecola", [1, 0])
def test_snapshot_blocks_bit_exact(mgp, require_gpu, usecola):
    N, box = 24, 64.0
    pos, vel, D, D2 = make_particles(N, box, 41)
    ids = (np.arange(N ** 3, dtype=np.uint64) * 7919 + (1 << 40)) % (1 << 44)        # 64-bit IDs beyond 2^32
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8, use_cola=usecola, sort_particles=0)
    pm.upload_particles(pos, vel, D, D2, ids)
    lengthfac, vfac, dDdy, dD2dy = 0.987654321, 123.456789 * 0.31, 0.7123, -0.2345
    sumxyz = np.array([0.0123, -0.0045, 0.0301])
    p, v, i = pm.pack_snapshot(lengthfac, vfac, sumxyz, dDdy, dD2dy)
    assert np.array_equal(i, ids)                                                  # order = storage order, never sorted here
    assert np.array_equal(p.view(np.uint32), (lengthfac * pos.astype(np.float64)).astype(np.float32).view(np.uint32))
    lpt = (D.astype(np.float64) * dDdy + D2.astype(np.float64) * dD2dy) * float(usecola)
    ref = (vfac * ((vel.astype(np.float64) - sumxyz[None, :]) + lpt)).astype(np.float32)
    assert np.array_equal(v.view(np.uint32), ref.view(np.uint32))
    pm.close()


def test_snapshot_after_steps_keeps_ids_with_their_particles(mgp, require_gpu):
    """After sorted steps the blocks are in the device's storage order: the same permutation for all three."""
    N, box = 16, 50.0
    pos, vel, D, D2 = make_particles(N, box, 42, clustered=True)
    pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8, sort_particles=1, deposit_mode=mgp.DEPOSIT_ATOMIC)
    pm.upload_particles(pos, vel, D, D2)
    pm.GetDisplacements()
    pm.Kick(0.5, 0.02, 1.3, -0.4)
    pm.Drift(0.5, 0.03, -0.01)
    got = pm.download_particles()
    p, v, i = pm.pack_snapshot(1.0, 1.0, np.zeros(3), 0.0, 0.0)
    assert np.array_equal(i, got["id"])
    assert np.array_equal(p, got["pos"]) and np.array_equal(v, got["vel"])
    pm.close()


def test_pinned_host_allocation_round_trip(mgp, require_gpu):
    L = mgp.load_library()
    ptr = L.mgp_alloc_host(1 << 20)
    assert ptr
    L.mgp_free_host(ptr)
