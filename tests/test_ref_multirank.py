"""The multi-process MPI stand-in (standins/shim_mpi_mp.c + shim_fft_mp.c, started by oracle/mprun.py) runs the
UNMODIFIED reference driver on several ranks -- x-slabs, halo exchanges, hop-by-hop particle migration, the
request / response exchange of the SCALEDEPENDENT displacement fields.  Pinned here against the one-rank run of the same
executable: in double precision every rank count must reproduce it bit for bit (reductions add in rank order, the
slab transforms run the same 1-D kernels on the same numbers), and every rank must hold exactly the particles of its
slab (auxPM.c:151-153).  This is the CPU baseline of bench.py and the multi-rank oracle of the slab path."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mprun  # noqa: E402
from test_dropin_driver import read_gadget  # noqa: E402


def _run(tmp, variant, model, K, N, nsteps, lcdm_growth):
    import bench
    wd = os.path.join(str(tmp), "%s_%d" % (variant, K))
    pf = bench.write_paramfile(wd, N, 100.0, model, nsteps, lcdm_growth=lcdm_growth)
    rc, out, errs = mprun.run([mprun.exe_path(variant), pf], K, scratch_mb=mprun.scratch_mb_for(N), timeout=900)
    assert rc == 0, (rc, out[-1500:], errs)
    snaps = sorted(glob.glob(os.path.join(wd, "output", "bench_z0p000.*")), key=lambda s: int(s.rsplit(".", 1)[1]))
    assert len(snaps) == K                                   # one snapshot file per task (main.c:915)
    pk = {os.path.basename(f): open(f).read() for f in glob.glob(os.path.join(wd, "output", "pofk*"))}
    return [read_gadget(s) for s in snaps], pk


@pytest.mark.parametrize("variant,model,lcdm_growth,ranks", [("lcdm", "fofr", 1, (2, 4, 8)), ("dgp", "dgp", 1, (4,)),
                                                             ("fofr", "fofr", 0, (2, 4))])
def test_ranks_reproduce_the_one_rank_run(tmp_path, variant, model, lcdm_growth, ranks):
    if not mprun.available(variant) or not mprun.available("lcdm"):
        pytest.skip("oracle/_ref/*_mp not built (needs /root/reference at build time)")
    N, nsteps, box = 16, 3, 100.0
    one, pk1 = _run(tmp_path, variant, model, 1, N, nsteps, lcdm_growth)
    pos1, vel1, id1 = one[0]
    o1 = np.argsort(id1)
    assert np.array_equal(id1[o1], np.arange(N ** 3, dtype=np.uint64))
    assert len(pk1) >= nsteps
    for K in ranks:
        parts, pk = _run(tmp_path, variant, model, K, N, nsteps, lcdm_growth)
        ids = np.concatenate([p[2] for p in parts])
        pos = np.concatenate([p[0] for p in parts])
        vel = np.concatenate([p[1] for p in parts])
        o = np.argsort(ids)
        assert np.array_equal(ids[o], id1[o1])                                  # nobody lost, nobody duplicated
        assert np.array_equal(pos[o].view(np.uint32), pos1[o1].view(np.uint32))  # bit-identical
        assert np.array_equal(vel[o].view(np.uint32), vel1[o1].view(np.uint32))
        assert pk == pk1                                                        # every in-step P(k) file, to the last digit
        for r, (p, _, _) in enumerate(parts):                                   # slab ownership after the final MoveParticles
            slab = (p[:, 0].astype(np.float64) * N / box).astype(np.int64)
            assert ((slab // (N // K)) == r).all()


def test_uneven_slabs_match_the_library_slab_rule(tmp_path):
    """16 planes on 3 ranks (6 + 6 + 4, the FFTW-MPI block distribution): what each rank of the reference actually holds
    after its MoveParticles is what the library's host-side slab helpers say (mg-picola-public_b200/slab.py, the rule
    mgp_create and the migration kernels implement), and the run reproduces the one-rank run bit for bit."""
    if not mprun.available("lcdm"):
        pytest.skip("oracle/_ref/*_mp not built (needs /root/reference at build time)")
    from mgpicola_b200 import slab
    N, box, K = 16, 100.0, 3
    one, pk1 = _run(tmp_path, "lcdm", "fofr", 1, N, 3, 1)
    parts, pk = _run(tmp_path, "lcdm", "fofr", K, N, 3, 1)
    assert pk == pk1
    ids = np.concatenate([p[2] for p in parts])
    pos = np.concatenate([p[0] for p in parts])
    o, o1 = np.argsort(ids), np.argsort(one[0][2])
    assert np.array_equal(pos[o].view(np.uint32), one[0][0][o1].view(np.uint32))
    for r, (p, _, _) in enumerate(parts):
        nx, x0, _, _ = slab.layout(N, N, K, r)
        planes = (p[:, 0].astype(np.float64) * N / box).astype(np.int64)
        assert planes.min() >= x0 and planes.max() < x0 + nx
        assert (slab.owner_of(p[:, 0], N, box, K) == r).all()
    assert [slab.layout(N, N, K, r)[:2] for r in range(K)] == [(6, 0), (6, 6), (4, 12)]
