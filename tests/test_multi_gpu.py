"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2`): the slab-decomposed path -- NCCL
particle migration, distributed FFT with hand-written pack/unpack transposes, density / force halos,
all-reduced sums and P(k) -- against the single-task oracle.  Ownership and IDs exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pm_oracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _port():
    """One rendezvous port per pytest-xdist worker: the tiny problems of this file can share the GPUs, so several tests
    may run at once (pytest -n 4)."""
    w = os.environ.get("PYTEST_XDIST_WORKER", "gw0")
    return 29517 + 7 * int("".join(ch for ch in w if ch.isdigit()) or 0)


def _run(world, tmp, nmesh, steps, model, gb, mode, extra=(), env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--nmesh", str(nmesh), "--steps", str(steps),
           "--model", model, "--gb", str(gb), "--mode", str(mode), "--out", str(tmp)] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [dict(np.load(os.path.join(tmp, "rank%d.npz" % k))) for k in range(world)]


def _oracle(nmesh, steps, model):
    import test_gpu_parity as T
    N, box, om = nmesh, 100.0, 0.267
    pos, vel, D, D2 = T.make_particles(N, box, 77, clustered=True)
    pks = []
    for it in range(steps):
        A = 0.5 + 0.1 * it
        mg = None
        if model == "fofr":
            pc, c, m2 = po.fofr_scalars(A, om, box, 1e-5, 1.0)
            mg = dict(omega=om, a=A, phi_crit=pc, coupling=c, massterm2=m2)
        elif model == "dgp":
            c, f0 = po.dgp_scalars(A, om, 1.2)
            mg = dict(coupling=c, dgp_fac0=f0, rsmooth=1.0)
        out = po.get_displacements(pos, N, N, box, model="none" if model == "lcdm" else model, mg=mg,
                                   pofk=dict(nbins=16, bintype=0, subtract_shotnoise=1, kmin_hmpc=0.0, kmax_hmpc=0.0))
        pks.append(np.stack(out["pofk"]))
        vel, _, sv = po.kick(vel, out["disp"], D, D2, out["sumDxyz"], om, 1, A, 0.02, 1.3, -0.4)
        pos = po.drift(pos, vel, D, D2, sv, box, 1, 0.5, 0.03, -0.01)
    return pos, vel, np.stack(pks)


# exchange engines of the slab transforms (fft.cu): the default (x-transform kernel that stores / loads peer memory itself;
# N = 32 is a power of two), the same kernel around a local staging buffer + copy-engine blocks over NVLink, the cuFFT 1-D
# plan + transpose kernel over peer memory, and pack / NCCL all-to-all / unpack
ENGINES = {"fused": {}, "dma": {"MGP_XFFT_DMA": "1"}, "transpose": {"MGP_XFFT": "0"}, "nccl": {"MGP_P2P": "0"}}
CASES = [(w, "fofr", 8, 0, "fused") for w in (2, 4, 8)] + [(w, "lcdm", 4, 2, "fused") for w in (2, 4, 8)] + \
        [(w, "dgp", 8, 1, "fused") for w in (2, 4, 8)] + \
        [(w, "fofr", 8, 0, e) for w in (2, 8) for e in ("dma", "transpose", "nccl")] + [(2, "lcdm", 4, 2, "dma")]


@pytest.mark.parametrize("world,model,gb,mode,engine", CASES)
def test_slab_decomposed_steps_match_oracle(require_gpu, tmp_path, world, model, gb, mode, engine):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    from mgpicola_b200 import slab
    N, steps, box = 32, 2, 100.0
    ranks = _run(world, tmp_path, N, steps, model, gb, mode, env=ENGINES[engine])
    pos, vel, pks = _oracle(N, steps, model)
    all_ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(all_ids), np.arange(N ** 3, dtype=np.uint64))          # nobody lost, nobody duplicated
    tolp = (2e-5 if gb == 8 else 2e-3) * box / N
    for k, r in enumerate(ranks):
        # ownership: exactly the reference rule applied to the positions this rank holds
        own = slab.owner_of(r["pos"][:, 0], N, box, world)
        assert (own == k).all()
        nx, x0, _, _ = slab.layout(N, N, world, k)
        assert (int(r["x0"]), int(r["nx"])) == (x0, nx)
        ids = r["id"].astype(np.int64)
        dp = np.abs(r["pos"].astype(np.float64) - pos[ids])
        dp = np.minimum(dp, box - dp)
        assert dp.max() < tolp
        assert np.abs(r["vel"] - vel[ids]).max() < (1e-5 if gb == 8 else 1e-3) * np.abs(vel).max()
        # P(k): every rank holds the all-reduced result
        for it in range(steps):
            p, kk, n = r["pks"][it]
            assert np.array_equal(n, pks[it][2])
            good = n > 0
            rel = np.abs(p[good] - pks[it][0][good]) / (np.abs(pks[it][0][good]) + (box / N) ** 3)
            assert rel.max() < (1e-10 if gb == 8 else 1e-4)
        assert int(r["launches"]) > 20


@pytest.mark.parametrize("world,mode,compact", [(2, 0, "1"), (2, 0, "0"), (2, 3, "1"), (8, 0, "1")])
def test_migration_between_sorts_matches_oracle(require_gpu, tmp_path, world, mode, compact):
    """The production particle path: sort_particles = 4 and the ATOMIC (or ROWS) deposit, five steps, so that MoveParticles
    closes the holes of the leavers by compaction (k_compact_lists / k_compact_move, MGP_COMPACT = 1) on the steps between
    two sorts -- unsorted order, shrinking counts, stale row offsets -- and, with MGP_COMPACT = 0, through a full sort
    every time.  Same oracle comparison as the sorted path: IDs a permutation, ownership, positions, velocities, P(k)."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    from mgpicola_b200 import slab
    N, steps, box = 32, 5, 100.0
    ranks = _run(world, tmp_path, N, steps, "fofr", 8, mode, extra=["--sort-interval", "4"], env={"MGP_COMPACT": compact})
    pos, vel, pks = _oracle(N, steps, "fofr")
    all_ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(all_ids), np.arange(N ** 3, dtype=np.uint64))
    for k, r in enumerate(ranks):
        assert (slab.owner_of(r["pos"][:, 0], N, box, world) == k).all()
        ids = r["id"].astype(np.int64)
        dp = np.abs(r["pos"].astype(np.float64) - pos[ids])
        dp = np.minimum(dp, box - dp)
        assert dp.max() < 5e-5 * box / N
        assert np.abs(r["vel"] - vel[ids]).max() < 2e-5 * np.abs(vel).max()
        for it in range(steps):
            p, kk, n = r["pks"][it]
            assert np.array_equal(n, pks[it][2])                     # mode counts: exact
            good = n > 0
            rel = np.abs(p[good] - pks[it][0][good]) / (np.abs(pks[it][0][good]) + (box / N) ** 3)
            assert rel.max() < 1e-9                                  # mass conservation shows here first


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_ic_matches_reference(require_gpu, tmp_path, world):
    """displacement_fields() on P slabs (transposed k-space, distributed FFTs, displacement halos) ==
    the reference's single-task result (golden fixture), every rank holding its Lagrangian planes."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ic_lcdm.npz")))
    N, box = int(g["N"]), float(g["box"])
    ranks = _run(world, tmp_path, N, 0, "lcdm", 8, 0, extra=["--ic"])
    ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(ids, g["id"])                       # rank order == Lagrangian plane order, exact IDs
    for nm, key in (("D", "ZA"), ("D2", "LPT")):
        got = np.concatenate([r[nm] for r in ranks])
        assert np.abs(got - g[key]).max() < 2e-6 * np.abs(g[key]).max(), nm
    pos = np.concatenate([r["pos"] for r in ranks])
    dp = np.abs(pos.astype(np.float64) - g["pos"])
    assert np.minimum(dp, box - dp).max() < 1e-5 * box / N


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("merged", [False, True])
def test_scale_dependent_run_on_slabs(require_gpu, tmp_path, world, merged):
    """SCALEDEPENDENT f(R) run on P slabs: distributed delta1_k / delta2_k, per-step displacement fields with the
    request / response fetch for particles that left their birth slab (2LPT.c:1784-1980) == the reference's
    single-task run (tests/golden/sd_fofr.npz)."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "sd_fofr.npz")))
    N, box = int(g["N"]), float(g["box"])
    # merged: the production settings as well (no sort on most steps, hole compaction, request lists rebuilt per order)
    ranks = _run(world, tmp_path, N, 0, "fofr", 8, 0, extra=["--sd"] + (["--merged", "--sort-interval", "4"] if merged else []))
    assert np.array_equal(np.concatenate([r["id0"] for r in ranks]), g["id0"])
    dp = np.abs(np.concatenate([r["pos0"] for r in ranks]).astype(np.float64) - g["pos0"])
    assert np.minimum(dp, box - dp).max() < 1e-5
    ids = np.concatenate([r["id"] for r in ranks]).astype(np.int64)
    assert np.array_equal(np.sort(ids), np.arange(N ** 3))
    ro = np.argsort(g["id1"])
    pos = np.concatenate([r["pos"] for r in ranks])
    vel = np.concatenate([r["vel"] for r in ranks])
    dp = np.abs(pos.astype(np.float64) - g["pos1"][ro][ids])
    assert np.minimum(dp, box - dp).max() < 2e-5 * box / N
    assert np.abs(vel - g["vel1"][ro][ids]).max() < 1e-5 * np.abs(g["vel1"]).max()
    moved = sum(int(((r["id"].astype(np.int64) // (N * N)) // (N // world) != k).sum()) for k, r in enumerate(ranks))
    assert moved > 0, "no particle left its birth slab: the remote fetch was not exercised"


@pytest.mark.parametrize("world", [2, 4, 8])
def test_halo_finder_on_ranks_matches_oracle(require_gpu, tmp_path, world):
    """mgp_fof_find on several ranks: the strip x <= dx_extra travels to the left neighbour over NCCL and the "already
    yours?" flags travel back (mm_main.c:363-373, mm_fof.c:417-437).  Every rank's catalogue must be what the restatement
    of MatchMaker gives for the same particles on the same number of tasks (pinned to the reference as a 2-rank program in
    tests/test_fof.py), bit for bit."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    import fof_case as fc
    N, box = 32, 100.0
    ranks = _run(world, tmp_path, N, 0, "lcdm", 8, 0, extra=["--fof"])
    c = fc.FOF_DEFAULTS
    tasks = [dict(pos=r["pos"], vel=r["vel"], D=r["D"], D2=r["D2"], local_p_start=int(r["p0"])) for r in ranks]
    orc, info = po.fof_halos(tasks, c["norm_pos"], c["norm_vel"], box * c["norm_pos"], c["dx_extra"], c["b_fof"], c["np_min"],
                             c["mass_part"], N, 1, c["dDdy"], c["dD2dy"])
    assert sum(h.size for h in orc) >= 40 and all(i["n_toleft"] > 0 for i in info)
    for k, r in enumerate(ranks):
        fc.assert_same_halos(r["halos"], orc[k], exact_vectors=False)
