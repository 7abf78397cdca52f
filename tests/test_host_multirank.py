"""Host-side multi-rank logic on CPU: world_size 2 over gloo (spawned with torch.multiprocessing),
plus pure-integer checks of the slab bookkeeping against the reference rule."""
import os
import socket

import numpy as np
import pytest


def test_layout_matches_reference_rule(mgp):
    from mgpicola_b200 import slab
    for N, Ns, P in [(128, 128, 1), (128, 128, 4), (64, 32, 8), (96, 64, 5), (256, 128, 3), (16, 16, 16)]:
        tot_nx, tot_np = 0, 0
        s2t = slab.slab_to_task(N, P)
        for r in range(P):
            nx, x0, npl, p0 = slab.layout(N, Ns, P, r)
            assert nx == int((s2t == r).sum())
            if nx:
                assert x0 == int(np.argmax(s2t == r))
            tot_nx += nx
            tot_np += npl
        assert tot_nx == N and tot_np == Ns           # disjoint cover of mesh planes and Lagrangian planes


def test_owner_rule_edges(mgp):
    from mgpicola_b200 import slab
    N, box, P = 64, 100.0, 4
    top = np.nextafter(np.float32(box), np.float32(0))
    h = np.float32(box / N)
    x = np.array([0.0, top, 16 * h, np.nextafter(16 * h, np.float32(0)), 32 * h, 47.999 * h], np.float32)
    X = (x.astype(np.float64) * (N / box)).astype(np.int64)
    assert list(slab.owner_of(x, N, box, P)) == [int(v) // 16 for v in X]
    assert slab.owner_of(np.array([top]), N, box, P)[0] == P - 1


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mgpicola_b200  # noqa: F401
    from mgpicola_b200 import dist as mdist
    from mgpicola_b200 import slab
    r, w, _ = mdist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    token = mdist.share_from_rank0(lambda: os.urandom(128))        # stands in for the ncclUniqueId
    assert len(token) == 128
    # every rank selects its particles from the same seeded set: disjoint cover, exact counts
    N, box = 32, 50.0
    rng = np.random.default_rng(7)
    pos = (rng.random((N ** 3, 3)) * box).astype(np.float32)
    own = slab.owner_of(pos[:, 0], N, box, world)
    mine = own == rank
    counts = np.zeros(world, dtype=np.int64)
    counts[rank] = mine.sum()
    counts = mdist.allreduce_sum_array(counts)
    assert counts.sum() == N ** 3
    assert mdist.allreduce_max(float(rank)) == world - 1
    tokens = mdist.allreduce_sum_array(np.frombuffer(token, dtype=np.uint8).astype(np.int64))
    assert np.array_equal(tokens, np.frombuffer(token, dtype=np.uint8).astype(np.int64) * world)   # same id everywhere
    mdist.barrier()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("%d" % counts[rank])


def test_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
