"""Developer tool (GPU): wall-clock of the section 8(f).4 entry points at a production-like size, beside the bytes they must move.

  python tools/fof_lc_probe.py [--n1d 256] [--lc-n1d 128] [--reps 3]

FoF halo finder (mgp_fof_find + mgp_fof_get): n1d^3 particles, ~30 % of them in Gaussian blobs with an m^-2 mass function
(20 .. 20000 members) over a uniform background, b = 0.2, np_min = 20, strip of one mean inter-particle distance.
Lightcone (mgp_lightcone_count, mgp_drift_lightcone): the scalars, tables and replicate list of the reference's own
Drift_Lightcone call on the test case (tests/lightcone_case.py, needs oracle/_ref), applied to lc_n1d^3 uniform particles.
Every call synchronises the stream before it returns, so time.perf_counter around it is the device time plus the launch
and copy overheads a caller sees.  No parity leg here: tests/test_fof.py and tests/test_lightcone.py are the checkers."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def clustered(n1d, box, seed, frac=0.3):
    rng = np.random.default_rng(seed)
    n = n1d ** 3
    ipd = box / n1d
    sizes = []
    left = int(frac * n)
    while left > 0:
        m = np.minimum((20.0 / rng.random(4096)).astype(np.int64), 20000)
        sizes.append(m)
        left -= int(m.sum())
    sizes = np.concatenate(sizes)
    keep = np.cumsum(sizes) <= int(frac * n)
    sizes = sizes[keep]
    centres = rng.uniform(0, box, (sizes.size, 3))
    sig = 0.12 * ipd * (sizes / 50.0) ** (1.0 / 3.0)
    nb = int(sizes.sum())
    blob = np.repeat(centres, sizes, axis=0) + rng.standard_normal((nb, 3)) * np.repeat(sig, sizes)[:, None]
    pos = np.concatenate([blob, rng.uniform(0, box, (n - nb, 3))])
    pos = np.mod(pos, box).astype(np.float32)
    pos[pos >= np.float32(box)] = 0.0
    pos = pos[rng.permutation(n)]
    vel = (rng.standard_normal((n, 3)) * 2.0).astype(np.float32)
    return pos, vel, sizes


def nrep_of(case):
    return int(np.asarray(case["reps"]).reshape(-1, 3).shape[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n1d", type=int, default=256)
    ap.add_argument("--lc-n1d", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--skip-lightcone", action="store_true")
    a = ap.parse_args()
    import mgpicola_b200 as mgp
    out = {}

    N = a.n1d
    box = 100.0 * N / 32.0
    t0 = time.perf_counter()
    pos, vel, sizes = clustered(N, box, 5)
    zero = np.zeros_like(pos)
    t_make = time.perf_counter() - t0
    pm = mgp.PM(N, N, box, grid_bytes=4, use_cola=0, sort_particles=0)
    pm.upload_particles(pos, vel, zero, zero)
    ts = []
    for _ in range(a.reps):
        l0 = pm.launch_count()
        t0 = time.perf_counter()
        h = pm.MatchMaker(1.0, 1.0, box, box / N, 0.2, 20, 1.0)
        ts.append(time.perf_counter() - t0)
        launches = pm.launch_count() - l0
    n = N ** 3
    out["fof"] = dict(n1d=N, particles=n, blobs=int(sizes.size), halos=int(h.size), in_halos=int(h["np"].sum()),
                      largest=int(h["np"][0]) if h.size else 0, ms=[round(1e3 * t, 3) for t in ts], launches=launches,
                      particles_per_s=n / min(ts), algorithmic_bytes_per_particle=116,
                      gbs=116.0 * n / min(ts) / 1e9, make_s=round(t_make, 2))
    print(json.dumps(out["fof"]), flush=True)
    pm.close()

    if not a.skip_lightcone:
        import lightcone_case as lcc
        case = lcc.reference_case(tempfile.mkdtemp(prefix="lcprobe"))
        if case is None:
            print(json.dumps({"lightcone": "oracle/_ref lightcone build missing"}))
        else:
            i, s = case["inputs"], case["scalars"]
            M = a.lc_n1d
            rng = np.random.default_rng(3)
            m = M ** 3
            k = rng.integers(0, i["pos"].shape[0], m)                    # the case's particles, resampled and jittered
            p = np.mod(i["pos"][k] + rng.uniform(0, i["box"] / 16, (m, 3)).astype(np.float32), np.float32(i["box"])).astype(np.float32)
            p[p >= np.float32(i["box"])] = 0.0
            pm = mgp.PM(M, M, i["box"], omega=case["omega"], grid_bytes=4, use_cola=i["use_cola"], sort_particles=0)
            pm.upload_particles(p, i["vel"][k], i["D"][k], i["D2"][k])
            tc = []
            for _ in range(a.reps):
                t0 = time.perf_counter()
                cnt = pm.lightcone_count(s, case["reps"], i["sumxyz"])
                tc.append(time.perf_counter() - t0)
            total = int(cnt.sum())
            before = pm.download_particles(("pos",))["pos"]
            t0 = time.perf_counter()
            rows = pm.Drift_Lightcone(s, case["reps"], i["sumxyz"], cap=int(cnt.max()), pinned=True)
            t_drift = time.perf_counter() - t0
            assert sum(r.shape[0] for r in rows) == total
            del rows
            pm.upload_particles(before, i["vel"][k], i["D"][k], i["D2"][k])     # the same step again, library call alone
            ls, keep = pm._lightcone_step(s, case["reps"], i["sumxyz"])
            cap = int(cnt.max())
            host = pm.L.mgp_alloc_host(24 * cap * nrep_of(case))
            c2 = np.zeros(nrep_of(case), np.uint64)
            t0 = time.perf_counter()
            pm._ck(pm.L.mgp_drift_lightcone(pm.ctx, mgp.C.byref(ls), cap, host, c2.ctypes.data))
            t_call = time.perf_counter() - t0
            pm.L.mgp_free_host(host)
            assert int(c2.sum()) == total
            nrep = int(np.asarray(case["reps"]).reshape(-1, 3).shape[0])
            out["lightcone"] = dict(n1d=M, particles=m, replicates=nrep, rows=total, count_ms=[round(1e3 * t, 3) for t in tc],
                                    count_pairs_per_s=m * nrep / min(tc), count_gbs=56.0 * m / min(tc) / 1e9,
                                    drift_ms_python_binding=round(1e3 * t_drift, 3),
                                    drift_ms_library_call_pinned_block=round(1e3 * t_call, 3), row_bytes=24 * total,
                                    rows_gbs=24.0 * total / t_call / 1e9)
            print(json.dumps(out["lightcone"]), flush=True)
            pm.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
