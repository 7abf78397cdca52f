"""Per-phase CUDA-event timings of one COLA step for each deposit strategy (developer tool; the
judged numbers come from bench.py).  Usage: python tools/phase_bench.py [N] [grid_bytes]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import mgpicola_b200 as mgp   # noqa: E402


def particles(N, box, clustered, seed=1):
    rng = np.random.default_rng(seed)
    n = N ** 3
    q = np.stack(np.meshgrid(*[np.arange(N, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    q = (q + 0.5) * np.float32(box / N)
    amp = 3.0 if clustered else 0.3     # displacement rms in cells
    # smooth large-scale displacement field => clustering similar to late-time COLA
    k = 2 * np.pi / box
    psi = np.zeros_like(q)
    for m in range(1, 6):
        ph = rng.random(3) * 2 * np.pi
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        arg = k * m * (q @ d) + ph[0]
        psi += (np.sin(arg)[:, None] * d[None, :] * (amp * box / N / m)).astype(np.float32)
    pos = np.mod(q + psi + rng.standard_normal(q.shape).astype(np.float32) * np.float32(0.1 * box / N), np.float32(box)).astype(np.float32)
    pos[pos >= np.float32(box)] = 0
    return pos, psi.astype(np.float32)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    gb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    box = 1.0 * N
    for clustered in (False, True):
        pos, psi = particles(N, box, clustered)
        vel = np.zeros_like(pos)
        for mode in (2, 1, 0):
            pm = mgp.PM(N, N, box, grid_bytes=gb, deposit_mode=mode, model=mgp.MODEL_FOFR, include_screening=1)
            pm.set_pofk(0, 0, 1, 0.0, 0.0)
            pm.upload_particles(pos, vel, psi, 0.1 * psi)
            s = pm.scalars(a=0.5, phi_crit=1e-5, coupling=1 / 3, massterm2=5.0, compute_pofk=1)
            for it in range(2):      # warm-up
                pm.GetDisplacements(s); pm.Kick(0.5, 0.01, 1.0, 1.0); pm.Drift(0.01, 0.001, 0.0)
            pm.set_phase_timing(True)
            pm.phase_times(reset=True)
            t0 = time.time()
            nst = 3
            for it in range(nst):
                pm.GetDisplacements(s); pm.Kick(0.5, 0.01, 1.0, 1.0); pm.Drift(0.01, 0.001, 0.0)
            wall = (time.time() - t0) / nst * 1e3
            ph = {k: round(v[0] / nst, 3) for k, v in pm.phase_times().items() if v[1]}
            print(json.dumps(dict(N=N, gb=gb, clustered=clustered, mode=mode, wall_ms=round(wall, 2), phases_ms=ph)), flush=True)
            pm.close()


if __name__ == "__main__":
    main()
