"""In-situ device timing of PtoMesh / MtoParticles pieces (CUDA events on the library's stream) for the deposit strategies.
Usage: python tools/deposit_probe.py [N] [mode]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import mgpicola_b200 as mgp   # noqa: E402
import bench                  # noqa: E402
from mgpicola_b200 import cosmology  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
box = bench.box_for(N)
pm = mgp.PM(N, N, box, omega=0.267, grid_bytes=8, deposit_mode=mode, sort_particles=4)
pm.set_pofk(64, 1, 1, 0.03, 2.0)
cos = cosmology.LCDM(0.267, 9.0)
pm.ic_generate(bench.amplitude_table(N, box), seed=5001)
pm.init_particles(cos.growth_D(0.1), cos.growth_D2(0.1))
stream = torch.cuda.ExternalStream(pm.stream)


def timed(fn, reps=3):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


pm.set_phase_timing(True)
for step in range(6):
    pm.phase_times(reset=True)
    t_p = timed(lambda: pm.PtoMesh(), 1)
    ph1 = {k: round(v[0], 3) for k, v in pm.phase_times(reset=True).items() if v[1]}
    pm.Forces()
    t_g = timed(lambda: pm.MtoParticles(), 1)
    ph2 = {k: round(v[0], 3) for k, v in pm.phase_times(reset=True).items() if v[1]}
    pm.Kick(0.1 + 0.03 * step, 0.02, 1.0, -0.4)
    pm.Drift(0.3, 0.03, -0.01)
    print("step %d: PtoMesh call %.3f ms %s | MtoParticles call %.3f ms %s" % (step, t_p[0], ph1, t_g[0], ph2), flush=True)
pm.close()
