"""PtoMesh repeated on the same particles: in-situ device time per call (CUDA events on the library's stream)."""
import subprocess
import sys

import torch

sys.path.insert(0, ".")
import mgpicola_b200 as mgp   # noqa: E402
import bench                  # noqa: E402
from mgpicola_b200 import cosmology  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
box = bench.box_for(N)
pm = mgp.PM(N, N, box, omega=0.267, grid_bytes=8, deposit_mode=mode, sort_particles=4)
cos = cosmology.LCDM(0.267, 9.0)
pm.ic_generate(bench.amplitude_table(N, box), seed=5001)
pm.init_particles(cos.growth_D(0.1), cos.growth_D2(0.1))
stream = torch.cuda.ExternalStream(pm.stream)
pm.set_phase_timing(True)
for it in range(8):
    pm.phase_times(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    pm.PtoMesh()
    e1.record(stream)
    torch.cuda.synchronize()
    ph = {k: round(v[0], 3) for k, v in pm.phase_times(reset=True).items() if v[1]}
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print("call %d: %.3f ms %s | %s" % (it, e0.elapsed_time(e1), ph, clk), flush=True)
pm.close()
