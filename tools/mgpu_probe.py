"""developer probe: wall time of every driver-level call of a COLA step on N ranks (torchrun), synchronised."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import mgpicola_b200 as mgp
from mgpicola_b200 import cosmology

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
nid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [mgp.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nid = ids[0]
N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.DEFAULT_NMESH
use_sd = int(sys.argv[2]) if len(sys.argv) > 2 else 1
box = bench.box_for(N)
cos = cosmology.LCDM(bench.OMEGA, bench.Z_INIT)
sd = cosmology.ScaleDependentGrowth(cos, box, N, "fofr") if use_sd else None
pm = mgp.PM(N, N, box, omega=bench.OMEGA, model=mgp.MODEL_FOFR, include_screening=1, grid_bytes=8, rank=rank, nranks=world,
            device=local, nccl_id=nid, deposit_mode=0, sort_particles=4, scale_dependent=use_sd)
pm.set_pofk(64, 1, 1, 0.03, 2.0)
A0 = 1.0 / (1.0 + bench.Z_INIT)
power = bench.amplitude_table(N, box)
if sd is not None:
    power = power * sd.pofk_ratio_by_k2()
pm.ic_generate(power, seed=5001)
if sd is not None:
    for o in (1, 2):
        pm.assign_displacment_field_to_particles(0, o, sd.table(0, o, A0))
pm.init_particles(cos.growth_D(A0), cos.growth_D2(A0))
st = bench.Stepper(pm, cos, "fofr", box, sd, "merged")


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    acc.setdefault(label, []).append((time.perf_counter() - t) * 1e3)


for it in range(6):
    acc = {}
    sc, A, dda, ddD, ddD2, dyyy, dD, dD2, tabs = st.pre[it]
    timed("MoveParticles", pm.MoveParticles, acc)
    timed("PtoMesh", lambda: pm.PtoMesh(sc), acc)
    timed("FifthForce", lambda: pm.ComputeFifthForce(sc), acc)
    timed("Forces", pm.Forces, acc)
    timed("MtoParticles", pm.MtoParticles, acc)
    if tabs is not None:
        timed("assign dD", lambda: pm.assign_displacement_fields_merged(3, tabs[0], tabs[1]), acc)
        timed("assign ddD", lambda: pm.assign_displacement_fields_merged(2, tabs[2], tabs[3]), acc)
    timed("Kick", lambda: pm.Kick(A, dda, ddD, ddD2), acc)
    timed("Drift", lambda: pm.Drift(dyyy, dD, dD2), acc)
    if rank == 0:
        print("step %d np=%d " % (it, pm.numpart) + "  ".join("%s %.2f" % (k, v[0]) for k, v in acc.items()) +
              "  | total %.2f" % sum(v[0] for v in acc.values()), flush=True)
# bench-like loop: no added synchronisation; host wall time per driver-level call, slow steps are itemised
def wall(label, fn, acc):
    t = time.perf_counter()
    fn()
    acc[label] = (time.perf_counter() - t) * 1e3


for it in range(6, 30):
    acc = {}
    sc, A, dda, ddD, ddD2, dyyy, dD, dD2, tabs = st.pre[it]
    wall("Move", pm.MoveParticles, acc)
    wall("PtoMesh", lambda: pm.PtoMesh(sc), acc)
    wall("Fifth", lambda: pm.ComputeFifthForce(sc), acc)
    wall("Forces", pm.Forces, acc)
    wall("MtoP", pm.MtoParticles, acc)
    if tabs is not None:
        wall("asg1", lambda: pm.assign_displacement_fields_merged(3, tabs[0], tabs[1]), acc)
        wall("asg2", lambda: pm.assign_displacement_fields_merged(2, tabs[2], tabs[3]), acc)
    wall("Kick", lambda: pm.Kick(A, dda, ddD, ddD2), acc)
    wall("Drift", lambda: pm.Drift(dyyy, dD, dD2), acc)
    tot = sum(acc.values())
    if rank == 0:
        print("it %d total %.2f  " % (it, tot) + ("  ".join("%s %.2f" % kv for kv in acc.items()) if tot > 16 else ""), flush=True)
pm.close()
