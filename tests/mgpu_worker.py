"""Worker for the multi-GPU parity tests: one process per GPU (torchrun), x-slab decomposition.
Every rank starts from a deliberately WRONG share of a seeded particle set (round-robin by index), so
the first MoveParticles migrates almost everything; then COLA steps are taken and each rank saves its
particles.  The parent test compares against the single-task oracle."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmesh", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--model", default="fofr")
    ap.add_argument("--gb", type=int, default=8)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--ic", action="store_true", help="generate the golden-fixture ICs on the ranks instead of stepping")
    ap.add_argument("--sd", action="store_true", help="scale-dependent run of tests/golden/sd_fofr.npz on the ranks")
    ap.add_argument("--merged", action="store_true")
    ap.add_argument("--fof", action="store_true", help="FoF halo finder on the ranks (tests/fof_case.py particles)")
    ap.add_argument("--sort-interval", type=int, default=1, help="sort_particles of the context (>= 3: the hole compaction of "
                    "MoveParticles runs between sorts)")
    a = ap.parse_args()
    import torch
    import mgpicola_b200 as mgp
    from mgpicola_b200 import dist as mdist
    from oracle import pm_oracle as po            # only for the per-step scalar helpers (host numbers)
    import test_gpu_parity as T
    rank, world, local = mdist.init_process_group("nccl")
    torch.cuda.set_device(local)
    if a.ic:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ic_lcdm.npz")))
        N, box = int(g["N"]), float(g["box"])
        nid = mdist.share_from_rank0(mgp.nccl_unique_id)
        pm = mgp.PM(N, N, box, grid_bytes=a.gb, rank=rank, nranks=world, device=local, nccl_id=nid)
        pm.ic_generate(g["power_by_k2"], seed=int(g["seed"]))
        pm.init_particles(float(g["Di"]), float(g["Di2"]))
        got = pm.download_particles()
        np.savez(os.path.join(a.out, "rank%d.npz" % rank), p0=pm.local_p_start, npl=pm.local_np, **got)
        mdist.barrier()
        pm.close()
        return
    if a.fof:
        import fof_case as fc
        N, box = a.nmesh, 100.0
        pos, vel, D, D2 = fc.make_particles(N, box, 3)
        mine = (np.arange(N ** 3) % world) == rank
        nid = mdist.share_from_rank0(mgp.nccl_unique_id)
        pm = mgp.PM(N, N, box, grid_bytes=a.gb, buffer=2.5, rank=rank, nranks=world, device=local, nccl_id=nid, sort_particles=1)
        pm.upload_particles(pos[mine], vel[mine], D[mine], D2[mine], np.arange(N ** 3, dtype=np.uint64)[mine])
        pm.MoveParticles()                  # every particle to the rank that owns its slab (auxPM.c:151-153)
        got = pm.download_particles()
        c = fc.FOF_DEFAULTS
        h = pm.MatchMaker(c["norm_pos"], c["norm_vel"], box * c["norm_pos"], c["dx_extra"], c["b_fof"], c["np_min"], c["mass_part"],
                          c["dDdy"], c["dD2dy"])
        np.savez(os.path.join(a.out, "rank%d.npz" % rank), halos=h, p0=pm.local_p_start, **got)
        mdist.barrier()
        pm.close()
        return
    if a.sd:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "sd_fofr.npz")))
        N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
        nid = mdist.share_from_rank0(mgp.nccl_unique_id)
        pm = mgp.PM(N, N, box, omega=om, model=mgp.MODEL_FOFR, include_screening=1, grid_bytes=a.gb, scale_dependent=1, buffer=2.5,
                    rank=rank, nranks=world, device=local, nccl_id=nid, deposit_mode=a.mode, sort_particles=a.sort_interval)
        c = g["pofk_cfg"]
        pm.set_pofk(int(c[0]), int(c[1]), int(c[2]), float(c[3]), float(c[4]))
        pm.ic_generate(g["power_by_k2"], seed=int(g["seed"]))            # keeps delta1_k / delta2_k, transposed slabs
        for idx, (ft, order) in enumerate([(0, 1), (0, 2), (1, 1), (1, 2)]):
            pm.assign_displacment_field_to_particles(ft, order, g["G_init"][idx])
        pm.init_particles(0.0, 0.0)
        init = pm.download_particles()
        pks = []
        for it, (A, AI, AF, AFF, dda, dyyy) in enumerate(g["steps"]):
            pc, cc, m2 = po.fofr_scalars(A, om, box, float(g["fofr0"]), float(g["nfofr"]))
            pm.GetDisplacements(pm.scalars(a=A, phi_crit=pc, coupling=cc, massterm2=m2, compute_pofk=1))
            pks.append(np.stack(pm.step_power_spectrum()))
            Tb = g["G_steps"][it]
            if a.merged:
                pm.assign_displacement_fields_merged(3, Tb[0], Tb[1])
                pm.assign_displacement_fields_merged(2, Tb[2], Tb[3])
            else:
                for idx, (ft, order) in enumerate([(3, 1), (3, 2), (2, 1), (2, 2)]):
                    pm.assign_displacment_field_to_particles(ft, order, Tb[idx])
            pm.Kick(A, dda, 0.0, 0.0)
            pm.Drift(dyyy, 0.0, 0.0)
        pm.MoveParticles()
        got = pm.download_particles(want=("pos", "vel", "id"))
        np.savez(os.path.join(a.out, "rank%d.npz" % rank), pks=np.stack(pks), id0=init["id"], pos0=init["pos"], **got)
        mdist.barrier()
        pm.close()
        return
    N, box, om = a.nmesh, 100.0, 0.267
    pos, vel, D, D2 = T.make_particles(N, box, 77, clustered=True)
    ids = np.arange(N ** 3, dtype=np.uint64)
    mine = (np.arange(N ** 3) % world) == rank
    nid = mdist.share_from_rank0(mgp.nccl_unique_id)
    mid = {"lcdm": mgp.MODEL_NONE, "fofr": mgp.MODEL_FOFR, "dgp": mgp.MODEL_DGP}[a.model]
    pm = mgp.PM(N, N, box, omega=om, model=mid, include_screening=1, grid_bytes=a.gb, deposit_mode=a.mode, buffer=2.5,
                rank=rank, nranks=world, device=local, nccl_id=nid, sort_particles=a.sort_interval)
    pm.set_pofk(16, 0, 1, 0.0, 0.0)
    pm.upload_particles(pos[mine], vel[mine], D[mine], D2[mine], ids[mine])
    pks = []
    for it in range(a.steps):
        A = 0.5 + 0.1 * it
        if a.model == "fofr":
            pc, c, m2 = po.fofr_scalars(A, om, box, 1e-5, 1.0)
            s = pm.scalars(a=A, phi_crit=pc, coupling=c, massterm2=m2, compute_pofk=1)
        elif a.model == "dgp":
            c, f0 = po.dgp_scalars(A, om, 1.2)
            s = pm.scalars(a=A, coupling=c, dgp_fac0=f0, rsmooth=1.0, compute_pofk=1)
        else:
            s = pm.scalars(a=A, compute_pofk=1)
        pm.GetDisplacements(s)
        pks.append(np.stack(pm.step_power_spectrum()))
        pm.Kick(A, 0.02, 1.3, -0.4)
        pm.Drift(0.5, 0.03, -0.01)
    pm.MoveParticles()                  # final ownership
    got = pm.download_particles()
    np.savez(os.path.join(a.out, "rank%d.npz" % rank), pks=np.stack(pks), x0=pm.local_x_start, nx=pm.local_nx,
             launches=pm.launch_count(), **got)
    mdist.barrier()
    pm.close()


if __name__ == "__main__":
    main()
