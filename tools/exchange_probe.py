"""torchrun probe of the slab exchange over peer memory: device time of each piece alone (mgp_debug_time_exchange)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mgpicola_b200 as mgp  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
nid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [mgp.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nid = ids[0]
else:
    os.environ["MGP_FORCE_SLAB"] = "1"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pm = mgp.PM(N, N, 100.0, grid_bytes=gb, rank=rank, nranks=world, device=local, nccl_id=nid)
slab = N * N * (N // 2 + 1) * 2 * gb / world            # complex bytes of one rank's slab
remote = slab * (world - 1) / world
names = ["flag barrier", "fused bwd (x-FFT + push)", "fused fwd (pull + x-FFT)", "transpose bwd (push)", "transpose fwd (push)",
         "copy-engine peer copy of the remote share", "DMA exchange bwd (strided blocks to every peer)",
         "DMA exchange fwd (strided blocks to every peer)", "batched c2r of the 3 force grids (pipeline)", "one local 2-D c2r"]
only = [int(x) for x in os.environ.get('PROBE_ONLY', '').split(',') if x]
for which, nm in enumerate(names):
    if only and which not in only:
        continue
    try:
        pm.debug_time_exchange(which, 2)
        ms = pm.debug_time_exchange(which, 6)
    except mgp.MgpError as e:
        if rank == 0:
            print("%-45s unavailable (%s)" % (nm, e))
        continue
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("N=%d P=%d gb=%d  %-45s %8.3f ms   slab %.0f MB -> %6.0f GB/s HBM-side (R+W), remote %.0f MB -> %5.0f GB/s per direction"
              % (N, world, gb, nm, t.item(), slab / 1e6, 2 * slab / t.item() / 1e6, remote / 1e6, remote / t.item() / 1e6), flush=True)
pm.close()
