"""torchrun probe of host <-> device copy bandwidth with pinned memory: every rank alone, then all ranks at once.
Explains (or not) the end-to-end numbers of bench.py on several ranks."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 512 << 20
t0 = time.perf_counter()
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
t_pin = time.perf_counter() - t0
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for r in range(world):          # one rank at a time
        sync()
        if r == rank:
            res[name + "_alone"] = nbytes / timed(fn) / 1e9
    sync()
    res[name + "_together"] = nbytes / timed(fn) / 1e9
    sync()
aff = sorted(os.sched_getaffinity(0))
out = [None] * world
if world > 1:
    dist.all_gather_object(out, (rank, t_pin, res, len(aff)))
else:
    out = [(rank, t_pin, res, len(aff))]
if rank == 0:
    for r, tp, rs, na in out:
        print("rank %d: pin 512 MiB %.2f s, cpus %d | " % (r, tp, na) + "  ".join("%s %.1f GB/s" % kv for kv in rs.items()), flush=True)
    for k in ("h2d_together", "d2h_together"):
        print("aggregate %s: %.1f GB/s" % (k, sum(o[2][k] for o in out)))
