"""One process per GPU: rank / world from the torchrun environment, NCCL unique id for the library's
own communicator handed round through torch.distributed (any backend; gloo on CPU in the tests)."""
import os


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend=None):
    """Initialises torch.distributed when WORLD_SIZE > 1; returns (rank, world, local_rank)."""
    rank, world, local = env_rank()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
    return rank, world, local


def share_from_rank0(make):
    """rank 0 evaluates make(); everybody returns its value (used for the 128-byte ncclUniqueId)."""
    rank, world, _ = env_rank()
    if world == 1:
        return make()
    import torch.distributed as dist
    box = [make() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def allreduce_max(value):
    rank, world, _ = env_rank()
    if world == 1:
        return value
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum_array(a):
    rank, world, _ = env_rank()
    if world == 1:
        return a
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(a.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def barrier():
    rank, world, _ = env_rank()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
