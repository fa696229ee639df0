"""Multi-GPU plumbing: envs are independent, so each rank owns a contiguous slab of envs and nothing is
exchanged per step. The only collective is one all-reduce of the small integer episode-stat vector per
rollout (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests). Reference analogue:
`num_destruction += sum(info)` and the final_rewards mean over envs, rl/train.py:81-88,158-165."""
import os


def shard(total_envs, rank=None, world_size=None):
    """Contiguous slab [first, first+count) of `total_envs` owned by `rank` (global env id = first + i, so
    seeds and synthetic action streams are shard invariant)."""
    if rank is None:
        rank = int(os.environ.get("RANK", "0"))
    if world_size is None:
        world_size = int(os.environ.get("WORLD_SIZE", "1"))
    base, rem = divmod(int(total_envs), world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_reduce_episode_stats(vec):
    """Sum the int64 episode-stat vector over ranks (index 20, maxVlner_max, is max-reduced). No-op when
    torch.distributed is not initialised. Integer sums make the result bitwise shard invariant."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vec
    mx = vec[20:21].clone()
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    vec[20:21] = mx
    return vec


def init_from_env(backend=None):
    """torch.distributed init from torchrun's environment (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world
