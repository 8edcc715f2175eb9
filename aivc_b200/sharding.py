"""Multi-GPU partition of a sequence: GOPs are independent units (each starts with its own I
frame and references nothing outside itself -- func_util/GOP_structure.py:27-137,
model_management.py:169-173), so rank r codes GOPs r, r+N, ... and only byte strings travel.
No data-path collective; the gather below moves the finished bitstreams (KB..MB) to rank 0."""
import torch.distributed as dist


def gops_of_rank(n_gops, rank, world):
    return list(range(rank, n_gops, world))


def gather_gop_bytes(mine, n_gops):
    """mine: {gop index: bytes} on every rank -> list of all GOP byte strings on rank 0
    (None elsewhere), in GOP order, ready for container.pack_video."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [mine[i] for i in range(n_gops)]
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if dist.get_rank() != 0:
        return None
    merged = {}
    for d in gathered:
        merged.update(d)
    return [merged[i] for i in range(n_gops)]
