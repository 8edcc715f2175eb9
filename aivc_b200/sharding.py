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


# ------------------------------------------------------------------------------------------
# Frame-level sharding inside ONE GOP (latency mode): frames of the same dependency level are
# independent given the reconstructions of earlier levels (gop.levels), so a level's frames are
# dealt round-robin to the ranks and every new reconstruction is broadcast (NCCL over NVLink:
# 3.1 MB of 8-bit 4:2:0 planes per 1080p frame) before the next level starts.  For '1_GOP_32'
# on 8 GPUs the critical path is 1+1+1+1+1+1+2 = 8 frame times instead of 33.
def _bcast_planes(planes, shapes, src, device):
    import torch
    flat = torch.cat([p.reshape(-1) for p in planes]) if planes is not None else \
        torch.empty(sum(shapes), dtype=torch.uint8, device=device)
    dist.broadcast(flat, src=src)
    out, pos = [], 0
    for n in shapes:
        out.append(flat[pos:pos + n])
        pos += n
    return tuple(out)


def encode_gop_frame_parallel(codec, frames, gop_struct, plane_sizes, device):
    """Every rank holds all source frames of the GOP.  Returns (bytes per frame on rank 0 / None
    elsewhere, reconstructions of all frames on every rank)."""
    from .gop import levels
    rank, world = dist.get_rank(), dist.get_world_size()
    rec, mine = {}, {}
    for level in levels(gop_struct):
        local = {}
        for i, f in enumerate(level):
            if i % world == rank:
                e = gop_struct[f]
                mine[f], local[f] = codec.encode_frame(frames[f], e['type'], rec.get(e['prev_ref']),
                                                       rec.get(e['next_ref']))
        for i, f in enumerate(level):           # one broadcast per new reference
            rec[f] = _bcast_planes(local.get(f), plane_sizes, i % world, device)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None, rec
    merged = {}
    for d in gathered:
        merged.update(d)
    return merged, rec


def decode_gop_frame_parallel(codec, frame_bytes, gop_struct, plane_sizes, device):
    """frame_bytes: {'frame_i': bytes} on every rank. Returns all reconstructions on every rank."""
    from .gop import levels
    rank, world = dist.get_rank(), dist.get_world_size()
    rec = {}
    for level in levels(gop_struct):
        local = {}
        for i, f in enumerate(level):
            if i % world == rank:
                e = gop_struct[f]
                local[f] = codec.decode_frame(frame_bytes[f], e['type'], rec.get(e['prev_ref']),
                                              rec.get(e['next_ref']))
        for i, f in enumerate(level):
            rec[f] = _bcast_planes(local.get(f), plane_sizes, i % world, device)
    return rec
