"""Multi-GPU sharding of circuit batches: one process per GPU, contiguous slices, no traffic
while simulating; one all_gather of norms / amplitudes at the end.
Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (host-logic tests)."""
from typing import Tuple


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of ``total`` batch members owned by ``rank``; the first
    ``total % world_size`` ranks hold one extra member."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_slices(local, total: int, group=None):
    """all_gather per-member results (first axis = local batch slice) into the full batch order.
    Slices may differ by one member between ranks, so they are padded to a common length."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    width = max(counts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if pad.is_complex():
        pad = torch.view_as_real(pad).contiguous()
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = [o[:c] for o, c in zip(out, counts)]
    full = torch.cat(parts, dim=0)
    if local.is_complex():
        full = torch.view_as_complex(full.contiguous())
    return full
