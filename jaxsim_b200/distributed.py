"""Multi-GPU plumbing: environments shard by rank, one collective at readback.

The reference has no multi-device path (no pmap/shard_map/collectives).
The natural B200 layout (BASELINE config 4) is one process per GPU, each
stepping a contiguous range of environments with ZERO communication per step; the only
collective is an ``all_gather`` of the state leaves when the caller wants the full batch
back (NCCL over NVLink on GPUs; the same code runs on ``gloo`` for the CPU tests).
"""

from __future__ import annotations

import torch
import torch.distributed as dist

STATE_LEAVES = (
    "_joint_positions", "_joint_velocities", "_base_quaternion", "_base_linear_velocity",
    "_base_angular_velocity", "_base_position",
)


def shard_range(global_batch: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous range ``[start, stop)`` of environments owned by ``rank``; the first
    ``global_batch % world_size`` ranks get one extra environment."""
    if not (0 <= rank < world_size) or global_batch < 0:
        raise ValueError((global_batch, rank, world_size))
    base, extra = divmod(global_batch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_state(data) -> torch.Tensor:
    """Flatten the state leaves (+ soft-contact state) of a batched data object to (B, k)."""
    B = data._base_quaternion.shape[0]
    parts = [getattr(data, k).reshape(B, -1) for k in STATE_LEAVES]
    if "tangential_deformation" in data.contact_state:
        parts.append(data.contact_state["tangential_deformation"].reshape(B, -1))
    return torch.cat(parts, dim=-1).contiguous()


def unpack_state(flat: torch.Tensor, n: int, nc: int | None) -> dict:
    """Inverse of :func:`pack_state`: dict of leaves (contact state only if ``nc`` given)."""
    widths = [n, n, 4, 3, 3, 3]
    out, o = {}, 0
    for k, w in zip(STATE_LEAVES, widths):
        out[k] = flat[:, o : o + w].contiguous()
        o += w
    if nc is not None:
        out["tangential_deformation"] = flat[:, o : o + 3 * nc].reshape(flat.shape[0], nc, 3).contiguous()
        o += 3 * nc
    assert o == flat.shape[1], (o, flat.shape)
    return out


def all_gather_state(data, group=None) -> torch.Tensor:
    """Gather the packed state of every rank's shard, in rank order: (sum_B, k).
    Shards may have different sizes (``shard_range``)."""
    flat = pack_state(data)
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=flat.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([flat.shape[0]], dtype=torch.int64, device=flat.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        full = torch.empty(world * sizes[0], flat.shape[1], dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(full, flat, group=group)
        return full
    mx = max(sizes)
    pad = torch.zeros(mx, flat.shape[1], dtype=flat.dtype, device=flat.device)
    pad[: flat.shape[0]] = flat
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
