"""Host-side logic of the multi-GPU path (one process per GPU, `torch.distributed` for the plumbing).

Frames are sharded across ranks; voxels are spatially owned (8^3-voxel bricks hashed over the ranks, the same
function libssm.so uses on the device: `ssm_voxel_owner`).  Inside libssm.so the point exchange is an NCCL
all-to-all; this module holds what stays on the host: the frame split, the NCCL unique-id bootstrap over the
process group, a numpy mirror of the ownership function (for routing host-resident clouds and for tests), and the
gather of the per-rank voxel tables into one map.
"""
from __future__ import annotations

import numpy as np

BRICK_SHIFT = 3
_M64 = (1 << 64) - 1


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frames for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def voxel_ijk(xyz: np.ndarray, leaf: float) -> np.ndarray:
    """pcl::VoxelGrid cell of each point: floor(fp32(coord) * fp32(1 / leaf)) evaluated in fp32."""
    inv = np.float32(1.0) / np.float32(leaf)
    return np.floor(xyz.astype(np.float32) * inv).astype(np.int32)


def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(33)
    x = (x * np.uint64(0xFF51AFD7ED558CCD)) & np.uint64(_M64)
    x ^= x >> np.uint64(33)
    x = (x * np.uint64(0xC4CEB9FE1A85EC53)) & np.uint64(_M64)
    x ^= x >> np.uint64(33)
    return x


def voxel_owner_np(ijk: np.ndarray, world: int) -> np.ndarray:
    """numpy mirror of ssm_voxel_owner (csrc/ssm_internal.cuh voxel_owner) for arrays of voxel coordinates."""
    if world <= 1:
        return np.zeros(len(ijk), np.int32)
    b = (ijk.astype(np.int64) >> BRICK_SHIFT) + (1 << 20)
    with np.errstate(over="ignore"):
        key = (b[:, 0].astype(np.uint64) & np.uint64(0xFFFFFFFF)) | ((b[:, 1].astype(np.uint64) & np.uint64(0xFFFFFFFF)) << np.uint64(21)) \
            | ((b[:, 2].astype(np.uint64) & np.uint64(0xFFFFFFFF)) << np.uint64(42))
        h = _mix64(key ^ np.uint64(0x9E3779B97F4A7C15))
    return (h % np.uint64(world)).astype(np.int32)


def bucket_points(xyz: np.ndarray, rgba: np.ndarray, label: np.ndarray, leaf: float, world: int):
    """Group a host-resident cloud by owning rank.  Returns (order, counts): `order` sorts points by owner."""
    owner = voxel_owner_np(voxel_ijk(xyz, leaf), world)
    order = np.argsort(owner, kind="stable")
    counts = np.bincount(owner, minlength=world).astype(np.int64)
    return order, counts


def init_comm(ctx, group=None, p2p: bool = True):
    """Create the NCCL communicator inside libssm.so for `ctx`: rank 0 makes the unique id, the process group
    broadcasts it (any backend), every rank calls ssm_comm_init."""
    import torch
    import torch.distributed as dist
    from .lib import Context
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    uid = torch.from_numpy(Context.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
    if dist.get_backend(group) == "nccl":
        uid = uid.cuda()
    dist.broadcast(uid, src=0, group=group)
    ctx.comm_init(uid.cpu().numpy(), rank, world)
    if p2p:
        # peer-memory routing: all-gather the CUDA IPC handles of the inboxes, map the peers
        mine = torch.from_numpy(ctx.comm_ipc_export())
        if dist.get_backend(group) == "nccl":
            mine = mine.cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        ctx.comm_ipc_connect(torch.stack(allh).cpu().numpy())


def merge_exports(parts: list[dict]) -> dict:
    """Union of per-rank voxel tables (disjoint by ownership), ordered by (k, j, i) like pcl::VoxelGrid output."""
    keys = ("ijk", "xyz", "rgba", "label", "count", "votes")
    out = {k: np.concatenate([p[k] for p in parts]) for k in keys if all(k in p for p in parts)}
    ijk = out["ijk"]
    order = np.lexsort((ijk[:, 0], ijk[:, 1], ijk[:, 2]))
    return {k: v[order] for k, v in out.items()}


def gather_map_native(ctx, group=None, sorted: bool = True, fields=None) -> dict | None:
    """The same through the C ABI alone (ssm_map_export_gathered: NCCL record gather + device finalize / sort on rank 0);
    the process group only adds up the sizes so that rank 0 can allocate."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return ctx.map_export(sorted=sorted, fields=fields)
    n = torch.tensor([ctx.map_size()], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        n = n.cuda()
    dist.all_reduce(n, group=group)
    return ctx.map_export_gathered(int(n.item()) if dist.get_rank(group) == 0 else None, sorted=sorted, fields=fields)


def gather_map(ctx, group=None) -> dict | None:
    """All ranks export their table; rank 0 returns the merged map, the others None."""
    import torch.distributed as dist
    part = ctx.map_export(sorted=False)
    world = dist.get_world_size(group)
    if world == 1:
        return merge_exports([part])
    parts = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(part, parts, dst=0, group=group)
    return merge_exports(parts) if parts is not None else None
