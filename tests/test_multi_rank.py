"""Multi-rank path.  CPU part: world_size-2 gloo processes exercise the host-side logic (frame split, ownership
function mirrored in numpy vs libssm.so's ssm_voxel_owner, bucketing + all-to-all + per-rank fusion == one map).
GPU part (-m gpu, needs >= 2 GPUs): the real thing over NCCL, 1-GPU table == union of the 2-GPU tables."""
import os
import socket

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import distributed as D
from semantic_slam_mapping_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_frames_partitions():
    for n in (0, 1, 7, 100, 4541):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in D.shard_frames(n, r, w)]
            assert got == list(range(n))
            sizes = [len(D.shard_frames(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_numpy_owner_matches_library():
    from semantic_slam_mapping_b200 import build, voxel_owner
    build.build()
    rng = np.random.default_rng(3)
    ijk = rng.integers(-5000, 5000, (400, 3)).astype(np.int32)
    ijk[:5] = [[0, 0, 0], [-1, -1, -1], [7, 8, -9], [(1 << 20) - 1, 0, 0], [-(1 << 20), 5, 5]]
    for w in (1, 2, 4, 8):
        want = np.array([voxel_owner(int(a), int(b), int(c), w) for a, b, c in ijk])
        assert (D.voxel_owner_np(ijk, w) == want).all()


def _cloud(seed, n=3000):
    rng = np.random.default_rng(seed)
    xyz = np.concatenate([rng.normal(0, 3, (n, 2)), rng.uniform(2, 30, (n, 1))], axis=1).astype(np.float32)
    return xyz, rng.integers(0, 1 << 24, n).astype(np.uint32), rng.integers(0, 12, n).astype(np.uint8)


def _gloo_worker(rank, world, port, leaf, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = D.shard_frames(6, rank, world)
        vm = oracle.VoxelMap(leaf, 12)
        for f in frames:
            xyz, rgba, lab = _cloud(100 + f)
            order, counts = D.bucket_points(xyz, rgba, lab, leaf, world)
            rec = np.concatenate([xyz[order].view(np.uint32), rgba[order][:, None], lab[order].astype(np.uint32)[:, None]], axis=1)
            send_counts = torch.from_numpy(counts)
            recv_counts = torch.empty(world, dtype=torch.int64)
            dist.all_to_all_single(recv_counts, send_counts)
            send = torch.from_numpy(rec.copy()).view(torch.int32).contiguous()      # uint32 payload travels as int32
            recv = torch.empty((int(recv_counts.sum()), 5), dtype=torch.int32)
            dist.all_to_all_single(recv, send, output_split_sizes=recv_counts.tolist(), input_split_sizes=counts.tolist())
            got = recv.numpy().view(np.uint32)
            # everything that arrived is mine
            assert (D.voxel_owner_np(D.voxel_ijk(got[:, :3].copy().view(np.float32), leaf), world) == rank).all()
            vm.insert(got[:, :3].copy().view(np.float32), got[:, 3].copy(), got[:, 4].astype(np.uint8))
        ex = vm.export()
        parts = [None] * world if rank == 0 else None
        dist.gather_object({"ijk": ex["ijk"], "xyz": ex["centroid"], "rgba": ex["rgba"], "label": ex["label"], "count": ex["count"],
                            "votes": ex["votes"]}, parts, dst=0)
        if rank == 0:
            q.put(D.merge_exports(parts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_routing_equals_single_map(world):
    """Host-side routing logic (frame split, ownership, bucketing, all-to-all, map merge) over 2 and 3 gloo ranks: the
    merged map is the single-process map, vote for vote (3 ranks: uneven frame shards, a non-power-of-two owner hash)."""
    import torch.multiprocessing as mp
    leaf = 0.1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, leaf, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vm = oracle.VoxelMap(leaf, 12)
    for f in range(6):
        vm.insert(*_cloud(100 + f))
    want = vm.export()
    assert (merged["ijk"] == want["ijk"]).all() and (merged["count"] == want["count"]).all()
    assert (merged["votes"] == want["votes"]).all() and (merged["label"] == want["label"]).all()
    assert (merged["rgba"] == want["rgba"]).all()
    assert np.allclose(merged["xyz"], want["centroid"], rtol=1e-5, atol=1e-5)


# ---- GPU: NCCL all-to-all inside libssm.so --------------------------------------------------------------------
def _nccl_worker(rank, world, port, q, p2p, split=0, n=6):
    if split:
        os.environ["SSM_TUNE3"] = str(split)     # sub-batch streams: SGBM per sub-batch, one routing exchange per batch
    import torch
    import torch.distributed as dist
    from semantic_slam_mapping_b200 import Context, Params
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        H, W, Dd = 96, 320, 64
        p = Params(num_disparities=Dd, max_width=W, max_height=H, max_batch=3, resolution=0.05, map_capacity=1 << 18)
        seq = synth.sequence(n, H, W, Dd, 12, seed=9)
        with Context(p, device=rank) as ctx:
            D.init_comm(ctx, p2p=p2p)
            mine = D.shard_frames(n, rank, world)
            sl = slice(mine.start, mine.stop)
            if split == 2:
                # streaming: route overlap on, two async batches (the second one's SGBM overlaps the first one's exchange)
                ctx.set_route_overlap(True)
                half = (mine.stop - mine.start + 1) // 2
                keep = []                                   # host buffers stay alive until synchronize()
                for a, b in ((mine.start, mine.start + half), (mine.start + half, mine.stop)):
                    pins = [np.ascontiguousarray(seq[k][a:b]) for k in ("left", "right", "semantic", "rgb", "pose")]
                    keep.append(pins)
                    ctx.pipeline_batch_host_async(*pins)
                ctx.synchronize()
            else:
                ctx.pipeline_batch_host(seq["left"][sl], seq["right"][sl], seq["semantic"][sl], seq["rgb"][sl], seq["pose"][sl])
            merged = D.gather_map(ctx)
            native = D.gather_map_native(ctx)     # the C-ABI gather (NCCL records -> rank 0 -> device finalize + sort)
            if rank == 0:
                assert set(native) == set(merged)
                for k in merged:
                    assert native[k].shape == merged[k].shape and (native[k] == merged[k]).all(), k
        if rank == 0:
            q.put(merged)
    finally:
        dist.destroy_process_group()


def _run_ranks_and_compare(world, p2p, split, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q, p2p, split, n)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    H, W, Dd = 96, 320, 64
    seq = synth.sequence(n, H, W, Dd, 12, seed=9)
    mp_ = oracle.MapParams()
    vm = oracle.VoxelMap(0.05, 12)
    for i in range(n):
        d = oracle.sgbm(seq["left"][i], seq["right"][i], oracle.SgbmParams(num_disparities=Dd))
        pc = oracle.generate_point_cloud(oracle.disparity_to_depth(d, mp_), seq["semantic"][i], seq["rgb"][i], mp_, seq["pose"][i])
        vm.insert(pc["xyz"], pc["rgba"], pc["label"])
    want = vm.export()
    assert (merged["ijk"] == want["ijk"]).all() and (merged["count"] == want["count"]).all()
    assert (merged["votes"] == want["votes"]).all() and (merged["label"] == want["label"]).all()
    ref = want["centroid_d"]
    assert (np.abs(merged["xyz"] - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("p2p,split", [(True, 0), (False, 0), (True, 2), (False, 3)])
def test_two_gpu_map_equals_oracle_map(p2p, split):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _run_ranks_and_compare(2, p2p, split, 6)


@pytest.mark.gpu
@pytest.mark.parametrize("world,p2p", [(4, True), (4, False), (8, True)])
def test_many_gpu_map_equals_oracle_map(world, p2p):
    """The spatially owned map over 4 / 8 ranks (every rank routes to every other one: one system-scope reservation per
    (CTA, owner), the NCCL all-to-all fallback) is the oracle's single map, vote for vote."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run under gpurun --gpus {world})")
    _run_ranks_and_compare(world, p2p, 0, 8 if world == 4 else 16)
