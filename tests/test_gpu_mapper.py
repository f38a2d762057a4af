"""-m gpu parity tests of the mapper half and the whole path vs the CPU oracle."""
import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import Context, Params, synth

pytestmark = pytest.mark.gpu


def _mp(p: Params):
    return oracle.MapParams(cx=p.cx, cy=p.cy, fx=p.fx, fy=p.fy, baseline=p.baseline, scale=p.scale, roix=p.roix,
                            roiy=p.roiy, roiz=p.roiz, max_distance=p.max_distance, palette_bgr=list(p.palette_bgr),
                            drop_mask=p.drop_mask, dynamic_mask=p.dynamic_mask, dilate_iterations=p.dilate_iterations,
                            colour_source=p.colour_source)


def _op(p: Params):
    return oracle.SgbmParams(num_disparities=p.num_disparities)


def _frame(H, W, D, seed, p):
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    _, sem = synth.label_mask(H, W, p.num_labels, seed, cell=16)
    rgb = np.repeat(L[..., None], 3, axis=-1)
    return L, R, sem, rgb


def _compare_maps(got, want, tol=1e-5):
    assert got["ijk"].shape == want["ijk"].shape
    assert (got["ijk"] == want["ijk"]).all()
    assert (got["count"] == want["count"]).all()
    assert (got["votes"] == want["votes"]).all()
    assert (got["label"] == want["label"]).all()
    assert (got["rgba"] == want["rgba"]).all()
    ref = want["centroid_d"]
    err = np.abs(got["xyz"].astype(np.float64) - ref)
    assert (err <= tol * np.maximum(np.abs(ref), 1.0)).all()      # 1e-5 relative (absolute below 1 m)
    err32 = np.abs(got["xyz"].astype(np.float64) - want["centroid"].astype(np.float64))
    assert (err32 <= 1e-4 * np.maximum(np.abs(ref), 1.0)).all()   # PCL-style fp32 sequential sums, looser


def test_depth_mask_cloud_bit_exact():
    H, W, D = 120, 400, 64
    p = Params(num_disparities=D, max_width=W, max_height=H)
    mp = _mp(p)
    L, R, sem, rgb = _frame(H, W, D, 3, p)
    disp = oracle.sgbm(L, R, _op(p))
    T = synth.poses(4, 1)[3]
    with Context(p) as ctx:
        depth = ctx.disparity_to_depth(disp)
        assert (depth == oracle.disparity_to_depth(disp, mp)).all()
        assert (ctx.semantic_motion_fuse(sem) == oracle.moving_mask(sem, mp)).all()
        pc = ctx.generate_point_cloud(depth, sem, rgb, T)
    want = oracle.generate_point_cloud(depth, sem, rgb, mp, T)
    assert len(want["xyz"]) > 1000
    assert pc["xyz"].shape == want["xyz"].shape
    assert (pc["xyz"].view(np.uint32) == want["xyz"].view(np.uint32)).all()   # fp32 coordinates bit for bit
    assert (pc["rgba"] == want["rgba"]).all() and (pc["label"] == want["label"]).all()


def test_reference_compiled_mapper_vectors(golden_dir):
    """tests/golden/mapper_ref.npz holds what the REFERENCE's own compiled FrameReader::next / Mapper::semantic_motion_fuse /
    Mapper::generatePointCloud / RGBDFrame::project2dTo3d produce (oracle/_ref/libref_mapper.so, generator
    tests/golden/make_golden_mapper.py): the CUDA path reproduces them bit for bit -- disparity at the reference's 80
    disparities, depth image, moving mask, camera-space cloud (identity pose), colours and the transformed cloud."""
    import os
    g = np.load(os.path.join(golden_dir, "mapper_ref.npz"))
    for name in ("frame0", "frame1", "adv"):
        sem, rgb, T = g[f"{name}/semantic"], g[f"{name}/rgb"], g[f"{name}/pose"]
        H, W = sem.shape[:2]
        p = Params(num_disparities=80, max_width=max(W, 96), max_height=H)
        with Context(p) as ctx:
            if name != "adv":
                disp = ctx.sgbm(g[f"{name}/left"], g[f"{name}/right"])
                assert (disp == g[f"{name}/disp"]).all()
                depth = ctx.disparity_to_depth(disp)
                assert (depth == g[f"{name}/depth"]).all()
            depth = g[f"{name}/depth"]
            assert (ctx.semantic_motion_fuse(sem) == g[f"{name}/mask"]).all()
            cam = ctx.generate_point_cloud(depth, sem, rgb, np.eye(4))
            world = ctx.generate_point_cloud(depth, sem, rgb, T)
        assert cam["xyz"].shape == g[f"{name}/xyz_cam"].shape
        assert (cam["xyz"].view(np.uint32) == g[f"{name}/xyz_cam"].view(np.uint32)).all()
        assert (world["xyz"].view(np.uint32) == g[f"{name}/xyz"].view(np.uint32)).all()
        assert (cam["rgba"] == g[f"{name}/rgba"]).all() and (world["rgba"] == g[f"{name}/rgba"]).all()


def test_semantic_colour_variant_and_unknown_colours():
    H, W, D = 60, 200, 32
    p = Params(num_disparities=D, max_width=W, max_height=H, colour_source=1)
    mp = _mp(p)
    rng = np.random.default_rng(0)
    depth = rng.integers(0, 45000, (H, W)).astype(np.uint16)
    _, sem = synth.label_mask(H, W, 12, 5, cell=8)
    sem[rng.random((H, W)) < 0.05] = (1, 2, 3)   # colours outside the palette -> label 255, still mapped
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    with Context(p) as ctx:
        pc = ctx.generate_point_cloud(depth, sem, rgb, np.eye(4))
    want = oracle.generate_point_cloud(depth, sem, rgb, mp, np.eye(4))
    assert (pc["xyz"].view(np.uint32) == want["xyz"].view(np.uint32)).all()
    assert (pc["rgba"] == want["rgba"]).all() and (pc["label"] == want["label"]).all()
    assert (pc["label"] == 255).any()


@pytest.mark.parametrize("leaf", [0.1, 0.05, 0.02])
def test_voxel_fusion_counts_votes_exact(leaf):
    H, W, D, n = 96, 320, 64, 4
    p = Params(num_disparities=D, max_width=W, max_height=H, resolution=leaf, map_capacity=1 << 18)
    mp = _mp(p)
    poses = synth.poses(n, 2)
    vm = oracle.VoxelMap(leaf, p.num_labels)
    with Context(p) as ctx:
        for i in range(n):
            L, R, sem, rgb = _frame(H, W, D, 40 + i, p)
            depth = oracle.disparity_to_depth(oracle.sgbm(L, R, _op(p)), mp)
            ctx.map_integrate_frame(depth, sem, rgb, poses[i])
            pc = oracle.generate_point_cloud(depth, sem, rgb, mp, poses[i])
            vm.insert(pc["xyz"], pc["rgba"], pc["label"])
        assert ctx.map_size() == len(vm)
        got = ctx.map_export(sorted=True)
        # idempotent clear + explicit-cloud integration gives the same table
        ctx.map_clear()
        assert ctx.map_size() == 0
    want = vm.export()
    assert want["count"].sum() > 5000
    _compare_maps(got, want)


def test_whole_path_batch_host_vs_oracle(tmp_path):
    H, W, D, B = 96, 320, 64, 3
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 18)
    mp = _mp(p)
    seq = synth.sequence(B, H, W, D, 12, seed=3)
    vm = oracle.VoxelMap(p.resolution, p.num_labels)
    disps = []
    for i in range(B):
        d = oracle.sgbm(seq["left"][i], seq["right"][i], _op(p))
        disps.append(d)
        depth = oracle.disparity_to_depth(d, mp)
        pc = oracle.generate_point_cloud(depth, seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i])
        vm.insert(pc["xyz"], pc["rgba"], pc["label"])
    with Context(p) as ctx:
        nvox, disp = ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"], want_disp=True)
        got = ctx.map_export()
        path = str(tmp_path / "map.pcd")
        ctx.map_save_pcd(path)
        assert ctx.kernel_launches() > 0
    assert int((disp != np.stack(disps)).sum()) == 0
    assert nvox == len(vm)
    _compare_maps(got, vm.export())
    head = open(path, "rb").read(200).decode("ascii", "ignore")
    assert "FIELDS x y z rgba" in head and f"POINTS {nvox}" in head


def test_table_full_is_reported(monkeypatch):
    """SSM_NO_GROW=1 pins the table at ssm_params.map_capacity: a full table (and spill list) is SSM_ERR_CAPACITY."""
    from semantic_slam_mapping_b200 import SsmError
    monkeypatch.setenv("SSM_NO_GROW", "1")
    p = Params(num_disparities=32, max_width=64, max_height=16, resolution=0.02, map_capacity=1024)
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-20, 20, (20000, 3)).astype(np.float32)
    with Context(p) as ctx:
        with pytest.raises(SsmError) as e:
            ctx.map_integrate_points(xyz, np.zeros(20000, np.uint32), np.zeros(20000, np.uint8))
        assert e.value.code == -4


def _random_cloud(rng, n, span):
    xyz = rng.uniform(-span, span, (n, 3)).astype(np.float32)
    xyz[: n // 4] = np.round(xyz[: n // 4] * 4) / 4            # many points sharing voxels
    return xyz, rng.integers(0, 1 << 24, n).astype(np.uint32), rng.integers(0, 12, n).astype(np.uint8)


@pytest.mark.parametrize("span,leaf", [(30.0, 0.05), (900.0, 0.02), (3.0, 0.1)])
def test_table_grows_and_device_export_orders_like_pcl(span, leaf):
    """A 1024-slot table takes 60 k random points: the spill list + growth steps keep every point (map == oracle), and the
    device radix sort orders the export by (k, j, i) over 1 to 7 digit passes (extent from 60 to 90 000 cells per axis)."""
    p = Params(num_disparities=32, max_width=64, max_height=64, resolution=leaf, map_capacity=1024)
    rng = np.random.default_rng(5)
    vm = oracle.VoxelMap(leaf, p.num_labels)
    with Context(p) as ctx:
        for _ in range(3):
            xyz, rgba, lab = _random_cloud(rng, 20000, span)
            ctx.map_integrate_points(xyz, rgba, lab)
            vm.insert(xyz, rgba & 0xffffff, lab)
        st = ctx.map_stats()
        got = ctx.map_export(sorted=True)
        raw = ctx.map_export(sorted=False)
        part = ctx.map_export(sorted=True, fields=("xyz", "label"))
        assert ctx.map_export_device_ms() > 0
    want = vm.export()
    assert st["grow_steps"] >= 1 and st["voxels"] == len(vm) and st["load_factor"] <= 0.5 and st["slots"] >= 2 * len(vm)
    assert st["max_probe"] < 1024 and st["table_bytes"] == 128 * st["slots"]
    _compare_maps(got, want)
    order = np.lexsort((raw["ijk"][:, 0], raw["ijk"][:, 1], raw["ijk"][:, 2]))
    for k in got:
        assert (raw[k][order] == got[k]).all()
    assert set(part) == {"xyz", "label"} and (part["xyz"] == got["xyz"]).all() and (part["label"] == got["label"]).all()


def test_streaming_growth_from_tiny_table():
    """Streaming entry point with a table far too small for the sequence: the pinned counter mirror drives the growth steps
    between batches, points that found no slot wait in the spill list; the final map is the oracle's, vote for vote."""
    import torch
    H, W, D, B, nb = 96, 320, 64, 2, 4
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.02, map_capacity=1024)
    mp = _mp(p)
    seq = synth.sequence(B * nb, H, W, D, 12, seed=31)
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in seq.items() if k != "label"}
    vm = oracle.VoxelMap(p.resolution, p.num_labels)
    for i in range(B * nb):
        d = oracle.sgbm(seq["left"][i], seq["right"][i], _op(p))
        pc = oracle.generate_point_cloud(oracle.disparity_to_depth(d, mp), seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i])
        vm.insert(pc["xyz"], pc["rgba"], pc["label"])
    assert len(vm) > 20 * 1024
    with Context(p) as ctx:
        for k in range(nb):
            sl = slice(k * B, (k + 1) * B)
            ctx.pipeline_batch_host_async(pin["left"][sl].numpy(), pin["right"][sl].numpy(), pin["semantic"][sl].numpy(),
                                          pin["rgb"][sl].numpy(), pin["pose"][sl].numpy(), None)
        ctx.synchronize()
        assert ctx.map_size() == len(vm)
        st = ctx.map_stats()
        got = ctx.map_export()
        ctx.map_reserve(1 << 20)
        assert ctx.map_stats()["slots"] == 1 << 20
        again = ctx.map_export()
    assert st["grow_steps"] >= 2
    _compare_maps(got, vm.export())
    for k in got:
        assert (again[k] == got[k]).all()


def test_export_into_pinned_arrays_and_empty_map():
    import torch
    p = Params(num_disparities=32, max_width=64, max_height=64, resolution=0.1, map_capacity=1 << 12)
    rng = np.random.default_rng(9)
    xyz, rgba, lab = _random_cloud(rng, 5000, 5.0)
    with Context(p) as ctx:
        empty = ctx.map_export()
        assert all(len(v) == 0 for v in empty.values())
        ctx.map_integrate_points(xyz, rgba, lab)
        n = ctx.map_size()
        pinned = {"xyz": torch.empty((n + 7, 3), dtype=torch.float32).pin_memory().numpy(),
                  "rgba": torch.empty(n + 7, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
        got = ctx.map_export(into=pinned)
        ref = ctx.map_export()
        with pytest.raises(ValueError):
            ctx.map_export(into={"xyz": np.empty((n - 1, 3), np.float32)})
    assert len(got["xyz"]) == n and (got["xyz"] == ref["xyz"]).all() and (got["rgba"] == ref["rgba"]).all()


def test_async_host_pipeline_matches_oracle():
    """Streaming entry point: three batches through the double-buffered staging; same map as the oracle."""
    import torch
    H, W, D, B, nb = 96, 320, 64, 2, 3
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 18)
    mp = _mp(p)
    seq = synth.sequence(B * nb, H, W, D, 12, seed=23)
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in seq.items() if k != "label"}
    res = torch.zeros(nb, dtype=torch.int32).pin_memory()
    vm = oracle.VoxelMap(p.resolution, p.num_labels)
    sizes = []
    for i in range(B * nb):
        d = oracle.sgbm(seq["left"][i], seq["right"][i], _op(p))
        pc = oracle.generate_point_cloud(oracle.disparity_to_depth(d, mp), seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i])
        vm.insert(pc["xyz"], pc["rgba"], pc["label"])
        if i % B == B - 1:
            sizes.append(len(vm))
    with Context(p) as ctx:
        for k in range(nb):
            sl = slice(k * B, (k + 1) * B)
            ctx.pipeline_batch_host_async(pin["left"][sl].numpy(), pin["right"][sl].numpy(), pin["semantic"][sl].numpy(),
                                          pin["rgb"][sl].numpy(), pin["pose"][sl].numpy(), res[k:k + 1])
        ctx.synchronize()
        got = ctx.map_export()
    assert res.tolist() == sizes
    _compare_maps(got, vm.export())


def test_split_batch_streams_match_oracle(monkeypatch):
    """SSM_TUNE3 = 2: the batch runs as two sub-batches on two streams."""
    monkeypatch.setenv("SSM_TUNE3", "2")
    H, W, D, B = 96, 320, 64, 5
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 18)
    mp = _mp(p)
    seq = synth.sequence(B, H, W, D, 12, seed=29)
    vm = oracle.VoxelMap(p.resolution, p.num_labels)
    disps = []
    for i in range(B):
        d = oracle.sgbm(seq["left"][i], seq["right"][i], _op(p))
        disps.append(d)
        pc = oracle.generate_point_cloud(oracle.disparity_to_depth(d, mp), seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i])
        vm.insert(pc["xyz"], pc["rgba"], pc["label"])
    with Context(p) as ctx:
        nvox, disp = ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"], want_disp=True)
        got = ctx.map_export()
    assert int((disp != np.stack(disps)).sum()) == 0 and nvox == len(vm)
    _compare_maps(got, vm.export())


def test_keyframe_cache_redraw_after_pose_update():
    """Mapper::viewer's periodic redraw (mapper.cpp:121-131) from cached camera-space clouds (:17-20, :90-91) after the
    pose graph rewrote the keyframe poses (pose_graph.cpp:253-260): the map equals one built from scratch."""
    H, W, D, N = 96, 320, 64, 6
    p = Params(num_disparities=D, max_width=W, max_height=H, resolution=0.05, map_capacity=1 << 18)
    mp = _mp(p)
    frames = []
    for i in range(N):
        L, R, sem, rgb = _frame(H, W, D, 40 + i, p)
        depth = oracle.disparity_to_depth(oracle.sgbm(L, R, _op(p)), mp)
        frames.append((depth, sem, rgb))
    poses0 = synth.poses(N, 1)
    # "optimised" poses: a different trajectory seed plus a small extra rotation / shift per frame
    poses1 = synth.poses(N, 2).copy()
    for i in range(N):
        a = 0.01 * (i + 1)
        Rz = np.array([[np.cos(a), -np.sin(a), 0, 0.03 * i], [np.sin(a), np.cos(a), 0, -0.02 * i], [0, 0, 1, 0.01], [0, 0, 0, 1]])
        poses1[i] = Rz @ poses1[i]

    def oracle_map(ids, poses):
        vm = oracle.VoxelMap(p.resolution, p.num_labels)
        for i in ids:
            pc = oracle.generate_point_cloud(*frames[i], mp, poses[i])
            vm.insert(pc["xyz"], pc["rgba"], pc["label"])
        return vm.export()

    with Context(p) as ctx:
        ids = [ctx.keyframe_add(*frames[i], poses0[i]) for i in range(N)]
        n_live, n_pts = ctx.keyframe_count()
        assert n_live == N and n_pts == sum(len(oracle.generate_point_cloud(*frames[i], mp, poses0[i])["xyz"]) for i in range(N))
        ctx.map_redraw()                                     # every keyframe, original poses
        _compare_maps(ctx.map_export(), oracle_map(range(N), poses0))
        for i in range(N):
            ctx.keyframe_set_pose(ids[i], poses1[i])
        ctx.map_redraw(ids[::2])                             # the reference redraws every second keyframe (i += 2)
        _compare_maps(ctx.map_export(), oracle_map(range(0, N, 2), poses1))
        ctx.map_integrate_keyframes(ids[1::2])               # incremental branch: += without clearing
        _compare_maps(ctx.map_export(), oracle_map(list(range(0, N, 2)) + list(range(1, N, 2)), poses1))
        ctx.keyframe_release(ids[0])
        assert ctx.keyframe_count()[0] == N - 1
        ctx.map_redraw()
        _compare_maps(ctx.map_export(), oracle_map(range(1, N), poses1))
        with pytest.raises(Exception):
            ctx.keyframe_set_pose(ids[0], poses1[0])         # released id


def test_full_size_map_equals_oracle_map():
    """Full KITTI-size frames (1241 x 376, 128 disparities) through the whole path: the fused map equals the oracle's voxel for
    voxel -- voxel set, counts, votes, majority labels, colours exactly, centroids within 1e-5 -- at 0.05 m and, re-fused, at 0.02 m
    and at the 0.1 m of BASELINE configs[0] (the reference's own mapper_resolution, parameters.txt:97)."""
    H, W, D, B = 376, 1241, 128, 4
    seq = synth.sequence(B, H, W, D, 12, seed=41)
    clouds = None
    for leaf in (0.05, 0.02, 0.1):
        p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=leaf, map_capacity=1 << 21)
        mp = _mp(p)
        if clouds is None:
            clouds = []
            for i in range(B):
                d = oracle.sgbm(seq["left"][i], seq["right"][i], _op(p))
                clouds.append((d, oracle.generate_point_cloud(oracle.disparity_to_depth(d, mp), seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i])))
        vm = oracle.VoxelMap(leaf, p.num_labels)
        for _, pc in clouds:
            vm.insert(pc["xyz"], pc["rgba"], pc["label"])
        with Context(p) as ctx:
            nvox, disp = ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"], want_disp=True)
            got = ctx.map_export()
        assert all(int((disp[i] != clouds[i][0]).sum()) == 0 for i in range(B))
        assert nvox == len(vm) > 50000
        _compare_maps(got, vm.export())


def test_map_is_independent_of_frame_order_and_batch_split():
    """All accumulators are integers, so the fused map does not depend on the order in which frames arrive or on how a sequence
    is cut into calls: one batch of six frames == the same frames reversed == three calls of two frames == six single-frame
    calls through the drop-in entry points (counts, votes, labels, colours and centroids bit for bit)."""
    H, W, D, B = 128, 416, 64, 6
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 18)
    seq = synth.sequence(B, H, W, D, 12, seed=77)
    keys = ("left", "right", "semantic", "rgb", "pose")

    def run(order, chunk):
        with Context(p) as ctx:
            for a in range(0, B, chunk):
                idx = order[a:a + chunk]
                ctx.pipeline_batch_host(*[np.ascontiguousarray(seq[k][idx]) for k in keys])
            return ctx.map_export()

    def run_single_frame_calls():
        with Context(p) as ctx:
            for i in range(B):
                depth = ctx.disparity_to_depth(ctx.sgbm(seq["left"][i], seq["right"][i]))
                ctx.map_integrate_frame(depth, seq["semantic"][i], seq["rgb"][i], seq["pose"][i])
            return ctx.map_export()

    fwd = list(range(B))
    ref = run(fwd, B)
    assert len(ref["count"]) > 10000
    for other in (run(fwd[::-1], B), run(fwd, 2), run([3, 0, 5, 1, 4, 2], 3), run_single_frame_calls()):
        for k in ("ijk", "count", "votes", "label", "rgba", "xyz"):
            assert other[k].shape == ref[k].shape and (other[k] == ref[k]).all(), k


def test_full_size_stress_properties_19_classes_2cm_voxels():
    """BASELINE configs[3]/[4] flavour at full KITTI size: 19-class palette, 0.02 m voxels, a batch through the whole path.
    Size-independent properties: every generated point lands in exactly one voxel (sum of counts == points), label votes
    never exceed counts, re-fusing the same batch doubles every count and leaves voxel membership unchanged, and frame 0's
    disparity equals the oracle's."""
    from semantic_slam_mapping_b200.params import CITYSCAPES19_BGR
    H, W, D, B, L = 376, 1241, 128, 6, 19
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.02, map_capacity=1 << 22,
               palette_bgr=list(CITYSCAPES19_BGR), drop_mask=1 << 0, dynamic_mask=(1 << 11) | (1 << 12))
    seq = synth.sequence(B, H, W, D, L, seed=23)             # 19-class masks rendered through the Cityscapes palette
    mp = _mp(p)
    with Context(p) as ctx:
        nvox, disp = ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"], want_disp=True)
        m1 = ctx.map_export()
        npts = 0
        for i in range(B):
            depth = ctx.disparity_to_depth(disp[i])
            npts += len(ctx.generate_point_cloud(depth, seq["semantic"][i], seq["rgb"][i], seq["pose"][i])["xyz"])
        assert nvox == len(m1["count"]) and nvox > 100000
        assert int(m1["count"].sum()) == npts
        assert (m1["votes"].sum(axis=1) <= m1["count"]).all() and int(m1["votes"].sum()) == npts   # every kept point has a palette label
        assert (m1["label"] < L).all() and not (m1["votes"][:, 0] > 0).any()                      # class 0 is dropped from the cloud
        ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"])
        m2 = ctx.map_export()
        assert (m2["ijk"] == m1["ijk"]).all() and (m2["count"] == 2 * m1["count"]).all() and (m2["votes"] == 2 * m1["votes"]).all()
        assert (m2["label"] == m1["label"]).all() and (m2["rgba"] == m1["rgba"]).all()
        assert np.allclose(m2["xyz"], m1["xyz"], rtol=0, atol=1e-6)
    want0 = oracle.sgbm(seq["left"][0], seq["right"][0], _op(p))
    assert int((disp[0] != want0).sum()) == 0
    pc0 = oracle.generate_point_cloud(oracle.disparity_to_depth(want0, mp), seq["semantic"][0], seq["rgb"][0], mp, seq["pose"][0])
    assert len(pc0["xyz"]) > 50000
