"""The mapper half of the path pinned to the REFERENCE's own compiled sources (oracle/_ref/libref_mapper.so = src/mapper.cpp,
src/rgbdframe.cpp, src/parameter_reader.cpp, src/stereo.cpp compiled untouched from /root/reference against oracle/cvstub +
oracle/refstub): FrameReader::next's disparity -> depth loop, RGBDFrame::project2dTo3d, Mapper::semantic_motion_fuse and
Mapper::generatePointCloud.  The C oracle's restatement must agree bit for bit -- live where the library can be built (this
container) and against the committed vectors tests/golden/mapper_ref.npz (generator make_golden_mapper.py) everywhere."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import synth
from semantic_slam_mapping_b200.params import SEGNET12_BGR


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_committed_vectors_match_the_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "mapper_ref.npz"))
    mp = oracle.MapParams()
    sp = oracle.SgbmParams(num_disparities=80)        # the reference's hard-coded count (src/stereo.cpp:18)
    for name in ("frame0", "frame1"):
        disp = oracle.sgbm(g[f"{name}/left"], g[f"{name}/right"], sp)
        assert (disp == g[f"{name}/disp"]).all()
        depth = oracle.disparity_to_depth(disp, mp)
        assert (depth == g[f"{name}/depth"]).all() and (depth > 0).sum() > 2000
    for name in ("frame0", "frame1", "adv"):
        depth, sem, rgb, T = g[f"{name}/depth"], g[f"{name}/semantic"], g[f"{name}/rgb"], g[f"{name}/pose"]
        assert (oracle.moving_mask(sem, mp) == g[f"{name}/mask"]).all()
        pc = oracle.generate_point_cloud(depth, sem, rgb, mp, T)
        assert len(pc["xyz"]) == len(g[f"{name}/xyz"]) > 500
        assert (_bits(pc["xyz_cam"]) == _bits(g[f"{name}/xyz_cam"])).all()      # project2dTo3d, the reference's own arithmetic
        assert (pc["rgba"] == g[f"{name}/rgba"]).all()                          # colour of the left image, b g r order
        assert (_bits(pc["xyz"]) == _bits(g[f"{name}/xyz"])).all()              # + transformPointCloud (stand-in's written definition)
    assert (g["adv/mask"] == 255).any() and (g["adv/mask"] == 0).any()


needs_ref = pytest.mark.skipif(oracle.ref_mapper() is None, reason="oracle/_ref/libref_mapper.so is not available (no /root/reference and no prebuilt file)")


@needs_ref
@pytest.mark.parametrize("seed,H,W", [(1, 64, 200), (2, 48, 161), (3, 96, 320)])
def test_live_reference_frame_reader_and_mapper(seed, H, W):
    mp = oracle.MapParams()
    L, R, _ = synth.stereo_pair(H, W, 80, 100 + seed)
    _, sem = synth.label_mask(H, W, 12, seed, cell=10)
    rgb = np.stack([L, np.roll(L, 5, axis=0), L // 2], axis=-1)
    depth, disp = oracle.ref_frame_next(L, R, rgb, sem, mp)
    assert oracle.ref_sgbm_params()["num_disparities"] == 80
    want_disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=80))
    assert (disp == want_disp).all()
    assert (depth == oracle.disparity_to_depth(want_disp, mp)).all()
    T = synth.poses(8, seed)[7]
    ref = oracle.ref_mapper_cloud(depth, sem, rgb, mp, T)
    pc = oracle.generate_point_cloud(depth, sem, rgb, mp, T)
    assert (ref["mask"] == oracle.moving_mask(sem, mp)).all()
    assert len(ref["xyz"]) == len(pc["xyz"]) > 300
    assert (_bits(ref["xyz_cam"]) == _bits(pc["xyz_cam"])).all() and (ref["rgba"] == pc["rgba"]).all()
    assert (_bits(ref["xyz"]) == _bits(pc["xyz"])).all()


@needs_ref
def test_live_reference_other_camera_and_depth_edge_values():
    """Another calibration (the commented KITTI-05 block of parameters.txt:44-48), a tighter ROI and max distance, random depth."""
    mp = oracle.MapParams(cx=601.8873, cy=183.1104, fx=707.0912, fy=707.0912, baseline=0.537904488, roix=12.0, roiy=3.0, roiz=25.0,
                          max_distance=17.5)
    rng = np.random.default_rng(5)
    H, W = 50, 120
    depth = rng.integers(0, 30000, (H, W)).astype(np.uint16)
    depth[3, :4] = [17500, 17501, 17499, 0]
    pal = np.asarray(SEGNET12_BGR, np.uint8)
    sem = pal[rng.integers(0, 12, (H, W))]
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    T = synth.poses(3, 9)[2]
    ref = oracle.ref_mapper_cloud(depth, sem, rgb, mp, T)
    pc = oracle.generate_point_cloud(depth, sem, rgb, mp, T)
    assert (ref["mask"] == oracle.moving_mask(sem, mp)).all()
    assert len(ref["xyz"]) == len(pc["xyz"])
    assert (_bits(ref["xyz_cam"]) == _bits(pc["xyz_cam"])).all() and (_bits(ref["xyz"]) == _bits(pc["xyz"])).all() and (ref["rgba"] == pc["rgba"]).all()
    L, R, _ = synth.stereo_pair(40, 180, 80, 77)
    sem2 = pal[rng.integers(0, 12, (40, 180))]
    rgb2 = np.repeat(L[..., None], 3, axis=-1)
    depth2, disp2 = oracle.ref_frame_next(L, R, rgb2, sem2, mp)
    assert (depth2 == oracle.disparity_to_depth(disp2, mp)).all() and (depth2 > 0).any()


@needs_ref
def test_reference_build_flags_do_not_change_the_glue():
    """The reference builds with -march=native -O3 (CMakeLists.txt:11), which leaves GCC's -ffp-contract=fast on; the canonical
    oracle build uses -ffp-contract=off.  The executed glue has no multiply-add a compiler could contract: both builds of the
    reference's own sources give identical depth images and camera-space clouds on this host."""
    if oracle.ref_mapper(native=True) is None:
        pytest.skip("native build unavailable")
    mp = oracle.MapParams()
    L, R, _ = synth.stereo_pair(64, 200, 80, 31)
    _, sem = synth.label_mask(64, 200, 12, 31, cell=10)
    rgb = np.repeat(L[..., None], 3, axis=-1)
    d0, _ = oracle.ref_frame_next(L, R, rgb, sem, mp)
    d1, _ = oracle.ref_frame_next(L, R, rgb, sem, mp, native=True)
    assert (d0 == d1).all()
    T = synth.poses(5, 2)[4]
    a = oracle.ref_mapper_cloud(d0, sem, rgb, mp, T)
    b = oracle.ref_mapper_cloud(d0, sem, rgb, mp, T, native=True)
    assert (_bits(a["xyz_cam"]) == _bits(b["xyz_cam"])).all() and (a["rgba"] == b["rgba"]).all() and (a["mask"] == b["mask"]).all()
