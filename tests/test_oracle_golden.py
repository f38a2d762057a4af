"""Pins the CPU oracle against the committed cv2-generated golden vectors (tests/golden/make_golden.py)
and, when cv2 is importable, against cv2 itself on fresh seeds.  CPU only."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import synth


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "sgbm_small.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def test_sgbm_small_golden(golden_dir):
    z, names = _cases(golden_dir)
    assert len(names) >= 8
    for n in names:
        D, bs, uniq, spw, spr, d12, cap = [int(v) for v in z[f"{n}/params"]]
        p = oracle.SgbmParams(num_disparities=D, block_size=bs, p1=4 * bs * bs, p2=32 * bs * bs, uniqueness_ratio=uniq,
                              speckle_window_size=spw, speckle_range=spr, disp12_max_diff=d12, pre_filter_cap=cap)
        got = oracle.sgbm(z[f"{n}/left"], z[f"{n}/right"], p)
        want = z[f"{n}/disp"]
        assert got.dtype == np.int16 and got.shape == want.shape
        assert int((got != want).sum()) == 0, n
        # SURVEY App. A output facts: columns [0, D) are always invalid
        assert (got[:, :D] == -16).all()


def test_sgbm_kitti_d128_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "sgbm_kitti_d128.npz"))
    L, R, _ = synth.stereo_pair(376, 1241, 128, 0)
    assert int(L.astype(np.int64).sum()) == int(z["left_sum"]) and int(R.astype(np.int64).sum()) == int(z["right_sum"])
    got = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=128))
    assert int((got != z["disp"]).sum()) == 0
    assert (got >= 0).mean() > 0.8  # the synthetic pair is well-textured


def test_stage_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "stages.npz"))
    assert (oracle.median3x3(z["median_in"]) == z["median_out"]).all()
    assert (oracle.filter_speckles(z["speckle_in"], -16, 20, 64) == z["speckle_out"]).all()
    mp = oracle.MapParams()
    # moving mask == dilate(dynamic-class mask, ones 3x3, 2 iterations): render the mask as a semantic image
    sem = np.zeros(z["dilate_in"].shape + (3,), np.uint8)
    sem[z["dilate_in"] == 255] = mp.palette_bgr[10]  # pedestrian
    sem[z["dilate_in"] == 0] = mp.palette_bgr[4]     # road
    assert (oracle.moving_mask(sem, mp) == z["dilate_out"]).all()


def test_sgbm_vs_cv2_fresh_seeds():
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    for H, W, D, seed in [(40, 130, 64, 11), (90, 260, 96, 12)]:
        L, R, _ = synth.stereo_pair(H, W, D, seed)
        p = oracle.SgbmParams(num_disparities=D)
        s = cv2.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=11, P1=p.p1, P2=p.p2, disp12MaxDiff=1,
                                  preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32,
                                  mode=cv2.STEREO_SGBM_MODE_SGBM)
        assert int((oracle.sgbm(L, R, p) != s.compute(L, R)).sum()) == 0


def test_sgbm_rejects_bad_arguments():
    L = np.zeros((10, 20), np.uint8)
    with pytest.raises(ValueError):
        oracle.sgbm(L, L, oracle.SgbmParams(num_disparities=24))  # not a multiple of 16
    with pytest.raises(ValueError):
        oracle.sgbm(L, L, oracle.SgbmParams(num_disparities=32))  # W <= D


def test_depth_glue_known_answers():
    """rgbdframe.cpp:85-116 on hand-computed pixels."""
    mp = oracle.MapParams()
    disp = np.full((4, 700), -16, np.int16)
    disp[1, 650] = 16 * 20      # 20 px -> z = f*b/20
    disp[2, 10] = 16 * 20       # x far outside roix=20? (10-607)*b/20 = -15.9 m -> inside
    disp[3, 600] = 0            # d == 0 skipped
    disp[0, 620] = 8            # half a pixel: z = f*b*16/8 = 765 m > roiz -> 0
    depth = oracle.disparity_to_depth(disp, mp)
    z = mp.fx * mp.baseline / 20.0
    assert depth[1, 650] == int(z * 1000.0)
    assert depth[2, 10] == int(z * 1000.0)
    assert depth[3, 600] == 0 and depth[0, 620] == 0
    assert (depth[disp == -16] == 0).all()


def test_point_cloud_and_voxels_small():
    mp = oracle.MapParams()
    H, W = 6, 8
    depth = np.zeros((H, W), np.uint16)
    depth[2, 3] = 5000
    depth[2, 4] = 5004
    depth[4, 1] = 41000  # > max_distance*scale -> dropped
    depth[5, 5] = 7000
    ids, sem = synth.label_mask(H, W, 12, 0, cell=2)
    sem[:] = mp.palette_bgr[4]
    sem[5, 5] = mp.palette_bgr[0]  # sky -> dropped
    rgb = np.full((H, W, 3), 90, np.uint8)
    T = np.eye(4)
    T[0, 3] = 1.5
    pc = oracle.generate_point_cloud(depth, sem, rgb, mp, T)
    assert pc["pix"].tolist() == [2 * W + 3, 2 * W + 4]
    z = np.float32(5000 / 1000.0)
    x = np.float32((3 - mp.cx) * float(z) / mp.fx)
    assert pc["xyz_cam"][0, 2] == z and pc["xyz_cam"][0, 0] == x
    assert pc["xyz"][0, 0] == np.float32(float(x) + 1.5)
    assert (pc["label"] == 4).all() and (pc["rgba"] == 0x5A5A5A).all()
    vm = oracle.VoxelMap(0.1, 12)
    vm.insert(pc["xyz"], pc["rgba"], pc["label"])
    ex = vm.export()
    assert ex["count"].sum() == 2 and (ex["votes"].sum(axis=1) == ex["count"]).all()
    assert (ex["label"] == 4).all()


def test_moving_mask_blocks_points():
    mp = oracle.MapParams()
    H, W = 9, 9
    depth = np.full((H, W), 6000, np.uint16)
    sem = np.empty((H, W, 3), np.uint8)
    sem[:] = mp.palette_bgr[4]
    sem[4, 4] = mp.palette_bgr[10]  # one pedestrian pixel -> 5x5 block masked
    pc = oracle.generate_point_cloud(depth, sem, sem, mp, np.eye(4))
    assert len(pc["pix"]) == H * W - 25
    m = oracle.moving_mask(sem, mp)
    assert m.sum() == 25 * 255 and m[2:7, 2:7].all()
