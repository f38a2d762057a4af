"""-m gpu: the DROP-IN under the reference's own class and call sites.  oracle/_ref/libdropin.so (prebuilt in the build container, it
travels with the snapshot) holds the reference's FrameReader::next (src/rgbdframe.cpp, compiled untouched) and the reference's
rgbd_tutor::Mapper class definition (include/mapper.h) with semantic_slam_mapping_b200/host/dropin/stereo.cpp + mapper.cpp in place of
src/stereo.cpp + src/mapper.cpp -- i.e. the reference's calDisparity_SGBM call (src/rgbdframe.cpp:82) and generatePointCloud land in
libssm.so.  Outputs must equal what the same harness produced over the reference's own sources (tests/golden/mapper_ref.npz,
cues_ref.npz) and the oracle."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drop():
    h = oracle.dropin()
    if h is None:
        pytest.skip("oracle/_ref/libdropin.so is not available")
    return h


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_reference_frame_reader_and_mapper_run_on_libssm(drop, golden_dir):
    g = np.load(os.path.join(golden_dir, "mapper_ref.npz"))
    mp = oracle.MapParams()
    for name in ("frame0", "frame1"):
        depth, disp = oracle.ref_frame_next(g[f"{name}/left"], g[f"{name}/right"], g[f"{name}/rgb"], g[f"{name}/semantic"], mp, handle=drop)
        assert (disp == g[f"{name}/disp"]).all()          # calDisparity_SGBM(cv::Mat) -> ssm_sgbm, 80 disparities
        assert (depth == g[f"{name}/depth"]).all()        # the reference's own depth loop on top of it
    for name in ("frame0", "frame1", "adv"):
        c = oracle.ref_mapper_cloud(g[f"{name}/depth"], g[f"{name}/semantic"], g[f"{name}/rgb"], mp, g[f"{name}/pose"], handle=drop)
        assert (c["mask"] == g[f"{name}/mask"]).all()
        assert c["xyz"].shape == g[f"{name}/xyz"].shape
        assert (_bits(c["xyz_cam"]) == _bits(g[f"{name}/xyz_cam"])).all() and (_bits(c["xyz"]) == _bits(g[f"{name}/xyz"])).all()
        assert (c["rgba"] == g[f"{name}/rgba"]).all()


def test_dense_cue_dropins_match_the_oracle(drop):
    H, W = 72, 256
    L, R, _ = synth.stereo_pair(H, W, 80, 5)
    disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=80))
    out = np.empty((H, W), np.int16)
    drop.ref_calDisparity_SGBM(L.ctypes.data, R.ctypes.data, W, H, out.ctypes.data)
    assert (out == disp).all()
    f, cx, cy, b = 718.856, 607.1928 * W / 1241, 185.2157 * H / 376, 0.532331858
    xyz = np.empty((H, W, 10), np.float32)
    drop.ref_triangulate10D(L.ctypes.data, disp.ctypes.data, W, H, f, cx, cy, b, 30000.0, -1000.0, 30000.0, xyz.ctypes.data)
    want = oracle.triangulate10d(L, disp, f, cx, cy, b)
    assert (xyz.view(np.uint32) == want.view(np.uint32)).all()
    roi = (20.0, -3.0, 40.0)
    drop.ref_correct3DPoints(xyz.ctypes.data, W, H, roi[0], roi[1], roi[2], 0.02, 0.0)
    want2 = oracle.correct_3d_points(want, roi, 0.02)
    assert (xyz.view(np.uint32) == want2.view(np.uint32)).all()
    m = np.empty((H, W), np.uint8)
    drop.ref_setImageROI(xyz.ctypes.data, W, H, m.ctypes.data)
    assert (m == oracle.set_image_roi(want2)).all()
