"""CPU tests of the dense-motion-cue oracle: the C restatement against vectors produced by the
reference's own src/stereo.cpp and src/uvdisparity.cpp (tests/golden/cues_ref.npz, generator make_golden_stereo.py),
against the live reference build where oracle/_ref/libref_stereo.so exists, and numpy restatements of the
U/V-disparity histograms."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import synth


def _eq(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                                 b.view(np.uint32) if b.dtype == np.float32 else b)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cues_ref.npz"))


def test_reference_sgbm_parameters(gold):
    # what calDisparity_SGBM (src/stereo.cpp:16-28) sets on cv::StereoSGBM == the oracle's defaults
    got = dict(zip([str(k) for k in gold["sgbm_params_keys"]], [int(v) for v in gold["sgbm_params_vals"]]))
    d = oracle.SgbmParams()
    for k, v in got.items():
        assert getattr(d, k) == v, k
    assert got["num_disparities"] == 80 and got["block_size"] == 11 and got["p1"] == 484 and got["p2"] == 3872


def test_triangulate_correct_roi_match_reference_vectors(gold):
    f, cx, cy, b = gold["cam"]
    xyz = oracle.triangulate10d(gold["left"], gold["disp"], f, cx, cy, b)
    assert _eq(xyz, gold["xyz"])                      # bit for bit, infinities included
    assert np.isinf(xyz[..., :3]).any() and (gold["disp"] == 0).any()
    cor = oracle.correct_3d_points(gold["xyz"], tuple(gold["roi"]), gold["pitch"][0], gold["pitch"][1])
    assert _eq(cor, gold["corrected"])
    assert _eq(oracle.set_image_roi(gold["corrected"]), gold["roi_mask"])
    assert 0 < (gold["roi_mask"] > 0).mean() < 1


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libref_stereo.so not built (no /root/reference on this machine)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_equals_live_reference(seed):
    H, W, D = 56, 200, 64
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=D))
    f, cx, cy, b = 700.0 + seed, W / 2 - 0.3, H / 2 + 0.6, 0.5
    want = oracle.ref_triangulate10d(L, disp, f, cx, cy, b)
    got = oracle.triangulate10d(L, disp, f, cx, cy, b)
    assert _eq(got, want)
    roi = (15.0, 2.0, 35.0)
    assert _eq(oracle.correct_3d_points(got, roi, 0.02 * seed, 0.0), oracle.ref_correct_3d_points(want, roi, 0.02 * seed, 0.0))
    cor = oracle.correct_3d_points(got, roi, 0.02 * seed)
    assert _eq(oracle.set_image_roi(cor), oracle.ref_set_image_roi(cor))
    # calDisparity_SGBM itself: the reference's parameter block over the oracle's SGBM == oracle defaults
    assert _eq(oracle.ref_cal_disparity_sgbm(L, R), oracle.sgbm(L, R, oracle.SgbmParams()))


def test_uv_disparity_match_reference_vectors(gold):
    """The U/V-disparity restatement against what the reference's own UVDisparity::calVDisparity / calUDisparity
    (src/uvdisparity.cpp, compiled against oracle/cvstub) produced: histograms, 8-bit maps, record channels 8 and 7."""
    out, vint, v8 = oracle.v_disparity(gold["uv_disp"], gold["uv_xyz"])
    assert _eq(vint, gold["v_int"]) and _eq(v8, gold["v_u8"]) and _eq(out, gold["v_xyz"])
    assert vint.shape[1] == 91 and vint[8, 0] >= 1            # the bin == v_cols spill of row 7 landed on row 8, bin 0
    assert vint.sum() == (gold["uv_disp"] > 0).sum() - 1      # the last row's spill is past the matrix
    out, uint_, u8 = oracle.u_disparity(gold["uv_disp"], gold["uv_corrected"], gold["uv_roi_mask"], gold["uv_ground"])
    assert _eq(uint_, gold["u_int"]) and _eq(u8, gold["u_u8"]) and _eq(out, gold["u_xyz"])
    assert uint_.sum() > 100 and (gold["u_xyz"][..., 7] > 0).any()


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libref_stereo.so not built (no /root/reference on this machine)")
@pytest.mark.parametrize("H,W,D,seed", [(56, 200, 64, 1), (37, 131, 32, 2), (120, 400, 128, 3), (90, 310, 80, 4)])
def test_uv_disparity_equals_live_reference(H, W, D, seed):
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=D))
    disp[1, D + 3] = 16 * (D - 2) + 9                         # a new maximum with fraction .5625: bin index == v_cols
    disp[H - 1, D + 5] = 16 * (D - 2) + 9
    xyz = oracle.ref_triangulate10d(L, disp, 718.856, W / 2.0, H / 2.0, 0.54)
    got, want = oracle.v_disparity(disp, xyz), oracle.ref_v_disparity(disp, xyz)
    assert all(_eq(a, b) for a, b in zip(got, want))
    cor = oracle.ref_correct_3d_points(want[0], (30000.0, -1000.0, 30000.0), 0.02)
    roi = oracle.ref_set_image_roi(cor)
    ground = (np.random.default_rng(seed).integers(0, 3, (H, W)) > 0).astype(np.uint8) * 255
    got, want = oracle.u_disparity(disp, cor, roi, ground), oracle.ref_u_disparity(disp, cor, roi, ground)
    assert all(_eq(a, b) for a, b in zip(got, want))


def _np_v_disparity(disp, W):
    H = disp.shape[0]
    mx = int(disp.max())
    v_cols = max(0, int(np.ceil(mx / 16)))
    flat = np.zeros(H * v_cols + v_cols + 1, np.int64)
    ii, jj = np.nonzero(disp > 0)
    dis = np.rint(disp[ii, jj].astype(np.float32) / np.float32(16)).astype(np.int64)     # round half to even
    ids = np.clip(dis, 0, v_cols)
    np.add.at(flat, ii * v_cols + ids, 1)
    vint = flat[: H * v_cols].reshape(H, v_cols)
    scale = np.float32(255) * np.float32(1.0) / np.float32(W)
    v8 = ((vint.astype(np.float32) * scale).astype(np.int64) & 255).astype(np.uint8)
    return vint.astype(np.int32), v8


def test_v_disparity_semantics():
    H, W, D = 48, 160, 64
    L, R, _ = synth.stereo_pair(H, W, D, 9)
    disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=D))
    disp[2, 9] = 16 * 20 + 8          # 20.5 -> 20: round half to even
    disp[5, 7] = 970                  # 60.625: the global maximum, v_cols = 61 and round() = 61: bin id == v_cols spills
    disp[H - 1, 3] = 970              # ... past the end of the matrix on the last row: dropped
    xyz = oracle.triangulate10d(L, disp, 700.0, 80.0, 24.0, 0.5)
    out, vint, v8 = oracle.v_disparity(disp, xyz)
    want_int, want8 = _np_v_disparity(disp, W)
    assert vint.shape == want_int.shape == (H, 61)
    assert (vint == want_int).all() and (v8 == want8).all()
    assert vint.sum() == (disp > 0).sum() - 1                 # only the last-row spill is lost
    assert vint[6, 0] >= 1                                    # row 5's spill landed on row 6, bin 0
    # channel 8: little-endian int of 4 bytes at flat byte offset v * v_cols + 4 * d of the 8-bit map (zero beyond the end)
    flat8 = np.concatenate([v8.reshape(-1), np.zeros(8 * 61 + 8, np.uint8)])
    d = np.rint(xyz[..., 5]).astype(np.int64)
    v = np.arange(H)[:, None].repeat(W, 1)
    off = v * 61 + 4 * np.maximum(d, 0)
    word = sum(flat8[off + q].astype(np.int64) << (8 * q) for q in range(4))
    word = np.where(word >= 2 ** 31, word - 2 ** 32, word)
    want_c8 = np.where(d > 0, word, 0).astype(np.float32)
    assert _eq(out[..., 8], want_c8)
    assert _eq(out[..., :8], xyz[..., :8])


def test_u_disparity_semantics():
    H, W, D = 48, 160, 64
    L, R, _ = synth.stereo_pair(H, W, D, 10)
    disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=D))
    xyz = oracle.correct_3d_points(oracle.triangulate10d(L, disp, 700.0, 80.0, 24.0, 0.5), (20.0, 3.0, 40.0), 0.02)
    roi = oracle.set_image_roi(xyz)
    ground = (np.random.default_rng(0).random((H, W)) < 0.7).astype(np.uint8) * 255
    out, uint_, u8 = oracle.u_disparity(disp, xyz, roi, ground)
    u_rows = int(np.ceil(disp.max() / 16)) + 1
    want = np.zeros((u_rows, W), np.int64)
    ii, jj = np.nonzero((disp > 0) & (roi > 0) & (ground > 0) & (disp // 16 > 0))
    np.add.at(want, (disp[ii, jj] // 16, jj), 1)
    assert uint_.shape == (u_rows, W) and (uint_ == want).all() and want.sum() > 100
    scale = np.float32(255) * np.float32(1.0) / np.float32(H)
    assert (u8 == ((want.astype(np.float32) * scale).astype(np.int64) & 255).astype(np.uint8)).all()
    d = np.rint(xyz[..., 5]).astype(np.int64)
    u = np.arange(W)[None, :].repeat(H, 0)
    want_c7 = np.where(d >= 0, u8[np.clip(d, 0, u_rows - 1), u], 0).astype(np.float32)
    assert _eq(out[..., 7], want_c7)


def test_empty_and_all_invalid_maps():
    H, W = 8, 24
    img = np.full((H, W), 7, np.uint8)
    disp = np.full((H, W), -16, np.int16)
    xyz = oracle.triangulate10d(img, disp, 700.0, 12.0, 4.0, 0.5)
    assert np.isinf(xyz[..., :3]).all()                       # every pixel equals the minimum
    out, vint, v8 = oracle.v_disparity(disp, xyz)
    assert vint.shape == (H, 0) and (out[..., 8] == 0).all()
    out, uint_, u8 = oracle.u_disparity(disp, xyz, np.ones((H, W), np.uint8), np.ones((H, W), np.uint8))
    assert uint_.shape == (1, W) and uint_.sum() == 0 and (out[..., 7] == 0).all()
