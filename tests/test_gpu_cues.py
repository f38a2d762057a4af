"""-m gpu parity tests of the dense motion cues through the C ABI vs the CPU oracle, which is itself
pinned to the reference's own src/stereo.cpp and src/uvdisparity.cpp (tests/test_oracle_cues.py).  Everything is compared bit for bit (fp32
words, infinities and NaNs included)."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import Context, Params, synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _eq(a, b):
    return a.shape == b.shape and np.array_equal(_bits(np.ascontiguousarray(a)), _bits(np.ascontiguousarray(b)))


def _case(H, W, D, seed):
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    disp = oracle.sgbm(L, R, oracle.SgbmParams(num_disparities=D))
    return L, disp


@pytest.fixture(scope="module")
def ctx():
    with Context(Params(num_disparities=64, max_width=512, max_height=256, max_batch=4, map_capacity=1 << 16)) as c:
        yield c


def test_golden_vectors_from_reference_source(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "cues_ref.npz"))
    f, cx, cy, b = g["cam"]
    xyz = ctx.triangulate10d(g["left"], g["disp"], f, cx, cy, b, roi=tuple(g["roi"]))
    assert _eq(xyz, g["xyz"])
    cor = ctx.correct_3d_points(g["xyz"], tuple(g["roi"]), g["pitch"][0], g["pitch"][1])
    assert _eq(cor, g["corrected"])
    assert _eq(ctx.set_image_roi(g["corrected"]), g["roi_mask"])
    # U/V-disparity histograms: vectors produced by the reference's own UVDisparity::calVDisparity / calUDisparity
    got, vint, v8 = ctx.v_disparity(g["uv_disp"], g["uv_xyz"])
    assert _eq(vint, g["v_int"]) and _eq(v8, g["v_u8"]) and _eq(got, g["v_xyz"])
    got, uint_, u8 = ctx.u_disparity(g["uv_disp"], g["uv_corrected"], g["uv_roi_mask"], g["uv_ground"])
    assert _eq(uint_, g["u_int"]) and _eq(u8, g["u_u8"]) and _eq(got, g["u_xyz"])


@pytest.mark.parametrize("H,W,D,seed", [(56, 200, 64, 1), (37, 131, 32, 2), (120, 400, 128, 3)])
def test_host_entry_points_match_oracle(ctx, H, W, D, seed):
    L, disp = _case(H, W, D, seed)
    disp[2, 9] = 16 * 20 + 8                                  # round half to even
    disp[5, 7] = 16 * (D + 3) + 10                            # global maximum whose bin id == v_cols (spill into the next row)
    disp[H - 1, 3] = 16 * (D + 3) + 10                        # ... dropped past the end of the matrix
    f, cx, cy, b = 707.09, W / 2 - 0.4, H / 2 + 0.3, 0.537
    roi = (12.0, 1.2, 30.0)
    xyz = ctx.triangulate10d(L, disp, f, cx, cy, b)
    want = oracle.triangulate10d(L, disp, f, cx, cy, b)
    assert _eq(xyz, want) and np.isinf(want[..., :3]).any()
    got_v, vint, v8 = ctx.v_disparity(disp, xyz)
    want_v, wint, w8 = oracle.v_disparity(disp, want)
    assert _eq(vint, wint) and _eq(v8, w8) and _eq(got_v, want_v) and wint.sum() > 0
    got_c = ctx.correct_3d_points(got_v, roi, 0.027, 0.0)
    want_c = oracle.correct_3d_points(want_v, roi, 0.027, 0.0)
    assert _eq(got_c, want_c)
    mask = ctx.set_image_roi(got_c)
    assert _eq(mask, oracle.set_image_roi(want_c)) and 0 < (mask > 0).mean() < 1
    ground = (np.random.default_rng(seed).random((H, W)) < 0.8).astype(np.uint8) * 255
    got_u, uint_, u8 = ctx.u_disparity(disp, got_c, mask, ground)
    want_u, wuint, wu8 = oracle.u_disparity(disp, want_c, mask, ground)
    assert _eq(uint_, wuint) and _eq(u8, wu8) and _eq(got_u, want_u) and wuint.sum() > 0


def test_all_invalid_and_zero_maps(ctx):
    H, W = 9, 40
    img = np.arange(H * W, dtype=np.uint8).reshape(H, W)
    for fill in (-16, 0):
        disp = np.full((H, W), fill, np.int16)
        xyz = ctx.triangulate10d(img, disp, 700.0, 20.0, 4.0, 0.5)
        assert _eq(xyz, oracle.triangulate10d(img, disp, 700.0, 20.0, 4.0, 0.5))
        out, vint, v8 = ctx.v_disparity(disp, xyz)
        assert vint.shape == (H, 0) and _eq(out, oracle.v_disparity(disp, xyz)[0])
        ones = np.ones((H, W), np.uint8)
        out, uint_, u8 = ctx.u_disparity(disp, xyz, ones, ones)
        want = oracle.u_disparity(disp, xyz, ones, ones)
        assert uint_.shape == (1, W) and _eq(out, want[0]) and _eq(uint_, want[1])


def test_device_batch_stages_match_oracle(ctx):
    import torch
    B, H, W, D = 3, 96, 320, 64
    imgs, disps = zip(*[_case(H, W, D, 10 + i) for i in range(B)])
    imgs, disps = np.stack(imgs), np.stack(disps)
    disps[1, 4, 4] = 16 * D + 9                               # frame 1: maximum whose bin id == v_cols
    f, cx, cy, b = 718.856, 160.2, 47.9, 0.532
    roi, pitch = (15.0, 1.5, 35.0), 0.021
    cap = D + 2
    v_stride, u_stride = H * cap + 4, cap * W
    dev = torch.device("cuda:0")
    d_img, d_disp = torch.from_numpy(imgs).to(dev), torch.from_numpy(disps).to(dev)
    d_xyz = torch.empty((B, H, W, 10), dtype=torch.float32, device=dev)
    d_vint = torch.empty((B, v_stride), dtype=torch.int32, device=dev)
    d_v8 = torch.zeros((B, v_stride), dtype=torch.uint8, device=dev)
    d_uint = torch.empty((B, u_stride), dtype=torch.int32, device=dev)
    d_u8 = torch.zeros((B, u_stride), dtype=torch.uint8, device=dev)
    d_roi = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    ground = (np.random.default_rng(3).random((B, H, W)) < 0.75).astype(np.uint8)
    d_ground = torch.from_numpy(ground).to(dev)
    torch.cuda.synchronize()
    ctx.motion_cues_stage1_device(d_img, d_disp, d_xyz, B, W, H, f, cx, cy, b, d_vint, d_v8, v_stride, cap)
    ctx.synchronize()
    xyz1 = d_xyz.cpu().numpy()
    ctx.motion_cues_stage2_device(d_disp, d_xyz, d_roi, B, W, H, roi, pitch, d_ground, d_uint, d_u8, u_stride, cap)
    ctx.synchronize()
    assert ctx.motion_cues_overflow() == 0
    xyz2, roi_mask = d_xyz.cpu().numpy(), d_roi.cpu().numpy()
    vint, uint_ = d_vint.cpu().numpy(), d_uint.cpu().numpy()
    for i in range(B):
        w0 = oracle.triangulate10d(imgs[i], disps[i], f, cx, cy, b)
        w1, wvint, wv8 = oracle.v_disparity(disps[i], w0)
        assert _eq(xyz1[i], w1)
        vc = wvint.shape[1]
        assert _eq(vint[i, : H * vc].reshape(H, vc), wvint)
        w2 = oracle.correct_3d_points(w1, roi, pitch)
        wmask = oracle.set_image_roi(w2)
        w3, wuint, wu8 = oracle.u_disparity(disps[i], w2, wmask, ground[i])
        assert _eq(roi_mask[i], wmask)
        ur = wuint.shape[0]
        assert _eq(uint_[i, : ur * W].reshape(ur, W), wuint)
        assert _eq(xyz2[i], w3)


def test_capacity_errors(ctx):
    from semantic_slam_mapping_b200.lib import SsmError
    import ctypes as C
    H, W = 16, 64
    disp = np.full((H, W), 16 * 50, np.int16)
    vi = np.zeros((H, 4), np.int32)
    vc = C.c_int(0)
    rc = ctx._L.ssm_v_disparity(ctx._h, disp.ctypes.data_as(C.c_void_p), W * 2, W, H, None, vi.ctypes.data_as(C.c_void_p), None, 4, C.byref(vc))
    assert rc == -4 and vc.value == 50                        # SSM_ERR_CAPACITY, size still reported
    with pytest.raises(SsmError):
        ctx.triangulate10d(np.zeros((0, 0), np.uint8), np.zeros((0, 0), np.int16), 1.0, 0.0, 0.0, 1.0)
