#!/usr/bin/env python
"""Regenerates tests/golden/cues_ref.npz by RUNNING THE REFERENCE'S OWN src/stereo.cpp and src/uvdisparity.cpp (compiled
where they lie by oracle/Makefile against oracle/cvstub into oracle/_ref/libref_stereo.so): calDisparity_SGBM's parameter
set, triangulate10D, correct3DPoints, setImageROI and UVDisparity::calVDisparity / calUDisparity outputs on small seeded
inputs.  /root/reference only exists in the
build container, so the vectors are committed; tests/test_oracle_cues.py checks the C oracle against them everywhere
and against the live reference build where oracle/_ref is present.

Run from the repo root:  python tests/golden/make_golden_stereo.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from semantic_slam_mapping_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert oracle.ref() is not None, "oracle/_ref/libref_stereo.so could not be built (no /root/reference?)"
    out = {}
    H, W, D = 40, 112, 80
    L, R, _ = synth.stereo_pair(H, W, D, 21)
    # the reference's entry point with its hard-coded parameters (src/stereo.cpp:16-28); cv::StereoSGBM itself is
    # un-vendored third-party code, so the functor forwards to the oracle's restatement
    disp = oracle.ref_cal_disparity_sgbm(L, R)
    out["sgbm_params_keys"] = np.array(sorted(oracle.ref_sgbm_params().keys()))
    out["sgbm_params_vals"] = np.array([oracle.ref_sgbm_params()[k] for k in sorted(oracle.ref_sgbm_params().keys())], np.int32)
    # a disparity map with every special value: 0, the minimum (-16), large and small disparities
    rng = np.random.default_rng(5)
    disp = disp.copy()
    disp[rng.random(disp.shape) < 0.05] = 0
    disp[3, 5:40] = np.arange(35) * 40 + 7
    cam = np.array([718.856, 52.3, 19.7, 0.532331858])      # f, cx, cy, b
    roi = np.array([20.0, 1.5, 40.0])
    pitch = np.array([0.031, -0.012])
    xyz = oracle.ref_triangulate10d(L, disp, *cam, roi=tuple(roi))
    cor = oracle.ref_correct_3d_points(xyz, tuple(roi), pitch[0], pitch[1])
    mask = oracle.ref_set_image_roi(cor)
    out.update(left=L, disp=disp, cam=cam, roi=roi, pitch=pitch, xyz=xyz, corrected=cor, roi_mask=mask)
    # U/V-disparity histograms (src/uvdisparity.cpp:277-366, 195-274) in the order UVDisparity::Process runs them: V on the
    # triangulated records, U on the corrected ones.  The map's maximum is 90.5625 px: v_cols = 91 and round() = 91, so that
    # pixel's bin index equals v_cols (spills into the next row; past the matrix on the last row); 20.5 rounds to 20.
    duv = disp.copy()
    duv[7, 9] = 16 * 90 + 9
    duv[H - 1, 4] = 16 * 90 + 9
    duv[2, 11] = 16 * 20 + 8
    xyz_uv = oracle.ref_triangulate10d(L, duv, *cam, roi=tuple(roi))
    v_xyz, v_int, v_u8 = oracle.ref_v_disparity(duv, xyz_uv)
    cor_uv = oracle.ref_correct_3d_points(v_xyz, tuple(roi), pitch[0], pitch[1])
    mask_uv = oracle.ref_set_image_roi(cor_uv)
    ground = (np.random.default_rng(6).random((H, W)) < 0.7).astype(np.uint8) * 255
    u_xyz, u_int, u_u8 = oracle.ref_u_disparity(duv, cor_uv, mask_uv, ground)
    out.update(uv_disp=duv, uv_xyz=xyz_uv, v_xyz=v_xyz, v_int=v_int, v_u8=v_u8, uv_corrected=cor_uv, uv_roi_mask=mask_uv,
               uv_ground=ground, u_xyz=u_xyz, u_int=u_int, u_u8=u_u8)
    np.savez_compressed(os.path.join(OUT, "cues_ref.npz"), **out)
    print("wrote cues_ref.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
