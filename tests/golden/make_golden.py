#!/usr/bin/env python
"""Regenerates tests/golden/*.npz with cv2 (opencv-python-headless 4.13) -- the only executable
copy of the third-party code the reference calls on this path (cv::StereoSGBM at
/root/reference src/stereo.cpp:13-30, cv::dilate at src/mapper.cpp:214; medianBlur/filterSpeckles
are internal stages of StereoSGBM).  The reference ships no golden vectors of its own (SURVEY.md
section 4), so these pin the oracle, and through it the CUDA path.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from semantic_slam_mapping_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)


def cv_sgbm(L, R, D, bs=11, uniq=10, spw=100, spr=32, d12=1, cap=63):
    s = cv2.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=bs, P1=4 * bs * bs, P2=32 * bs * bs,
                              disp12MaxDiff=d12, preFilterCap=cap, uniquenessRatio=uniq, speckleWindowSize=spw,
                              speckleRange=spr, mode=cv2.STEREO_SGBM_MODE_SGBM)
    return s.compute(L, R)


def main():
    cases = {}
    # (name, H, W, D, seed, bs, uniq, speckle window)
    specs = [
        ("small_d32", 48, 96, 32, 1, 11, 10, 100),
        ("mid_d80_refdefault", 100, 300, 80, 3, 11, 10, 100),   # D=80 is the reference's own setting
        ("mid_d128", 120, 400, 128, 4, 11, 10, 100),
        ("bs5_u0_d48", 64, 160, 48, 2, 5, 0, 0),
        ("bs7_u15_d64", 72, 200, 64, 6, 7, 15, 50),
        ("odd_shape_d16", 37, 53, 16, 8, 3, 10, 100),
    ]
    for name, H, W, D, seed, bs, uniq, spw in specs:
        L, R, _ = synth.stereo_pair(H, W, D, seed)
        cases[name] = dict(left=L, right=R, disp=cv_sgbm(L, R, D, bs, uniq, spw),
                           params=np.array([D, bs, uniq, spw, 32, 1, 63], np.int32))
    rng = np.random.default_rng(5)
    L = rng.integers(0, 256, (60, 200), dtype=np.uint8)
    R = rng.integers(0, 256, (60, 200), dtype=np.uint8)
    cases["noise_saturating_d64"] = dict(left=L, right=R, disp=cv_sgbm(L, R, 64),
                                         params=np.array([64, 11, 10, 100, 32, 1, 63], np.int32))
    flat = np.full((40, 120), 77, np.uint8)
    cases["constant_image_d32"] = dict(left=flat, right=flat, disp=cv_sgbm(flat, flat, 32),
                                       params=np.array([32, 11, 10, 100, 32, 1, 63], np.int32))
    np.savez_compressed(os.path.join(OUT, "sgbm_small.npz"),
                        **{f"{k}/{f}": v for k, c in cases.items() for f, v in c.items()})

    # KITTI-shaped full frame at BASELINE.json's 128 disparities: inputs are regenerated from the seed
    L, R, _ = synth.stereo_pair(376, 1241, 128, 0)
    np.savez_compressed(os.path.join(OUT, "sgbm_kitti_d128.npz"), disp=cv_sgbm(L, R, 128),
                        left_sum=np.int64(L.astype(np.int64).sum()), right_sum=np.int64(R.astype(np.int64).sum()))

    # Cityscapes-shaped full frame at BASELINE.json configs[3]: 2048 x 1024, 256 disparities (inputs regenerated from the seed)
    L, R, _ = synth.stereo_pair(1024, 2048, 256, 0)
    np.savez_compressed(os.path.join(OUT, "sgbm_cityscapes_d256.npz"), disp=cv_sgbm(L, R, 256),
                        left_sum=np.int64(L.astype(np.int64).sum()), right_sum=np.int64(R.astype(np.int64).sum()))

    # stage-level vectors
    img = rng.integers(-16, 2048, (50, 70)).astype(np.int16)
    img[rng.random(img.shape) < 0.3] = -16
    sp = img.copy()
    cv2.filterSpeckles(sp, -16, 20, 64)
    mask = (rng.random((45, 61)) < 0.03).astype(np.uint8) * 255
    np.savez_compressed(os.path.join(OUT, "stages.npz"), median_in=img, median_out=cv2.medianBlur(img, 3),
                        speckle_in=img, speckle_out=sp, dilate_in=mask,
                        dilate_out=cv2.dilate(mask, np.ones((3, 3), np.uint8), iterations=2))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
