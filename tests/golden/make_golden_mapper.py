#!/usr/bin/env python
"""Regenerates tests/golden/mapper_ref.npz: outputs of the REFERENCE's own compiled FrameReader::next (disparity -> depth,
src/rgbdframe.cpp:85-116), Mapper::semantic_motion_fuse (src/mapper.cpp:189-216) and Mapper::generatePointCloud with
RGBDFrame::project2dTo3d (src/mapper.cpp:12-94, include/rgbdframe.h:63-75) on small synthetic frames, produced through
oracle/_ref/libref_mapper.so (the reference's .cpp files compiled from /root/reference against oracle/cvstub + oracle/refstub;
needs /root/reference, i.e. the build container).  Run from the repo root:  python tests/golden/make_golden_mapper.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from semantic_slam_mapping_b200 import synth  # noqa: E402
from semantic_slam_mapping_b200.params import SEGNET12_BGR  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases():
    """(name, left, right, semantic, rgb, pose): the reference's calDisparity_SGBM hard-codes 80 disparities (src/stereo.cpp:18)."""
    H, W, D = 72, 256, 80
    poses = synth.poses(6, 4)
    for k, seed in enumerate((11, 12)):
        L, R, _ = synth.stereo_pair(H, W, D, seed)
        _, sem = synth.label_mask(H, W, 12, seed, cell=12)
        rgb = np.stack([L, np.roll(L, 3, axis=1), 255 - L], axis=-1)      # three different channels: the colour tag is b, g, r of THIS image
        yield f"frame{k}", L, R, sem, rgb, poses[2 + 3 * k]


def adversarial():
    """A depth image that exercises every filter of generatePointCloud directly: zeros, values around max_distance * scale, every
    palette colour (dropped / dynamic / kept classes) and colours outside the palette, dynamic pixels at the image border."""
    rng = np.random.default_rng(7)
    H, W = 40, 96
    depth = rng.integers(0, 45000, (H, W)).astype(np.uint16)
    depth[rng.random((H, W)) < 0.2] = 0
    depth[0, :8] = [39999, 40000, 40001, 1, 65535, 0, 40000, 2]
    pal = np.asarray(SEGNET12_BGR, np.uint8)
    sem = pal[rng.integers(0, 12, (H // 4, W // 4))].repeat(4, axis=0).repeat(4, axis=1)
    sem[rng.random((H, W)) < 0.03] = (1, 2, 3)
    sem[0, 0] = pal[10]; sem[H - 1, W - 1] = pal[11]; sem[H // 2, 0] = pal[10]
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    T = np.array([[0.36, 0.48, -0.8, 1.5], [-0.8, 0.6, 0.0, -2.25], [0.48, 0.64, 0.6, 100.125], [0, 0, 0, 1]], np.float64)
    return depth, sem, rgb, T


def main():
    if oracle.ref_mapper() is None:
        raise SystemExit("oracle/_ref/libref_mapper.so cannot be built here (no /root/reference)")
    mp = oracle.MapParams()
    out = {}
    for name, L, R, sem, rgb, T in cases():
        depth, disp = oracle.ref_frame_next(L, R, rgb, sem, mp)
        c = oracle.ref_mapper_cloud(depth, sem, rgb, mp, T)
        out.update({f"{name}/left": L, f"{name}/right": R, f"{name}/semantic": sem, f"{name}/rgb": rgb, f"{name}/pose": T,
                    f"{name}/disp": disp, f"{name}/depth": depth, f"{name}/mask": c["mask"], f"{name}/xyz_cam": c["xyz_cam"],
                    f"{name}/xyz": c["xyz"], f"{name}/rgba": c["rgba"]})
    depth, sem, rgb, T = adversarial()
    c = oracle.ref_mapper_cloud(depth, sem, rgb, mp, T)
    out.update({"adv/depth": depth, "adv/semantic": sem, "adv/rgb": rgb, "adv/pose": T, "adv/mask": c["mask"], "adv/xyz_cam": c["xyz_cam"],
                "adv/xyz": c["xyz"], "adv/rgba": c["rgba"]})
    np.savez_compressed(os.path.join(OUT, "mapper_ref.npz"), **out)
    print("wrote mapper_ref.npz:", {k: v.shape for k, v in out.items() if k.endswith("xyz")})


if __name__ == "__main__":
    main()
