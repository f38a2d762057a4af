#!/usr/bin/env python
"""Latency of the drop-in single-frame calls (host buffers in, host buffers out), the way the reference uses them:
calDisparity_SGBM per frame on the main thread (src/rgbdframe.cpp:82), then depth + cloud + voxel fusion per keyframe.
usage: python scripts/bench_single_frame.py [--disparities 80] [--reps 30]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from semantic_slam_mapping_b200 import Context, Params, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--disparities", type=int, default=80)
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    H, W, D = 376, 1241, a.disparities
    seq = synth.sequence(4, H, W, D, 12, seed=3)
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=1, resolution=0.05, map_capacity=1 << 22)
    with Context(p) as ctx:
        for i in range(3):
            d = ctx.sgbm(seq["left"][i % 4], seq["right"][i % 4])
        t = time.perf_counter()
        for i in range(a.reps):
            d = ctx.sgbm(seq["left"][i % 4], seq["right"][i % 4])
        sgbm_ms = (time.perf_counter() - t) / a.reps * 1e3
        depth = ctx.disparity_to_depth(d)
        ctx.map_integrate_frame(depth, seq["semantic"][0], seq["rgb"][0], seq["pose"][0])
        t = time.perf_counter()
        for i in range(a.reps):
            depth = ctx.disparity_to_depth(d)
            ctx.map_integrate_frame(depth, seq["semantic"][i % 4], seq["rgb"][i % 4], seq["pose"][i % 4])
        map_ms = (time.perf_counter() - t) / a.reps * 1e3
    print(json.dumps({"shape": [H, W], "disparities": D, "calDisparity_SGBM_ms": round(sgbm_ms, 3), "frames_per_s": round(1e3 / sgbm_ms, 1),
                      "depth_plus_map_integrate_ms": round(map_ms, 3)}))


if __name__ == "__main__":
    main()
