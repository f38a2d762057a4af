#!/usr/bin/env python
"""BASELINE.json configs[3] (Cityscapes-shaped 2048x1024 stereo, 256 disparities, 19 classes) on one GPU: device-resident
frames/s and per-stage times.  A parity-test configuration, not the headline bench line (bench.py measures configs[1]);
kept to track where the D = 256 code paths stand.  --kitti D runs the KITTI shape with D disparities instead (D = 80 is
the value hard-coded in the reference's src/stereo.cpp:18).
usage: python scripts/bench_cityscapes.py [--batch 8] [--steps 5] [--kitti 80]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from semantic_slam_mapping_b200 import Context, synth
    from semantic_slam_mapping_b200.params import cityscapes_params
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kitti", type=int, default=0, help="KITTI shape 1241x376 with this many disparities instead of configs[3]")
    a = ap.parse_args()
    H, W, D, B = 1024, 2048, 256, a.batch
    if a.kitti:
        from semantic_slam_mapping_b200 import Params
        H, W, D = 376, 1241, a.kitti
        p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 24)
        seq = synth.sequence(B, H, W, D, 12, seed=5, distinct=2)
    else:
        p = cityscapes_params(max_batch=B, resolution=0.05, map_capacity=1 << 24)
        seq = synth.sequence(B, H, W, D, 19, seed=5, distinct=2)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in seq.items() if k != "label"}
    st = torch.cuda.Stream(device=dev)
    with Context(p) as ctx:
        def step():
            ctx.pipeline_batch_device(d["left"], d["right"], d["semantic"], d["rgb"], d["pose"], B, W, H, stream=st.cuda_stream)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ctx.set_stage_timing(True)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        stages = ctx.stage_times_ms()
        ctx.set_stage_timing(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.steps):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        N = (W - D) * H * D
        print(json.dumps({"workload": f"{W}x{H}, {D} disparities", "batch": B, "ms_per_step": ms,
                          "frames_per_s": B / (ms * 1e-3), "stage_ms": {k: round(v, 3) for k, v in stages.items()},
                          "alg_GBps": {"cost": round(2 * N * B / (stages["cost"] * 1e6), 1), "vertical": round(4 * N * B / (stages["vertical"] * 1e6), 1),
                                       "horizontal": round(4 * N * B / (stages["horizontal"] * 1e6), 1)},
                          "voxels": ctx.map_size()}))


if __name__ == "__main__":
    main()
