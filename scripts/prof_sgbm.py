#!/usr/bin/env python
"""Micro-benchmark for kernel work: B frames through ssm_pipeline_batch_device with the per-stage CUDA-event timers (one stream, no sub-batches below 64 frames).
usage: python scripts/prof_sgbm.py [--batch 33] [--reps 10] [--disparities 128] [--width 1241] [--height 376]
Run it under ncu for a capture of single kernels (scripts/gpu_prof.sh); env knobs (SSM_NO_COST_TMA, SSM_TUNEx ...) apply."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from semantic_slam_mapping_b200 import Context, Params, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=33)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--disparities", type=int, default=128)
    ap.add_argument("--width", type=int, default=1241)
    ap.add_argument("--height", type=int, default=376)
    ap.add_argument("--distinct", type=int, default=6)
    a = ap.parse_args()
    H, W, D, B = a.height, a.width, a.disparities, a.batch
    seq = synth.sequence(B, H, W, D, 12, seed=100, distinct=a.distinct)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(np.ascontiguousarray(seq[k])).to(dev) for k in ("left", "right", "semantic", "rgb")}
    dP = torch.from_numpy(np.ascontiguousarray(seq["pose"]).astype(np.float64)).to(dev)
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 24)
    import time
    with Context(p) as ctx:
        def step():
            ctx.pipeline_batch_device(d["left"], d["right"], d["semantic"], d["rgb"], dP, B, W, H)
        for _ in range(2):
            step()
        ctx.synchronize()
        ctx.set_stage_timing(True)
        t = time.perf_counter()
        for _ in range(a.reps):
            step()
        ctx.synchronize()
        wall = (time.perf_counter() - t) / a.reps * 1e3
        st = ctx.stage_times_ms()
        ctx.set_stage_timing(False)
    print(json.dumps({"shape": [H, W], "D": D, "batch": B, "ms_per_batch_wall": round(wall, 4), "frames_per_s": round(B / wall * 1e3, 1),
                      "stage_ms": {k: round(v, 4) for k, v in st.items()},
                      "env": {k: v for k, v in os.environ.items() if k.startswith("SSM_")}}))


if __name__ == "__main__":
    main()
