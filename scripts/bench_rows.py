#!/usr/bin/env python
"""Device-time micro-benchmark of the SURVEY 8f rows built so far (not the headline metric; bench.py is):
dense motion cues (stage 1 + stage 2), label production, keyframe redraw.  CUDA events on the launching stream, inputs
resident, KITTI shape.  Prints one JSON object.  usage: python scripts/bench_rows.py [--batch 33] [--reps 20]"""
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
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    H, W, D, B = 376, 1241, 128, a.batch
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=0.05, map_capacity=1 << 23)
    seq = synth.sequence(B, H, W, D, 12, seed=3, distinct=4)
    dev = torch.device("cuda:0")
    out = {"shape": [H, W], "batch": B, "reps": a.reps}
    with Context(p) as ctx:
        st = torch.cuda.Stream(device=dev)
        s = st.cuda_stream
        d_left = torch.from_numpy(seq["left"]).to(dev)
        d_right = torch.from_numpy(seq["right"]).to(dev)
        d_disp = torch.empty((B, H, W), dtype=torch.int16, device=dev)
        ctx.sgbm_batch_device(d_left, d_right, d_disp, B, W, H, stream=s)
        torch.cuda.synchronize()

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.reps):
                fn()
            e1.record(st)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.reps

        # dense motion cues
        cap = D + 2
        v_stride, u_stride = H * cap + 4, cap * W
        d_xyz = torch.empty((B, H, W, 10), dtype=torch.float32, device=dev)
        d_vint = torch.empty((B, v_stride), dtype=torch.int32, device=dev)
        d_v8 = torch.empty((B, v_stride), dtype=torch.uint8, device=dev)
        d_uint = torch.empty((B, u_stride), dtype=torch.int32, device=dev)
        d_u8 = torch.empty((B, u_stride), dtype=torch.uint8, device=dev)
        d_roi = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        f, cx, cy, b = p.fx, p.cx, p.cy, p.baseline
        ms1 = timed(lambda: ctx.motion_cues_stage1_device(d_left, d_disp, d_xyz, B, W, H, f, cx, cy, b, d_vint, d_v8, v_stride, cap, stream=s))
        ms2 = timed(lambda: ctx.motion_cues_stage2_device(d_disp, d_xyz, d_roi, B, W, H, (20.0, 1.5, 40.0), 0.02, None, d_uint, d_u8, u_stride, cap, stream=s))
        npix = B * H * W
        out["motion_cues"] = {
            "stage1_ms": ms1, "stage2_ms": ms2, "frames_per_s": B / ((ms1 + ms2) * 1e-3),
            # algorithmic bytes: stage 1 reads disp (2) + grey (1), writes the 40-byte record; stage 2 reads + writes the record
            # (channel 7 pass included: 2 x 40 more), reads disp twice, writes the roi mask
            "stage1_alg_GBps": npix * (2 + 2 + 1 + 40) / (ms1 * 1e6), "stage2_alg_GBps": npix * (40 * 4 + 2 + 1) / (ms2 * 1e6),
        }
        # label production 480x360 -> frame size
        idx = torch.from_numpy(np.random.default_rng(0).integers(0, 12, (B, 360, 480)).astype(np.uint8)).to(dev)
        d_sem = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
        lut = np.zeros((256, 3), np.uint8)
        lut[:12] = np.array(p.palette_bgr, np.uint8)
        msl = timed(lambda: ctx.labels_from_indices_batch_device(idx, d_sem, B, 480, 360, W, H, lut, None, stream=s))
        out["labels"] = {"ms": msl, "frames_per_s": B / (msl * 1e-3), "alg_GBps": (B * 360 * 480 + npix * 3) / (msl * 1e6)}
        # keyframe redraw: 16 cached keyframes re-fused under new poses
        mp = None
        kids = []
        for i in range(16):
            depth = ctx.disparity_to_depth(d_disp[i % B].cpu().numpy())
            kids.append(ctx.keyframe_add(depth, seq["semantic"][i % B], seq["rgb"][i % B], seq["pose"][i % B]))
        import time
        ctx.map_redraw(kids)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.map_redraw(kids)
        dt = (time.perf_counter() - t0) / 5
        n_live, n_pts = ctx.keyframe_count()
        out["redraw"] = {"keyframes": n_live, "points": n_pts, "ms": dt * 1e3, "points_per_s": n_pts / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
