#!/usr/bin/env python
"""PNG ingest: frames/s of decoding the four images of a KITTI-shaped frame (left and right as grey from
colour files, the colour image and the label image as BGR) into device buffers -- (a) inflate on the GPU (one warp per file;
the host only walks chunks and checks CRCs), (b) zlib inflate on all host cores -- both followed by GPU un-filtering /
conversion -- next to cv2.imdecode on a thread pool of the same size (the reference's cv::imread path).
usage: python scripts/bench_ingest.py [--frames 264] [--reps 3]"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    import torch
    from semantic_slam_mapping_b200 import Context, Params, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=264)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    H, W, B = 376, 1241, a.frames
    seq = synth.sequence(4, H, W, 128, 12, seed=3, distinct=4)
    rng = np.random.default_rng(0)
    NP = min(B, 16)                                    # distinct files (encoding is slow); the batch cycles through them
    photo = [np.clip(seq["rgb"][i % 4].astype(int) + rng.integers(-6, 7, (H, W, 3)), 0, 255).astype(np.uint8) for i in range(NP)]
    enc = lambda x: cv2.imencode(".png", x)[1].tobytes()
    left = [enc(p) for p in photo]
    right = [enc(np.roll(p, 7, axis=1)) for p in photo]
    sem = [enc(seq["semantic"][i % 4]) for i in range(NP)]
    left, right, sem = [[x[i % NP] for i in range(B)] for x in (left, right, sem)]
    rgb = left
    cores = os.cpu_count()
    dev = torch.device("cuda:0")
    d_grey = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
    d_bgr = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    with Context(Params(num_disparities=128, max_width=W, max_height=H, max_batch=1, map_capacity=1 << 16)) as ctx:
        streams = [torch.cuda.Stream() for _ in range(4)]   # the four image kinds of a frame batch decode side by side
        def ours(threads):
            ctx.png_decode_batch_device(left, W, H, False, d_grey[0], host_threads=threads, stream=streams[0].cuda_stream)
            ctx.png_decode_batch_device(right, W, H, False, d_grey[1], host_threads=threads, stream=streams[1].cuda_stream)
            ctx.png_decode_batch_device(rgb, W, H, True, d_bgr[0], host_threads=threads, stream=streams[2].cuda_stream)
            ctx.png_decode_batch_device(sem, W, H, True, d_bgr[1], host_threads=threads, stream=streams[3].cuda_stream)
        res = {}
        for name, threads in (("gpu_inflate", 0), ("host_inflate", cores)):
            ours(threads)
            ctx.png_batch_wait()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(a.reps):
                ours(threads)
            ctx.png_batch_wait()
            torch.cuda.synchronize()
            res[name] = (time.perf_counter() - t) / a.reps
        # the GPU decoder alone: one colour batch, wall time from call to completion
        ctx.png_batch_wait()
        torch.cuda.synchronize()
        t = time.perf_counter()
        ctx.png_decode_batch_device(rgb, W, H, True, d_bgr[0], host_threads=0, stream=streams[0].cuda_stream)
        ctx.png_batch_wait()
        colour_batch_ms = (time.perf_counter() - t) * 1e3
    with ThreadPoolExecutor(cores) as ex:
        def ref():
            jobs = [(p, cv2.IMREAD_GRAYSCALE) for p in left] + [(p, cv2.IMREAD_GRAYSCALE) for p in right] + \
                   [(p, cv2.IMREAD_COLOR) for p in rgb] + [(p, cv2.IMREAD_COLOR) for p in sem]
            return list(ex.map(lambda j: cv2.imdecode(np.frombuffer(j[0], np.uint8), j[1]), jobs))
        ref()
        t = time.perf_counter()
        for _ in range(a.reps):
            ref()
        ref_s = (time.perf_counter() - t) / a.reps
    print(json.dumps({"frames": B, "images_per_frame": 4, "host_threads": cores, "png_bytes_per_frame": int(sum(map(len, left + right + rgb + sem)) / B),
                      "ours_gpu_inflate_frames_per_s": round(B / res["gpu_inflate"], 1), "ours_host_inflate_frames_per_s": round(B / res["host_inflate"], 1),
                      "cv2_thread_pool_frames_per_s": round(B / ref_s, 1),
                      "gpu_colour_batch_ms": round(colour_batch_ms, 2), "gpu_colour_images_per_s": round(B / colour_batch_ms * 1e3, 1),
                      "note": "ours: DEFLATE on the GPU (one warp per file) or zlib on host threads, then un-filter + convert on the GPU, output resident in HBM; "
                              "cv2: full decode on host threads, output in host memory"}))


if __name__ == "__main__":
    main()
