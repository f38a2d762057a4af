#!/usr/bin/env python
"""bench.py -- throughput of the dense stereo-to-semantic-map path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (SGM disparity -> depth -> labelled cloud -> voxel-hash fusion) over one
batch of B synthetic KITTI-shaped frames (1241x376, 128 disparities, 12-class masks, poses) per GPU.
Workload (config.workload): BASELINE.json configs[1] -- a 100-frame synthetic KITTI-shaped sequence, 0.05 m voxels.
Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM (device entry point, CUDA events,
max over ranks); `e2e` = the same through the host entry point ssm_pipeline_batch_host with pinned host buffers
(H2D of every step's inputs and the D2H of the step's result inside the timed region).
`--impl reference` times the reference's CPU implementation of the path instead: cv2.StereoSGBM (the library call
src/stereo.cpp:13-30 makes) + the C oracle's restatement of the depth/cloud/VoxelGrid glue, one process per host core.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, D, LABELS, LEAF = 1241, 376, 128, 12, 0.05
SEQ_FRAMES = 100
METRIC = "semantic-map frames/s at 1241x376, 128 disparities (SGM disparity + labelled cloud + voxel fusion)"


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# communicator), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


# ------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(points_per_frame: float) -> dict:
    """SURVEY.md section 8d per-frame algorithmic bytes of each stage (materialised cost volume formulation)."""
    N = (W - D) * H * D
    px = W * H
    return {
        "cost": 2 * px + 2 * N,                 # K1: read L,R; write C
        "vertical": 2 * N + 2 * N,              # K2: read C, write S_v
        "horizontal": 2 * N + 2 * N + 2 * px,   # K3: read C, S_v; write disp1 records
        "select": 2 * px + 2 * 2 * px,          # K4 + K5: L-R check, median (fused with the band-local part of K6 since r1 v10)
        "post": 4 * 2 * px,                     # K6: speckle filter (K5 + K6 = 6 passes over the int16 image in SURVEY 8d)
        "points": 2 * px + 3 * px + 3 * px + 20 * points_per_frame,   # K7
        "fuse": 20 * points_per_frame + 64 * points_per_frame,       # K8 upper bound: one record RMW per point
        "total": 10 * N + 30 * px + 84 * points_per_frame,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        """Summary of the samples that arrived inside [t_begin, t_end] (host clock); all samples if none did."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r[1:] for r in self.rows if t_begin <= r[0] <= t_end + 0.06]
        rows = inside if inside else [r[1:] for r in self.rows]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm: cv2.StereoSGBM + oracle glue, one worker process per host core
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n = args
    import cv2
    import oracle
    from semantic_slam_mapping_b200 import synth
    cv2.setNumThreads(1)
    sg = cv2.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=11, P1=4 * 121, P2=32 * 121, disp12MaxDiff=1,
                               preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32,
                               mode=cv2.STEREO_SGBM_MODE_SGBM)
    mp = oracle.MapParams()
    seq = synth.sequence(n, H, W, D, LABELS, seed=seed, distinct=min(n, 2))
    t0 = time.perf_counter()
    clouds = []
    for i in range(n):
        disp = sg.compute(seq["left"][i], seq["right"][i])
        depth = oracle.disparity_to_depth(disp, mp)
        clouds.append(oracle.generate_point_cloud(depth, seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i]))
    return time.perf_counter() - t0, [(c["xyz"], c["rgba"], c["label"]) for c in clouds]


def cpu_reference_step(pool, cores: int, frames_per_core: int, seed: int) -> tuple[float, int]:
    """One bounded sample: cores*frames_per_core frames through the CPU path; returns (seconds, frames).
    Frame stages run one process per core (inputs are generated inside each worker before its clock starts, so
    the time is max-over-workers of the compute part); the voxel merge is single-threaded, as pcl::VoxelGrid is."""
    import oracle
    res = pool.map(_cpu_worker, [(seed * 1000 + i, frames_per_core) for i in range(cores)])
    t_frames = max(t for t, _ in res)
    t0 = time.perf_counter()
    vm = oracle.VoxelMap(LEAF, LABELS)
    for _, clouds in res:
        for xyz, rgba, lab in clouds:
            vm.insert(xyz, rgba, lab)
    vm.export()
    return t_frames + (time.perf_counter() - t0), cores * frames_per_core


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    fpc = max(1, args.ref_frames_per_core)
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(args.warmup):
            cpu_reference_step(pool, cores, 1, 900 + i)
        t = 0.0
        frames = 0
        for i in range(args.steps):
            dt, n = cpu_reference_step(pool, cores, fpc, i)
            t += dt
            frames += n
    fps = frames / t
    sample = f"{args.steps} steps x {cores * fpc} frames (1241x376, D=128): cv2 4.13 StereoSGBM MODE_SGBM single-thread per process + C-oracle depth/cloud glue, one process per core; single-threaded VoxelGrid-style merge"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic KITTI-shaped sequence, 1241x376, 128 disparities, 12 classes, 0.05 m voxels",
                   "frames_per_step": cores * fpc},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from semantic_slam_mapping_b200 import Context, Params, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    nb = max(1, min(args.input_batches, SEQ_FRAMES // B))   # distinct input batches cycled through the steps
    p = Params(num_disparities=D, max_width=W, max_height=H, max_batch=B, resolution=LEAF, map_capacity=args.map_capacity)
    ctx = Context(p, device=local_rank)
    if world > 1:
        from semantic_slam_mapping_b200 import distributed as ssm_dist
        ssm_dist.init_comm(ctx, p2p=not args.no_p2p)   # NCCL communicator + (default) peer-memory inboxes over NVLink
        ctx.set_route_overlap(not args.no_route_overlap)   # a batch's point exchange overlaps the next batch's SGBM

    # synthetic sequence: this rank's frames (frame batches are sharded per GPU; poses continue across ranks)
    n_frames = nb * B
    seq = synth.sequence(n_frames, H, W, D, LABELS, seed=11 + rank, distinct=min(n_frames, args.distinct))
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in seq.items() if k != "label"}
    devb = {k: v.to(dev) for k, v in pin.items()}
    tstream = torch.cuda.Stream(device=dev)      # the stream every kernel of the timed region is launched on
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device(i):
        j = (i % nb) * B
        ctx.pipeline_batch_device(devb["left"][j:j + B], devb["right"][j:j + B], devb["semantic"][j:j + B], devb["rgb"][j:j + B],
                                  devb["pose"][j:j + B], B, W, H, stream=stream)

    def step_host(i):
        j = (i % nb) * B
        nvox, _ = ctx.pipeline_batch_host(pin["left"][j:j + B].numpy(), pin["right"][j:j + B].numpy(), pin["semantic"][j:j + B].numpy(),
                                          pin["rgb"][j:j + B].numpy(), pin["pose"][j:j + B].numpy())
        return nvox

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- per-stage device times (a separate short pass: stage events serialise the sub-batch streams) ------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    ctx.set_stage_timing(True)
    for i in range(max(3, min(args.steps, 5))):
        step_device(i)
    barrier()
    stage = ctx.stage_times_ms()
    ctx.set_stage_timing(False)

    # ---- device-resident throughput ("value") --------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                     # nvidia-smi needs a moment to start: launch it before the warm-up
    ctx.map_clear()
    for i in range(args.warmup):
        step_device(i)
    ctx.map_clear()
    barrier()
    t_begin = time.perf_counter()
    launches0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for i in range(args.steps):
        step_device(i)
    if world > 1:
        ctx.synchronize()      # route overlap: the last batch's exchange runs on the library's route stream -- inside the timed region
    e1.record(tstream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.kernel_launches() - launches0
    n_vox = ctx.map_size()
    clk = clocks.stop(t_begin, time.perf_counter()) if rank == 0 else None

    # points per frame (for the algorithmic-byte model): one compact cloud on the first frame's disparity
    disp0 = ctx.sgbm(seq["left"][0], seq["right"][0])
    pts = len(ctx.generate_point_cloud(ctx.disparity_to_depth(disp0), seq["semantic"][0], seq["rgb"][0], seq["pose"][0])["xyz"])

    # ---- end to end through the host entry point ("e2e") ----------------------------------------------------
    # Every step copies its inputs from pinned host memory (H2D inside the timed region) and copies the step's result
    # (the map size after the batch) back to pinned host memory.  The streaming entry point stages inputs through two
    # device buffer sets, so the copy of step k+1 overlaps the kernels of step k.
    result = torch.zeros(args.steps + 8, dtype=torch.int32).pin_memory()

    def step_host(i, slot):
        j = (i % nb) * B
        ctx.pipeline_batch_host_async(pin["left"][j:j + B].numpy(), pin["right"][j:j + B].numpy(), pin["semantic"][j:j + B].numpy(),
                                      pin["rgb"][j:j + B].numpy(), pin["pose"][j:j + B].numpy(), result[slot:slot + 1])

    ctx.map_clear()
    for i in range(min(args.warmup, 3)):
        step_host(i, args.steps + i)
    ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i, i)
    ctx.synchronize()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert int(result[args.steps - 1]) > 0 and ctx.map_size() > 0      # results arrived; no capacity error

    total_frames = args.steps * B * world
    value = total_frames / (dev_ms * 1e-3)
    e2e = total_frames / e2e_s
    if rank != 0:
        return

    # ---- roofline of the dominant stage ------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ab = algorithmic_bytes(pts)
    kernels = {"cost": "k_prefilter8 + k_cost_fused", "vertical": "k_vertical3 (cluster kernel)", "horizontal": "k_hfwd + k_hrev",
               "select": "k_select_fused (records -> L-R check -> median -> band-local speckle components)", "post": "k_cc_merge_bands + k_cc_count_roots + k_cc_apply_bands", "points": "k_depth + k_labels + k_moving_mask",
               "fuse": "k_points_fuse" if world == 1 else "k_points_p2p + barrier + k_fuse_list"}
    dom = max(stage, key=stage.get)
    achieved = ab[dom] * B / (stage[dom] * 1e-3) / 1e9
    stage_roof = {k: {"ms_per_step": round(v, 4), "alg_GBps": round(ab[k] * B / (v * 1e-3) / 1e9, 1) if v > 0 else None}
                  for k, v in stage.items()}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get(dom) is not None:      # measured for launches of _frames_per_launch frames; a launch of the stage pass has B
            traffic = round(tj[dom] * B / tj.get("_frames_per_launch", 33))

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_step(pool, cores, 1, 777)
            dt, n = cpu_reference_step(pool, cores, args.ref_frames_per_core, 778)
        cpu = {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n} frames (1241x376, D=128): cv2 4.13 StereoSGBM (single-thread per process) + C-oracle glue, "
                         f"one process per core, single-threaded voxel merge; {dt:.1f} s wall"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": "configs[1]: 100-frame synthetic KITTI-shaped stereo sequence with poses, 1241x376, 128 disparities, "
                               "12-class masks, 0.05 m voxel map", "frames_per_step_per_gpu": B, "distinct_input_batches": nb,
                   "l2": "each step streams > 2 GB of cost volumes through HBM (>> 126 MB L2); input batches cycle",
                   "voxels_in_map_rank0": n_vox, "points_per_frame": pts,
                   "parallelism": (f"frames sharded over {world} GPU(s); voxel hash spatially owned; points routed to the owner by "
                                   + ("NCCL send/recv all-to-all" if args.no_p2p else "peer-memory stores over NVLink fused into the point kernel + NCCL barrier"))
                   if world > 1 else "1 GPU"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * (2 * W * H + 6 * W * H + 128), "d2h_bytes_per_step": 4,
                "api": "ssm_pipeline_batch_host_async (pinned host buffers, double-buffered staging) + ssm_synchronize"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": f"stage '{dom}': {kernels[dom]}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "frames_per_launch": B, "peak_source": peak_src,
                     "note": "achieved = SURVEY 8d algorithmic bytes of the stage x frames per launch / CUDA-event time of the stage; "
                             "traffic = ncu dram bytes of the stage's kernels (profiles/dominant_kernel_traffic.json), scaled to the same frames per launch",
                     "whole_path": {"alg_bytes_per_frame": ab["total"], "achieved": ab["total"] * total_frames / world / (dev_ms * 1e-3) / 1e9,
                                    "frac": ab["total"] * total_frames / world / (dev_ms * 1e-3) / 1e9 / peak},
                     "stages": stage_roof},
        "cpu_baseline": cpu,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=66, help="frames per step per GPU (66 = two sub-batches of 33, each one full wave of 4-CTA clusters on 132 SMs)")
    ap.add_argument("--input-batches", type=int, default=3)
    ap.add_argument("--no-route-overlap", action="store_true", help="N > 1: keep the point exchange on the pipeline stream")
    ap.add_argument("--distinct", type=int, default=10, help="distinct synthetic images generated on the host")
    ap.add_argument("--map-capacity", type=int, default=1 << 24)
    ap.add_argument("--ref-frames-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL send/recv all-to-all instead of peer-memory routing")
    args = ap.parse_args()
    capture_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
