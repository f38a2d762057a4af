#!/usr/bin/env python
"""bench.py -- throughput of the dense stereo-to-semantic-map path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4] [--batch B] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (SGM disparity -> depth -> labelled cloud -> voxel-hash fusion) over one batch of
B synthetic frames per GPU.  --config selects the BASELINE.json workload (default 1 = configs[1], the one the metric is
quoted on):
  1  configs[1]: 100-frame KITTI-shaped sequence (1241x376, 128 disparities, 12 classes), 0.05 m voxels.  The 100 stereo
     pairs / label masks are streamed cyclically while the trajectory keeps advancing, so every step inserts new voxels.
  2  configs[2]: the 4541-frame KITTI-00-length sequence sharded over the GPUs (strong scaling: total work fixed; --steps
     defaults to one pass over the rank's shard).
  3  configs[3]: Cityscapes-shaped 2048x1024, 256 disparities, 19 classes.
  4  configs[4]: large-map stress, 0.02 m voxels, 20 000 frames over 8 GPUs (2500 per GPU; --steps defaults to the
     rank's share), all points routed to their owning rank.
Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM (device entry point, CUDA events, max
over ranks); `e2e` = the same through the host entry point ssm_pipeline_batch_host_async with pinned host buffers (H2D
of every step's inputs and the D2H of the step's result inside the timed region); `e2e_with_export` additionally
finalizes, orders and copies the whole fused map (xyz, rgba, label per voxel -- what Mapper::viewer hands to
viewer.showCloud on every update, src/mapper.cpp:154-159) to pinned host memory after every step.
`parity` (outside the timed regions): every rank's first batch is recomputed on the CPU (cv2.StereoSGBM -- the library
call the reference makes -- or the C oracle's SGBM, then the oracle's depth / cloud / VoxelGrid restatement); the
disparities of all its frames and the N-rank map gathered through ssm_map_export_gathered must match: disparity, voxel
set, counts, votes and majority labels exactly, centroids within 1e-5 relative.  A mismatch makes the run exit non-zero.
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

METRIC = "semantic-map frames/s at 1241x376, 128 disparities (SGM disparity + labelled cloud + voxel fusion)"

CONFIGS = {
    1: dict(W=1241, H=376, D=128, labels=12, leaf=0.05, batch=66, seq_frames=100, scaling="weak", capacity=1 << 28, steps=30,
            workload="configs[1]: 100-frame synthetic KITTI-shaped stereo sequence with poses, 1241x376, 128 disparities, 12-class masks, "
                     "0.05 m voxel map; the 100 frames are streamed cyclically while the trajectory keeps advancing (every step inserts new voxels)"),
    2: dict(W=1241, H=376, D=128, labels=12, leaf=0.05, batch=66, seq_frames=4541, scaling="strong", capacity=1 << 29, steps=None,
            workload="configs[2]: 4541-frame KITTI-00-length synthetic sequence, batched frames sharded across the GPUs, spatially owned voxel "
                     "hash, 1241x376, 128 disparities, 12 classes, 0.05 m voxels (100 distinct stereo pairs, 4541 distinct poses, map never cleared)"),
    3: dict(W=2048, H=1024, D=256, labels=19, leaf=0.05, batch=14, seq_frames=100, scaling="weak", capacity=1 << 26, steps=20, distinct=24,
            workload="configs[3]: Cityscapes-shaped 2048x1024 stereo, 256 disparities, 19-class labels, 0.05 m voxels"),
    4: dict(W=1241, H=376, D=128, labels=12, leaf=0.02, batch=66, seq_frames=2500, scaling="weak", capacity=1 << 29, steps=None,
            workload="configs[4]: large-map stress, 0.02 m voxels, 2500 frames per GPU (20 000 over 8 GPUs), label-histogram fusion with all "
                     "points routed to their owning rank; 1241x376, 128 disparities, 12 classes"),
}


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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(cfg: dict, points_per_frame: float) -> dict:
    """SURVEY.md section 8d per-frame algorithmic bytes of each stage (materialised cost volume formulation)."""
    W, H, D = cfg["W"], cfg["H"], cfg["D"]
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
# the synthetic sequence: `distinct` stereo pairs / masks (seeded by frame index) cycled under an advancing trajectory
# ------------------------------------------------------------------------------------------------------------
def make_params(cfg: dict, B: int, capacity: int):
    from semantic_slam_mapping_b200 import Params
    from semantic_slam_mapping_b200.params import cityscapes_params
    if cfg["labels"] == 19:
        return cityscapes_params(max_batch=B, resolution=cfg["leaf"], map_capacity=capacity)
    return Params(num_disparities=cfg["D"], max_width=cfg["W"], max_height=cfg["H"], max_batch=B, resolution=cfg["leaf"],
                  map_capacity=capacity)


def oracle_map_params(p):
    import oracle
    return oracle.MapParams(cx=p.cx, cy=p.cy, fx=p.fx, fy=p.fy, baseline=p.baseline, scale=p.scale, roix=p.roix, roiy=p.roiy,
                            roiz=p.roiz, max_distance=p.max_distance, palette_bgr=list(p.palette_bgr), drop_mask=p.drop_mask,
                            dynamic_mask=p.dynamic_mask, dilate_iterations=p.dilate_iterations, colour_source=p.colour_source)


def frame_images(cfg: dict, image_seed: int):
    """The stereo pair, semantic image and rgb image of pool entry `image_seed` (numpy only: same bytes on every host)."""
    from semantic_slam_mapping_b200 import synth
    L, R, _ = synth.stereo_pair(cfg["H"], cfg["W"], cfg["D"], image_seed)
    _, sem = synth.label_mask(cfg["H"], cfg["W"], cfg["labels"], image_seed)
    return L, R, sem, np.repeat(L[..., None], 3, axis=-1)


def image_seed(rank: int, k: int) -> int:
    return 1100003 * (11 + rank) + k


# ------------------------------------------------------------------------------------------------------------
# CPU path: cv2.StereoSGBM (or the C oracle's SGBM) + oracle glue.  Used by the reference arm, the cpu_baseline leg and the
# parity gate (as the checker, outside every timed region).
# ------------------------------------------------------------------------------------------------------------
def _cpu_sgbm(cfg: dict):
    try:
        import cv2
        cv2.setNumThreads(1)
        sg = cv2.StereoSGBM_create(minDisparity=0, numDisparities=cfg["D"], blockSize=11, P1=4 * 121, P2=32 * 121, disp12MaxDiff=1,
                                   preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32,
                                   mode=cv2.STEREO_SGBM_MODE_SGBM)
        return lambda L, R: sg.compute(L, R)
    except ImportError:
        import oracle
        sp = oracle.SgbmParams(num_disparities=cfg["D"])
        return lambda L, R: oracle.sgbm(L, R, sp)


def _cpu_frames_worker(args):
    """Frames (image seeds + poses) through the CPU path.  Returns (seconds of compute, [per-frame (disp | None, cloud)])."""
    cfg, pdict, seeds, poses, want_disp = args
    import oracle
    sgbm = _cpu_sgbm(cfg)
    p = make_params(cfg, 1, 1024)
    for k, v in pdict.items():
        setattr(p, k, v)
    mp = oracle_map_params(p)
    frames = [frame_images(cfg, s) for s in seeds]
    t0 = time.perf_counter()
    out = []
    for (L, R, sem, rgb), T in zip(frames, poses):
        disp = sgbm(L, R)
        depth = oracle.disparity_to_depth(disp, mp)
        c = oracle.generate_point_cloud(depth, sem, rgb, mp, T)
        out.append((disp if want_disp else None, (c["xyz"], c["rgba"], c["label"])))
    return time.perf_counter() - t0, out


def cpu_reference_step(pool, cfg: dict, cores: int, frames_per_core: int, seed: int) -> tuple[float, int]:
    """One bounded sample: cores*frames_per_core frames through the CPU path; returns (seconds, frames).
    Frame stages run one process per core (inputs are generated inside each worker before its clock starts, so
    the time is max-over-workers of the compute part); the voxel merge is single-threaded, as pcl::VoxelGrid is."""
    import oracle
    from semantic_slam_mapping_b200 import synth
    n = cores * frames_per_core
    poses = synth.poses(n, seed)
    jobs = [(cfg, {}, [image_seed(seed, (i * frames_per_core + k) % 2) for k in range(frames_per_core)],
             poses[i * frames_per_core:(i + 1) * frames_per_core], False) for i in range(cores)]
    res = pool.map(_cpu_frames_worker, jobs)
    t_frames = max(t for t, _ in res)
    t0 = time.perf_counter()
    vm = oracle.VoxelMap(cfg["leaf"], cfg["labels"])
    for _, frames in res:
        for _, (xyz, rgba, lab) in frames:
            vm.insert(xyz, rgba, lab)
    vm.export()
    return t_frames + (time.perf_counter() - t0), n


def run_reference(args, cfg, rank: int, world: int, pool, cores: int):
    if rank != 0:
        return
    fpc = max(1, args.ref_frames_per_core)
    for i in range(args.warmup):
        cpu_reference_step(pool, cfg, cores, 1, 900 + i)
    t = 0.0
    frames = 0
    for i in range(args.steps):
        dt, n = cpu_reference_step(pool, cfg, cores, fpc, i)
        t += dt
        frames += n
    fps = frames / t
    sample = (f"{args.steps} steps x {cores * fpc} frames ({cfg['W']}x{cfg['H']}, D={cfg['D']}): cv2 4.13 StereoSGBM MODE_SGBM single-thread per "
              f"process + C-oracle depth/cloud glue, one process per core; single-threaded VoxelGrid-style merge")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "frames_per_step": cores * fpc},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# parity gate
# ------------------------------------------------------------------------------------------------------------
def _pack_ijk(ijk: np.ndarray) -> np.ndarray:
    b = ijk.astype(np.int64) + (1 << 20)
    return b[:, 0] | (b[:, 1] << 21) | (b[:, 2] << 42)


def parity_gate(ctx, cfg, p, rank, world, pool, procs, host_batch, seeds, poses, dist):
    """This rank's first batch on the GPU (through the host entry point) and on the CPU; disparities compared per rank, the
    gathered N-rank map compared on rank 0.  Returns the parity dict on rank 0 (None elsewhere); all ranks learn `ok`."""
    import oracle
    from semantic_slam_mapping_b200 import distributed as ssm_dist
    B = len(seeds)
    t0 = time.perf_counter()
    # CPU side first (the pool is idle during the GPU part anyway)
    per = -(-B // procs)
    jobs = [(cfg, {}, seeds[i:i + per], poses[i:i + per], True) for i in range(0, B, per)]
    async_res = pool.map_async(_cpu_frames_worker, jobs)
    ctx.map_clear()
    nvox, disp = ctx.pipeline_batch_host(*host_batch, want_disp=True)
    if world > 1:
        ctx.synchronize()
    got = ssm_dist.gather_map_native(ctx) if world > 1 else ctx.map_export()
    res = async_res.get()
    frames = [f for _, fr in res for f in fr]
    disp_mismatch = sum(int((disp[i] != frames[i][0]).sum()) for i in range(B))
    vm = oracle.VoxelMap(cfg["leaf"], cfg["labels"])
    for _, (xyz, rgba, lab) in frames:
        vm.insert(xyz, rgba, lab)
    ex = vm.export()
    part = {"key": _pack_ijk(ex["ijk"]), "count": ex["count"].astype(np.int64), "votes": ex["votes"].astype(np.int64),
            "sum": ex["centroid_d"] * ex["count"][:, None].astype(np.float64), "disp_mismatch": disp_mismatch, "points": int(ex["count"].sum())}
    if world > 1:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0)
    else:
        parts = [part]
    out = None
    if rank == 0:
        key = np.concatenate([q["key"] for q in parts])
        uniq, inv = np.unique(key, return_inverse=True)
        cnt = np.zeros(len(uniq), np.int64)
        np.add.at(cnt, inv, np.concatenate([q["count"] for q in parts]))
        votes = np.zeros((len(uniq), cfg["labels"]), np.int64)
        np.add.at(votes, inv, np.concatenate([q["votes"] for q in parts]))
        sums = np.zeros((len(uniq), 3), np.float64)
        np.add.at(sums, inv, np.concatenate([q["sum"] for q in parts]))
        # pcl::VoxelGrid order == ascending (k, j, i) == ascending packed key (k in the top bits): the same order as the export
        gkey = _pack_ijk(got["ijk"])
        same_set = len(gkey) == len(uniq) and bool((gkey == uniq).all())
        if same_set:
            count_mm = int((got["count"].astype(np.int64) != cnt).sum())
            vote_mm = int((got["votes"].astype(np.int64) != votes).sum())
            best = np.where(votes.max(axis=1) > 0, votes.argmax(axis=1), 255)
            label_mm = int((got["label"].astype(np.int64) != best).sum())
            ref = sums / cnt[:, None]
            err = np.abs(got["xyz"].astype(np.float64) - ref) / np.maximum(np.abs(ref), 1.0)
            cent_max = float(err.max()) if len(err) else 0.0
        else:
            count_mm = vote_mm = label_mm = -1
            cent_max = float("inf")
        dm = sum(q["disp_mismatch"] for q in parts)
        ok = same_set and count_mm == 0 and vote_mm == 0 and label_mm == 0 and dm == 0 and cent_max <= 1e-5
        out = {"ok": bool(ok), "nranks": world, "frames": B * world, "disp_mismatch": dm, "voxel_set_equal": bool(same_set),
               "voxels": int(len(uniq)), "points": int(sum(q["points"] for q in parts)), "count_mismatch": count_mm, "vote_mismatch": vote_mm,
               "label_mismatch": label_mm, "centroid_max_rel_err": cent_max, "centroid_tol": 1e-5,
               "checker": "cv2 4.13 StereoSGBM + C oracle glue/VoxelGrid (CPU), map gathered through ssm_map_export_gathered" if world > 1 else
                          "cv2 4.13 StereoSGBM + C oracle glue/VoxelGrid (CPU), map through ssm_map_export",
               "seconds": round(time.perf_counter() - t0, 2)}
    ctx.map_clear()
    return out


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, cfg, rank: int, world: int, local_rank: int, pool, cores: int):
    import torch
    import torch.distributed as dist
    from semantic_slam_mapping_b200 import Context, synth

    W, H, D, LABELS = cfg["W"], cfg["H"], cfg["D"], cfg["labels"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch or cfg["batch"]
    strong = cfg["scaling"] == "strong"
    # frames of this rank: strong scaling shards a fixed sequence, weak scaling gives every rank seq_frames of its own
    if strong:
        from semantic_slam_mapping_b200.distributed import shard_frames
        mine = shard_frames(cfg["seq_frames"], rank, world)
        my_first, my_frames, total_seq = mine.start, len(mine), cfg["seq_frames"]
    else:
        my_first, my_frames, total_seq = rank * cfg["seq_frames"], cfg["seq_frames"], world * cfg["seq_frames"]
    steps = args.steps if args.steps else (-(-my_frames // B) if cfg["steps"] is None else cfg["steps"])
    if strong:
        steps = min(steps, -(-((cfg["seq_frames"] + world - 1) // world) // B))
    if world > 1:   # every rank runs the same number of steps (the exchange is collective)
        t = torch.tensor([steps], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steps = int(t.item())
    # initial table size: enough for the whole run at a load factor below one half, so that no growth step (a re-insertion of
    # every record, tests/test_gpu_mapper.py::test_streaming_growth_from_tiny_table) falls into the timed region; a smaller
    # --map-capacity shows the growth path (config.map_rank0.grow_steps)
    p = make_params(cfg, B, args.map_capacity or (cfg["capacity"] // world if strong else cfg["capacity"]))
    ctx = Context(p, device=local_rank)
    if world > 1:
        from semantic_slam_mapping_b200 import distributed as ssm_dist
        ssm_dist.init_comm(ctx, p2p=not args.no_p2p)   # NCCL communicator + (default) peer-memory inboxes over NVLink
        ctx.set_route_overlap(not args.no_route_overlap)   # a batch's point exchange overlaps the next batch's SGBM

    # ---- inputs: a pool of `distinct` frames (cyclic, with B wrap-around copies so that every batch is one slice) ----------
    distinct = max(1, min(args.distinct or cfg.get("distinct", 100), max(my_frames, B)))
    t_gen = time.perf_counter()
    seeds = [image_seed(rank, k) for k in range(distinct)]
    imgs = [frame_images(cfg, s) for s in seeds]
    idx = [k % distinct for k in range(distinct + B)]
    host = {"left": np.stack([imgs[k][0] for k in idx]), "right": np.stack([imgs[k][1] for k in idx]),
            "semantic": np.stack([imgs[k][2] for k in idx]), "rgb": np.stack([imgs[k][3] for k in idx])}
    del imgs
    n_poses = (args.warmup + steps + 2) * B
    # one trajectory over all ranks' frames; with weak scaling the ranks' blocks follow each other, and every pass over a
    # rank's block continues where the last rank's block ended (new ground on every step)
    traj_len = max(total_seq, world * n_poses) + world * n_poses
    traj = synth.poses(traj_len, 11)

    def pose_block(step: int) -> np.ndarray:
        """Poses of this rank's batch `step` (warm-up steps use negative numbers and the tail of the trajectory).  Strong
        scaling: slots past the end of the rank's shard get a NaN pose -- their points are not finite and are dropped like
        pcl's is_dense == false rule drops them, so a filler frame in a collective step leaves the map untouched."""
        if strong:
            if step < 0:
                return np.ascontiguousarray(traj[[traj_len - 1 - ((-step) * B + i) for i in range(B)]])
            out = np.full((B, 4, 4), np.nan)
            n = max(0, min(B, my_frames - step * B))
            out[:n] = traj[my_first + step * B: my_first + step * B + n]
            return out
        first = (step * world + rank) * B if step >= 0 else traj_len - ((-step) * world + rank + 1) * B
        return np.ascontiguousarray(traj[first:first + B])
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
    devb = {k: v.to(dev) for k, v in pin.items()}
    pose_steps = np.stack([pose_block(i) for i in range(-args.warmup, steps)])       # [warmup + steps][B][4][4]
    pin_pose = torch.from_numpy(pose_steps).pin_memory()
    dev_pose = pin_pose.to(dev)
    log(f"[rank {rank}] inputs: {distinct} distinct frames, {steps} steps x {B} frames, generated in {time.perf_counter() - t_gen:.1f} s")
    tstream = torch.cuda.Stream(device=dev)      # the stream every kernel of the timed region is launched on
    stream = tstream.cuda_stream
    assert stream != 0

    def frames_in_step(i: int) -> int:
        """Frames of step i that belong to the sequence (strong scaling: the rank's last batch may be partial)."""
        if not strong or i < 0:
            return B
        return max(0, min(B, my_frames - i * B))

    def slot(i: int) -> int:
        return (i * B) % distinct if i >= 0 else ((-i) * B) % distinct

    def step_device(i):
        n = max(frames_in_step(i), 1)     # the exchange is collective: a rank past the end of its shard takes part with one filler frame (NaN pose)
        j = slot(i)
        ctx.pipeline_batch_device(devb["left"][j:j + n], devb["right"][j:j + n], devb["semantic"][j:j + n], devb["rgb"][j:j + n],
                                  dev_pose[i + args.warmup][:n], n, W, H, stream=stream)

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

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- parity gate (outside every timed region) -------------------------------------------------------------------
    parity = None
    if not args.no_parity:
        j = slot(0)
        hb = [host[k][j:j + B] for k in ("left", "right", "semantic", "rgb")] + [pose_steps[args.warmup]]
        parity = parity_gate(ctx, cfg, p, rank, world, pool, max(1, cores // world), hb, [seeds[(j + k) % distinct] for k in range(B)],
                             pose_steps[args.warmup], dist)
        if rank == 0:
            log(f"parity: {parity}")

    # ---- per-stage device times (a separate short pass: stage events serialise the sub-batch streams) ------------
    # (a first untimed pass over the same steps lets the voxel hash grow to the size these steps need: a growth step allocates,
    # which synchronises the device in the middle of whichever stage is running)
    n_stage = max(1, min(steps, 5))
    for timed in (False, True):
        ctx.map_clear()
        for i in range(-args.warmup, 0):
            step_device(i)
        barrier()
        ctx.set_stage_timing(timed)
        for i in range(n_stage):
            step_device(i)
        barrier()
    stage = ctx.stage_times_ms()
    ctx.set_stage_timing(False)

    # ---- device-resident throughput ("value") --------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                     # nvidia-smi needs a moment to start: launch it before the warm-up
    ctx.map_clear()
    for i in range(-args.warmup, 0):
        step_device(i)
    barrier()
    ctx.map_clear()                        # the warm-up frames lie elsewhere on the trajectory: the timed map starts empty
    barrier()
    t_begin = time.perf_counter()
    launches0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for i in range(steps):
        step_device(i)
    if world > 1:
        ctx.synchronize()      # route overlap: the last batch's exchange runs on the library's route stream -- inside the timed region
    e1.record(tstream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.kernel_launches() - launches0
    n_vox = ctx.map_size()
    stats = ctx.map_stats()
    clk = clocks.stop(t_begin, time.perf_counter()) if rank == 0 else None
    my_total = sum(max(frames_in_step(i), 0) for i in range(steps))
    total_frames = int(round(sum_over_ranks(float(my_total))))
    total_vox = int(round(sum_over_ranks(float(n_vox))))

    # points per frame (for the algorithmic-byte model): one compact cloud on the first frame's disparity
    disp0 = ctx.sgbm(host["left"][0], host["right"][0])
    pts = len(ctx.generate_point_cloud(ctx.disparity_to_depth(disp0), host["semantic"][0], host["rgb"][0], pose_steps[0][0])["xyz"])

    # ---- end to end through the host entry point ("e2e") ----------------------------------------------------
    # Every step copies its inputs from pinned host memory (H2D inside the timed region) and copies the step's result
    # (the map size after the batch) back to pinned host memory.  The streaming entry point stages inputs through two
    # device buffer sets, so the copy of step k+1 overlaps the kernels of step k.
    result = torch.zeros(steps + args.warmup + 8, dtype=torch.int32).pin_memory()

    def step_host(i, rslot):
        n = max(frames_in_step(i), 1)
        j = slot(i)
        ctx.pipeline_batch_host_async(pin["left"][j:j + n].numpy(), pin["right"][j:j + n].numpy(), pin["semantic"][j:j + n].numpy(),
                                      pin["rgb"][j:j + n].numpy(), pin_pose[i + args.warmup][:n].numpy(), result[rslot:rslot + 1])

    def e2e_run(export_every: int):
        ctx.map_clear()
        for i in range(-min(args.warmup, 3), 0):
            step_host(i, steps - i)
        ctx.synchronize()
        ctx.map_clear()
        barrier()
        exported = 0
        t0 = time.perf_counter()
        for i in range(steps):
            step_host(i, i)
            if export_every and (i + 1) % export_every == 0:
                m = ctx.map_export(sorted=True, into=export_pin)     # K9 finalize + radix sort on the device, D2H into pinned arrays
                exported = len(m["rgba"])
        ctx.synchronize()
        barrier()
        return max_over_ranks(time.perf_counter() - t0), exported

    e2e_s, _ = e2e_run(0)
    assert int(result[steps - 1]) > 0 and ctx.map_size() > 0      # results arrived; no capacity error
    e2e = total_frames / e2e_s

    # ---- e2e with the map leaving the GPU after every step (Mapper::viewer shows the fused cloud on every update) ----------
    exp = None
    if not args.no_export:
        cap = int(n_vox * 1.05) + 1024
        export_pin = {"xyz": torch.empty((cap, 3), dtype=torch.float32).pin_memory().numpy(),
                      "rgba": torch.empty(cap, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
                      "label": torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()}
        every = max(1, args.export_every)
        exp_s, exported = e2e_run(every)
        ctx.map_export(sorted=True, into=export_pin)
        exp = {"value": total_frames / exp_s, "unit": "frames/s", "export_every_steps": every, "voxels_last_export_rank0": int(exported),
               "d2h_bytes_last_export_rank0": int(exported) * 17, "export_device_ms_last": round(ctx.map_export_device_ms(), 4),
               "api": "ssm_pipeline_batch_host_async + ssm_map_export(sorted, xyz + rgba + label) into pinned arrays, per rank"}

    value = total_frames / (dev_ms * 1e-3)
    if rank != 0:
        return True

    # ---- roofline of the dominant stage ------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ab = algorithmic_bytes(cfg, pts)
    kernels = {"cost": "k_prefilter_tab + k_cost_tma (TMA-fed fused cost kernel)", "vertical": "k_vertical3 (cluster kernel)", "horizontal": "k_hfwd + k_hrev",
               "select": "k_select_fused (records -> L-R check -> median -> band-local speckle components)", "post": "k_cc_merge_bands + k_cc_count_roots + k_cc_apply_bands", "points": "k_depth + k_labels + k_moving_mask",
               "fuse": "k_points_fuse" if world == 1 else "k_points_p2p + k_flag_barrier + k_fuse_list"}
    dom = max(stage, key=stage.get)
    achieved = ab[dom] * B / (stage[dom] * 1e-3) / 1e9
    stage_roof = {k: {"ms_per_step": round(v, 4), "alg_GBps": round(ab[k] * B / (v * 1e-3) / 1e9, 1) if v > 0 else None,
                      "frac": round(ab[k] * B / (v * 1e-3) / 1e9 / peak, 4) if v > 0 else None}
                  for k, v in stage.items()}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath) and args.config == 1:
        tj = json.load(open(tpath))
        if tj.get(dom) is not None:      # measured for launches of _frames_per_launch frames; a launch of the stage pass has B
            traffic = round(tj[dom] * B / tj.get("_frames_per_launch", 33))

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_reference_step(pool, cfg, cores, 1, 777)
        dt, n = cpu_reference_step(pool, cfg, cores, args.ref_frames_per_core, 778)
        cpu = {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n} frames ({W}x{H}, D={D}): cv2 4.13 StereoSGBM (single-thread per process) + C-oracle glue, "
                         f"one process per core, single-threaded voxel merge; {dt:.1f} s wall"}

    frames_per_rank_step = total_frames / world / steps
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "config_index": args.config, "frames_per_step_per_gpu": B, "frames_total": total_frames,
                   "distinct_frames_per_gpu": distinct,
                   "l2": "each step streams > 2 GB of cost volumes through HBM (>> 126 MB L2); the input pool is cycled",
                   "voxels_in_map": total_vox, "voxels_in_map_rank0": n_vox, "points_per_frame": pts,
                   "map_rank0": {"slots": stats["slots"], "load_factor": round(stats["load_factor"], 4), "mean_probe": round(stats["mean_probe"], 3),
                                 "max_probe": stats["max_probe"], "grow_steps": stats["grow_steps"], "table_GB": round(stats["table_bytes"] / 1e9, 2)},
                   "parallelism": (f"frames sharded over {world} GPU(s); voxel hash spatially owned; points routed to the owner by "
                                   + ("NCCL send/recv all-to-all" if args.no_p2p else "peer-memory stores over NVLink fused into the point kernel + arrival flags in the peers' inbox headers (no collective)"))
                   if world > 1 else "1 GPU"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(round(frames_per_rank_step * (2 * W * H + 6 * W * H + 128))), "d2h_bytes_per_step": 4,
                "api": "ssm_pipeline_batch_host_async (pinned host buffers, double-buffered staging) + ssm_synchronize"},
        "e2e_with_export": exp,
        "parity": parity,
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
    return parity is None or parity["ok"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 30 for config 1; one pass over the rank's frames for configs 2 and 4)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="BASELINE.json configs[k]")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (default 66 = two sub-batches of 33, each one full wave of 4-CTA clusters on 132 SMs; 14 = two waves of seven 16-CTA clusters for config 3)")
    ap.add_argument("--no-route-overlap", action="store_true", help="N > 1: keep the point exchange on the pipeline stream")
    ap.add_argument("--distinct", type=int, default=0, help="distinct synthetic frames generated on the host per GPU (default 100; 24 for config 3)")
    ap.add_argument("--map-capacity", type=int, default=0, help="initial voxel hash slots (the table doubles when half full)")
    ap.add_argument("--ref-frames-per-core", type=int, default=2)
    ap.add_argument("--export-every", type=int, default=1, help="e2e_with_export: export the map after every k-th step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-export", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL send/recv all-to-all instead of peer-memory routing")
    args = ap.parse_args()
    capture_stdout()
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the CPU worker pool (reference arm, cpu_baseline leg, parity checker) is forked before CUDA / NCCL are touched
    import multiprocessing as mp
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    nproc = cores if (args.impl == "reference" or world == 1) else max(1, cores // world)
    if args.impl == "reference" and rank != 0:
        return
    pool = mp.get_context("fork").Pool(nproc)
    try:
        if args.impl == "reference":
            if not args.steps:
                args.steps = 5
            run_reference(args, cfg, rank, world, pool, cores)
            return
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import torch
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ok = True
        try:
            ok = run_ours(args, cfg, rank, world, local_rank, pool, cores)
        finally:
            if world > 1:
                import torch.distributed as dist
                dist.destroy_process_group()
        if not ok:
            log("PARITY MISMATCH: the GPU path differs from the CPU checker (see the parity block of the JSON line)")
            sys.exit(3)
    finally:
        pool.terminate()


if __name__ == "__main__":
    main()
