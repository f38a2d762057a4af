"""The C++ host layer (semantic_slam_mapping_b200/host: calDisparity_SGBM + Mapper mirrors of the reference's
stereo.h / mapper.h) driven like experiment/exp_mapping.cpp drives the reference, checked against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import build, synth


def test_host_layer_builds_and_links():
    exe = build.build_host_driver()
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libssm.so" in out and "opencv" not in out and "pcl" not in out


@pytest.mark.gpu
def test_cpp_host_path_matches_oracle(tmp_path):
    exe = build.build_host_driver()
    n, H, W, D, leaf = 5, 96, 320, 64, 0.1
    seq = synth.sequence(n, H, W, D, 12, seed=17)
    fin, fout, pcd = tmp_path / "in.bin", tmp_path / "out.bin", tmp_path / "map.pcd"
    with open(fin, "wb") as f:
        f.write(struct.pack("<4id", n, H, W, D, leaf))
        for i in range(n):
            for k in ("left", "right", "semantic", "rgb"):
                f.write(np.ascontiguousarray(seq[k][i]).tobytes())
            f.write(np.ascontiguousarray(seq["pose"][i], np.float64).tobytes())
    r = subprocess.run([exe, str(fin), str(fout), str(pcd)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "points in global map" in r.stdout and "Mapping cost time" in r.stdout and "Map saved!" in r.stdout
    raw = open(fout, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off)
        off += a.nbytes
        return a

    mp = oracle.MapParams()
    vm = oracle.VoxelMap(leaf, 12)
    first_pc = None
    for i in range(n):
        disp = take(np.int16, H * W).reshape(H, W)
        want = oracle.sgbm(seq["left"][i], seq["right"][i], oracle.SgbmParams(num_disparities=D))
        assert int((disp != want).sum()) == 0
        pc = oracle.generate_point_cloud(oracle.disparity_to_depth(want, mp), seq["semantic"][i], seq["rgb"][i], mp, seq["pose"][i])
        first_pc = first_pc or pc
        vm.insert(pc["xyz"], pc["rgba"], pc["label"])
    want = vm.export()
    nv = int(take(np.int64, 1)[0])
    assert nv == len(vm) and nv > 1000
    ijk = take(np.int32, nv * 3).reshape(nv, 3)
    xyz = take(np.float32, nv * 3).reshape(nv, 3)
    rgba, count = take(np.uint32, nv), take(np.uint32, nv)
    votes = take(np.uint32, nv * 12).reshape(nv, 12)
    label = take(np.uint8, nv)
    assert (ijk == want["ijk"]).all() and (count == want["count"]).all() and (votes == want["votes"]).all()
    assert (label == want["label"]).all() and (rgba == want["rgba"]).all()
    ref = want["centroid_d"]
    assert (np.abs(xyz - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0)).all()
    npts = int(take(np.int64, 1)[0])
    assert npts == len(first_pc["xyz"])
    pts = take(np.uint32, npts * 4).reshape(npts, 4)
    assert (pts[:, :3] == first_pc["xyz"].view(np.uint32)).all() and (pts[:, 3] == first_pc["rgba"]).all()
    # dense motion cues through the C++ mirrors (triangulate10D, UVDisparity::calVDisparity, correct3DPoints, setImageROI,
    # UVDisparity::calUDisparity) on keyframe 0
    v_cols, u_rows = [int(v) for v in take(np.int32, 2)]
    got_xyz = take(np.float32, H * W * 10).reshape(H, W, 10)
    got_roi = take(np.uint8, H * W).reshape(H, W)
    got_vint = take(np.int32, H * v_cols).reshape(H, v_cols)
    got_u8 = take(np.uint8, u_rows * W).reshape(u_rows, W)
    disp0 = oracle.sgbm(seq["left"][0], seq["right"][0], oracle.SgbmParams(num_disparities=D))
    grey = np.ascontiguousarray(seq["rgb"][0][..., 0])
    ground = np.zeros((H, W), np.uint8)
    ground[H // 2:] = 255
    x = oracle.triangulate10d(grey, disp0, mp.fx, mp.cx, mp.cy, mp.baseline)
    x, wvint, _ = oracle.v_disparity(disp0, x)
    x = oracle.correct_3d_points(x, (15.0, 1.5, 35.0), 0.02, 0.0)
    wroi = oracle.set_image_roi(x)
    x, _, wu8 = oracle.u_disparity(disp0, x, wroi, ground)
    assert wvint.shape == (H, v_cols) and wu8.shape == (u_rows, W)
    assert (got_vint == wvint).all() and (got_u8 == wu8).all() and (got_roi == wroi).all()
    assert np.array_equal(got_xyz.view(np.uint32), x.view(np.uint32))
    head = open(pcd, "rb").read(300).decode("ascii", "ignore")
    assert f"POINTS {nv}" in head
