"""-m gpu parity tests of the label-production kernel through the C ABI: cv2 4.13 golden vectors and the numpy oracle."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import Context, Params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with Context(Params(num_disparities=64, max_width=512, max_height=256, max_batch=2, map_capacity=1 << 16)) as c:
        yield c


def test_golden_vectors(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "labels_cv2.npz"))
    lut = g["lut"]
    sem, raw = ctx.labels_from_indices(g["small_idx"], 125, 38, lut)
    assert (sem == g["small_sem"]).all() and (raw == g["small_raw"]).all()
    sem, raw = ctx.labels_from_indices(g["full_idx"], 97, 61, lut)
    assert (sem == g["full_sem_up"]).all() and (raw == g["full_raw_up"]).all()
    sem, raw = ctx.labels_from_indices(g["full_idx"], 17, 13, lut)
    assert (sem == g["full_sem_down"]).all() and (raw == g["full_raw_down"]).all()
    sem, raw = ctx.labels_from_indices(g["segnet_idx"], 1241, 376, lut)            # the reference's 0002.png at KITTI size
    assert (raw == g["segnet_raw"]).all() and (sem == lut[g["segnet_raw"]]).all()


@pytest.mark.parametrize("sw,sh,dw,dh", [(480, 360, 1241, 376), (480, 360, 2048, 1024), (33, 21, 300, 7), (1, 1, 40, 30), (64, 48, 64, 48)])
def test_matches_oracle(ctx, sw, sh, dw, dh):
    rng = np.random.default_rng(sw + dh)
    idx = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
    lut = rng.integers(0, 256, (256, 3)).astype(np.uint8)
    sem, raw = ctx.labels_from_indices(idx, dw, dh, lut)
    wsem, wraw = oracle.labels_from_indices(idx, dw, dh, lut)
    assert (raw == wraw).all() and (sem == wsem).all()


def test_device_batch_feeds_the_mapper(ctx):
    import torch
    B, sw, sh, dw, dh = 3, 480, 360, 400, 120
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 12, (B, sh, sw)).astype(np.uint8)
    lut = np.zeros((256, 3), np.uint8)
    lut[:12] = np.array(ctx.params.palette_bgr, np.uint8)
    dev = torch.device("cuda:0")
    d_idx = torch.from_numpy(idx).to(dev)
    d_sem = torch.empty((B, dh, dw, 3), dtype=torch.uint8, device=dev)
    d_raw = torch.empty((B, dh, dw), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.labels_from_indices_batch_device(d_idx, d_sem, B, sw, sh, dw, dh, lut, d_raw)
    ctx.synchronize()
    for i in range(B):
        wsem, wraw = oracle.labels_from_indices(idx[i], dw, dh, lut)
        assert (d_raw[i].cpu().numpy() == wraw).all() and (d_sem[i].cpu().numpy() == wsem).all()
    # the produced colour image is what Mapper::semantic_motion_fuse consumes
    mp = oracle.MapParams()
    sem0 = d_sem[0].cpu().numpy()
    assert (ctx.semantic_motion_fuse(sem0) == oracle.moving_mask(sem0, mp)).all()
