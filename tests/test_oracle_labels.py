"""CPU tests of the label-production oracle: the numpy restatement of cv::resize(INTER_LINEAR) + cv::LUT
against the committed cv2 4.13 vectors (tests/golden/labels_cv2.npz, one of them the reference's own 0002.png mask) and,
where cv2 is importable, against cv2 live on fresh seeds."""
import os

import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "labels_cv2.npz"))


def test_oracle_matches_cv2_vectors(gold):
    lut = gold["lut"]
    sem, raw = oracle.labels_from_indices(gold["small_idx"], 125, 38, lut)
    assert (sem == gold["small_sem"]).all() and (raw == gold["small_raw"]).all()
    sem, raw = oracle.labels_from_indices(gold["full_idx"], 97, 61, lut)
    assert (sem == gold["full_sem_up"]).all() and (raw == gold["full_raw_up"]).all()
    sem, raw = oracle.labels_from_indices(gold["full_idx"], 17, 13, lut)
    assert (sem == gold["full_sem_down"]).all() and (raw == gold["full_raw_down"]).all()


def test_reference_segnet_mask_at_kitti_size(gold):
    # /root/reference/0002.png as class indices, 480x360 -> 1241x376 (the shipped pipeline's label step)
    sem, raw = oracle.labels_from_indices(gold["segnet_idx"], 1241, 376, gold["lut"])
    assert (raw == gold["segnet_raw"]).all()
    assert (sem == gold["lut"][gold["segnet_raw"]]).all()
    # interpolating indices creates in-between ids along class borders (reference behaviour, reproduced)
    assert set(np.unique(raw)) - set(np.unique(gold["segnet_idx"]))


def test_identity_and_edge_shapes():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (9, 13)).astype(np.uint8)
    assert (oracle.resize_linear_u8(img, 13, 9) == img).all()
    one = np.array([[7]], np.uint8)
    assert (oracle.resize_linear_u8(one, 5, 4) == 7).all()


def test_live_against_cv2():
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    rng = np.random.default_rng(7)
    for (sw, sh, dw, dh) in [(480, 360, 1241, 376), (480, 360, 2048, 1024), (64, 48, 200, 77), (100, 80, 37, 29), (31, 17, 311, 99)]:
        img = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        assert (cv2.resize(img, (dw, dh)) == oracle.resize_linear_u8(img, dw, dh)).all(), (sw, sh, dw, dh)
