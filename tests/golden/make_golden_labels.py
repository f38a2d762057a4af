#!/usr/bin/env python
"""Regenerates tests/golden/labels_cv2.npz with cv2 4.13 (cv2.resize default INTER_LINEAR + cv2.LUT): the label
production step of /root/reference experiment/segnet.cpp:121-135.  One case is the reference's own 480x360 SegNet mask
(/root/reference/0002.png, converted to class indices through the palette) resized to the KITTI frame size.

Run from the repo root:  python tests/golden/make_golden_labels.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from semantic_slam_mapping_b200.params import SEGNET12_BGR  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)


def lut_table():
    lut = np.zeros((256, 3), np.uint8)
    for i, c in enumerate(SEGNET12_BGR):
        lut[i] = c
    for i in range(len(SEGNET12_BGR), 256):
        lut[i] = (i, 255 - i, (7 * i) & 255)        # the rest of the table is arbitrary but fixed
    return lut


def run(idx, dw, dh, lut):
    three = np.repeat(idx[..., None], 3, axis=-1)                       # segnet.cpp:121-131: the index in all three channels
    resized = cv2.resize(three, (dw, dh))                               # :134
    sem = cv2.LUT(resized, lut.reshape(1, 256, 3))                      # :135 (per-channel tables)
    raw = cv2.cvtColor(resized, cv2.COLOR_BGR2GRAY)                     # rgbdframe.cpp:133
    return sem, raw


def main():
    lut = lut_table()
    out = {"lut": lut}
    rng = np.random.default_rng(3)
    small = rng.integers(0, 12, (36, 48)).astype(np.uint8)
    out["small_idx"] = small
    out["small_sem"], out["small_raw"] = run(small, 125, 38, lut)
    full = rng.integers(0, 256, (30, 40)).astype(np.uint8)                # every table entry, down- and up-scaling
    out["full_idx"] = full
    out["full_sem_up"], out["full_raw_up"] = run(full, 97, 61, lut)
    out["full_sem_down"], out["full_raw_down"] = run(full, 17, 13, lut)
    ref_png = "/root/reference/0002.png"
    if os.path.exists(ref_png):
        img = cv2.imread(ref_png)
        idx = np.full(img.shape[:2], 255, np.uint8)
        for i, c in enumerate(SEGNET12_BGR):
            idx[(img == np.array(c, np.uint8)).all(-1)] = i
        assert (idx != 255).all()
        out["segnet_idx"] = idx
        sem, raw = run(idx, 1241, 376, lut)
        out["segnet_raw"] = raw                                           # sem = lut[raw]; stored once
        assert (sem == lut[raw]).all()
    np.savez_compressed(os.path.join(OUT, "labels_cv2.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
