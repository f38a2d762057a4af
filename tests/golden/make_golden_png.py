#!/usr/bin/env python
"""Regenerates tests/golden/png_cases.npz: small PNG files (all five scanline filters, grey / grey + alpha / RGB / RGBA /
palette, split IDAT, one written by cv2 itself) together with what cv2 4.13 decodes them to (IMREAD_GRAYSCALE and
IMREAD_COLOR) -- the library call the reference makes for every image it reads (src/rgbdframe.cpp:45-78, 138-180).
Run from the repo root:  python tests/golden/make_golden_png.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_png import _cases  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for name, png in _cases():
        arr = np.frombuffer(png, np.uint8)
        out[f"{name}/png"] = arr
        out[f"{name}/grey"] = cv2.imdecode(arr, cv2.IMREAD_GRAYSCALE)
        out[f"{name}/bgr"] = cv2.imdecode(arr, cv2.IMREAD_COLOR)
    np.savez_compressed(os.path.join(OUT, "png_cases.npz"), **out)
    print("wrote png_cases.npz:", sorted({k.split('/')[0] for k in out}))


if __name__ == "__main__":
    main()
