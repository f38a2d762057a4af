"""CPU tests of the PNG-ingest restatement: oracle.png_decode against cv2 4.13 (the library whose imread the
reference calls, src/rgbdframe.cpp:45-78, 138-180) on files written by cv2 and by a test encoder that uses all five scanline
filters and every supported colour type; committed vectors in tests/golden/png_cases.npz (generator make_golden_png.py)."""
import os

import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")


def _cases(seed=0):
    rng = np.random.default_rng(seed)
    H, W = 23, 37
    smooth = (np.add.outer(np.arange(H) * 5, np.arange(W) * 3) % 256).astype(np.uint8)
    yield "grey", oracle.png_encode(smooth ^ rng.integers(0, 8, (H, W), dtype=np.uint8), 0)
    yield "rgb", oracle.png_encode(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), 2)
    yield "rgba", oracle.png_encode(rng.integers(0, 256, (H, W, 4), dtype=np.uint8), 6)
    yield "grey_alpha", oracle.png_encode(rng.integers(0, 256, (H, W, 2), dtype=np.uint8), 4)
    pal = rng.integers(0, 256, (12, 3), dtype=np.uint8)
    yield "palette", oracle.png_encode(rng.integers(0, 12, (H, W), dtype=np.uint8), 3, palette=pal)
    yield "paeth_only_split_idat", oracle.png_encode(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), 2, filters=[4], idat_split=97)
    yield "one_pixel", oracle.png_encode(np.array([[[7, 200, 31]]], np.uint8), 2)
    ok, buf = cv2.imencode(".png", rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
    yield "written_by_cv2", buf.tobytes()


@pytest.mark.parametrize("name,png", list(_cases()))
def test_png_decode_matches_cv2(name, png):
    arr = np.frombuffer(png, np.uint8)
    for colour in (False, True):
        want = cv2.imdecode(arr, cv2.IMREAD_COLOR if colour else cv2.IMREAD_GRAYSCALE)
        got = oracle.png_decode(png, colour)
        assert got.shape == want.shape and np.array_equal(got, want), (name, colour)


def test_committed_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "png_cases.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) >= 8
    for n in names:
        png = g[f"{n}/png"].tobytes()
        assert np.array_equal(oracle.png_decode(png, False), g[f"{n}/grey"]), n
        assert np.array_equal(oracle.png_decode(png, True), g[f"{n}/bgr"]), n
