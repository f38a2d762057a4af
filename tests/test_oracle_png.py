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
    # files with a significant gamma: libpng converts colour -> grey in linear light (cv::imread(path, 0) of such a file)
    import struct
    mixed = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    mixed[:5] = mixed[:5, :, :1]                              # pixels with R == G == B bypass the gamma tables
    yield "rgb_gama_45455", oracle.png_encode(mixed, 2, extra_chunks=[(b"gAMA", struct.pack(">I", 45455))])
    yield "rgb_gama_220000", oracle.png_encode(mixed, 2, extra_chunks=[(b"gAMA", struct.pack(">I", 220000))])
    yield "rgb_gama_96000_not_significant", oracle.png_encode(mixed, 2, extra_chunks=[(b"gAMA", struct.pack(">I", 96000))])
    yield "rgba_srgb_overrides_gama", oracle.png_encode(rng.integers(0, 256, (H, W, 4), dtype=np.uint8), 6,
                                                        extra_chunks=[(b"gAMA", struct.pack(">I", 100000)), (b"sRGB", b"\x00")])
    yield "palette_gama_55555", oracle.png_encode(rng.integers(0, 12, (H, W), dtype=np.uint8), 3, palette=pal,
                                                  extra_chunks=[(b"gAMA", struct.pack(">I", 55555))])
    yield "grey_gama_45455", oracle.png_encode(smooth, 0, extra_chunks=[(b"gAMA", struct.pack(">I", 45455))])


def _broken():
    """Files cv2 refuses (imdecode returns None): a CRC mismatch in IDAT / IHDR, a scanline filter type above 4."""
    import struct
    import zlib
    good = oracle.png_encode(np.arange(60, dtype=np.uint8).reshape(6, 10), 0)
    i = good.index(b"IDAT")
    n = struct.unpack(">I", good[i - 4:i])[0]
    bad_crc = bytearray(good)
    bad_crc[i + 4 + n] ^= 0xff                                # the stored CRC itself: the compressed data stay valid
    yield "idat_crc", bytes(bad_crc)
    raw = b"".join(bytes([5 if y == 2 else 0]) + bytes(range(10 * y, 10 * y + 10)) for y in range(6))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    yield "filter_type_5", (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 10, 6, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw))
                            + chunk(b"IEND", b""))
    bad_ihdr = bytearray(good)
    bad_ihdr[8 + 8 + 13] ^= 0x01
    yield "ihdr_crc", bytes(bad_ihdr)


@pytest.mark.parametrize("name,png", list(_broken()))
def test_png_files_cv2_refuses_are_errors(name, png):
    arr = np.frombuffer(png, np.uint8)
    assert cv2.imdecode(arr, cv2.IMREAD_GRAYSCALE) is None, name
    with pytest.raises(ValueError):
        oracle.png_decode(png, False)


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
