"""-m gpu parity tests of the PNG ingest through the C ABI: host inflate + GPU un-filtering / conversion
against cv2 4.13's decode (committed vectors and live) and against the oracle's restatement."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import Context, Params, synth
from semantic_slam_mapping_b200.lib import SsmError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with Context(Params(num_disparities=64, max_width=320, max_height=96, max_batch=4, map_capacity=1 << 18, resolution=0.05)) as c:
        yield c


def test_committed_vectors_from_cv2(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "png_cases.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        png = g[f"{n}/png"].tobytes()
        assert np.array_equal(ctx.png_decode(png, False), g[f"{n}/grey"]), n
        assert np.array_equal(ctx.png_decode(png, True), g[f"{n}/bgr"]), n
        w, h, ch = ctx.png_info(png)
        assert (h, w) == g[f"{n}/grey"].shape and ch in (1, 3)


@pytest.mark.parametrize("ctype,filters", [(2, None), (2, [1]), (2, [3]), (2, [4]), (0, None), (6, [4, 3]), (4, [1, 2]), (3, None)])
def test_kitti_size_batches_match_oracle_and_cv2(ctx, ctype, filters):
    """Batches at the KITTI frame size (odd width), every filter type, every colour type, into device buffers."""
    import torch
    cv2 = pytest.importorskip("cv2")
    H, W, B = 376, 1241, 3
    rng = np.random.default_rng(ctype * 7 + (filters[0] if filters else 0))
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    pngs = []
    for b in range(B):
        base = synth.stereo_pair(H, W, 64, 300 + b)[0]
        img = np.stack([np.roll(base, 3 * c, axis=1) for c in range(ch)], axis=-1) if ch > 1 else base
        if ctype == 3:
            img = base % 19
        pngs.append(oracle.png_encode(img, ctype, filters=filters, palette=rng.integers(0, 256, (19, 3), dtype=np.uint8), level=1))
    for colour in (False, True):
        d_out = torch.empty((B, H, W, 3) if colour else (B, H, W), dtype=torch.uint8, device="cuda")
        ctx.png_decode_batch_device(pngs, W, H, colour, d_out, host_threads=2)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        for b in range(B):
            want = cv2.imdecode(np.frombuffer(pngs[b], np.uint8), cv2.IMREAD_COLOR if colour else cv2.IMREAD_GRAYSCALE)
            assert np.array_equal(got[b], want), (ctype, filters, colour, b)
    assert np.array_equal(oracle.png_decode(pngs[0], True)[:8], got[0][:8])      # the restatement agrees too (a few rows: it is slow)


def test_decoded_frames_feed_the_pipeline(ctx):
    """PNG files in, map out: the decoded device images are exactly the arrays the pipeline would have been given."""
    import torch
    cv2 = pytest.importorskip("cv2")
    H, W, D, B = 96, 320, 64, 2
    seq = synth.sequence(B, H, W, D, 12, seed=77)
    enc = lambda a: cv2.imencode(".png", a)[1].tobytes()
    dev = torch.device("cuda")
    d_left = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    d_right = torch.empty_like(d_left)
    d_sem = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    d_rgb = torch.empty_like(d_sem)
    ctx.png_decode_batch_device([enc(seq["left"][b]) for b in range(B)], W, H, False, d_left)
    ctx.png_decode_batch_device([enc(seq["right"][b]) for b in range(B)], W, H, False, d_right)
    ctx.png_decode_batch_device([enc(seq["semantic"][b]) for b in range(B)], W, H, True, d_sem)
    ctx.png_decode_batch_device([enc(seq["rgb"][b]) for b in range(B)], W, H, True, d_rgb)
    torch.cuda.synchronize()
    assert np.array_equal(d_left.cpu().numpy(), seq["left"]) and np.array_equal(d_sem.cpu().numpy(), seq["semantic"])
    d_pose = torch.from_numpy(np.ascontiguousarray(seq["pose"])).to(dev)
    ctx.map_clear()
    ctx.pipeline_batch_device(d_left, d_right, d_sem, d_rgb, d_pose, B, W, H)
    n_png = ctx.map_size()
    ctx.map_clear()
    n_raw, _ = ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"], want_disp=True)
    assert n_png == n_raw > 1000


def test_files_cv2_refuses_are_errors(ctx):
    """CRC mismatch in a critical chunk, scanline filter type 5: cv2.imdecode returns None, the C ABI returns an error."""
    from test_oracle_png import _broken
    for name, png in _broken():
        for colour in (False, True):
            with pytest.raises(SsmError):
                ctx.png_decode(png, colour)


def test_rejected_files(ctx):
    import struct
    import zlib
    good = oracle.png_encode(np.zeros((4, 5), np.uint8), 0)
    with pytest.raises(SsmError):
        ctx.png_decode(b"not a png at all, just some bytes that are long enough to look", False)
    with pytest.raises(SsmError):
        ctx.png_decode(good[:40], False)                                     # truncated
    ihdr16 = good[:8 + 8] + struct.pack(">IIBBBBB", 5, 4, 16, 0, 0, 0, 0)
    bad = ihdr16 + struct.pack(">I", zlib.crc32(ihdr16[12:]) & 0xffffffff) + good[8 + 8 + 13 + 4:]
    with pytest.raises(SsmError):
        ctx.png_decode(bad, False)                                           # 16 bits per sample
    import torch
    d = torch.empty((1, 8, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(SsmError):
        ctx.png_decode_batch_device([good], 8, 8, False, d)                  # size differs from the batch's frame size
