"""-m gpu tests of the GPU DEFLATE decoder behind the PNG ingest (k_inflate, one warp per stream): bit-exact with zlib on
streams of every block type, and the error behaviour of inflate() for corrupt streams."""
import struct
import zlib

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


def _payloads():
    rng = np.random.default_rng(2024)
    img = synth.stereo_pair(96, 320, 64, 5)[0]
    out = {
        "empty": b"",
        "one": b"x",
        "random": rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),                 # literals, long codes
        "zeros": bytes(100000),                                                            # distance 1, length 258 runs
        "period3": bytes([1, 2, 3]) * 30000,                                               # overlapping copies, dist < length
        "text": (b"the quick brown fox jumps over the lazy dog. " * 3000),
        "image": img.tobytes(),
        "sub_filtered": (np.diff(img.astype(np.int16), axis=1, prepend=0) & 255).astype(np.uint8).tobytes(),
        "skewed": rng.choice(np.arange(256, dtype=np.uint8), 200000, p=np.r_[0.5, np.full(255, 0.5 / 255)]).tobytes(),
        "far": rng.integers(0, 256, 32768, dtype=np.uint8).tobytes() * 3,                  # matches at the full 32 KB distance
        "geometric": rng.geometric(0.02, 150000).clip(0, 255).astype(np.uint8).tobytes(),  # > 10-bit codes in the literal table
    }
    return out


def _compress(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, flush_every=0):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    if flush_every <= 0:
        return co.compress(data) + co.flush()
    out = b""
    for i in range(0, len(data), flush_every):
        out += co.compress(data[i:i + flush_every]) + co.flush(zlib.Z_FULL_FLUSH if (i // flush_every) % 2 else zlib.Z_SYNC_FLUSH)
    return out + co.flush()


def test_every_block_type_matches_zlib(ctx):
    streams, want = [], []
    for name, data in _payloads().items():
        for kw in (dict(level=0), dict(level=1), dict(level=6), dict(level=9), dict(level=6, strategy=zlib.Z_FIXED),
                   dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY), dict(level=6, strategy=zlib.Z_RLE), dict(level=6, wbits=9),
                   dict(level=6, flush_every=4099), dict(level=0, flush_every=1000)):
            streams.append(_compress(data, **kw))
            want.append(data)
    assert len(streams) % 2 == 0
    for m in (len(streams), len(streams) - 1):      # even counts run the lean kernel, odd ones the literal-pair-table kernel
        outs, status = ctx.zlib_inflate_batch(streams[:m], [len(w) for w in want[:m]])
        assert status.tolist() == [0] * m
        for o, w in zip(outs, want):
            assert o.tobytes() == w


def test_output_window_smaller_than_the_stream_is_filled_and_the_rest_ignored(ctx):
    data = _payloads()["image"]
    z = _compress(data)
    outs, status = ctx.zlib_inflate_batch([z, z, z], [len(data) - 1, 1000, 0])
    assert status.tolist() == [0, 0, 0]
    assert outs[0].tobytes() == data[:-1] and outs[1].tobytes() == data[:1000]


def test_corrupt_streams_are_errors(ctx):
    data = _payloads()["text"]
    z = bytearray(_compress(data))
    n = len(data)
    bad_adler = bytes(z[:-1]) + bytes([z[-1] ^ 1])
    bad_header = bytes([z[0] ^ 0x0f]) + bytes(z[1:])
    truncated = bytes(z[:len(z) // 2])
    stored = _compress(data[:5000], level=0)
    bad_stored = bytearray(stored)
    bad_stored[5] ^= 0xff                                       # NLEN no longer the complement of LEN
    reserved_block = bytes(z[:2]) + bytes([0x07]) + bytes(z[3:])   # BTYPE = 3
    short = _compress(data[:100])
    streams = [bytes(z), bad_adler, bad_header, truncated, bytes(bad_stored), reserved_block, short]
    sizes = [n, n, n, n, 5000, n, 200]
    outs, status = ctx.zlib_inflate_batch(streams, sizes)
    assert status[0] == 0 and outs[0].tobytes() == data
    assert status[1] == 7 and status[2] == 1 and status[4] == 2 and status[5] == 2 and status[6] == 8
    assert status[3] != 0
    for s, m in zip(streams[1:], sizes[1:]):                    # zlib agrees that each of them is an error (or short)
        d = zlib.decompressobj()
        try:
            got = d.decompress(s)
            assert len(got) < m or not d.eof
        except zlib.error:
            pass


def test_random_bit_flips_never_disagree_with_zlib(ctx):
    """A stream with a flipped bit either fails in both decoders or decodes to the same bytes in both."""
    rng = np.random.default_rng(77)
    data = _payloads()["sub_filtered"][:20000]
    z = _compress(data, level=6)
    streams = []
    for _ in range(200):
        b = bytearray(z)
        i = int(rng.integers(2, len(b) - 4))
        b[i] ^= 1 << int(rng.integers(0, 8))
        streams.append(bytes(b))
    outs, status = ctx.zlib_inflate_batch(streams, [len(data)] * len(streams))
    outs1, status1 = ctx.zlib_inflate_batch(streams[:-1], [len(data)] * (len(streams) - 1))   # the other kernel variant
    assert status1.tolist() == status[:-1].tolist() and all(a.tobytes() == b.tobytes() for a, b in zip(outs1, outs))
    n_ok = 0
    for s, o, st in zip(streams, outs, status):
        # the reference semantics: inflate() with all the input and avail_out = the expected size (what the host path does)
        try:
            got = zlib.decompressobj().decompress(s, len(data))
            ok = len(got) == len(data)
        except zlib.error:
            ok = False
        assert (st == 0) == ok, f"GPU status {st}, zlib {'accepts' if ok else 'rejects'}"
        if ok:
            n_ok += 1
            assert o.tobytes() == got
    assert 0 < n_ok < len(streams)


@pytest.mark.parametrize("ctype,filters,level", [(0, None, 6), (2, [4], 9), (2, [1], 1), (6, [4, 3], 6), (3, None, 6), (4, [1, 2], 0)])
def test_png_batches_through_the_gpu_decoder_match_cv2(ctx, ctype, filters, level):
    import torch
    cv2 = pytest.importorskip("cv2")
    H, W, B = 376, 1241, 5
    rng = np.random.default_rng(ctype * 11 + level)
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    pngs = []
    for b in range(B):
        base = synth.stereo_pair(H, W, 64, 400 + b)[0]
        img = np.stack([np.roll(base, 3 * c, axis=1) for c in range(ch)], axis=-1) if ch > 1 else base
        if ctype == 3:
            img = base % 19
        pngs.append(oracle.png_encode(img, ctype, filters=filters, palette=rng.integers(0, 256, (19, 3), dtype=np.uint8), level=level,
                                      idat_split=8192 if b % 2 else 0))
    for colour in (False, True):
        d_out = torch.empty((B, H, W, 3) if colour else (B, H, W), dtype=torch.uint8, device="cuda")
        ctx.png_decode_batch_device(pngs, W, H, colour, d_out, host_threads=0)
        ctx.png_batch_wait()
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        for b in range(B):
            want = cv2.imdecode(np.frombuffer(pngs[b], np.uint8), cv2.IMREAD_COLOR if colour else cv2.IMREAD_GRAYSCALE)
            assert np.array_equal(got[b], want), (ctype, filters, colour, b)


def test_png_files_written_by_cv2_decode_on_the_gpu(ctx):
    import torch
    cv2 = pytest.importorskip("cv2")
    H, W, B = 96, 320, 4
    seq = synth.sequence(B, H, W, 64, 12, seed=9)
    for arr, colour in ((seq["left"], False), (seq["semantic"], True)):
        pngs = [cv2.imencode(".png", arr[b], [cv2.IMWRITE_PNG_COMPRESSION, 1 + 2 * b])[1].tobytes() for b in range(B)]
        d_out = torch.empty(arr.shape, dtype=torch.uint8, device="cuda")
        ctx.png_decode_batch_device(pngs, W, H, colour, d_out, host_threads=0)
        ctx.png_batch_wait()
        assert np.array_equal(d_out.cpu().numpy(), arr)


def test_gpu_decoder_reports_corrupt_png_streams(ctx):
    import torch
    H, W = 96, 320
    img = synth.stereo_pair(H, W, 64, 3)[0]
    good = oracle.png_encode(img, 0, level=6)
    # flip a byte inside the IDAT payload and repair the chunk CRC, so only the zlib layer can notice
    pos = good.index(b"IDAT")
    ln = struct.unpack(">I", good[pos - 4:pos])[0]
    body = bytearray(good[pos + 4:pos + 4 + ln])
    body[ln // 2] ^= 0x55
    bad = good[:pos + 4] + bytes(body) + struct.pack(">I", zlib.crc32(b"IDAT" + bytes(body)) & 0xffffffff) + good[pos + 8 + ln:]
    d_out = torch.empty((2, H, W), dtype=torch.uint8, device="cuda")
    ctx.png_decode_batch_device([good, bad], W, H, False, d_out, host_threads=0)
    with pytest.raises(SsmError):
        ctx.png_batch_wait()
    ctx.png_decode_batch_device([good, good], W, H, False, d_out, host_threads=0)   # the context keeps working
    ctx.png_batch_wait()
    assert np.array_equal(d_out.cpu().numpy()[1], img)
    bad_filter = oracle.png_encode(img, 0, filters=[7], level=6)
    ctx.png_decode_batch_device([bad_filter], W, H, False, d_out, host_threads=0)
    with pytest.raises(SsmError):
        ctx.png_batch_wait()
