"""-m gpu parity tests of the stereo half: CUDA path (through the C ABI) vs the CPU oracle and the cv2 goldens."""
import os

import numpy as np
import pytest

import oracle
from semantic_slam_mapping_b200 import Context, Params, synth

pytestmark = pytest.mark.gpu


def _params(D, W, H, **kw):
    return Params(num_disparities=D, max_width=W, max_height=H, **kw)


def _oparams(p: Params):
    return oracle.SgbmParams(num_disparities=p.num_disparities, block_size=p.block_size, p1=p.p1, p2=p.p2,
                             disp12_max_diff=p.disp12_max_diff, pre_filter_cap=p.pre_filter_cap,
                             uniqueness_ratio=p.uniqueness_ratio, speckle_window_size=p.speckle_window_size,
                             speckle_range=p.speckle_range)


@pytest.mark.parametrize("H,W,D,seed", [(48, 96, 32, 1), (64, 160, 48, 2), (100, 300, 80, 3), (120, 400, 128, 4),
                                         (37, 53, 16, 8), (70, 350, 256, 9)])
def test_stages_bit_exact_vs_oracle(H, W, D, seed):
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    p = _params(D, W, H)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        C = ctx.debug_volume("C", W, H)
        Sf = ctx.debug_volume("S", W, H)   # the device keeps S_f = sat(L0+L1+L2+L3); the final S lives in registers only
        raw = ctx.debug_volume("disp_raw", W, H)
        med = ctx.debug_volume("disp_median", W, H)
    assert int((C != vols["C"]).sum()) == 0, "matching cost"
    # the device keeps S_v = sat(L1+L2+L3) (checkpointed horizontal sweep) or S_f = S_v + L0 (one-kernel sweep)
    assert int((Sf != vols["Sv"]).sum()) == 0 or int((Sf != vols["Sf"]).sum()) == 0, "aggregated cost"
    assert int((raw != vols["disp_raw"]).sum()) == 0, "WTA / uniqueness / sub-pixel / L-R check"
    assert int((med != vols["disp_median"]).sum()) == 0, "median"
    assert int((got != want).sum()) == 0, "speckle filter / final disparity"


def test_small_goldens_from_cv2(golden_dir):
    z = np.load(os.path.join(golden_dir, "sgbm_small.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    for n in names:
        D, bs, uniq, spw, spr, d12, cap = [int(v) for v in z[f"{n}/params"]]
        L, R = z[f"{n}/left"], z[f"{n}/right"]
        p = _params(D, L.shape[1], L.shape[0], block_size=bs, p1=4 * bs * bs, p2=32 * bs * bs, uniqueness_ratio=uniq,
                    speckle_window_size=spw, speckle_range=spr, disp12_max_diff=d12, pre_filter_cap=cap)
        with Context(p) as ctx:
            got = ctx.sgbm(L, R)
        assert int((got != z[f"{n}/disp"]).sum()) == 0, n


def test_kitti_d128_golden_from_cv2(golden_dir):
    z = np.load(os.path.join(golden_dir, "sgbm_kitti_d128.npz"))
    L, R, _ = synth.stereo_pair(376, 1241, 128, 0)
    with Context(_params(128, 1241, 376)) as ctx:
        got = ctx.sgbm(L, R)
    assert got.shape == (376, 1241) and got.dtype == np.int16
    assert int((got != z["disp"]).sum()) == 0


def test_cityscapes_d256_golden_from_cv2(golden_dir):
    """BASELINE configs[3] at full size: one 2048 x 1024 frame, 256 disparities (16-CTA clusters in the vertical kernel, 8 disparities
    per lane and 32-byte winner records in the horizontal sweep), against cv2 4.13's output committed by tests/golden/make_golden.py."""
    z = np.load(os.path.join(golden_dir, "sgbm_cityscapes_d256.npz"))
    L, R, _ = synth.stereo_pair(1024, 2048, 256, 0)
    assert int(L.astype(np.int64).sum()) == int(z["left_sum"]) and int(R.astype(np.int64).sum()) == int(z["right_sum"])
    with Context(_params(256, 2048, 1024)) as ctx:
        got = ctx.sgbm(L, R)
        import torch
        dL = torch.from_numpy(np.stack([L, L])).cuda()
        dR = torch.from_numpy(np.stack([R, R])).cuda()
    assert got.shape == (1024, 2048) and int((got != z["disp"]).sum()) == 0
    with Context(_params(256, 2048, 1024, max_batch=2)) as ctx:      # the batched shape of the bench (smaller clusters)
        dD = torch.empty((2, 1024, 2048), dtype=torch.int16, device="cuda")
        ctx.sgbm_batch_device(dL, dR, dD, 2, 2048, 1024)
        torch.cuda.synchronize()
        both = dD.cpu().numpy()
    assert int((both[0] != z["disp"]).sum()) == 0 and int((both[1] != z["disp"]).sum()) == 0


def test_reference_default_d80_full_frame():
    """The reference's own setting (src/stereo.cpp:18: 80 disparities) at KITTI size, vs the oracle."""
    L, R, _ = synth.stereo_pair(376, 1241, 80, 21)
    p = _params(80, 1241, 376)
    want = oracle.sgbm(L, R, _oparams(p))
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
    assert int((got != want).sum()) == 0


def test_batch_equals_single_frames():
    import torch
    H, W, D, B = 96, 320, 64, 5
    p = _params(D, W, H, max_batch=B)
    Ls, Rs = zip(*[synth.stereo_pair(H, W, D, 100 + i)[:2] for i in range(B)])
    dL = torch.from_numpy(np.stack(Ls)).cuda()
    dR = torch.from_numpy(np.stack(Rs)).cuda()
    dD = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    with Context(p) as ctx:
        ctx.sgbm_batch_device(dL, dR, dD, B, W, H, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        batch = dD.cpu().numpy()
        for i in range(B):
            assert int((ctx.sgbm(Ls[i], Rs[i]) != batch[i]).sum()) == 0
    op = _oparams(p)
    assert int((oracle.sgbm(Ls[2], Rs[2], op) != batch[2]).sum()) == 0


def test_noise_and_constant_inputs():
    rng = np.random.default_rng(5)
    L = rng.integers(0, 256, (60, 200), dtype=np.uint8)
    R = rng.integers(0, 256, (60, 200), dtype=np.uint8)
    p = _params(64, 200, 60)
    with Context(p) as ctx:
        assert int((ctx.sgbm(L, R) != oracle.sgbm(L, R, _oparams(p))).sum()) == 0
        flat = np.full((60, 200), 77, np.uint8)
        assert int((ctx.sgbm(flat, flat) != oracle.sgbm(flat, flat, _oparams(p))).sum()) == 0


def test_error_behaviour():
    from semantic_slam_mapping_b200 import SsmError
    with pytest.raises(SsmError):
        Context(Params(num_disparities=24))
    with pytest.raises(SsmError):
        Context(Params(block_size=13))
    with Context(_params(32, 100, 50)) as ctx:
        with pytest.raises(SsmError):
            ctx.sgbm(np.zeros((60, 100), np.uint8), np.zeros((60, 100), np.uint8))  # taller than the context
        with pytest.raises(SsmError):
            ctx.sgbm(np.zeros((10, 30), np.uint8), np.zeros((10, 30), np.uint8))    # W <= D


@pytest.mark.parametrize("min_cluster", [1, 2, 3, 4, 5, 6, 8])
@pytest.mark.parametrize("H,W,D,seed", [(40, 101, 32, 31), (33, 21, 16, 32), (50, 333, 128, 33), (24, 270, 256, 34), (30, 200, 80, 35)])
def test_vertical_cluster_kernel_all_cluster_sizes(monkeypatch, min_cluster, H, W, D, seed):
    """The three top-down paths in one cluster launch: strips of unequal / zero width, 1..8 CTAs per frame."""
    monkeypatch.setenv("SSM_MIN_CLUSTER", str(min_cluster))
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    p = _params(D, W, H)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        Sf = ctx.debug_volume("S", W, H)
    assert int((Sf != vols["Sv"]).sum()) == 0 or int((Sf != vols["Sf"]).sum()) == 0
    assert int((got != want).sum()) == 0


def test_legacy_per_direction_kernels_still_exact(monkeypatch):
    monkeypatch.setenv("SSM_LEGACY_VERTICAL", "1")
    L, R, _ = synth.stereo_pair(64, 240, 96, 41)
    p = _params(96, 240, 64)
    with Context(p) as ctx:
        assert int((ctx.sgbm(L, R) != oracle.sgbm(L, R, _oparams(p))).sum()) == 0


def test_vertical_cluster_kernel_batched_frames():
    import torch
    H, W, D, B = 60, 500, 128, 7
    p = _params(D, W, H, max_batch=B)
    Ls, Rs = zip(*[synth.stereo_pair(H, W, D, 300 + i)[:2] for i in range(B)])
    dL = torch.from_numpy(np.stack(Ls)).cuda()
    dR = torch.from_numpy(np.stack(Rs)).cuda()
    dD = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    with Context(p) as ctx:
        ctx.sgbm_batch_device(dL, dR, dD, B, W, H, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    op = _oparams(p)
    for i in (0, 3, 6):
        assert int((oracle.sgbm(Ls[i], Rs[i], op) != dD[i].cpu().numpy()).sum()) == 0


@pytest.mark.parametrize("H,W,D,bs,seed", [(45, 140, 16, 11, 51), (40, 120, 32, 11, 52), (52, 170, 64, 5, 53), (130, 333, 128, 11, 54),
                                            (61, 300, 256, 11, 55), (36, 150, 80, 11, 56), (20, 90, 48, 3, 57), (25, 200, 128, 1, 58)])
def test_fused_cost_kernel_matches_oracle_cost_volume(H, W, D, bs, seed):
    """Fused pixel-cost + box-sum kernel (every tile width, run-time and compile-time window, band edges)."""
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    p = _params(D, W, H, block_size=bs, p1=4 * bs * bs, p2=32 * bs * bs)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        C = ctx.debug_volume("C", W, H)
    assert int((C != vols["C"]).sum()) == 0
    assert int((got != want).sum()) == 0


@pytest.mark.parametrize("H,W,D,seed", [(130, 333, 128, 54), (36, 150, 80, 56), (47, 421, 112, 59)])
def test_in_kernel_table_cost_path_still_exact(monkeypatch, H, W, D, seed):
    """SSM_NO_COST_TMA=1: k_cost_fused (tables built inside the kernel) on the shapes that k_cost_tma serves by default."""
    monkeypatch.setenv("SSM_NO_COST_TMA", "1")
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    p = _params(D, W, H)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        C = ctx.debug_volume("C", W, H)
    assert int((C != vols["C"]).sum()) == 0 and int((got != want).sum()) == 0


@pytest.mark.parametrize("H,W,D,seed", [(33, 161, 128, 91), (70, 500, 128, 92), (64, 257, 96, 93), (41, 1241, 128, 94), (12, 140, 128, 95)])
def test_tma_table_cost_kernel_edges(H, W, D, seed):
    """k_prefilter_tab + k_cost_tma: widths that leave a partial last tile / partial last quad, a single short band, the padded margin."""
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    p = _params(D, W, H)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        C = ctx.debug_volume("C", W, H)
        # a second call with a smaller frame re-uses the table buffer with another pitch
        L2, R2, _ = synth.stereo_pair(H - 3, W - 9, D, seed + 100)
        got2 = ctx.sgbm(L2, R2)
    assert int((C != vols["C"]).sum()) == 0 and int((got != want).sum()) == 0
    assert int((got2 != oracle.sgbm(L2, R2, _oparams(p))).sum()) == 0


def test_legacy_two_kernel_cost_path_still_exact(monkeypatch):
    monkeypatch.setenv("SSM_LEGACY_COST", "1")
    L, R, _ = synth.stereo_pair(64, 240, 64, 61)
    p = _params(64, 240, 64)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        C = ctx.debug_volume("C", W=240, h=64) if False else ctx.debug_volume("C", 240, 64)
    assert int((C != vols["C"]).sum()) == 0 and int((got != want).sum()) == 0


@pytest.mark.parametrize("uniq", [0, 5, 10, 30, 99, 100])
def test_uniqueness_ratio_sweep(uniq):
    """The arithmetic uniqueness threshold of the checkpointed sweep against the oracle for several ratios."""
    L, R, _ = synth.stereo_pair(50, 200, 64, 71)
    rng = np.random.default_rng(7)
    R = np.clip(R.astype(int) + rng.integers(-12, 13, R.shape), 0, 255).astype(np.uint8)   # ambiguous matches
    p = _params(64, 200, 50, uniqueness_ratio=uniq)
    with Context(p) as ctx:
        assert int((ctx.sgbm(L, R) != oracle.sgbm(L, R, _oparams(p))).sum()) == 0


def test_legacy_one_kernel_horizontal_sweep_still_exact(monkeypatch):
    monkeypatch.setenv("SSM_LEGACY_HSWEEP", "1")
    L, R, _ = synth.stereo_pair(64, 300, 128, 81)
    p = _params(128, 300, 64)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        Sf = ctx.debug_volume("S", 300, 64)
    assert int((Sf != vols["Sf"]).sum()) == 0 and int((got != want).sum()) == 0


def test_cityscapes_width_256_disparities_band():
    """BASELINE configs[3] shape class: 2048 px wide, 256 disparities (NR = 4 path: generic kernels / 16-CTA clusters)."""
    H, W, D = 24, 2048, 256
    L, R, _ = synth.stereo_pair(H, W, D, 91)
    p = _params(D, W, H)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
    assert int((got != oracle.sgbm(L, R, _oparams(p))).sum()) == 0


@pytest.mark.parametrize("rows", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("H,W,D,seed", [(61, 233, 64, 101), (40, 300, 128, 102), (19, 150, 32, 103)])
def test_fused_selection_band_heights(monkeypatch, rows, H, W, D, seed):
    """Fused selection kernel (records -> L-R check -> median -> band-local speckle components) for several band
    heights: odd widths, bands that do not divide H, components that cross several band borders."""
    monkeypatch.setenv("SSM_SELECT_ROWS", str(rows))
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    rng = np.random.default_rng(seed)
    R = np.clip(R.astype(int) + rng.integers(-10, 11, R.shape), 0, 255).astype(np.uint8)   # speckles and rejected pixels
    p = _params(D, W, H)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        raw = ctx.debug_volume("disp_raw", W, H)
        med = ctx.debug_volume("disp_median", W, H)
    assert int((raw != vols["disp_raw"]).sum()) == 0, "L-R checked disparity"
    assert int((med != vols["disp_median"]).sum()) == 0, "median"
    assert int((got != want).sum()) == 0, "speckle filter"


@pytest.mark.parametrize("spw,spr", [(0, 0), (20, 1), (400, 2), (100, 32)])
def test_fused_selection_speckle_parameters(spw, spr):
    L, R, _ = synth.stereo_pair(90, 260, 64, 111)
    rng = np.random.default_rng(3)
    R = np.clip(R.astype(int) + rng.integers(-14, 15, R.shape), 0, 255).astype(np.uint8)
    p = _params(64, 260, 90, speckle_window_size=spw, speckle_range=spr)
    with Context(p) as ctx:
        assert int((ctx.sgbm(L, R) != oracle.sgbm(L, R, _oparams(p))).sum()) == 0


def test_legacy_separate_selection_kernels_still_exact(monkeypatch):
    monkeypatch.setenv("SSM_LEGACY_SELECT", "1")
    L, R, _ = synth.stereo_pair(64, 300, 128, 121)
    p = _params(128, 300, 64)
    with Context(p) as ctx:
        assert int((ctx.sgbm(L, R) != oracle.sgbm(L, R, _oparams(p))).sum()) == 0


@pytest.mark.parametrize("no_pad", [False, True])
@pytest.mark.parametrize("H,W,D,seed", [(60, 330, 80, 131), (45, 260, 96, 132), (38, 300, 112, 133), (52, 180, 48, 134), (30, 420, 192, 135)])
def test_padded_disparity_layouts(monkeypatch, no_pad, H, W, D, seed):
    """D = 80 (the reference's src/stereo.cpp:18), 96, 112 run on the 128-disparity kernels, D = 48 on the 64-disparity ones and
    D = 192 on the 256-disparity ones (lanes at d >= D switched off); SSM_NO_PAD=1 keeps the exact-D layout.  Both bit-exact,
    volumes included."""
    if no_pad:
        monkeypatch.setenv("SSM_NO_PAD", "1")
    L, R, _ = synth.stereo_pair(H, W, D, seed)
    p = _params(D, W, H)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        C = ctx.debug_volume("C", W, H)
        Sv = ctx.debug_volume("S", W, H)
    assert int((C != vols["C"]).sum()) == 0, "matching cost"
    assert int((Sv != vols["Sv"]).sum()) == 0 or int((Sv != vols["Sf"]).sum()) == 0, "aggregated cost"
    assert int((got != want).sum()) == 0


def test_padded_layout_batched_frames():
    """Batched frames through the whole pipeline at the reference's D = 80 (padded layout, sub-batch offsets in layout cells)."""
    H, W, D, B = 40, 210, 80, 3
    seq = synth.sequence(B, H, W, D, 12, seed=141)
    p = _params(D, W, H, max_batch=B, map_capacity=1 << 16)
    with Context(p) as ctx:
        _, got = ctx.pipeline_batch_host(seq["left"], seq["right"], seq["semantic"], seq["rgb"], seq["pose"], want_disp=True)
    for i in range(B):
        assert int((got[i] != oracle.sgbm(seq["left"][i], seq["right"][i], _oparams(p))).sum()) == 0


@pytest.mark.parametrize("uniq", [0, 10, 99])
@pytest.mark.parametrize("rows", [2, 8])
def test_256_disparities_checkpointed_sweep_and_fused_selection(monkeypatch, uniq, rows):
    """BASELINE configs[3] disparity count on the checkpointed horizontal sweep (8 disparities per lane, 32-byte winner
    records) and the fused selection kernel: ambiguous matches, winners at lane borders, several uniqueness ratios."""
    monkeypatch.setenv("SSM_SELECT_ROWS", str(rows))
    H, W, D = 36, 400, 256
    L, R, _ = synth.stereo_pair(H, W, D, 151 + uniq)
    rng = np.random.default_rng(uniq)
    R = np.clip(R.astype(int) + rng.integers(-12, 13, R.shape), 0, 255).astype(np.uint8)
    p = _params(D, W, H, uniqueness_ratio=uniq)
    want, vols = oracle.sgbm(L, R, _oparams(p), want_volumes=True)
    with Context(p) as ctx:
        got = ctx.sgbm(L, R)
        Sv = ctx.debug_volume("S", W, H)
        raw = ctx.debug_volume("disp_raw", W, H)
    assert int((Sv != vols["Sv"]).sum()) == 0, "S_v (the checkpointed sweep keeps the three top-down paths' sum)"
    assert int((raw != vols["disp_raw"]).sum()) == 0, "WTA / uniqueness / sub-pixel / L-R check"
    assert int((got != want).sum()) == 0


def test_legacy_one_kernel_sweep_256_disparities(monkeypatch):
    monkeypatch.setenv("SSM_LEGACY_HSWEEP", "1")
    L, R, _ = synth.stereo_pair(30, 380, 256, 161)
    p = _params(256, 380, 30)
    with Context(p) as ctx:
        assert int((ctx.sgbm(L, R) != oracle.sgbm(L, R, _oparams(p))).sum()) == 0
