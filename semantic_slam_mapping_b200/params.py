"""`ssm_params` (include/ssm.h) mirrored for ctypes, with the reference's defaults.

SGBM values: /root/reference src/stereo.cpp:16-28.  Camera / mapper values: parameters.txt:37-41,50-54,63,97-98.
Palette: src/mapper.cpp:42-54,206-208 (12-class SegNet driving palette, BGR).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

SSM_MAX_LABELS = 20
LABEL_UNKNOWN = 255

SEGNET12_NAMES = ["sky", "building", "pole", "road_marking", "road", "pavement", "tree", "sign_symbol",
                  "fence", "car", "pedestrian", "cyclist"]
SEGNET12_BGR = [
    (128, 128, 128), (0, 0, 128), (128, 192, 192), (0, 69, 255), (128, 64, 128), (222, 40, 60),
    (0, 128, 128), (128, 128, 192), (128, 64, 64), (128, 0, 64), (0, 64, 64), (192, 128, 0),
]
# Cityscapes trainId palette (BGR) for the 19-class config of BASELINE.json
CITYSCAPES19_BGR = [
    (128, 64, 128), (232, 35, 244), (70, 70, 70), (156, 102, 102), (153, 153, 190), (153, 153, 153),
    (30, 170, 250), (0, 220, 220), (35, 142, 107), (152, 251, 152), (180, 130, 70), (60, 20, 220),
    (0, 0, 255), (142, 0, 0), (70, 0, 0), (100, 60, 0), (100, 80, 0), (230, 0, 0), (32, 11, 119),
]


class CParams(C.Structure):
    """Binary layout of `struct ssm_params`."""

    _fields_ = [
        ("min_disparity", C.c_int), ("num_disparities", C.c_int), ("block_size", C.c_int),
        ("p1", C.c_int), ("p2", C.c_int), ("disp12_max_diff", C.c_int), ("pre_filter_cap", C.c_int),
        ("uniqueness_ratio", C.c_int), ("speckle_window_size", C.c_int), ("speckle_range", C.c_int),
        ("cx", C.c_double), ("cy", C.c_double), ("fx", C.c_double), ("fy", C.c_double),
        ("baseline", C.c_double), ("scale", C.c_double),
        ("roix", C.c_double), ("roiy", C.c_double), ("roiz", C.c_double),
        ("resolution", C.c_double), ("max_distance", C.c_double),
        ("num_labels", C.c_int), ("palette_bgr", (C.c_uint8 * 3) * SSM_MAX_LABELS),
        ("drop_mask", C.c_uint32), ("dynamic_mask", C.c_uint32),
        ("dilate_iterations", C.c_int), ("colour_source", C.c_int),
        ("max_width", C.c_int), ("max_height", C.c_int), ("max_batch", C.c_int),
        ("map_capacity", C.c_uint64),
    ]


@dataclass
class Params:
    # stereo.cpp:16-28
    min_disparity: int = 0
    num_disparities: int = 80
    block_size: int = 11
    p1: int = 4 * 11 * 11
    p2: int = 32 * 11 * 11
    disp12_max_diff: int = 1
    pre_filter_cap: int = 63
    uniqueness_ratio: int = 10
    speckle_window_size: int = 100
    speckle_range: int = 32
    # parameters.txt:37-41,50-54,63
    cx: float = 607.1928
    cy: float = 185.2157
    fx: float = 718.8560
    fy: float = 718.8560
    baseline: float = 0.532331858
    scale: float = 1000.0
    roix: float = 20.0
    roiy: float = 5.0
    roiz: float = 40.0
    # parameters.txt:97-98
    resolution: float = 0.1
    max_distance: float = 40.0
    palette_bgr: list = field(default_factory=lambda: list(SEGNET12_BGR))
    drop_mask: int = (1 << 0) | (1 << 2) | (1 << 11)   # sky, pole, cyclist (mapper.cpp:41-55)
    dynamic_mask: int = (1 << 10) | (1 << 11)          # pedestrian, cyclist (mapper.cpp:206-208)
    dilate_iterations: int = 2
    colour_source: int = 0
    max_width: int = 1241
    max_height: int = 376
    max_batch: int = 1
    map_capacity: int = 1 << 22

    @property
    def num_labels(self) -> int:
        return len(self.palette_bgr)

    def c(self) -> CParams:
        p = CParams()
        for name, _ in CParams._fields_:
            if name in ("palette_bgr", "num_labels"):
                continue
            setattr(p, name, getattr(self, name))
        p.num_labels = self.num_labels
        for i, (b, g, r) in enumerate(self.palette_bgr):
            p.palette_bgr[i][0], p.palette_bgr[i][1], p.palette_bgr[i][2] = b, g, r
        return p


def cityscapes_params(**kw) -> Params:
    """BASELINE.json configs[3]: 2048x1024, 256 disparities, 19 classes (calibration: SURVEY section 8d)."""
    base = dict(num_disparities=256, cx=1096.98, cy=513.137, fx=2262.52, fy=2262.52, baseline=0.209,
                palette_bgr=list(CITYSCAPES19_BGR), drop_mask=(1 << 10), dynamic_mask=(1 << 11) | (1 << 12),
                max_width=2048, max_height=1024)
    base.update(kw)
    return Params(**base)
