"""ctypes loader for the CPU oracle (oracle/ssm_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never from the product package
``semantic_slam_mapping_b200`` (tests/test_boundary.py greps for that).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssm_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/ssm_oracle.c with gcc (a few seconds)."""
    src = os.path.join(_HERE, "ssm_oracle.c")
    hdr = os.path.join(_HERE, "ssm_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libssm_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _SgbmParams(C.Structure):
    _fields_ = [
        ("num_disparities", C.c_int),
        ("block_size", C.c_int),
        ("p1", C.c_int),
        ("p2", C.c_int),
        ("disp12_max_diff", C.c_int),
        ("pre_filter_cap", C.c_int),
        ("uniqueness_ratio", C.c_int),
        ("speckle_window_size", C.c_int),
        ("speckle_range", C.c_int),
        ("legacy_p2_form", C.c_int),
    ]


class _MapParams(C.Structure):
    _fields_ = [
        ("cx", C.c_double),
        ("cy", C.c_double),
        ("fx", C.c_double),
        ("fy", C.c_double),
        ("baseline", C.c_double),
        ("scale", C.c_double),
        ("roix", C.c_double),
        ("roiy", C.c_double),
        ("roiz", C.c_double),
        ("max_distance", C.c_double),
        ("num_labels", C.c_int),
        ("palette_bgr", (C.c_uint8 * 3) * 32),
        ("drop_mask", C.c_uint32),
        ("dynamic_mask", C.c_uint32),
        ("dilate_iterations", C.c_int),
        ("colour_source", C.c_int),
    ]


@dataclass
class SgbmParams:
    """Defaults = src/stereo.cpp:16-28 of the reference."""

    num_disparities: int = 80
    block_size: int = 11
    p1: int = 4 * 11 * 11
    p2: int = 32 * 11 * 11
    disp12_max_diff: int = 1
    pre_filter_cap: int = 63
    uniqueness_ratio: int = 10
    speckle_window_size: int = 100
    speckle_range: int = 32
    legacy_p2_form: int = 0

    def c(self) -> _SgbmParams:
        return _SgbmParams(
            self.num_disparities, self.block_size, self.p1, self.p2, self.disp12_max_diff, self.pre_filter_cap,
            self.uniqueness_ratio, self.speckle_window_size, self.speckle_range, self.legacy_p2_form,
        )


# 12-class SegNet palette, BGR (src/mapper.cpp:42-54,206-208)
SEGNET12_BGR = [
    (128, 128, 128),  # 0 sky
    (0, 0, 128),      # 1 building
    (128, 192, 192),  # 2 pole
    (0, 69, 255),     # 3 road marking
    (128, 64, 128),   # 4 road
    (222, 40, 60),    # 5 pavement
    (0, 128, 128),    # 6 tree
    (128, 128, 192),  # 7 sign symbol
    (128, 64, 64),    # 8 fence
    (128, 0, 64),     # 9 car
    (0, 64, 64),      # 10 pedestrian
    (192, 128, 0),    # 11 cyclist
]


@dataclass
class MapParams:
    """Defaults = parameters.txt:37-41,50-54,63,97-98 and the shipped mapper.cpp class sets."""

    cx: float = 607.1928
    cy: float = 185.2157
    fx: float = 718.8560
    fy: float = 718.8560
    baseline: float = 0.532331858
    scale: float = 1000.0
    roix: float = 20.0
    roiy: float = 5.0
    roiz: float = 40.0
    max_distance: float = 40.0
    palette_bgr: list = field(default_factory=lambda: list(SEGNET12_BGR))
    drop_mask: int = (1 << 0) | (1 << 2) | (1 << 11)   # sky, pole, cyclist  (mapper.cpp:41-55)
    dynamic_mask: int = (1 << 10) | (1 << 11)          # pedestrian, cyclist (mapper.cpp:206-208)
    dilate_iterations: int = 2
    colour_source: int = 0

    def c(self) -> _MapParams:
        m = _MapParams()
        for k in ("cx", "cy", "fx", "fy", "baseline", "scale", "roix", "roiy", "roiz", "max_distance"):
            setattr(m, k, float(getattr(self, k)))
        m.num_labels = len(self.palette_bgr)
        for i, (b, g, r) in enumerate(self.palette_bgr):
            m.palette_bgr[i][0], m.palette_bgr[i][1], m.palette_bgr[i][2] = b, g, r
        m.drop_mask = self.drop_mask
        m.dynamic_mask = self.dynamic_mask
        m.dilate_iterations = self.dilate_iterations
        m.colour_source = self.colour_source
        return m


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.oracle_sgbm.restype = C.c_int
        L.oracle_sgbm.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.POINTER(_SgbmParams), vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
        L.oracle_median3x3_s16.argtypes = [vp, vp, C.c_int, C.c_int]
        L.oracle_filter_speckles.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_disparity_to_depth.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_MapParams), vp]
        L.oracle_moving_mask.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_MapParams), vp]
        L.oracle_generate_point_cloud.restype = C.c_int
        L.oracle_generate_point_cloud.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.POINTER(_MapParams), vp, vp, vp, vp, vp, vp]
        L.oracle_map_create.restype = vp
        L.oracle_map_create.argtypes = [C.c_double, C.c_int]
        L.oracle_map_destroy.argtypes = [vp]
        L.oracle_map_clear.argtypes = [vp]
        L.oracle_map_insert.argtypes = [vp, vp, vp, vp, C.c_int]
        L.oracle_map_size.restype = C.c_int64
        L.oracle_map_size.argtypes = [vp]
        L.oracle_map_export.restype = C.c_int64
        L.oracle_map_export.argtypes = [vp] * 8
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sgbm(left: np.ndarray, right: np.ndarray, params: SgbmParams, want_volumes: bool = False):
    """calDisparity_SGBM (src/stereo.cpp:11-38).  Returns disp int16 [H][W] (x16, invalid -16);
    with want_volumes also a dict of C, S, disp_raw (pre-median), disp_median (pre-speckle)."""
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    H, W = left.shape
    D = params.num_disparities
    disp = np.empty((H, W), np.int16)
    vols = {}
    Cv = Sv = raw = med = Sf = Svert = None
    if want_volumes:
        Cv = np.empty((H, W - D, D), np.int16)
        Sv = np.empty((H, W - D, D), np.int16)
        Sf = np.empty((H, W - D, D), np.int16)
        Svert = np.empty((H, W - D, D), np.int16)
        raw = np.empty((H, W), np.int16)
        med = np.empty((H, W), np.int16)
        vols = {"C": Cv, "S": Sv, "Sf": Sf, "Sv": Svert, "disp_raw": raw, "disp_median": med}
    cp = params.c()
    rc = lib().oracle_sgbm(_p(left), _p(right), W, H, W, C.byref(cp), _p(disp), W, _p(Cv), _p(Sv), _p(raw), _p(med), _p(Sf), _p(Svert))
    if rc != 0:
        raise ValueError(f"oracle_sgbm rejected the arguments (rc={rc})")
    return (disp, vols) if want_volumes else disp


def median3x3(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.int16)
    out = np.empty_like(img)
    lib().oracle_median3x3_s16(_p(img), _p(out), img.shape[1], img.shape[0])
    return out


def filter_speckles(img: np.ndarray, new_val: int, max_size: int, max_diff: int) -> np.ndarray:
    out = np.ascontiguousarray(img, np.int16).copy()
    lib().oracle_filter_speckles(_p(out), out.shape[1], out.shape[0], new_val, max_size, max_diff)
    return out


def disparity_to_depth(disp: np.ndarray, mp: MapParams) -> np.ndarray:
    disp = np.ascontiguousarray(disp, np.int16)
    depth = np.empty(disp.shape, np.uint16)
    cp = mp.c()
    lib().oracle_disparity_to_depth(_p(disp), disp.shape[1], disp.shape[0], C.byref(cp), _p(depth))
    return depth


def moving_mask(semantic_bgr: np.ndarray, mp: MapParams) -> np.ndarray:
    sem = np.ascontiguousarray(semantic_bgr, np.uint8)
    H, W = sem.shape[:2]
    out = np.empty((H, W), np.uint8)
    cp = mp.c()
    lib().oracle_moving_mask(_p(sem), W, H, C.byref(cp), _p(out))
    return out


def generate_point_cloud(depth, semantic_bgr, rgb_bgr, mp: MapParams, T):
    """Mapper::generatePointCloud (src/mapper.cpp:12-94).  Returns dict of row-major-ordered arrays."""
    depth = np.ascontiguousarray(depth, np.uint16)
    sem = np.ascontiguousarray(semantic_bgr, np.uint8)
    rgb = np.ascontiguousarray(rgb_bgr, np.uint8)
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    H, W = depth.shape
    n = H * W
    xyz = np.empty((n, 3), np.float32)
    cam = np.empty((n, 3), np.float32)
    rgba = np.empty(n, np.uint32)
    lab = np.empty(n, np.uint8)
    pix = np.empty(n, np.int32)
    cp = mp.c()
    k = lib().oracle_generate_point_cloud(_p(depth), _p(sem), _p(rgb), W, H, C.byref(cp), _p(T), _p(xyz), _p(cam), _p(rgba), _p(lab), _p(pix))
    return {"xyz": xyz[:k].copy(), "xyz_cam": cam[:k].copy(), "rgba": rgba[:k].copy(), "label": lab[:k].copy(), "pix": pix[:k].copy()}


class VoxelMap:
    """pcl::VoxelGrid<PointXYZRGBA> over the union of inserted clouds + per-voxel label votes."""

    def __init__(self, leaf: float, num_labels: int):
        self.num_labels = num_labels
        self._h = lib().oracle_map_create(float(leaf), num_labels)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_map_destroy(self._h)
            self._h = None

    def clear(self):
        lib().oracle_map_clear(self._h)

    def insert(self, xyz, rgba, label):
        xyz = np.ascontiguousarray(xyz, np.float32)
        rgba = np.ascontiguousarray(rgba, np.uint32)
        label = np.ascontiguousarray(label, np.uint8)
        lib().oracle_map_insert(self._h, _p(xyz), _p(rgba), _p(label), xyz.shape[0])

    def __len__(self):
        return int(lib().oracle_map_size(self._h))

    def export(self):
        n = len(self)
        L = self.num_labels
        out = {
            "ijk": np.empty((n, 3), np.int32),
            "centroid": np.empty((n, 3), np.float32),
            "centroid_d": np.empty((n, 3), np.float64),
            "rgba": np.empty(n, np.uint32),
            "count": np.empty(n, np.uint32),
            "votes": np.empty((n, L), np.uint32),
            "label": np.empty(n, np.uint8),
        }
        lib().oracle_map_export(self._h, _p(out["ijk"]), _p(out["centroid"]), _p(out["centroid_d"]), _p(out["rgba"]),
                                _p(out["count"]), _p(out["votes"]), _p(out["label"]))
        return out


# ---- dense motion cues ---------------------------------------------------------------------------
def _cue_sigs(L):
    vp = C.c_void_p
    d = C.c_double
    L.oracle_triangulate10d.argtypes = [vp, vp, C.c_int, C.c_int, d, d, d, d, vp]
    L.oracle_correct_3d_points.argtypes = [vp, C.c_int, C.c_int, d, d, d, d, d]
    L.oracle_set_image_roi.argtypes = [vp, C.c_int, C.c_int, vp]
    L.oracle_v_disparity.restype = C.c_int
    L.oracle_v_disparity.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.oracle_u_disparity.restype = C.c_int
    L.oracle_u_disparity.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    return L


def triangulate10d(img, disp, f, cx, cy, b) -> np.ndarray:
    """triangulate10D (src/stereo.cpp:41-118): [H][W][10] fp32."""
    img = np.ascontiguousarray(img, np.uint8)
    disp = np.ascontiguousarray(disp, np.int16)
    H, W = disp.shape
    xyz = np.empty((H, W, 10), np.float32)
    _cue_sigs(lib()).oracle_triangulate10d(_p(img), _p(disp), W, H, f, cx, cy, b, _p(xyz))
    return xyz


def correct_3d_points(xyz, roi, pitch1, pitch2=0.0) -> np.ndarray:
    """correct3DPoints (src/stereo.cpp:127-181); returns a corrected copy."""
    out = np.ascontiguousarray(xyz, np.float32).copy()
    H, W = out.shape[:2]
    _cue_sigs(lib()).oracle_correct_3d_points(_p(out), W, H, roi[0], roi[1], roi[2], pitch1, pitch2)
    return out


def set_image_roi(xyz) -> np.ndarray:
    """setImageROI (src/stereo.cpp:183-192)."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    H, W = xyz.shape[:2]
    m = np.empty((H, W), np.uint8)
    _cue_sigs(lib()).oracle_set_image_roi(_p(xyz), W, H, _p(m))
    return m


def v_disparity(disp, xyz, cap_cols=1024):
    """UVDisparity::calVDisparity (src/uvdisparity.cpp:277-366).  Returns (xyz with channel 8 filled, v_dis_int, v_dis)."""
    disp = np.ascontiguousarray(disp, np.int16)
    out = np.ascontiguousarray(xyz, np.float32).copy()
    H, W = disp.shape
    vi = np.zeros(H * cap_cols, np.int32)
    v8 = np.zeros(H * cap_cols, np.uint8)
    vc = _cue_sigs(lib()).oracle_v_disparity(_p(disp), W, H, _p(out), _p(vi), _p(v8), cap_cols)
    if vc < 0:
        raise ValueError("v-disparity wider than cap_cols")
    return out, vi[: H * vc].reshape(H, vc).copy(), v8[: H * vc].reshape(H, vc).copy()


def u_disparity(disp, xyz, roi_mask, ground_mask, cap_rows=1025):
    """UVDisparity::calUDisparity (src/uvdisparity.cpp:195-274).  Returns (xyz with channel 7 filled, u_dis_int, u_dis)."""
    disp = np.ascontiguousarray(disp, np.int16)
    out = np.ascontiguousarray(xyz, np.float32).copy()
    roi_mask = np.ascontiguousarray(roi_mask, np.uint8)
    ground_mask = np.ascontiguousarray(ground_mask, np.uint8)
    H, W = disp.shape
    ui = np.zeros(cap_rows * W, np.int32)
    u8 = np.zeros(cap_rows * W, np.uint8)
    ur = _cue_sigs(lib()).oracle_u_disparity(_p(disp), W, H, _p(out), _p(roi_mask), _p(ground_mask), _p(ui), _p(u8), cap_rows)
    if ur < 0:
        raise ValueError("u-disparity taller than cap_rows")
    return out, ui[: ur * W].reshape(ur, W).copy(), u8[: ur * W].reshape(ur, W).copy()


# ---- the reference's own src/stereo.cpp, compiled against oracle/cvstub (oracle/_ref/libref_stereo.so) ---------------
_REF_PATH = os.path.join(_HERE, "_ref", "libref_stereo.so")
_REFERENCE_ROOT = os.environ.get("SSM_REFERENCE_ROOT", "/root/reference")
_ref = None


def build_ref() -> str | None:
    """Compile the reference's src/stereo.cpp where it lies (only possible where /root/reference exists; the GPU box uses
    the prebuilt file that travels with the snapshot).  Returns the path, or None when neither source nor binary exists."""
    src = os.path.join(_REFERENCE_ROOT, "src", "stereo.cpp")
    if os.path.exists(src):
        deps = [src, os.path.join(_REFERENCE_ROOT, "src", "uvdisparity.cpp"), os.path.join(_HERE, "ref_stereo_wrap.cpp"),
                os.path.join(_HERE, "ref_uv_wrap.cpp"), os.path.join(_HERE, "cvstub", "cvstub.hpp"), os.path.join(_HERE, "cvstub", "cvstub_more.hpp")]
        if (not os.path.exists(_REF_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(_REF_PATH) for f in deps):
            subprocess.check_call(["make", "-C", _HERE, "-B", "_ref/libref_stereo.so", f"REF={_REFERENCE_ROOT}"], stdout=subprocess.DEVNULL)
    return _REF_PATH if os.path.exists(_REF_PATH) else None


_SGBM_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.c_int, C.c_int, C.c_void_p)
_ref_seen_params = {}


@_SGBM_CB
def _sgbm_cb(l, r, w, h, nd, bs, p1, p2, d12, cap, uniq, spw, spr, out):
    # cv::StereoSGBM is un-vendored third-party code: the reference's calDisparity_SGBM runs over the oracle's restatement
    _ref_seen_params.update(num_disparities=nd, block_size=bs, p1=p1, p2=p2, disp12_max_diff=d12, pre_filter_cap=cap,
                            uniqueness_ratio=uniq, speckle_window_size=spw, speckle_range=spr)
    left = np.ctypeslib.as_array(C.cast(l, C.POINTER(C.c_uint8)), (h, w))
    right = np.ctypeslib.as_array(C.cast(r, C.POINTER(C.c_uint8)), (h, w))
    d = sgbm(left, right, SgbmParams(**_ref_seen_params))
    C.memmove(out, d.ctypes.data, d.nbytes)


def ref():
    """ctypes handle on oracle/_ref/libref_stereo.so, or None when it is not available."""
    global _ref
    if _ref is None:
        path = build_ref()
        if path is None:
            return None
        R = C.CDLL(path)
        vp, d = C.c_void_p, C.c_double
        R.ref_set_sgbm.argtypes = [_SGBM_CB]
        R.ref_calDisparity_SGBM.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        R.ref_triangulate10D.argtypes = [vp, vp, C.c_int, C.c_int, d, d, d, d, d, d, d, vp]
        R.ref_correct3DPoints.argtypes = [vp, C.c_int, C.c_int, d, d, d, d, d]
        R.ref_setImageROI.argtypes = [vp, C.c_int, C.c_int, vp]
        R.ref_calVDisparity.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int]
        R.ref_calVDisparity.restype = C.c_int
        R.ref_calUDisparity.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int]
        R.ref_calUDisparity.restype = C.c_int
        R.ref_set_sgbm(_sgbm_cb)
        _ref = R
    return _ref


def ref_sgbm_params() -> dict:
    """The cv::StereoSGBM fields the reference's calDisparity_SGBM set on its last call (src/stereo.cpp:16-28)."""
    return dict(_ref_seen_params)


def ref_cal_disparity_sgbm(left, right) -> np.ndarray:
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    H, W = left.shape
    out = np.empty((H, W), np.int16)
    ref().ref_calDisparity_SGBM(_p(left), _p(right), W, H, _p(out))
    return out


def ref_triangulate10d(img, disp, f, cx, cy, b, roi=(30000.0, -1000.0, 30000.0)) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    disp = np.ascontiguousarray(disp, np.int16)
    H, W = disp.shape
    xyz = np.empty((H, W, 10), np.float32)
    ref().ref_triangulate10D(_p(img), _p(disp), W, H, f, cx, cy, b, roi[0], roi[1], roi[2], _p(xyz))
    return xyz


def ref_correct_3d_points(xyz, roi, pitch1, pitch2=0.0) -> np.ndarray:
    out = np.ascontiguousarray(xyz, np.float32).copy()
    H, W = out.shape[:2]
    ref().ref_correct3DPoints(_p(out), W, H, roi[0], roi[1], roi[2], pitch1, pitch2)
    return out


def ref_set_image_roi(xyz) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, np.float32)
    H, W = xyz.shape[:2]
    m = np.empty((H, W), np.uint8)
    ref().ref_setImageROI(_p(xyz), W, H, _p(m))
    return m


# ---- the mapper half from the reference's own compiled sources (oracle/_ref/libref_mapper.so, oracle/ref_mapper_wrap.cpp) ----
_REF_MAPPER = {}


def build_ref_mapper(native: bool = False) -> str | None:
    """Compile src/mapper.cpp, src/rgbdframe.cpp, src/parameter_reader.cpp and src/stereo.cpp where they lie (needs
    /root/reference; elsewhere the prebuilt file that travelled with the snapshot is used).  native=True: the reference's own
    flags (-march=native -O3, so GCC's default -ffp-contract=fast) instead of the canonical -ffp-contract=off."""
    name = "libref_mapper_native.so" if native else "libref_mapper.so"
    path = os.path.join(_HERE, "_ref", name)
    src = os.path.join(_REFERENCE_ROOT, "src", "mapper.cpp")
    if os.path.exists(src):
        deps = [src, os.path.join(_REFERENCE_ROOT, "src", "rgbdframe.cpp"), os.path.join(_REFERENCE_ROOT, "include", "rgbdframe.h"),
                os.path.join(_HERE, "ref_mapper_wrap.cpp"), os.path.join(_HERE, "ref_stereo_wrap.cpp"), os.path.join(_HERE, "Makefile"),
                os.path.join(_HERE, "cvstub", "cvstub.hpp"), os.path.join(_HERE, "cvstub", "cvstub_more.hpp"),
                os.path.join(_HERE, "refstub", "prelude.hpp"), os.path.join(_HERE, "refstub", "pcl", "common", "transforms.h")]
        if (not os.path.exists(path)) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in deps):
            subprocess.check_call(["make", "-C", _HERE, "-B", f"_ref/{name}", f"REF={_REFERENCE_ROOT}"], stdout=subprocess.DEVNULL)
    return path if os.path.exists(path) else None


def ref_mapper(native: bool = False):
    """ctypes handle on oracle/_ref/libref_mapper[_native].so, or None when it is not available."""
    if native not in _REF_MAPPER:
        path = build_ref_mapper(native)
        if path is None:
            return None
        R = C.CDLL(path)
        vp = C.c_void_p
        R.ref_set_sgbm.argtypes = [_SGBM_CB]
        R.ref_frame_next.argtypes = [C.c_char_p, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
        R.ref_frame_next.restype = C.c_int
        R.ref_mapper_cloud.argtypes = [C.c_char_p, vp, vp, vp, C.c_int, C.c_int, vp, C.c_double, vp, vp, vp, vp, vp, C.c_int]
        R.ref_mapper_cloud.restype = C.c_int
        R.ref_set_sgbm(_sgbm_cb)
        _REF_MAPPER[native] = R
    return _REF_MAPPER[native]


def build_dropin() -> str | None:
    """oracle/_ref/libdropin.so: the harness of ref_mapper_wrap.cpp / ref_stereo_wrap.cpp over the product's DROP-IN sources
    (semantic_slam_mapping_b200/host/dropin/stereo.cpp, mapper.cpp) compiled against the reference's own headers in place of its
    src/stereo.cpp and src/mapper.cpp.  Building it is the type check of the drop-in; running it needs a GPU (it calls libssm.so)."""
    path = os.path.join(_HERE, "_ref", "libdropin.so")
    if os.path.exists(os.path.join(_REFERENCE_ROOT, "include", "mapper.h")):
        drop = os.path.join(_HERE, "..", "semantic_slam_mapping_b200", "host", "dropin")
        deps = [os.path.join(drop, f) for f in ("stereo.cpp", "mapper.cpp", "dropin.hpp")] + [
            os.path.join(_HERE, "ref_mapper_wrap.cpp"), os.path.join(_HERE, "ref_stereo_wrap.cpp"), os.path.join(_HERE, "Makefile"),
            os.path.join(_HERE, "..", "include", "ssm.h"), os.path.join(_HERE, "..", "semantic_slam_mapping_b200", "libssm.so")]
        if (not os.path.exists(path)) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in deps if os.path.exists(f)):
            subprocess.check_call(["make", "-C", _HERE, "-B", "_ref/libdropin.so", f"REF={_REFERENCE_ROOT}"], stdout=subprocess.DEVNULL)
    return path if os.path.exists(path) else None


def dropin():
    """ctypes handle on oracle/_ref/libdropin.so (same entry points as ref_mapper() and ref()), or None."""
    if "dropin" not in _REF_MAPPER:
        path = build_dropin()
        if path is None:
            return None
        R = C.CDLL(path)
        vp, d = C.c_void_p, C.c_double
        R.ref_frame_next.argtypes = [C.c_char_p, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
        R.ref_frame_next.restype = C.c_int
        R.ref_mapper_cloud.argtypes = [C.c_char_p, vp, vp, vp, C.c_int, C.c_int, vp, C.c_double, vp, vp, vp, vp, vp, C.c_int]
        R.ref_mapper_cloud.restype = C.c_int
        R.ref_calDisparity_SGBM.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        R.ref_triangulate10D.argtypes = [vp, vp, C.c_int, C.c_int, d, d, d, d, d, d, d, vp]
        R.ref_correct3DPoints.argtypes = [vp, C.c_int, C.c_int, d, d, d, d, d]
        R.ref_setImageROI.argtypes = [vp, C.c_int, C.c_int, vp]
        _REF_MAPPER["dropin"] = R
    return _REF_MAPPER["dropin"]


def _cam9(mp: MapParams) -> np.ndarray:
    return np.array([mp.cx, mp.cy, mp.fx, mp.fy, mp.baseline, mp.scale, mp.roix, mp.roiy, mp.roiz], np.float64)


def ref_frame_next(left, right, rgb_bgr, semantic_bgr, mp: MapParams, native: bool = False, handle=None):
    """The reference's FrameReader::next (src/rgbdframe.cpp:34-191) on one stereo frame: (depth u16, disparity i16).  The
    images are served to its cv::imread calls from memory; cv::StereoSGBM runs the C oracle's SGBM with the parameter block
    the reference's calDisparity_SGBM sets (80 disparities)."""
    import tempfile
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    rgb = np.ascontiguousarray(rgb_bgr, np.uint8)
    sem = np.ascontiguousarray(semantic_bgr, np.uint8)
    H, W = left.shape
    depth = np.empty((H, W), np.uint16)
    disp = np.empty((H, W), np.int16)
    cam = _cam9(mp)
    with tempfile.TemporaryDirectory() as tmp:
        rc = (handle or ref_mapper(native)).ref_frame_next(tmp.encode(), _p(left), _p(right), _p(rgb), _p(sem), W, H, _p(cam), _p(depth), _p(disp))
    if rc != 0:
        raise RuntimeError("the reference's FrameReader::next returned no frame")
    return depth, disp


def ref_mapper_cloud(depth, semantic_bgr, rgb_bgr, mp: MapParams, T, native: bool = False, handle=None):
    """The reference's Mapper::semantic_motion_fuse + Mapper::generatePointCloud (src/mapper.cpp:12-94, 189-216) with
    RGBDFrame::project2dTo3d (include/rgbdframe.h:63-75) on one frame.  Returns dict(mask, xyz_cam, xyz, rgba): xyz_cam / rgba
    are the frame's cached camera-space cloud (the reference's own loop), xyz the cloud after the stand-in's
    pcl::transformPointCloud (written definition)."""
    import tempfile
    depth = np.ascontiguousarray(depth, np.uint16)
    sem = np.ascontiguousarray(semantic_bgr, np.uint8)
    rgb = np.ascontiguousarray(rgb_bgr, np.uint8)
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    H, W = depth.shape
    n = H * W
    mask = np.empty((H, W), np.uint8)
    cam_xyz = np.empty((n, 3), np.float32)
    xyz = np.empty((n, 3), np.float32)
    rgba = np.empty(n, np.uint32)
    cam = _cam9(mp)
    with tempfile.TemporaryDirectory() as tmp:
        k = (handle or ref_mapper(native)).ref_mapper_cloud(tmp.encode(), _p(depth), _p(sem), _p(rgb), W, H, _p(cam), float(mp.max_distance), _p(T),
                                                _p(mask), _p(cam_xyz), _p(xyz), _p(rgba), n)
    if k < 0:
        raise RuntimeError("generatePointCloud returned clouds of different sizes for two poses")
    return {"mask": mask, "xyz_cam": cam_xyz[:k].copy(), "xyz": xyz[:k].copy(), "rgba": rgba[:k].copy()}


# ---- label production before the path: experiment/segnet.cpp:131-146, src/rgbdframe.cpp:118-136 --------
def resize_linear_tables(dst: int, src: int, clamp: bool):
    """Source index and the two 11-bit fixed-point weights of cv::resize(INTER_LINEAR) for every destination coordinate
    (OpenCV imgproc resize, un-vendored third-party; pinned bit-exactly against cv2 4.13 by tests/test_oracle_labels.py).
    x tables clamp at the borders (weight 0 on the outside sample); y tables do not -- the row INDEX is clamped later."""
    scale = 1.0 / (dst / src)                                  # double, as 1. / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo, hi = s < 0, s >= src - 1
        f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
        s = np.where(lo, 0, np.where(hi, src - 1, s)).astype(np.int32)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)    # saturate_cast<short>: round half to even
    w1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return s, w0, w1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv::resize(img, dsize=(dw, dh)) with the default INTER_LINEAR on a single-channel 8-bit image (the 3-channel index
    image of experiment/segnet.cpp:121-134 carries the same value in every channel)."""
    img = np.ascontiguousarray(img, np.uint8)
    sh, sw = img.shape
    xi, xa0, xa1 = resize_linear_tables(dw, sw, True)
    yi, yb0, yb1 = resize_linear_tables(dh, sh, False)
    I = img.astype(np.int32)
    rows = I[:, xi] * xa0 + I[:, np.minimum(xi + 1, sw - 1)] * xa1              # horizontal pass, 11 fractional bits
    r0, r1 = rows[np.clip(yi, 0, sh - 1)], rows[np.clip(yi + 1, 0, sh - 1)]
    out = (((yb0[:, None] * (r0 >> 4)) >> 16) + ((yb1[:, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def labels_from_indices(index_img: np.ndarray, dw: int, dh: int, lut_bgr: np.ndarray):
    """experiment/segnet.cpp:131-135 (= the commented online path src/rgbdframe.cpp:130-135): argmax index image ->
    cv::resize to the frame size -> cv::LUT through the 256-entry BGR colour table.  Returns (semantic BGR [dh][dw][3],
    raw index image [dh][dw] = cvtColor(BGR2GRAY) of the resized index image, whose channels are equal)."""
    raw = resize_linear_u8(index_img, dw, dh)
    lut = np.ascontiguousarray(lut_bgr, np.uint8).reshape(256, 3)
    return lut[raw], raw


def ref_v_disparity(disp, xyz, cap_cols=1024):
    """The reference's own UVDisparity::calVDisparity (src/uvdisparity.cpp:277-366); same returns as v_disparity()."""
    disp = np.ascontiguousarray(disp, np.int16)
    out = np.ascontiguousarray(xyz, np.float32).copy()
    H, W = disp.shape
    vi = np.zeros(H * cap_cols, np.int32)
    v8 = np.zeros(H * cap_cols, np.uint8)
    vc = ref().ref_calVDisparity(_p(disp), W, H, _p(out), _p(vi), _p(v8), cap_cols)
    if vc < 0:
        raise ValueError("v-disparity wider than cap_cols")
    return out, vi[: H * vc].reshape(H, vc).copy(), v8[: H * vc].reshape(H, vc).copy()


def ref_u_disparity(disp, xyz, roi_mask, ground_mask, cap_rows=1025):
    """The reference's own UVDisparity::calUDisparity (src/uvdisparity.cpp:195-274); same returns as u_disparity()."""
    disp = np.ascontiguousarray(disp, np.int16)
    out = np.ascontiguousarray(xyz, np.float32).copy()
    roi_mask = np.ascontiguousarray(roi_mask, np.uint8)
    ground_mask = np.ascontiguousarray(ground_mask, np.uint8)
    H, W = disp.shape
    ui = np.zeros(cap_rows * W, np.int32)
    u8 = np.zeros(cap_rows * W, np.uint8)
    ur = ref().ref_calUDisparity(_p(disp), W, H, _p(out), _p(roi_mask), _p(ground_mask), _p(ui), _p(u8), cap_rows)
    if ur < 0:
        raise ValueError("u-disparity taller than cap_rows")
    return out, ui[: ur * W].reshape(ur, W).copy(), u8[: ur * W].reshape(ur, W).copy()


# ---- PNG ingest: restatement of what cv::imread does with an 8-bit non-interlaced PNG --------------
def png_decode(png: bytes, colour: bool) -> np.ndarray:
    """cv2.imdecode(png, IMREAD_COLOR / IMREAD_GRAYSCALE) for 8-bit, non-interlaced PNG files (grey, grey + alpha, RGB,
    RGBA, palette): chunk parsing, inflate, the five scanline filters of the PNG specification, palette expansion, alpha
    stripping, BGR order; grey from colour as libpng's rgb_to_gray with OpenCV's coefficients,
    (9797 R + 19234 G + 3737 B) >> 15 -- in linear light through libpng's gamma_to_1 / gamma_from_1 tables when the file
    carries a gAMA chunk outside 1 +- 0.05 or an sRGB chunk.  A CRC mismatch in a critical chunk or a scanline filter type
    above 4 raises ValueError (cv2 returns None for such files).  Pinned against cv2 4.13 in tests/test_oracle_png.py."""
    import math
    import struct
    import zlib
    assert png[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG file"
    pos, idat, pal, ihdr, gama, srgb = 8, [], None, None, 0, False
    while pos + 12 <= len(png):
        (n,), typ = struct.unpack(">I", png[pos:pos + 4]), png[pos + 4:pos + 8]
        data = png[pos + 8:pos + 8 + n]
        if not (typ[0] & 0x20) and struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0] != (zlib.crc32(typ + data) & 0xffffffff):
            raise ValueError("PNG chunk CRC mismatch")
        if typ == b"gAMA" and not idat and n == 4 and 16 <= struct.unpack(">I", data)[0] <= 625000000:
            gama = struct.unpack(">I", data)[0]
        elif typ == b"sRGB" and not idat:
            srgb = True
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"PLTE":
            pal = np.frombuffer(data, np.uint8).reshape(-1, 3)
        elif typ == b"IDAT":
            idat.append(data)
        elif typ == b"IEND":
            break
        pos += 12 + n
    W, H, depth, ctype, _, _, interlace = ihdr
    if depth != 8 or interlace != 0:
        raise ValueError("only 8-bit non-interlaced PNG files")
    bpp = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(H, W * bpp + 1)
    if (raw[:, 0] > 4).any():
        raise ValueError("bad PNG scanline filter type")
    img = np.zeros((H, W * bpp), np.int64)
    prev = np.zeros(W * bpp, np.int64)
    for y in range(H):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int64)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:
            cur = line.copy()
            for c in range(bpp):
                cur[c::bpp] = np.cumsum(line[c::bpp]) & 255
        else:
            cur = np.zeros(W * bpp, np.int64)
            for i in range(W * bpp):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
        img[y] = cur
        prev = cur
    px = img.reshape(H, W, bpp)
    if ctype in (2, 6):
        r, g, b = px[..., 0], px[..., 1], px[..., 2]
    elif ctype == 3:
        table = np.zeros((256, 3), np.int64)
        table[: len(pal)] = pal
        rgb = table[px[..., 0]]
        r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    else:
        r = g = b = px[..., 0]
    if colour:
        return np.stack([b, g, r], axis=-1).astype(np.uint8)
    plain = (9797 * r + 19234 * g + 3737 * b) >> 15
    file_gamma = 45455 if srgb else gama
    if ctype in (2, 3, 6) and file_gamma and not 95000 <= file_gamma <= 105000:
        # libpng 1.6 png_build_gamma_table + png_do_rgb_to_gray (screen gamma defaults to the reciprocal of the file gamma)
        def recip(a):
            return int(math.floor(1e10 / a + .5))

        def table(gm):
            t = np.arange(256, dtype=np.int64)
            if gm < 95000 or gm > 105000:
                for i in range(1, 255):
                    t[i] = int(math.floor(255 * math.pow(i / 255., gm * .00001) + .5))
            return t
        to1, from1 = table(recip(file_gamma)), table(recip(recip(file_gamma)))
        lin = from1[(9797 * to1[r] + 19234 * to1[g] + 3737 * to1[b] + 16384) >> 15]
        plain = np.where((r == g) & (r == b), plain, lin)
    return plain.astype(np.uint8)


def png_encode(img: np.ndarray, colour_type: int, filters=None, palette=None, level: int = 6, idat_split: int = 0,
               extra_chunks=()) -> bytes:
    """Test helper: an 8-bit non-interlaced PNG of `img` ([H][W] samples for types 0 / 3, [H][W][2|3|4] for 4 / 2 / 6) with a
    chosen filter type per row (list or None = cycle through all five), so that every un-filter path is exercised."""
    import struct
    import zlib
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape[:2]
    bpp = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[colour_type]
    rows = img.reshape(H, W * bpp).astype(np.int64)
    out = bytearray()
    prev = np.zeros(W * bpp, np.int64)
    for y in range(H):
        ft = filters[y % len(filters)] if filters is not None else y % 5
        cur = rows[y]
        left = np.concatenate([np.zeros(bpp, np.int64), cur[:-bpp]])
        upleft = np.concatenate([np.zeros(bpp, np.int64), prev[:-bpp]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - prev
        elif ft == 3:
            f = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            f = cur - pred
        out.append(ft)
        out += bytes((f & 255).astype(np.uint8))
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    z = zlib.compress(bytes(out), level)
    parts = [z] if idat_split <= 0 else [z[i:i + idat_split] for i in range(0, len(z), idat_split)]
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, colour_type, 0, 0, 0))
    for t, d in extra_chunks:      # ancillary chunks in front of PLTE / IDAT, e.g. (b"gAMA", struct.pack(">I", 45455))
        png += chunk(t, d)
    if colour_type == 3:
        png += chunk(b"PLTE", bytes(np.ascontiguousarray(palette, np.uint8).reshape(-1)))
    for p in parts:
        png += chunk(b"IDAT", p)
    return png + chunk(b"IEND", b"")
