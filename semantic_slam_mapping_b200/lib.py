"""ctypes binding of libssm.so (include/ssm.h).  No CPU fallback: loading or creating a context fails
loudly when the CUDA extension or a B200 is missing."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .params import CParams, Params

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssm.so")

STAGES = ("cost", "vertical", "horizontal", "select", "post", "points", "fuse")

# every symbol include/ssm.h declares (tests/test_boundary.py checks the .so exports all of them)
SYMBOLS = [
    "ssm_default_params", "ssm_create", "ssm_destroy", "ssm_last_error", "ssm_version", "ssm_kernel_launches",
    "ssm_set_stage_timing", "ssm_stage_time_ms", "ssm_sgbm", "ssm_sgbm_batch_device", "ssm_debug_copy_volume",
    "ssm_disparity_to_depth", "ssm_semantic_motion_fuse", "ssm_generate_point_cloud", "ssm_map_integrate_frame",
    "ssm_map_integrate_points", "ssm_map_clear", "ssm_map_size", "ssm_map_export", "ssm_map_save_pcd",
    "ssm_pipeline_batch_device", "ssm_pipeline_batch_host", "ssm_pipeline_batch_host_async", "ssm_synchronize", "ssm_comm_get_unique_id",
    "ssm_comm_init", "ssm_comm_ipc_export", "ssm_comm_ipc_connect", "ssm_comm_destroy", "ssm_voxel_owner",
    "ssm_triangulate10d", "ssm_correct_3d_points", "ssm_set_image_roi", "ssm_v_disparity", "ssm_u_disparity",
    "ssm_set_route_overlap", "ssm_keyframe_add", "ssm_keyframe_set_pose", "ssm_keyframe_release", "ssm_keyframe_count", "ssm_map_redraw",
    "ssm_map_integrate_keyframes",
    "ssm_labels_from_indices", "ssm_labels_from_indices_batch_device",
    "ssm_motion_cues_stage1_device", "ssm_motion_cues_stage2_device", "ssm_motion_cues_overflow",
    "ssm_png_info", "ssm_png_decode_batch_device", "ssm_png_decode", "ssm_png_batch_wait", "ssm_zlib_inflate_batch",
    "ssm_map_export_device_ms", "ssm_map_export_gathered", "ssm_map_reserve", "ssm_map_stats",
]


class SsmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libssm error {code}: {msg}")
        self.code = code


class _VoxelExport(C.Structure):
    _fields_ = [("ijk", C.c_void_p), ("xyz", C.c_void_p), ("rgba", C.c_void_p), ("label", C.c_void_p),
                ("count", C.c_void_p), ("votes", C.c_void_p)]


class _MapStats(C.Structure):
    _fields_ = [("slots", C.c_uint64), ("voxels", C.c_uint64), ("load_factor", C.c_double), ("mean_probe", C.c_double),
                ("max_probe", C.c_uint64), ("grow_steps", C.c_uint64), ("table_bytes", C.c_uint64)]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree libssm.so.  Raises if it has not been built (`python -m semantic_slam_mapping_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, sz, u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64
    L.ssm_default_params.argtypes = [C.POINTER(CParams)]
    L.ssm_default_params.restype = None
    L.ssm_create.argtypes = [C.POINTER(CParams), i, C.POINTER(vp)]
    L.ssm_destroy.argtypes = [vp]
    L.ssm_destroy.restype = None
    L.ssm_last_error.restype = C.c_char_p
    L.ssm_version.restype = C.c_char_p
    L.ssm_kernel_launches.argtypes = [vp]
    L.ssm_kernel_launches.restype = u64
    L.ssm_set_stage_timing.argtypes = [vp, i]
    L.ssm_stage_time_ms.argtypes = [vp, i, C.POINTER(C.c_float)]
    L.ssm_sgbm.argtypes = [vp, vp, vp, i, i, sz, vp, sz]
    L.ssm_png_info.argtypes = [vp, sz, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.ssm_png_decode_batch_device.argtypes = [vp, i, C.POINTER(vp), C.POINTER(sz), i, i, i, vp, i, vp]
    L.ssm_png_decode.argtypes = [vp, vp, sz, i, vp, sz, C.POINTER(i), C.POINTER(i)]
    L.ssm_png_batch_wait.argtypes = [vp]
    L.ssm_zlib_inflate_batch.argtypes = [vp, i, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz), C.POINTER(i)]
    L.ssm_sgbm_batch_device.argtypes = [vp, i, vp, vp, i, i, vp, vp]
    L.ssm_debug_copy_volume.argtypes = [vp, i, i, vp, sz]
    L.ssm_disparity_to_depth.argtypes = [vp, vp, i, i, sz, vp, sz]
    L.ssm_semantic_motion_fuse.argtypes = [vp, vp, i, i, sz, vp, sz]
    L.ssm_generate_point_cloud.argtypes = [vp, vp, vp, vp, i, i, vp, vp, vp, vp, i, C.POINTER(i)]
    L.ssm_map_integrate_frame.argtypes = [vp, vp, vp, vp, i, i, vp]
    L.ssm_map_integrate_points.argtypes = [vp, vp, vp, vp, i]
    L.ssm_map_clear.argtypes = [vp]
    L.ssm_map_size.argtypes = [vp, C.POINTER(u64)]
    L.ssm_map_export.argtypes = [vp, C.POINTER(_VoxelExport), u64, i, C.POINTER(u64)]
    L.ssm_map_save_pcd.argtypes = [vp, C.c_char_p]
    L.ssm_map_export_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.ssm_map_export_gathered.argtypes = [vp, C.POINTER(_VoxelExport), u64, i, C.POINTER(u64)]
    L.ssm_map_reserve.argtypes = [vp, u64]
    L.ssm_map_stats.argtypes = [vp, C.POINTER(_MapStats)]
    L.ssm_pipeline_batch_device.argtypes = [vp, i, vp, vp, vp, vp, vp, i, i, vp, vp]
    L.ssm_pipeline_batch_host.argtypes = [vp, i, vp, vp, vp, vp, vp, i, i, vp, C.POINTER(u64)]
    L.ssm_pipeline_batch_host_async.argtypes = [vp, i, vp, vp, vp, vp, vp, i, i, vp]
    L.ssm_synchronize.argtypes = [vp]
    L.ssm_comm_get_unique_id.argtypes = [vp]
    L.ssm_comm_init.argtypes = [vp, vp, i, i]
    L.ssm_comm_ipc_export.argtypes = [vp, vp]
    L.ssm_comm_ipc_connect.argtypes = [vp, vp, i]
    L.ssm_comm_destroy.argtypes = [vp]
    L.ssm_voxel_owner.argtypes = [C.c_int32, C.c_int32, C.c_int32, i]
    d = C.c_double
    L.ssm_labels_from_indices.argtypes = [vp, vp, sz, i, i, i, i, vp, vp, sz, vp, sz]
    L.ssm_labels_from_indices_batch_device.argtypes = [vp, i, vp, i, i, i, i, vp, vp, vp, vp]
    L.ssm_set_route_overlap.argtypes = [vp, i]
    L.ssm_keyframe_add.argtypes = [vp, vp, vp, vp, i, i, vp, C.POINTER(i)]
    L.ssm_keyframe_set_pose.argtypes = [vp, i, vp]
    L.ssm_keyframe_release.argtypes = [vp, i]
    L.ssm_keyframe_count.argtypes = [vp, C.POINTER(i), C.POINTER(u64)]
    L.ssm_map_redraw.argtypes = [vp, vp, i]
    L.ssm_map_integrate_keyframes.argtypes = [vp, vp, i]
    L.ssm_triangulate10d.argtypes = [vp, vp, sz, vp, sz, i, i, d, d, d, d, d, d, d, vp]
    L.ssm_correct_3d_points.argtypes = [vp, vp, i, i, d, d, d, d, d]
    L.ssm_set_image_roi.argtypes = [vp, vp, i, i, vp, sz]
    L.ssm_v_disparity.argtypes = [vp, vp, sz, i, i, vp, vp, vp, i, C.POINTER(i)]
    L.ssm_u_disparity.argtypes = [vp, vp, sz, i, i, vp, vp, vp, vp, vp, i, C.POINTER(i)]
    L.ssm_motion_cues_stage1_device.argtypes = [vp, i, vp, vp, i, i, d, d, d, d, vp, vp, vp, sz, i, vp]
    L.ssm_motion_cues_stage2_device.argtypes = [vp, i, vp, i, i, vp, d, d, d, d, vp, vp, vp, vp, sz, i, vp]
    L.ssm_motion_cues_overflow.argtypes = [vp, C.POINTER(i)]
    _lib = L
    return L


def _ptr(a):
    """numpy array -> host pointer; torch tensor / int -> raw (device) pointer."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())


def voxel_owner(i: int, j: int, k: int, nranks: int) -> int:
    return int(load().ssm_voxel_owner(i, j, k, nranks))


class Context:
    """One `ssm_ctx` (one per GPU)."""

    def __init__(self, params: Params | None = None, device: int = 0):
        self.params = params or Params()
        self._L = load()
        self._h = C.c_void_p()
        cp = self.params.c()
        self._check(self._L.ssm_create(C.byref(cp), device, C.byref(self._h)))
        self.device = device

    def _check(self, rc: int):
        if rc != 0:
            raise SsmError(rc, self._L.ssm_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.ssm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- stereo.h ---------------------------------------------------------------------------------------
    def sgbm(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        if left.ndim != 2 or left.shape != right.shape:
            raise ValueError("left/right must be 2-D uint8 images of equal shape")
        h, w = left.shape
        disp = np.empty((h, w), np.int16)
        self._check(self._L.ssm_sgbm(self._h, _ptr(left), _ptr(right), w, h, w, _ptr(disp), w * 2))
        return disp

    def sgbm_batch_device(self, d_left, d_right, d_disp, batch: int, w: int, h: int, stream=None):
        self._check(self._L.ssm_sgbm_batch_device(self._h, batch, _ptr(d_left), _ptr(d_right), w, h, _ptr(d_disp),
                                                  C.c_void_p(stream) if stream else None))

    def debug_volume(self, which: str, w: int, h: int, batch_index: int = 0) -> np.ndarray:
        D = self.params.num_disparities
        ids = {"C": 0, "S": 1, "disp_raw": 2, "disp_median": 3}
        shape = (h, w - D, D) if which in ("C", "S") else (h, w)
        out = np.empty(shape, np.int16)
        self._check(self._L.ssm_debug_copy_volume(self._h, ids[which], batch_index, _ptr(out), out.nbytes))
        return out

    # -- glue ---------------------------------------------------------------------------------------------
    def disparity_to_depth(self, disp: np.ndarray) -> np.ndarray:
        disp = np.ascontiguousarray(disp, np.int16)
        h, w = disp.shape
        depth = np.empty((h, w), np.uint16)
        self._check(self._L.ssm_disparity_to_depth(self._h, _ptr(disp), w, h, w * 2, _ptr(depth), w * 2))
        return depth

    # -- mapper.h -----------------------------------------------------------------------------------------
    def semantic_motion_fuse(self, semantic_bgr: np.ndarray) -> np.ndarray:
        sem = np.ascontiguousarray(semantic_bgr, np.uint8)
        h, w = sem.shape[:2]
        mask = np.empty((h, w), np.uint8)
        self._check(self._L.ssm_semantic_motion_fuse(self._h, _ptr(sem), w, h, w * 3, _ptr(mask), w))
        return mask

    def generate_point_cloud(self, depth, semantic_bgr, rgb_bgr, T):
        depth = np.ascontiguousarray(depth, np.uint16)
        sem = np.ascontiguousarray(semantic_bgr, np.uint8)
        rgb = np.ascontiguousarray(rgb_bgr, np.uint8)
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        h, w = depth.shape
        n = h * w
        xyz = np.empty((n, 3), np.float32)
        rgba = np.empty(n, np.uint32)
        label = np.empty(n, np.uint8)
        k = C.c_int(0)
        self._check(self._L.ssm_generate_point_cloud(self._h, _ptr(depth), _ptr(sem), _ptr(rgb), w, h, _ptr(T), _ptr(xyz),
                                                     _ptr(rgba), _ptr(label), n, C.byref(k)))
        return {"xyz": xyz[: k.value].copy(), "rgba": rgba[: k.value].copy(), "label": label[: k.value].copy()}

    def map_integrate_frame(self, depth, semantic_bgr, rgb_bgr, T):
        depth = np.ascontiguousarray(depth, np.uint16)
        sem = np.ascontiguousarray(semantic_bgr, np.uint8)
        rgb = np.ascontiguousarray(rgb_bgr, np.uint8)
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        h, w = depth.shape
        self._check(self._L.ssm_map_integrate_frame(self._h, _ptr(depth), _ptr(sem), _ptr(rgb), w, h, _ptr(T)))

    def map_integrate_points(self, xyz, rgba, label):
        xyz = np.ascontiguousarray(xyz, np.float32)
        rgba = np.ascontiguousarray(rgba, np.uint32)
        label = np.ascontiguousarray(label, np.uint8)
        self._check(self._L.ssm_map_integrate_points(self._h, _ptr(xyz), _ptr(rgba), _ptr(label), xyz.shape[0]))

    def keyframe_add(self, depth, semantic_bgr, rgb_bgr, T) -> int:
        depth = np.ascontiguousarray(depth, np.uint16)
        sem = np.ascontiguousarray(semantic_bgr, np.uint8)
        rgb = np.ascontiguousarray(rgb_bgr, np.uint8)
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        h, w = depth.shape
        kid = C.c_int(-1)
        self._check(self._L.ssm_keyframe_add(self._h, _ptr(depth), _ptr(sem), _ptr(rgb), w, h, _ptr(T), C.byref(kid)))
        return int(kid.value)

    def keyframe_set_pose(self, kid: int, T):
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        self._check(self._L.ssm_keyframe_set_pose(self._h, kid, _ptr(T)))

    def keyframe_release(self, kid: int):
        self._check(self._L.ssm_keyframe_release(self._h, kid))

    def keyframe_count(self):
        n, pts = C.c_int(0), C.c_uint64(0)
        self._check(self._L.ssm_keyframe_count(self._h, C.byref(n), C.byref(pts)))
        return int(n.value), int(pts.value)

    def map_redraw(self, ids=None):
        a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self._L.ssm_map_redraw(self._h, _ptr(a), 0 if a is None else len(a)))

    def map_integrate_keyframes(self, ids=None):
        a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self._L.ssm_map_integrate_keyframes(self._h, _ptr(a), 0 if a is None else len(a)))

    def map_clear(self):
        self._check(self._L.ssm_map_clear(self._h))

    def map_size(self) -> int:
        n = C.c_uint64(0)
        self._check(self._L.ssm_map_size(self._h, C.byref(n)))
        return int(n.value)

    _EXPORT_KEYS = ("ijk", "xyz", "rgba", "label", "count", "votes")

    def _export_arrays(self, n: int, fields):
        L = self.params.num_labels
        shapes = {"ijk": ((n, 3), np.int32), "xyz": ((n, 3), np.float32), "rgba": ((n,), np.uint32), "label": ((n,), np.uint8),
                  "count": ((n,), np.uint32), "votes": ((n, L), np.uint32)}
        return {k: np.empty(*shapes[k]) for k in self._EXPORT_KEYS if fields is None or k in fields}

    def map_export(self, sorted: bool = True, fields=None, into: dict | None = None) -> dict:
        """K9 on the device (finalize + radix sort into pcl::VoxelGrid's order) and one D2H per requested array.
        fields: subset of ("ijk", "xyz", "rgba", "label", "count", "votes"), default all.  into: caller-owned (e.g. pinned)
        arrays of at least map_size() rows; the returned dict holds views of the first n rows."""
        if into is not None:
            out = into
            cap = min(len(v) for v in out.values())
        else:
            cap = self.map_size()
            out = self._export_arrays(cap, fields)
        ex = _VoxelExport(*[_ptr(out.get(k)) for k in self._EXPORT_KEYS])
        got = C.c_uint64(0)
        self._check(self._L.ssm_map_export(self._h, C.byref(ex), cap, 1 if sorted else 0, C.byref(got)))
        n = int(got.value)
        if n > cap:
            raise ValueError(f"export arrays hold {cap} voxels, the map has {n}")
        return {k: v[:n] for k, v in out.items()}

    def map_export_gathered(self, total_voxels: int, sorted: bool = True, fields=None) -> dict | None:
        """Collective: every rank calls it; rank 0 (ssm_comm_init's rank) gets the union of the ranks' tables, the others None.
        total_voxels: capacity of rank 0's arrays (e.g. the sum of the ranks' map_size())."""
        rank0 = total_voxels is not None and total_voxels >= 0
        out = self._export_arrays(total_voxels, fields) if rank0 else None
        ex = _VoxelExport(*[_ptr(out.get(k)) for k in self._EXPORT_KEYS]) if rank0 else None
        got = C.c_uint64(0)
        self._check(self._L.ssm_map_export_gathered(self._h, C.byref(ex) if rank0 else None, total_voxels if rank0 else 0,
                                                    1 if sorted else 0, C.byref(got)))
        if not rank0:
            return None
        n = int(got.value)
        if n > total_voxels:
            raise ValueError(f"export arrays hold {total_voxels} voxels, the gathered map has {n}")
        return {k: v[:n] for k, v in out.items()}

    def map_export_device_ms(self) -> float:
        ms = C.c_float(0)
        self._check(self._L.ssm_map_export_device_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def map_reserve(self, slots: int):
        self._check(self._L.ssm_map_reserve(self._h, slots))

    def map_stats(self) -> dict:
        st = _MapStats()
        self._check(self._L.ssm_map_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _MapStats._fields_}

    def map_save_pcd(self, path: str):
        self._check(self._L.ssm_map_save_pcd(self._h, path.encode()))

    # -- PNG ingest (FrameReader::next's cv::imread calls, src/rgbdframe.cpp:45-78, 138-180) ------------------------------
    def png_info(self, png: bytes):
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        buf = np.frombuffer(png, np.uint8)
        self._check(self._L.ssm_png_info(_ptr(buf), buf.size, C.byref(w), C.byref(h), C.byref(ch)))
        return w.value, h.value, ch.value

    def png_decode(self, png: bytes, colour: bool) -> np.ndarray:
        """== cv2.imdecode(png, IMREAD_COLOR if colour else IMREAD_GRAYSCALE): inflate on the host, the rest on the GPU."""
        w, h, _ = self.png_info(png)
        buf = np.frombuffer(png, np.uint8)
        out = np.empty((h, w, 3) if colour else (h, w), np.uint8)
        self._check(self._L.ssm_png_decode(self._h, _ptr(buf), buf.size, 1 if colour else 0, _ptr(out), out.nbytes, None, None))
        return out

    def png_decode_batch_device(self, pngs, w: int, h: int, colour: bool, d_out, host_threads: int = 0, stream=None):
        """`pngs`: list of bytes objects; d_out: device buffer [len(pngs)][h][w]([3]) u8."""
        bufs = [np.frombuffer(p, np.uint8) for p in pngs]
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * n)(*[b.size for b in bufs])
        self._check(self._L.ssm_png_decode_batch_device(self._h, n, ptrs, sizes, w, h, 1 if colour else 0, _ptr(d_out), host_threads,
                                                        C.c_void_p(stream) if stream else None))

    def png_batch_wait(self):
        """Blocks until the queued ingest batches have run; raises on a corrupt stream of a GPU-inflated batch."""
        self._check(self._L.ssm_png_batch_wait(self._h))

    def zlib_inflate_batch(self, streams, out_sizes):
        """The GPU DEFLATE decoder on its own: list of zlib streams (bytes) -> (list of uint8 arrays, status array)."""
        bufs = [np.frombuffer(s, np.uint8) if len(s) else np.zeros(0, np.uint8) for s in streams]
        n = len(bufs)
        outs = [np.zeros(int(m), np.uint8) for m in out_sizes]
        status = np.zeros(n, np.int32)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * n)(*[b.size for b in bufs])
        optrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        osizes = (C.c_size_t * n)(*[o.size for o in outs])
        self._check(self._L.ssm_zlib_inflate_batch(self._h, n, ptrs, sizes, optrs, osizes, status.ctypes.data_as(C.POINTER(C.c_int))))
        return outs, status

    # -- label production (experiment/segnet.cpp:121-135) ---------------------------------------------------------------
    def labels_from_indices(self, index_img, dw: int, dh: int, lut_bgr, want_raw: bool = True):
        idx = np.ascontiguousarray(index_img, np.uint8)
        lut = np.ascontiguousarray(lut_bgr, np.uint8).reshape(256, 3)
        sh, sw = idx.shape
        sem = np.empty((dh, dw, 3), np.uint8)
        raw = np.empty((dh, dw), np.uint8) if want_raw else None
        self._check(self._L.ssm_labels_from_indices(self._h, _ptr(idx), sw, sw, sh, dw, dh, _ptr(lut), _ptr(sem), dw * 3, _ptr(raw), dw))
        return sem, raw

    def labels_from_indices_batch_device(self, d_index, d_sem, batch, sw, sh, dw, dh, lut_bgr, d_raw=None, stream=None):
        lut = np.ascontiguousarray(lut_bgr, np.uint8).reshape(256, 3)
        self._check(self._L.ssm_labels_from_indices_batch_device(self._h, batch, _ptr(d_index), sw, sh, dw, dh, _ptr(lut), _ptr(d_sem), _ptr(d_raw),
                                                                 C.c_void_p(stream) if stream else None))

    # -- dense motion cues (stereo.h triangulate10D / correct3DPoints / setImageROI, UVDisparity::cal[UV]Disparity) --------
    def triangulate10d(self, img, disp, f, cx, cy, b, roi=(30000.0, -1000.0, 30000.0)) -> np.ndarray:
        img = np.ascontiguousarray(img, np.uint8)
        disp = np.ascontiguousarray(disp, np.int16)
        h, w = disp.shape
        xyz = np.empty((h, w, 10), np.float32)
        self._check(self._L.ssm_triangulate10d(self._h, _ptr(img), w, _ptr(disp), w * 2, w, h, f, cx, cy, b, roi[0], roi[1], roi[2], _ptr(xyz)))
        return xyz

    def correct_3d_points(self, xyz, roi, pitch1, pitch2=0.0) -> np.ndarray:
        out = np.ascontiguousarray(xyz, np.float32).copy()
        h, w = out.shape[:2]
        self._check(self._L.ssm_correct_3d_points(self._h, _ptr(out), w, h, roi[0], roi[1], roi[2], pitch1, pitch2))
        return out

    def set_image_roi(self, xyz) -> np.ndarray:
        xyz = np.ascontiguousarray(xyz, np.float32)
        h, w = xyz.shape[:2]
        m = np.empty((h, w), np.uint8)
        self._check(self._L.ssm_set_image_roi(self._h, _ptr(xyz), w, h, _ptr(m), w))
        return m

    def v_disparity(self, disp, xyz=None):
        """Returns (xyz with channel 8 filled or None, v_dis_int [H][v_cols], v_dis [H][v_cols])."""
        disp = np.ascontiguousarray(disp, np.int16)
        h, w = disp.shape
        out = None if xyz is None else np.ascontiguousarray(xyz, np.float32).copy()
        vc = C.c_int(0)
        self._check(self._L.ssm_v_disparity(self._h, _ptr(disp), w * 2, w, h, None, None, None, 0, C.byref(vc)))   # size query
        n = vc.value
        vi = np.zeros((h, n), np.int32)
        v8 = np.zeros((h, n), np.uint8)
        self._check(self._L.ssm_v_disparity(self._h, _ptr(disp), w * 2, w, h, _ptr(out), _ptr(vi), _ptr(v8), n, C.byref(vc)))
        return out, vi, v8

    def u_disparity(self, disp, xyz, roi_mask, ground_mask):
        """Returns (xyz with channel 7 filled or None, u_dis_int [u_rows][W], u_dis [u_rows][W])."""
        disp = np.ascontiguousarray(disp, np.int16)
        roi_mask = np.ascontiguousarray(roi_mask, np.uint8)
        ground_mask = np.ascontiguousarray(ground_mask, np.uint8)
        h, w = disp.shape
        out = None if xyz is None else np.ascontiguousarray(xyz, np.float32).copy()
        ur = C.c_int(0)
        self._check(self._L.ssm_u_disparity(self._h, _ptr(disp), w * 2, w, h, None, _ptr(roi_mask), _ptr(ground_mask), None, None, 0, C.byref(ur)))
        n = ur.value
        ui = np.zeros((n, w), np.int32)
        u8 = np.zeros((n, w), np.uint8)
        self._check(self._L.ssm_u_disparity(self._h, _ptr(disp), w * 2, w, h, _ptr(out), _ptr(roi_mask), _ptr(ground_mask), _ptr(ui), _ptr(u8), n, C.byref(ur)))
        return out, ui, u8

    def motion_cues_stage1_device(self, d_img, d_disp, d_xyz, batch, w, h, f, cx, cy, b, d_v_int=None, d_v8=None, hist_stride=0, cap_cols=0, stream=None):
        self._check(self._L.ssm_motion_cues_stage1_device(self._h, batch, _ptr(d_img), _ptr(d_disp), w, h, f, cx, cy, b, _ptr(d_xyz), _ptr(d_v_int),
                                                          _ptr(d_v8), hist_stride, cap_cols, C.c_void_p(stream) if stream else None))

    def motion_cues_stage2_device(self, d_disp, d_xyz, d_roi_mask, batch, w, h, roi, pitch1, d_ground=None, d_u_int=None, d_u8=None, hist_stride=0,
                                  cap_rows=0, stream=None):
        self._check(self._L.ssm_motion_cues_stage2_device(self._h, batch, _ptr(d_disp), w, h, _ptr(d_xyz), roi[0], roi[1], roi[2], pitch1, _ptr(d_ground),
                                                          _ptr(d_roi_mask), _ptr(d_u_int), _ptr(d_u8), hist_stride, cap_rows,
                                                          C.c_void_p(stream) if stream else None))

    def motion_cues_overflow(self) -> int:
        f = C.c_int(0)
        self._check(self._L.ssm_motion_cues_overflow(self._h, C.byref(f)))
        return int(f.value)

    # -- whole path ---------------------------------------------------------------------------------------
    def pipeline_batch_host(self, left, right, semantic, rgb, poses, want_disp: bool = False):
        """left/right [B][H][W] u8, semantic/rgb [B][H][W][3] u8, poses [B][4][4] f64 (host or pinned)."""
        b, h, w = left.shape
        disp = np.empty((b, h, w), np.int16) if want_disp else None
        nvox = C.c_uint64(0)
        self._check(self._L.ssm_pipeline_batch_host(self._h, b, _ptr(left), _ptr(right), _ptr(semantic), _ptr(rgb), _ptr(poses),
                                                    w, h, _ptr(disp), C.byref(nvox)))
        return int(nvox.value), disp

    def pipeline_batch_host_async(self, left, right, semantic, rgb, poses, n_voxels_pinned=None):
        """Enqueue one batch (pinned host arrays; keep them alive until synchronize()).  n_voxels_pinned: a pinned
        uint32 torch tensor / numpy array of one element that receives the map size after this batch."""
        b, h, w = left.shape
        self._check(self._L.ssm_pipeline_batch_host_async(self._h, b, _ptr(left), _ptr(right), _ptr(semantic), _ptr(rgb), _ptr(poses),
                                                          w, h, _ptr(n_voxels_pinned)))

    def pipeline_batch_device(self, d_left, d_right, d_sem, d_rgb, d_poses, batch: int, w: int, h: int, d_disp=None, stream=None):
        self._check(self._L.ssm_pipeline_batch_device(self._h, batch, _ptr(d_left), _ptr(d_right), _ptr(d_sem), _ptr(d_rgb),
                                                      _ptr(d_poses), w, h, _ptr(d_disp), C.c_void_p(stream) if stream else None))

    def set_route_overlap(self, on: bool):
        self._check(self._L.ssm_set_route_overlap(self._h, 1 if on else 0))

    def synchronize(self):
        self._check(self._L.ssm_synchronize(self._h))

    # -- instrumentation ------------------------------------------------------------------------------------
    def kernel_launches(self) -> int:
        return int(self._L.ssm_kernel_launches(self._h))

    def set_stage_timing(self, on: bool):
        self._check(self._L.ssm_set_stage_timing(self._h, 1 if on else 0))

    def stage_times_ms(self) -> dict:
        out = {}
        for i, name in enumerate(STAGES):
            ms = C.c_float(0)
            self._check(self._L.ssm_stage_time_ms(self._h, i, C.byref(ms)))
            out[name] = float(ms.value)
        return out

    # -- multi-GPU --------------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> np.ndarray:
        L = load()
        uid = np.zeros(128, np.uint8)
        rc = L.ssm_comm_get_unique_id(_ptr(uid))
        if rc != 0:
            raise SsmError(rc, L.ssm_last_error().decode())
        return uid

    def comm_init(self, unique_id: np.ndarray, rank: int, nranks: int):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        self._check(self._L.ssm_comm_init(self._h, _ptr(uid), rank, nranks))

    def comm_ipc_export(self) -> np.ndarray:
        h = np.zeros(64, np.uint8)
        self._check(self._L.ssm_comm_ipc_export(self._h, _ptr(h)))
        return h

    def comm_ipc_connect(self, handles: np.ndarray):
        handles = np.ascontiguousarray(handles, np.uint8)
        self._check(self._L.ssm_comm_ipc_connect(self._h, _ptr(handles), handles.shape[0]))
