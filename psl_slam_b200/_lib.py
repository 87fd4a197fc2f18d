"""ctypes binding of libpsl_frontend.so (the CUDA product library, include/psl_frontend.h).

There is no fallback: if the library is missing or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PSL_FRONTEND_SO") or os.path.join(_HERE, "libpsl_frontend.so")  # override: development A/B builds

PSL_OK, PSL_E_INVALID, PSL_E_CUDA, PSL_E_CAPACITY, PSL_E_INTERNAL = 0, -1, -2, -3, -4

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


class PslError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"psl error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("max_batch", C.c_int32), ("orb_nfeatures", C.c_int32), ("orb_scale_factor", C.c_float),
                ("orb_nlevels", C.c_int32), ("orb_ini_th_fast", C.c_int32), ("orb_min_th_fast", C.c_int32),
                ("orb_max_candidates", C.c_int32), ("chunk_frames", C.c_int32), ("line_nfeatures", C.c_int32),
                ("line_scale_factor", C.c_float), ("line_nlevels", C.c_int32), ("line_min_length", C.c_float),
                ("line_chunk_frames", C.c_int32), ("line_max_raw", C.c_int32)]


KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                          ("response", "<f4"), ("size", "<f4"), ("start_x", "<f4"), ("start_y", "<f4"),
                          ("end_x", "<f4"), ("end_y", "<f4"), ("s_oct_x", "<f4"), ("s_oct_y", "<f4"),
                          ("e_oct_x", "<f4"), ("e_oct_y", "<f4"), ("line_length", "<f4"),
                          ("num_pixels", "<i4")])  # psl_keyline, 68 B
QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"),
                        ("u_right", "<f4"), ("angle", "<f4"), ("flags", "<u4")])  # psl_proj_query
LINE_QUERY_DTYPE = np.dtype([("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4"), ("radius", "<f4"),
                             ("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"), ("length", "<f4"),
                             ("normal", "<f8", (3,)), ("flags", "<u4"), ("pad_", "<u4")])  # psl_line_query, 72 B
FUSE_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("u_right", "<f4"), ("radius", "<f4"), ("pred_level", "<i4"),
                             ("flags", "<u4")])  # psl_fuse_query
LINE_FUSE_QUERY_DTYPE = np.dtype([("u1", "<f4"), ("v1", "<f4"), ("u2", "<f4"), ("v2", "<f4"), ("radius", "<f4"),
                                  ("pred_level", "<i4"), ("flags", "<u4")])  # psl_line_fuse_query, 28 B
JUNCTION_DTYPE = np.dtype([("l1", "<i4"), ("l2", "<i4"), ("cross2d_x", "<f4"), ("cross2d_y", "<f4"),
                           ("cross3d", "<f8", (3,))])  # psl_line_junction, 40 B
POSE_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("u_right", "<f4"), ("inv_sigma2", "<f4"), ("xw", "<f4"),
                             ("yw", "<f4"), ("zw", "<f4"), ("flags", "<u4")])  # psl_pose_point, 32 B
POSE_LIL_DTYPE = np.dtype([("line1", "<f8", (6,)), ("line2", "<f8", (6,)), ("cross", "<f8", (3,)), ("obs1", "<f8", (3,)),
                           ("obs2", "<f8", (3,)), ("ins", "<f8", (2,)), ("flags", "<u4"), ("pad_", "<u4")])  # psl_pose_lil, 192 B
Q_VALID, Q_CLAIMS = 1, 2


class Distortion(C.Structure):  # psl_distortion
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("k1", C.c_float),
                ("k2", C.c_float), ("p1", C.c_float), ("p2", C.c_float), ("k3", C.c_float)]


class FrameView(C.Structure):  # psl_frame_view
    _fields_ = [("n", C.c_int32), ("kps_un", C.c_void_p), ("u_right", C.c_void_p), ("desc", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float)]


class LineFrameView(C.Structure):  # psl_line_frame_view
    _fields_ = [("n", C.c_int32), ("kl_un", C.c_void_p), ("ldesc", C.c_void_p), ("lineeq", C.c_void_p),
                ("lines3d", C.c_void_p), ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float),
                ("max_y", C.c_float), ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float)]


def make_line_frame_view(kl_un, ldesc, lineeq, lines3d, bounds):
    """Build a psl_line_frame_view over numpy arrays (kept alive by the returned tuple)."""
    kl_un = np.ascontiguousarray(kl_un, KEYLINE_DTYPE)
    ldesc = np.ascontiguousarray(ldesc, np.uint8)
    lineeq = np.ascontiguousarray(lineeq, np.float64)
    l3 = None if lines3d is None else np.ascontiguousarray(lines3d, np.float64)
    min_x, min_y, max_x, max_y = (np.float32(b) for b in bounds)
    fv = LineFrameView(len(kl_un), kl_un.ctypes.data, ldesc.ctypes.data, lineeq.ctypes.data,
                       None if l3 is None else l3.ctypes.data, min_x, min_y, max_x, max_y,
                       np.float32(64) / np.float32(max_x - min_x), np.float32(48) / np.float32(max_y - min_y))
    return fv, (kl_un, ldesc, lineeq, l3)


class MatchParams(C.Structure):  # psl_match_params
    _fields_ = [("mode", C.c_int32), ("th_dist", C.c_int32), ("nn_ratio", C.c_float),
                ("check_orientation", C.c_int32)]


class Camera(C.Structure):  # psl_camera
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("depth_factor", C.c_float)]


class TrackParams(C.Structure):  # psl_track_params
    _fields_ = [("th", C.c_float), ("nn_ratio", C.c_float), ("check_orientation", C.c_int32),
                ("th_dist", C.c_int32)]


class FrontendOut(C.Structure):  # psl_frontend_out
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("n", C.c_void_p), ("u_right", C.c_void_p),
                ("z", C.c_void_p), ("assign", C.c_void_p), ("nmatches", C.c_void_p), ("cap", C.c_int32),
                ("line_cap", C.c_int32), ("kl", C.c_void_p), ("ldesc", C.c_void_p), ("lineeq", C.c_void_p),
                ("nl", C.c_void_p), ("line_assign", C.c_void_p), ("line_nmatches", C.c_void_p)]


class KeyFrameView(C.Structure):  # psl_keyframe_view
    _fields_ = [("n", C.c_int32), ("kps_un", C.c_void_p), ("u_right", C.c_void_p), ("desc", C.c_void_p),
                ("has_mappoint", C.c_void_p)]


def make_keyframe_view(kps_un, u_right, desc, has_mappoint):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
    ur = np.ascontiguousarray(u_right, np.float32)
    desc = np.ascontiguousarray(desc, np.uint8)
    mp = np.ascontiguousarray(has_mappoint, np.uint8)
    return KeyFrameView(len(kps_un), kps_un.ctypes.data, ur.ctypes.data, desc.ctypes.data, mp.ctypes.data), (kps_un, ur, desc, mp)


class FeatureVector(C.Structure):  # psl_feature_vector
    _fields_ = [("n_nodes", C.c_int32), ("node_id", C.c_void_p), ("offs", C.c_void_p), ("idx", C.c_void_p)]


def make_frame_view(kps_un: np.ndarray, u_right, desc: np.ndarray, bounds):
    """Build a psl_frame_view over numpy arrays (kept alive by the returned tuple)."""
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    min_x, min_y, max_x, max_y = (np.float32(b) for b in bounds)
    fv = FrameView(len(kps_un), kps_un.ctypes.data, None if ur is None else ur.ctypes.data, desc.ctypes.data,
                   min_x, min_y, max_x, max_y, np.float32(64) / np.float32(max_x - min_x),
                   np.float32(48) / np.float32(max_y - min_y))  # Frame.cc:163-164
    return fv, (kps_un, ur, desc)


def make_feature_vector(node_id, offs, idx):
    node_id = np.ascontiguousarray(node_id, np.uint32)
    offs = np.ascontiguousarray(offs, np.int32)
    idx = np.ascontiguousarray(idx, np.uint32)
    return FeatureVector(len(node_id), node_id.ctypes.data, offs.ctypes.data, idx.ctypes.data), (node_id, offs, idx)


# every symbol include/psl_frontend.h declares (checked by tests/test_abi.py)
EXPORTS = ["psl_default_config", "psl_create", "psl_destroy", "psl_last_error", "psl_stream", "psl_sync",
           "psl_orb_tables", "psl_orb_extract", "psl_orb_extract_batch", "psl_orb_extract_batch_dev", "psl_debug_fetch", "psl_profile_enable",
           "psl_profile_read", "psl_launch_count", "psl_descriptor_distance", "psl_hamming_knn2",
           "psl_match_projection", "psl_match_bow", "psl_track_orb_batch", "psl_track_orb_batch_dev",
           "psl_line_extract", "psl_line_extract_batch", "psl_line_extract_batch_dev", "psl_line_match_nnr",
           "psl_line_search_geom", "psl_line_frame_bf_match", "psl_line_search_double", "psl_line_match_projection",
           "psl_plane_assoc", "psl_track_frontend_batch", "psl_track_frontend_batch_dev", "psl_convert_rgbd",
           "psl_convert_rgbd_dev", "psl_match_triangulation", "psl_match_fuse", "psl_line_search_triangulation", "psl_line_fuse",
           "psl_undistort_keypoints", "psl_undistort_keypoints_dev", "psl_image_bounds", "psl_plane_hypotheses",
           "psl_lines_3d", "psl_lines_3d_dev", "psl_match_bow_kf", "psl_match_sim3", "psl_match_initialization",
           "psl_line_junctions", "psl_line_junctions_dev", "psl_line_search_triangulation_new",
           "psl_pose_optimization", "psl_pose_optimization_dev", "psl_pose_optimization_lil",
           "psl_pose_optimization_lil_dev", "psl_track_rgbd_batch_dev", "psl_track_rgbd_batch",
           "psl_track_rgbd_batch_begin", "psl_track_rgbd_batch_end", "psl_track_pose_batch_dev"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                              "there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        L.psl_last_error.restype = C.c_char_p
        L.psl_last_error.argtypes = [C.c_void_p]
        L.psl_stream.restype = C.c_void_p
        L.psl_stream.argtypes = [C.c_void_p]
        L.psl_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.psl_destroy.argtypes = [C.c_void_p]
        L.psl_destroy.restype = None
        L.psl_sync.argtypes = [C.c_void_p]
        L.psl_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        _i, _p, _l = C.c_int32, C.c_void_p, C.c_int64
        L.psl_orb_extract.argtypes = [_p, _p, _i, _i, _i, _p, _p, _i, _p]
        L.psl_orb_extract_batch.argtypes = [_p, _p, _i, _i, _i, _i, _l, _p, _p, _i, _p]
        L.psl_orb_extract_batch_dev.argtypes = [_p, _p, _i, _i, _i, _i, _l, _p, _p, _i, _p]
        L.psl_profile_enable.argtypes = [_p, _i]
        L.psl_profile_read.argtypes = [_p, _p, _p]
        L.psl_launch_count.argtypes = [_p]
        L.psl_launch_count.restype = C.c_int64
        L.psl_descriptor_distance.argtypes = [_p, _p, _p, _i, _p]
        L.psl_hamming_knn2.argtypes = [_p, _p, _i, _p, _i, _p, _p]
        L.psl_match_projection.argtypes = [_p, _p, _p, _p, _i, _p, _p, _p, _p]
        L.psl_match_bow.argtypes = [_p, _p, _p, _p, _i, _p, _p, _p, _i, _p, C.c_float, _i, _i, _p, _p]
        L.psl_track_orb_batch_dev.argtypes = [_p, _p, _i, _l, _p, _i, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p,
                                              _p, _p, _i]
        L.psl_track_orb_batch.argtypes = [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i]
        L.psl_line_extract.argtypes = [_p, _p, _i, _i, _i, _p, _p, _p, _p, _i, _p]
        L.psl_line_extract_batch.argtypes = [_p, _p, _i, _i, _i, _i, _l, _p, _p, _p, _p, _i, _p]
        L.psl_line_extract_batch_dev.argtypes = [_p, _p, _i, _i, _i, _i, _l, _p, _p, _p, _p, _i, _p]
        _f = C.c_float
        L.psl_line_match_nnr.argtypes = [_p, _p, _i, _p, _i, _f, _p, _p]
        L.psl_line_search_geom.argtypes = [_p, _p, _p, _p, _i, _p, _p, _i, _p, _f, _p, _p]
        L.psl_line_frame_bf_match.argtypes = [_p, _p, _i, _p, _i, _f, _f, _p]
        L.psl_line_search_double.argtypes = [_p, _p, _i, _p, _i, _f, _f, _p, _p]
        L.psl_line_match_projection.argtypes = [_p, _p, _p, _p, _i, _p, _i, _f, _p, _p]
        L.psl_plane_assoc.argtypes = [_p, _p, _p, _i, _p, _p, _p, _i, _f, _f, _i, _p, _p]
        L.psl_track_frontend_batch_dev.argtypes = [_p, _p, _i, _l, _p, _i, _l, _i, _i, _i, _p, _p, _p, _f, _p]
        L.psl_track_frontend_batch.argtypes = [_p, _p, _p, _i, _i, _i, _p, _p, _p, _f, _p]
        L.psl_convert_rgbd.argtypes = [_p, _p, _i, _i, _p, _p, _f, _p, _i, _i, _i]
        L.psl_convert_rgbd_dev.argtypes = [_p, _p, _i, _i, _i, _l, _p, _i, _l, _p, _i, _l, _f, _p, _i, _i, _i]
        L.psl_match_triangulation.argtypes = [_p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _i, _i, _i, _i, _p, _p]
        L.psl_match_fuse.argtypes = [_p, _p, _p, _p, _i, _p, _i, _i, _p, _p]
        L.psl_line_search_triangulation.argtypes = [_p, _p, _p, _i, _p, _p, _i, _f, _f, _i, _p, _p]
        L.psl_undistort_keypoints.argtypes = [_p, _p, _i, _p, _p]
        L.psl_undistort_keypoints_dev.argtypes = [_p, _p, _p, _i, _i, _p, _p]
        L.psl_image_bounds.argtypes = [_p, _i, _i, _p, _p]
        L.psl_lines_3d.argtypes = [_p, _p, _i, _p, _i, _i, _f, _f, _f, _f, C.c_uint32, _p, _p]
        L.psl_lines_3d_dev.argtypes = [_p, _p, _p, _i, _i, _p, _i, _i, _i, _l, _f, _f, _f, _f, C.c_uint32, _p, _p]
        L.psl_plane_hypotheses.argtypes = [_p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _i, _p]
        L.psl_line_fuse.argtypes = [_p, _p, _i, _p, _i, _p, _p, _i, _f, _i, _p, _p]
        L.psl_debug_fetch.argtypes = [_p, _i, _i, _i, _p, _l, _p]
        L.psl_match_bow_kf.argtypes = [_p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _f, _i, _i, _p, _p]
        L.psl_match_sim3.argtypes = [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p]
        L.psl_match_initialization.argtypes = [_p, _p, _p, _i, _p, _p, _i, _f, _i, _i, _p, _p]
        L.psl_line_search_triangulation_new.argtypes = [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _p, _f, _f, _i, _p, _p]
        L.psl_pose_optimization.argtypes = [_p, _p, _p, _i, _f, _f, _f, _f, _f, _p, _p, _p]
        L.psl_pose_optimization_dev.argtypes = [_p, _p, _p, _p, _i, _i, _f, _f, _f, _f, _f, _p, _p, _p]
        L.psl_track_rgbd_batch_dev.argtypes = [_p, _p, _i, _i, _i, _l, _p, _i, _l, _i, _i, _i, _p, _p, _p, _f, _p]
        L.psl_track_rgbd_batch.argtypes = [_p, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p, _f, _p]
        L.psl_track_rgbd_batch_begin.argtypes = [_p, _p, _i, _i, _p, _i, _i, _i, _p]
        L.psl_track_rgbd_batch_end.argtypes = [_p, _p, _p, _f, _p]
        L.psl_track_pose_batch_dev.argtypes = [_p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p]
        L.psl_pose_optimization_lil.argtypes = [_p, _p, _p, _i, _p, _i, _f, _f, _f, _f, _f, _p, _p, _p, _p]
        L.psl_pose_optimization_lil_dev.argtypes = [_p, _p, _p, _p, _i, _p, _p, _i, _i, _f, _f, _f, _f, _f, _p, _p, _p, _p]
        L.psl_line_junctions.argtypes = [_p, _p, _p, _i, _i, _i, _f, _f, _p, _p, _i, _p, _p]
        L.psl_line_junctions_dev.argtypes = [_p, _p, _p, _i, _i, _p, _i, _i, _f, _f, _p, _p, _i, _p, _p]
        _lib = L
    return _lib


def default_config() -> Config:
    cfg = Config()
    lib().psl_default_config(C.byref(cfg))
    return cfg
