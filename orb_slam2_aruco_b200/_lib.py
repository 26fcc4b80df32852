"""ctypes binding of libb200slam.so (the C-ABI declared in include/b200slam.h).

There is no CPU fallback: if the CUDA extension has not been built, importing callers get a loud
RuntimeError, and every compute entry point fails with B200_ENODEV on a box without a B200.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200slam.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
MARKER_DTYPE = np.dtype([("id", "<i4"), ("xy", "<f4", (8,))])
POSE_DTYPE = np.dtype([("rvec", "<f4", (3,)), ("tvec", "<f4", (3,)), ("err1", "<f4"), ("rvec2", "<f4", (3,)), ("tvec2", "<f4", (3,)), ("err2", "<f4")])
assert KP_DTYPE.itemsize == 28 and MARKER_DTYPE.itemsize == 36 and POSE_DTYPE.itemsize == 56

OK, EINVAL, ENODEV, ECUDA, ECAPACITY, ENOMEM = 0, -1, -2, -3, -4, -5
_NAMES = {0: "B200_OK", -1: "B200_EINVAL", -2: "B200_ENODEV", -3: "B200_ECUDA", -4: "B200_ECAPACITY", -5: "B200_ENOMEM"}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (_NAMES.get(code, str(code)), msg))
        self.code = code


_lib = None
vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200slam.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(orb_slam2_aruco_b200/build.py). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.b200_last_error.restype = C.c_char_p
        L.b200_launch_count.restype = i64
        L.b200_orb_create.argtypes = [C.POINTER(vp), i32, f32, i32, i32, i32, i32, i32, i32, i32]
        L.b200_orb_destroy.argtypes = [vp]
        L.b200_orb_max_keypoints.argtypes = [vp]
        L.b200_orb_get_level_info.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b200_orb_extract.argtypes = [vp, vp, i32, i32, i32, i64, i64, vp, vp, vp, vp]
        L.b200_orb_extract_host.argtypes = [vp, vp, i32, i32, i32, i64, i64, vp, vp, vp]
        L.b200_frontend_host.argtypes = [vp, vp, vp, i32, i32, i32, i64, i64, vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp]
        L.b200_aruco_check.argtypes = [vp, vp]
        L.b200_orb_set_profile.argtypes = [vp, i32]
        L.b200_orb_get_stage_ms.argtypes = [vp, vp]
        L.b200_orb_get_stage_frames.argtypes = [vp]
        L.b200_orb_get_pyramid.argtypes = [vp, i32, i32, vp, vp, vp]
        L.b200_orb_get_candidates.argtypes = [vp, i32, i32, vp, i32]
        L.b200_match_bf.argtypes = [vp, vp, i32, vp, vp, vp, i32, i32, f32, i32, i32, f32, vp, vp, i32, vp]
        L.b200_match_bf_kp.argtypes = [vp, vp, i32, vp, vp, vp, i32, i32, f32, i32, i32, f32, vp, vp, i32, vp]
        L.b200_match_bf_host.argtypes = [vp, vp, i32, vp, vp, vp, i32, i32, f32, i32, i32, f32, vp, vp, i32]
        L.b200_hamming_matrix_host.argtypes = [vp, i32, vp, i32, vp, i32]
        L.b200_match_by_projection_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, i32, i32, vp, i32]
        L.b200_match_by_bow_host.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, i32, f32, i32, i32, vp, i32]
        L.b200_distinctive_descriptors_host.argtypes = [vp, vp, i32, vp, vp, i32]
        L.b200_match_for_initialization_host.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, i32, f32, i32, vp, i32]
        L.b200_match_for_triangulation_host.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, i32]
        L.b200_match_kf_radius_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, i32, C.c_double, vp, vp, i32]
        L.b200_kf_project_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, vp, vp, i32, vp, vp, vp, i32]
        L.b200_kf_search_points_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, vp, vp, vp, i32, C.c_double, vp, vp, vp, vp, vp, i32]
        L.b200_match_candidates_host.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, i32]
        L.b200_aruco_create.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32, i32, i32]
        L.b200_aruco_destroy.argtypes = [vp]
        L.b200_aruco_max_markers.argtypes = [vp]
        L.b200_aruco_detect.argtypes = [vp, vp, i32, i32, i32, i64, i64, vp, vp, vp]
        L.b200_aruco_debug.argtypes = [vp, i32, vp, vp, vp, i32]
        L.b200_aruco_detect_host.argtypes = [vp, vp, i32, i32, i32, i64, i64, vp, vp]
        L.b200_aruco_get_contour.argtypes = [vp, i32, i32, vp, i32]
        L.b200_aruco_get_contours.argtypes = [vp, i32, i32, vp, vp, i32]
        L.b200_aruco_pose.argtypes = [vp, vp, i32, i32, f32, vp, vp, i32, vp]
        L.b200_aruco_pose_host.argtypes = [vp, i32, f32, vp, vp, i32]
        L.b200_aruco_detect_frame_host.argtypes = [vp, vp, i32, i32, i64, vp, vp, f32, vp, vp, vp, vp, i32]
        L.b200_voc_create.argtypes = [C.POINTER(vp), i32, i32, i32, vp, vp, vp, vp, i32]
        L.b200_voc_destroy.argtypes = [vp]
        L.b200_voc_num_words.argtypes = [vp]
        L.b200_voc_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
        L.b200_voc_transform_host.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.b200_frame_undistort.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp]
        L.b200_frame_undistort_points_host.argtypes = [vp, i32, vp, vp, i32]
        L.b200_frame_image_bounds.argtypes = [i32, i32, vp, vp, i32]
        L.b200_frame_assign_grid.argtypes = [vp, vp, i32, i32, vp, vp, vp, i32, vp]
        L.b200_frame_features_in_area.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, vp]
        L.b200_keyframe_features_in_area.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, vp]
        L.b200_collate_unique_id.argtypes = [vp]
        L.b200_collate_create.argtypes = [C.POINTER(vp), vp, i32, i32, i32]
        L.b200_collate_destroy.argtypes = [vp]
        L.b200_collate_rank.argtypes = [vp]
        L.b200_collate_world.argtypes = [vp]
        L.b200_collate_gather.argtypes = [vp, i32, vp, vp, vp, i32, vp]
        L.b200_collate_allgather.argtypes = [vp, i32, vp, vp, vp, vp]
        L.b200_collate_traffic.argtypes = [vp, vp, vp]
        L.b200_frontend_collate_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise B200Error(rc, lib().b200_last_error().decode(errors="replace"))
    return rc


def ptr(a):
    """void* of a numpy array or a torch tensor (device or host), or an int address"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch.Tensor


def launch_count():
    return int(lib().b200_launch_count())
