"""ctypes view of the C ABI in include/linemod_b200.h (liblinemod_b200.so, built in-tree by csrc/Makefile).

This is the binding a maintainer of the reference would write for a Python caller; the C++ caller's binding is the
header-only facade include/linemod_b200.hpp.  There is no fallback: if the shared library is
missing the import fails loudly, and lm_create fails when no CUDA device is usable.
"""
import ctypes as C
import os
import weakref

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblinemod_b200.so")

LM_OK, LM_E_INVALID, LM_E_CUDA, LM_E_IO, LM_E_NOTFOUND, LM_E_STATE = 0, -1, -2, -3, -4, -5
LM_8UC3, LM_16UC1, LM_8UC1 = 0, 1, 2
LM_COLOR_GRADIENT, LM_DEPTH_NORMAL = 0, 1
LM_MAX_MODALITIES, LM_MAX_LEVELS, LM_MAX_FEATURES = 4, 4, 63


class LmImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("type", C.c_int32),
                ("step", C.c_size_t)]


class LmModalityDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("weak_threshold", C.c_float), ("strong_threshold", C.c_float),
                ("distance_threshold", C.c_int32), ("difference_threshold", C.c_int32),
                ("extract_threshold", C.c_int32), ("num_features", C.c_int32)]


class LmQuery(C.Structure):
    _fields_ = [("threshold", C.c_float), ("class_ids", C.POINTER(C.c_char_p)), ("n_ids", C.c_int)]


class LmRect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class LmCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_double), ("fy", C.c_double),
                ("near_", C.c_double), ("far_", C.c_double)]


class LmRendererParams(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("angle_step", C.c_int32), ("radius_min", C.c_double), ("radius_max", C.c_double),
                ("radius_step", C.c_double), ("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_double),
                ("fy", C.c_double), ("near_", C.c_double), ("far_", C.c_double)]


POSE_DTYPE = np.dtype([("R", "<f8", (3, 3)), ("T", "<f8", (3,)), ("K", "<f4", (3, 3)), ("pad", "<u4"), ("D", "<f8"),
                       ("ori_dist", "<f8"), ("rect", [("x", "<i4"), ("y", "<i4"), ("width", "<i4"), ("height", "<i4")])])


class LmViewSphere(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("angle_min", C.c_int32), ("angle_max", C.c_int32), ("angle_step", C.c_int32),
                ("radius_min", C.c_float), ("radius_max", C.c_float), ("radius_step", C.c_float)]


MATCH_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("template_id", "<i4"), ("class_index", "<i4"),
                        ("similarity", "<f4")])
RAW_DTYPE = np.dtype([("order_key", "<u4"), ("coarse_pos", "<u4"), ("x", "<i4"), ("y", "<i4"), ("score", "<u4"),
                      ("nf", "<u4"), ("template_id", "<i4"), ("class_index", "<i4")])
HDR_DTYPE = np.dtype([("width", "<i4"), ("height", "<i4"), ("pyramid_level", "<i4"), ("num_features", "<i4")])
RESULT_HEADER_BYTES = 16

# every symbol include/linemod_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "lm_create", "lm_create_from_yaml", "lm_write_yaml", "lm_create_from_cache", "lm_write_cache", "lm_read_classes", "lm_write_classes", "lm_destroy",
    "lm_last_error", "lm_alloc_pinned", "lm_free_pinned", "lm_device", "lm_pyramid_levels", "lm_get_T",
    "lm_num_modalities", "lm_get_modality", "lm_num_classes", "lm_num_templates", "lm_class_id", "lm_get_templates",
    "lm_modality_process", "lm_qpyramid_destroy", "lm_qpyramid_levels", "lm_qpyramid_size", "lm_qpyramid_quantize", "lm_qpyramid_extract",
    "lm_add_template", "lm_add_template_from_quantized", "lm_add_synthetic_template", "lm_match", "lm_match_multi", "lm_match_batch", "lm_match_batch_multi", "lm_free_matches",
    "lm_stream_open", "lm_stream_push", "lm_stream_pop", "lm_stream_in_flight", "lm_stream_close",
    "lm_match_device", "lm_match_device_multi", "lm_match_device_multi_lane", "lm_device_result_region", "lm_copy_result_block", "lm_match_device_stream", "lm_finalize_raw", "lm_finalize_gathered", "lm_upload_images", "lm_set_shard", "lm_set_similarity_lut", "lm_get_similarity_lut",
    "lm_set_normal_lut", "lm_get_normal_lut", "lm_load_normal_lut_file", "lm_debug_fetch", "lm_build_front", "lm_level_geometry",
    "lm_mesh_create", "lm_mesh_load_stl", "lm_mesh_num_triangles", "lm_mesh_get_triangles", "lm_mesh_destroy", "lm_view_count", "lm_view_params",
    "lm_view_pose", "lm_render_views", "lm_add_templates_batch", "lm_train_views", "lm_depth_diff_batch",
    "lm_write_renderer_params", "lm_read_renderer_params", "lm_free_poses",
    "lm_group_create", "lm_group_create_grid", "lm_group_destroy", "lm_group_size", "lm_group_mode", "lm_group_member", "lm_group_set_option",
    "lm_group_match_batch_multi", "lm_group_match",
    "lm_cluster_matches", "lm_free_clusters", "lm_debug_coarse_map", "lm_debug_presort", "lm_last_timings", "lm_last_work", "lm_set_option",
]

_lib = None


class LinemodError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("linemod_b200 error %d: %s" % (code, message))
        self.code = code


def lib():
    """Loads liblinemod_b200.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C %s` (or __graft_entry__.build()); "
                          "there is no CPU fallback" % (LIB_PATH, os.path.join(_PKG, "csrc")))
    L = C.CDLL(LIB_PATH)
    vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
    L.lm_create.argtypes = [C.POINTER(C.c_int32), ci, C.POINTER(LmModalityDesc), ci, C.POINTER(vp)]
    L.lm_create_from_yaml.argtypes = [cp, C.POINTER(vp)]
    L.lm_write_yaml.argtypes = [vp, cp]
    L.lm_create_from_cache.argtypes = [cp, C.POINTER(vp)]
    L.lm_write_cache.argtypes = [vp, cp]
    L.lm_read_classes.argtypes = [vp, C.POINTER(cp), ci, cp]
    L.lm_write_classes.argtypes = [vp, cp]
    L.lm_destroy.argtypes = [vp]
    L.lm_destroy.restype = None
    L.lm_last_error.restype = cp
    L.lm_alloc_pinned.argtypes = [C.c_size_t]
    L.lm_alloc_pinned.restype = vp
    L.lm_free_pinned.argtypes = [vp]
    L.lm_free_pinned.restype = None
    for n in ("lm_device", "lm_pyramid_levels", "lm_num_modalities", "lm_num_classes"):
        getattr(L, n).argtypes = [vp]
    L.lm_get_T.argtypes = [vp, ci]
    L.lm_get_modality.argtypes = [vp, ci, C.POINTER(LmModalityDesc)]
    L.lm_num_templates.argtypes = [vp, cp]
    L.lm_class_id.argtypes = [vp, ci]
    L.lm_class_id.restype = cp
    L.lm_get_templates.argtypes = [vp, cp, ci, vp, vp]
    L.lm_add_template.argtypes = [vp, C.POINTER(LmImage), ci, cp, C.POINTER(LmImage), C.POINTER(LmRect)]
    L.lm_add_template_from_quantized.argtypes = [vp, C.POINTER(LmImage), C.POINTER(vp), cp, C.POINTER(LmImage),
                                                 C.POINTER(LmRect)]
    L.lm_add_synthetic_template.argtypes = [vp, cp, ci, vp, vp]
    L.lm_modality_process.argtypes = [C.POINTER(LmModalityDesc), C.POINTER(LmImage), C.POINTER(LmImage), ci, vp, C.POINTER(vp)]
    L.lm_qpyramid_destroy.argtypes = [vp]
    L.lm_qpyramid_destroy.restype = None
    L.lm_qpyramid_levels.argtypes = [vp]
    L.lm_qpyramid_size.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    L.lm_qpyramid_quantize.argtypes = [vp, ci, C.POINTER(LmImage)]
    L.lm_qpyramid_extract.argtypes = [vp, ci, vp, vp]
    L.lm_match.argtypes = [vp, C.POINTER(LmImage), ci, C.c_float, C.POINTER(cp), ci, C.POINTER(LmImage), ci,
                           C.POINTER(LmImage), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lm_match_multi.argtypes = [vp, C.POINTER(LmImage), ci, C.POINTER(LmQuery), ci, C.POINTER(LmImage), ci,
                                 C.POINTER(LmImage), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lm_match_device_multi.argtypes = [vp, C.POINTER(vp), ci, ci, ci, C.POINTER(LmQuery), ci, vp, C.POINTER(vp),
                                        C.POINTER(C.c_size_t)]
    L.lm_match_device_multi_lane.argtypes = [vp, ci, C.POINTER(vp), ci, ci, ci, C.POINTER(LmQuery), ci, vp, C.POINTER(vp),
                                             C.POINTER(C.c_size_t)]
    L.lm_device_result_region.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(ci)]
    L.lm_copy_result_block.argtypes = [vp, ci, vp, C.c_size_t, vp]
    L.lm_match_device_stream.argtypes = [vp, C.POINTER(vp), ci, ci, ci, ci, C.POINTER(LmQuery), ci, C.POINTER(vp), ci, vp,
                                         C.c_size_t]
    L.lm_match_batch.argtypes = [vp, C.POINTER(LmImage), ci, ci, C.c_float, C.POINTER(cp), ci, C.POINTER(vp),
                                 C.POINTER(C.c_size_t)]
    L.lm_match_batch_multi.argtypes = [vp, C.POINTER(LmImage), ci, ci, C.POINTER(LmQuery), ci, C.POINTER(vp),
                                       C.POINTER(C.c_size_t)]
    L.lm_free_matches.argtypes = [vp]
    L.lm_stream_open.argtypes = [vp, C.POINTER(LmQuery), ci, C.POINTER(vp)]
    L.lm_stream_push.argtypes = [vp, C.POINTER(LmImage), ci, ci]
    L.lm_stream_pop.argtypes = [vp, ci, ci, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(ci)]
    L.lm_stream_in_flight.argtypes = [vp]
    L.lm_stream_close.argtypes = [vp]
    L.lm_stream_close.restype = None
    L.lm_free_matches.restype = None
    L.lm_match_device.argtypes = [vp, C.POINTER(vp), ci, ci, ci, C.c_float, C.POINTER(cp), ci, vp, C.POINTER(vp),
                                  C.POINTER(C.c_size_t)]
    L.lm_finalize_raw.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lm_finalize_gathered.argtypes = [vp, vp, ci, ci, C.c_size_t, C.c_size_t, C.c_uint32, ci, C.POINTER(vp),
                                       C.POINTER(C.c_size_t), vp]
    L.lm_upload_images.argtypes = [vp, C.POINTER(LmImage), ci, C.POINTER(vp), vp]
    L.lm_set_shard.argtypes = [vp, ci, ci]
    for n in ("lm_set_similarity_lut", "lm_get_similarity_lut", "lm_set_normal_lut", "lm_get_normal_lut"):
        getattr(L, n).argtypes = [vp, vp]
    L.lm_load_normal_lut_file.argtypes = [vp, cp]
    L.lm_mesh_create.argtypes = [vp, ci, C.POINTER(vp)]
    L.lm_mesh_load_stl.argtypes = [cp, C.POINTER(vp)]
    L.lm_mesh_num_triangles.argtypes = [vp]
    L.lm_mesh_get_triangles.argtypes = [vp, vp]
    L.lm_mesh_destroy.argtypes = [vp]
    L.lm_mesh_destroy.restype = None
    L.lm_view_count.argtypes = [C.POINTER(LmViewSphere)]
    L.lm_view_params.argtypes = [C.POINTER(LmViewSphere), ci, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]
    L.lm_view_pose.argtypes = [vp, vp, vp, vp]
    L.lm_render_views.argtypes = [vp, vp, C.POINTER(LmCamera), vp, vp, ci, vp, vp, vp, vp]
    L.lm_add_templates_batch.argtypes = [vp, C.POINTER(LmImage), C.POINTER(LmImage), ci, ci, cp, vp, vp]
    L.lm_train_views.argtypes = [vp, vp, C.POINTER(LmCamera), vp, vp, ci, cp, vp, vp, vp, vp]
    L.lm_depth_diff_batch.argtypes = [vp, C.POINTER(LmImage), vp, C.POINTER(LmCamera), vp, vp, vp, vp, ci, vp]
    L.lm_write_renderer_params.argtypes = [cp, vp, C.c_size_t, C.POINTER(LmRendererParams)]
    L.lm_read_renderer_params.argtypes = [cp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(LmRendererParams)]
    L.lm_free_poses.argtypes = [vp]
    L.lm_free_poses.restype = None
    L.lm_group_create.argtypes = [vp, C.POINTER(C.c_int), ci, ci, C.POINTER(vp)]
    L.lm_group_create_grid.argtypes = [vp, C.POINTER(C.c_int), ci, ci, C.POINTER(vp)]
    L.lm_group_destroy.argtypes = [vp]
    L.lm_group_destroy.restype = None
    L.lm_group_size.argtypes = [vp]
    L.lm_group_mode.argtypes = [vp]
    L.lm_group_member.argtypes = [vp, ci]
    L.lm_group_member.restype = vp
    L.lm_group_set_option.argtypes = [vp, cp, ci]
    L.lm_group_match_batch_multi.argtypes = [vp, C.POINTER(LmImage), ci, ci, C.POINTER(LmQuery), ci, C.POINTER(vp),
                                             C.POINTER(C.c_size_t)]
    L.lm_group_match.argtypes = [vp, C.POINTER(LmImage), ci, C.c_float, C.POINTER(cp), ci, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lm_cluster_matches.argtypes = [vp, C.c_size_t, vp, vp, C.c_size_t, vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp)]
    L.lm_free_clusters.argtypes = [vp, vp]
    L.lm_free_clusters.restype = None
    L.lm_debug_fetch.argtypes = [vp, ci, ci, ci, vp]
    L.lm_debug_fetch.restype = C.c_long
    L.lm_build_front.argtypes = [vp, C.POINTER(LmImage), ci, C.POINTER(LmImage), ci]
    L.lm_level_geometry.argtypes = [vp, ci, C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]
    L.lm_debug_coarse_map.argtypes = [vp, cp, ci, vp]
    L.lm_debug_presort.argtypes = [vp, vp]
    L.lm_debug_presort.restype = C.c_long
    L.lm_last_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(ci)]
    L.lm_last_work.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.lm_set_option.argtypes = [vp, cp, ci]
    _lib = L
    return L


def last_error():
    return lib().lm_last_error().decode(errors="replace")


def check(rc):
    if rc < 0:
        raise LinemodError(rc, last_error())
    return rc


def image(a):
    """numpy array -> (LmImage, keepalive).  Rows may be strided (an ROI view); pixels within a row must be packed."""
    if a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3:
        t, px = LM_8UC3, 3
    elif a.dtype == np.uint16 and a.ndim == 2:
        t, px = LM_16UC1, 2
    elif a.dtype == np.uint8 and a.ndim == 2:
        t, px = LM_8UC1, 1
    else:
        raise TypeError("unsupported image dtype/shape %s %s" % (a.dtype, a.shape))
    if a.strides[1] != px or (a.ndim == 3 and a.strides[2] != 1) or a.strides[0] < a.shape[1] * px:
        a = np.ascontiguousarray(a)
    return LmImage(a.ctypes.data, a.shape[0], a.shape[1], t, a.strides[0]), a


def query_array(queries):
    """[(threshold, [class ids])] -> (LmQuery array, keepalive)."""
    keep = []
    arr = (LmQuery * max(1, len(queries)))()
    for i, (thr, ids) in enumerate(queries):
        enc = [c.encode() for c in ids]
        ca = (C.c_char_p * max(1, len(enc)))(*enc)
        keep.append((enc, ca))
        arr[i].threshold = thr
        arr[i].class_ids = ca
        arr[i].n_ids = len(enc)
    return arr, keep


def image_array(images):
    keep = [image(a) for a in images]
    arr = (LmImage * max(1, len(keep)))(*[k[0] for k in keep])
    return arr, keep


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked memory from lm_alloc_pinned (freed when the array is collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().lm_alloc_pinned(n)
    if not p:
        raise MemoryError("lm_alloc_pinned(%d) failed: %s" % (n, last_error()))
    buf = (C.c_uint8 * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, lib().lm_free_pinned, p)
    return arr
