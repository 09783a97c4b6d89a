"""Template generation at scale and the depth hypothesis check (N4) over the C ABI.

Mirrors the trainer loop of /root/reference/src/renderer.cpp:239-329: a mesh (`Renderer3d(stl_file)`), camera
parameters (`set_parameters`), the view sphere (`RendererIterator`) and, per view, render + `addTemplate`.  Rendering,
quantisation and feature extraction all run in the library's CUDA kernels; nothing here computes on pixels.
"""
import ctypes as C

import numpy as np

from ._capi import LmCamera, LmRendererParams, LmViewSphere, POSE_DTYPE, check, image, image_array, lib

RECT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("width", "<i4"), ("height", "<i4")])


class Mesh:
    """Triangle mesh in the object frame, metres (the reference loads config/stl/*.stl)."""

    def __init__(self, triangles=None, _handle=None):
        if _handle is not None:
            self._h = _handle
            return
        tri = np.ascontiguousarray(triangles, np.float32).reshape(-1, 3, 3)
        h = C.c_void_p()
        check(lib().lm_mesh_create(tri.ctypes.data, len(tri), C.byref(h)))
        self._h = h

    @classmethod
    def load_stl(cls, path):
        h = C.c_void_p()
        check(lib().lm_mesh_load_stl(str(path).encode(), C.byref(h)))
        return cls(_handle=h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().lm_mesh_destroy(h)

    def __len__(self):
        return lib().lm_mesh_num_triangles(self._h)

    @property
    def triangles(self):
        out = np.zeros((len(self), 3, 3), np.float32)
        if len(self):
            check(lib().lm_mesh_get_triangles(self._h, out.ctypes.data))
        return out


def camera(width=640, height=480, fx=535.566011, fy=537.168115, near=0.1, far=1000.0):
    """Renderer3d::set_parameters; defaults are the reference's Carmine values (renderer.cpp:205-213)."""
    return LmCamera(width, height, fx, fy, near, far)


class ViewSphere:
    """RendererIterator: n_points on a golden-spiral sphere x in-plane angles x radii (renderer.cpp:242-246)."""

    def __init__(self, n_points=150, angle_step=10, radius_min=0.5, radius_max=1.0, radius_step=0.1, angle_min=-80,
                 angle_max=80):
        self.c = LmViewSphere(n_points, angle_min, angle_max, angle_step, radius_min, radius_max, radius_step)

    def __len__(self):
        return check(lib().lm_view_count(C.byref(self.c)))

    def view(self, index):
        """-> (T[3], up[3], radius, sphere point, in-plane angle in degrees)"""
        T, up = np.zeros(3), np.zeros(3)
        r, pt, ang = C.c_float(), C.c_int32(), C.c_int32()
        check(lib().lm_view_params(C.byref(self.c), index, T.ctypes.data, up.ctypes.data, C.byref(r), C.byref(pt), C.byref(ang)))
        return T, up, r.value, pt.value, ang.value

    def views(self, indices=None):
        """-> (T[n,3], up[n,3]) for the given view indices (default: all, in iteration order)"""
        idx = range(len(self)) if indices is None else indices
        got = [self.view(i)[:2] for i in idx]
        return (np.array([g[0] for g in got]).reshape(-1, 3), np.array([g[1] for g in got]).reshape(-1, 3))


def view_pose(T, up):
    """Object pose in the camera frame for a view: Pc = R @ Po + t."""
    T, up = np.ascontiguousarray(T, np.float64), np.ascontiguousarray(up, np.float64)
    R, t = np.zeros((3, 3)), np.zeros(3)
    check(lib().lm_view_pose(T.ctypes.data, up.ctypes.data, R.ctypes.data, t.ctypes.data))
    return R, t


def _views(T, up):
    T = np.ascontiguousarray(T, np.float64).reshape(-1, 3)
    up = np.ascontiguousarray(up, np.float64).reshape(-1, 3)
    assert len(T) == len(up)
    return T, up


def render_views(det, mesh, cam, T, up, want=("bgr", "depth", "mask")):
    """lm_render_views -> dict of the requested images ([n, rows, cols(, 3)]) + "rects" (RECT_DTYPE)."""
    T, up = _views(T, up)
    n, rows, cols = len(T), cam.height, cam.width
    out = {}
    if "bgr" in want:
        out["bgr"] = np.zeros((n, rows, cols, 3), np.uint8)
    if "depth" in want:
        out["depth"] = np.zeros((n, rows, cols), np.uint16)
    if "mask" in want:
        out["mask"] = np.zeros((n, rows, cols), np.uint8)
    rects = np.zeros(n, RECT_DTYPE)
    ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
    check(lib().lm_render_views(det._h, mesh._h, C.byref(cam), T.ctypes.data, up.ctypes.data, n, ptr("bgr"), ptr("depth"),
                                ptr("mask"), rects.ctypes.data))
    out["rects"] = rects
    return out


def train_views(det, mesh, cam, T, up, class_id, centre_depth=False):
    """lm_train_views: render + addTemplate per view on the GPU -> (template_ids[n], bounding_boxes[n], mask_rects[n]
    [, centre depth in mm [n]])."""
    T, up = _views(T, up)
    n = len(T)
    tids = np.full(n, -1, np.int32)
    bbs, rects = np.zeros(n, RECT_DTYPE), np.zeros(n, RECT_DTYPE)
    centre = np.zeros(n, np.uint16)
    check(lib().lm_train_views(det._h, mesh._h, C.byref(cam), T.ctypes.data, up.ctypes.data, n, class_id.encode(),
                               tids.ctypes.data, bbs.ctypes.data, rects.ctypes.data, centre.ctypes.data if centre_depth else None))
    return (tids, bbs, rects, centre) if centre_depth else (tids, bbs, rects)


def add_templates_batch(det, views, class_id):
    """lm_add_templates_batch: views = [(sources, mask)] -> (template_ids[n], bounding_boxes[n]); == addTemplate per view."""
    n = len(views)
    M = lib().lm_num_modalities(det._h)
    flat, masks = [], []
    for sources, mask in views:
        assert len(sources) == M
        flat.extend(sources)
        masks.append(mask)
    sarr, keep1 = image_array(flat)
    marr, keep2 = image_array(masks)
    tids = np.full(max(n, 1), -1, np.int32)
    bbs = np.zeros(max(n, 1), RECT_DTYPE)
    check(lib().lm_add_templates_batch(det._h, sarr, marr, n, M, class_id.encode(), tids.ctypes.data, bbs.ctypes.data))
    return tids[:n], bbs[:n]


def depth_diff(det, scene_depth, mesh, cam, T, up, xs, ys):
    """lm_depth_diff_batch: rgbdDetector::depth_diff of n hypotheses (template view T/up laid at (x, y)) -> f64[n] metres."""
    T, up = _views(T, up)
    xs, ys = np.ascontiguousarray(xs, np.int32), np.ascontiguousarray(ys, np.int32)
    assert len(xs) == len(ys) == len(T)
    simg, keep = image(scene_depth)
    out = np.zeros(len(T), np.float64)
    check(lib().lm_depth_diff_batch(det._h, C.byref(simg), mesh._h, C.byref(cam), T.ctypes.data, up.ctypes.data,
                                    xs.ctypes.data, ys.ctypes.data, len(T), out.ctypes.data))
    return out


def write_renderer_params(path, poses, params):
    """writeLinemodTemplateParams (renderer.cpp:72-123): poses = POSE_DTYPE array, params = LmRendererParams."""
    poses = np.ascontiguousarray(poses, POSE_DTYPE)
    check(lib().lm_write_renderer_params(str(path).encode(), poses.ctypes.data, len(poses), C.byref(params)))


def read_renderer_params(path):
    """readLinemodTemplateParams (rgbdDetector.cpp:1681-1749) -> (POSE_DTYPE array, LmRendererParams)."""
    p, n, params = C.c_void_p(), C.c_size_t(), LmRendererParams()
    check(lib().lm_read_renderer_params(str(path).encode(), C.byref(p), C.byref(n), C.byref(params)))
    out = np.zeros(n.value, POSE_DTYPE)
    if n.value:
        C.memmove(out.ctypes.data, p.value, n.value * POSE_DTYPE.itemsize)
        lib().lm_free_poses(p)
    return out, params


def poses_for_views(T, up, cam, radii, rects, centre_depth_mm):
    """The trainer's per-template records (renderer.cpp:262-318) for views (T, up): R = lm_view_pose's rotation, T = minus
    the camera position, K from the camera, D = radius - centre depth, Ori_dist = radius, Rect = the render rectangle."""
    T, up = _views(T, up)
    out = np.zeros(len(T), POSE_DTYPE)
    for i in range(len(T)):
        out["R"][i] = view_pose(T[i], up[i])[0]
        out["T"][i] = -T[i]
        out["K"][i] = np.array([[cam.fx, 0, cam.width / 2.0], [0, cam.fy, cam.height / 2.0], [0, 0, 1]], np.float32)
        out["ori_dist"][i] = radii[i]
        out["D"][i] = float(radii[i]) - float(np.float32(centre_depth_mm[i]) / np.float32(1000.0))
        out["rect"][i] = tuple(rects[i])
    return out
