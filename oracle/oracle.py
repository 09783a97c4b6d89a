"""ctypes binding of the CPU oracle (oracle/linemod_oracle.cpp).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.  PARITY UNPINNED: see the header of
linemod_oracle.cpp -- the reference's hot path (OpenCV 2.4.x objdetect/linemod.cpp, entered at
/root/reference/src/rgbdDetector.cpp:31-34) is not vendored and cannot be built here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblinemod_oracle.so")

TYPE_8UC3, TYPE_16UC1, TYPE_8UC1 = 0, 1, 2
COLOR_GRADIENT, DEPTH_NORMAL = 0, 1


class OrcImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("type", C.c_int32),
                ("step", C.c_size_t)]


class OrcModality(C.Structure):
    _fields_ = [("type", C.c_int32), ("weak_threshold", C.c_float), ("strong_threshold", C.c_float),
                ("distance_threshold", C.c_int32), ("difference_threshold", C.c_int32),
                ("extract_threshold", C.c_int32), ("num_features", C.c_int32)]


MATCH_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("template_id", "<i4"), ("class_index", "<i4"),
                        ("similarity", "<f4")])
RAW_DTYPE = np.dtype([("order_key", "<u4"), ("coarse_pos", "<u4"), ("x", "<i4"), ("y", "<i4"), ("score", "<u4"),
                      ("nf", "<u4"), ("template_id", "<i4"), ("class_index", "<i4")])
CAND_DTYPE = np.dtype([("class_index", "<i4"), ("template_id", "<i4"), ("pos", "<i4"), ("raw", "<i4")])


class OrcViewSphere(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("angle_min", C.c_int32), ("angle_max", C.c_int32), ("angle_step", C.c_int32),
                ("radius_min", C.c_float), ("radius_max", C.c_float), ("radius_step", C.c_float)]


class OrcCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_double), ("fy", C.c_double),
                ("near_", C.c_double), ("far_", C.c_double)]


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++ only; no GPU, no reference sources needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("linemod_oracle.cpp", "render_oracle.cpp")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(C.c_int32), C.c_int, C.POINTER(OrcModality), C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_max_threads.restype = C.c_int
        for n in ("orc_set_similarity_lut", "orc_get_similarity_lut", "orc_set_normal_lut", "orc_get_normal_lut"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_add_template.restype = C.c_int
        L.orc_add_template.argtypes = [C.c_void_p, C.POINTER(OrcImage), C.c_int, C.c_char_p, C.POINTER(OrcImage),
                                       C.POINTER(C.c_int32)]
        L.orc_modality_process.restype = C.c_int
        L.orc_modality_process.argtypes = [C.c_void_p, C.c_int, C.POINTER(OrcImage), C.POINTER(OrcImage), C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.orc_add_synthetic_template.restype = C.c_int
        L.orc_add_synthetic_template.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_num_classes.argtypes = [C.c_void_p]
        L.orc_num_templates.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_class_id.restype = C.c_char_p
        L.orc_class_id.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_template.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_build_front.argtypes = [C.c_void_p, C.POINTER(OrcImage), C.c_int, C.POINTER(OrcImage), C.c_int]
        L.orc_match_only.restype = C.c_long
        L.orc_match_only.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_char_p), C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]
        L.orc_match.restype = C.c_long
        L.orc_match.argtypes = [C.c_void_p, C.POINTER(OrcImage), C.c_int, C.c_float, C.POINTER(C.c_char_p), C.c_int,
                                C.POINTER(OrcImage), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_last_presort.restype = C.c_long
        L.orc_last_presort.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_last_raw.restype = C.c_long
        L.orc_last_raw.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_last_candidates.restype = C.c_long
        L.orc_last_candidates.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_coarse_map.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        L.orc_debug_fetch.restype = C.c_long
        L.orc_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_level_geometry.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]
        L.orc_sort_unique.restype = C.c_long
        L.orc_sort_unique.argtypes = [C.c_void_p, C.c_long]
        L.orc_view_count.argtypes = [C.POINTER(OrcViewSphere)]
        L.orc_view_params.argtypes = [C.POINTER(OrcViewSphere), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_float)]
        L.orc_view_list.argtypes = [C.POINTER(OrcViewSphere), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_look_at.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_fast.argtypes = [C.c_void_p, C.c_int]
        L.orc_train_views.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcCamera), C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_char_p, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.POINTER(OrcCamera), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
        L.orc_depth_diff.restype = C.c_double
        L.orc_depth_diff.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 6
        L.orc_prim_phase_deg.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_prim_cg_quantize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _img(a):
    """numpy array (possibly a strided view with contiguous rows) -> (OrcImage, keepalive)."""
    if a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3:
        t = TYPE_8UC3
    elif a.dtype == np.uint16 and a.ndim == 2:
        t = TYPE_16UC1
    elif a.dtype == np.uint8 and a.ndim == 2:
        t = TYPE_8UC1
    else:
        raise TypeError("unsupported image %s %s" % (a.dtype, a.shape))
    if a.strides[1] != a.itemsize * (3 if t == TYPE_8UC3 else 1):
        a = np.ascontiguousarray(a)
    return OrcImage(a.ctypes.data, a.shape[0], a.shape[1], t, a.strides[0]), a


def _img_array(images):
    keep = [_img(a) for a in images]
    arr = (OrcImage * max(1, len(keep)))(*[k[0] for k in keep])
    return arr, keep


def color_gradient(weak=10.0, num_features=63, strong=55.0):
    return OrcModality(COLOR_GRADIENT, weak, strong, 2000, 50, 2, num_features)


def depth_normal(distance=2000, difference=50, num_features=63, extract=2):
    return OrcModality(DEPTH_NORMAL, 10.0, 55.0, distance, difference, extract, num_features)


class Stage:
    QUANTIZED, SPREAD, RESPONSE, LINEAR, MAGNITUDE, QUANT_RAW = range(6)


class OracleDetector:
    """CPU restatement of cv::linemod::Detector (the surface used at /root/reference/src/renderer.cpp:179-185,308
    and src/rgbdDetector.cpp:31-34)."""

    def __init__(self, modalities=None, T=(5, 8)):
        if modalities is None:
            modalities = [color_gradient(), depth_normal()]
        self.modalities = list(modalities)
        self.T = list(T)
        Ta = (C.c_int32 * len(self.T))(*self.T)
        Ma = (OrcModality * len(self.modalities))(*self.modalities)
        self._h = C.c_void_p(lib().orc_create(Ta, len(self.T), Ma, len(self.modalities)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # -- configuration
    def set_threads(self, n):
        lib().orc_set_threads(self._h, int(n))

    def set_fast(self, on=True):
        """Front end by the CPU baseline's SSE / threaded routines (linemod_fast.inc) instead of the plain restatement;
        results are identical (tests/test_oracle_fast.py)."""
        lib().orc_set_fast(self._h, int(bool(on)))

    @staticmethod
    def max_threads():
        return lib().orc_max_threads()

    def set_similarity_lut(self, lut):
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        assert lut.size == 256
        lib().orc_set_similarity_lut(self._h, lut.ctypes.data)

    def similarity_lut(self):
        out = np.empty(256, np.uint8)
        lib().orc_get_similarity_lut(self._h, out.ctypes.data)
        return out

    def set_normal_lut(self, lut):
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        assert lut.size == 8000
        lib().orc_set_normal_lut(self._h, lut.ctypes.data)

    def normal_lut(self):
        out = np.empty(8000, np.uint8)
        lib().orc_get_normal_lut(self._h, out.ctypes.data)
        return out

    def _err(self):
        return lib().orc_last_error(self._h).decode()

    # -- templates
    def add_template(self, sources, class_id, mask=None):
        arr, keep = _img_array(sources)
        bb = (C.c_int32 * 4)()
        mimg = None
        if mask is not None:
            m, mk = _img(mask)
            mimg = C.pointer(m)
        r = lib().orc_add_template(self._h, arr, len(sources), class_id.encode(), mimg, bb)
        if r == -2:
            raise ValueError(self._err())
        return r, tuple(bb)

    def modality_process(self, m, src, mask, level):
        """Modality::process(src, mask) of modality m, `level` pyrDown()s, then quantize() and extractTemplate()
        -> (quantized u8 image, ok, (width, height, pyramid_level, features[n,3]))."""
        simg, sk = _img(src)
        mimg = None
        if mask is not None:
            mm, mk = _img(mask)
            mimg = C.pointer(mm)
        r, c = src.shape[0] >> level, src.shape[1] >> level
        q = np.zeros((r, c), np.uint8)
        hdr = np.zeros(4, np.int32)
        feats = np.zeros((63, 3), np.int32)
        rc = lib().orc_modality_process(self._h, m, C.byref(simg), mimg, level, q.ctypes.data, hdr.ctypes.data, feats.ctypes.data)
        if rc == -2:
            raise ValueError(self._err())
        return q, bool(rc), (int(hdr[0]), int(hdr[1]), int(hdr[2]), feats[:int(hdr[3])].copy())

    def train_views(self, triangles, cam, T, up, class_id):
        """The trainer's loop (render + addTemplate per view, det threads at a time, templates in view order).
        -> int32 template ids per view (-1: some level lacked candidates)."""
        tri = np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
        T, up = np.ascontiguousarray(T, np.float64).reshape(-1, 3), np.ascontiguousarray(up, np.float64).reshape(-1, 3)
        tids = np.zeros(len(T), np.int32)
        rc = lib().orc_train_views(self._h, tri.ctypes.data, len(tri), C.byref(cam), T.ctypes.data, up.ctypes.data, len(T),
                                   class_id.encode(), tids.ctypes.data)
        if rc < 0:
            raise RuntimeError("orc_train_views failed: " + self._err())
        return tids

    def add_synthetic_template(self, class_id, templates):
        """templates: list (L*M) of (width, height, pyramid_level, features[n,3] int)."""
        hdr = np.array([[t[0], t[1], t[2], len(t[3])] for t in templates], np.int32)
        feats = np.concatenate([np.asarray(t[3], np.int32).reshape(-1, 3) for t in templates]).astype(np.int32)
        feats = np.ascontiguousarray(feats)
        r = lib().orc_add_synthetic_template(self._h, class_id.encode(), len(templates), hdr.ctypes.data,
                                             feats.ctypes.data)
        if r < 0:
            raise ValueError(self._err())
        return r

    def class_ids(self):
        return [lib().orc_class_id(self._h, i).decode() for i in range(lib().orc_num_classes(self._h))]

    def num_templates(self, class_id=None):
        return lib().orc_num_templates(self._h, class_id.encode() if class_id is not None else None)

    def get_template(self, class_id, template_id):
        n_t = len(self.T) * len(self.modalities)
        hdr = np.zeros((n_t, 4), np.int32)
        total = lib().orc_get_template(self._h, class_id.encode(), template_id, hdr.ctypes.data, None)
        if total < 0:
            raise KeyError((class_id, template_id))
        feats = np.zeros((max(total, 1), 3), np.int32)
        lib().orc_get_template(self._h, class_id.encode(), template_id, hdr.ctypes.data, feats.ctypes.data)
        out, k = [], 0
        for i in range(n_t):
            nf = int(hdr[i, 3])
            out.append((int(hdr[i, 0]), int(hdr[i, 1]), int(hdr[i, 2]), feats[k:k + nf].copy()))
            k += nf
        return out

    # -- matching
    def build_front(self, sources, masks=()):
        arr, keep = _img_array(sources)
        marr, mkeep = _img_array(masks)
        if lib().orc_build_front(self._h, arr, len(sources), marr, len(masks)) != 0:
            raise ValueError(self._err())

    @staticmethod
    def _ids(class_ids):
        ids = [c.encode() for c in class_ids]
        return (C.c_char_p * max(1, len(ids)))(*ids), len(ids)

    def _take(self, n, out):
        if n < 0:
            raise ValueError(self._err())
        res = np.empty(n, MATCH_DTYPE)
        if n:
            C.memmove(res.ctypes.data, out, n * MATCH_DTYPE.itemsize)
        lib().orc_free(out)
        return res

    def match_only(self, threshold, class_ids=(), keep_candidates=False):
        ids, n_ids = self._ids(class_ids)
        out = C.c_void_p()
        n = lib().orc_match_only(self._h, threshold, ids, n_ids, int(keep_candidates), C.byref(out))
        return self._take(n, out)

    def match(self, sources, threshold, class_ids=(), masks=(), keep_candidates=False):
        arr, keep = _img_array(sources)
        marr, mkeep = _img_array(masks)
        ids, n_ids = self._ids(class_ids)
        out = C.c_void_p()
        n = lib().orc_match(self._h, arr, len(sources), threshold, ids, n_ids, marr, len(masks),
                            int(keep_candidates), C.byref(out))
        return self._take(n, out)

    def last_presort(self):
        n = lib().orc_last_presort(self._h, None)
        res = np.empty(n, MATCH_DTYPE)
        lib().orc_last_presort(self._h, res.ctypes.data)
        return res

    def last_raw(self):
        """Survivors of the last match(keep_candidates=True) as lm_raw_match-shaped records (emission order)."""
        n = lib().orc_last_raw(self._h, None)
        res = np.empty(n, RAW_DTYPE)
        lib().orc_last_raw(self._h, res.ctypes.data)
        return res

    def last_candidates(self):
        n = lib().orc_last_candidates(self._h, None)
        res = np.empty(n, CAND_DTYPE)
        lib().orc_last_candidates(self._h, res.ctypes.data)
        return res

    def geometry(self, level, modality=0):
        g = (C.c_int32 * 5)()
        ps = C.c_size_t()
        if lib().orc_level_geometry(self._h, level, modality, g, C.byref(ps)) != 0:
            raise ValueError("no front end built")
        return dict(rows=g[0], cols=g[1], T=g[2], W=g[3], H=g[4], plane_stride=ps.value)

    def coarse_map(self, class_id, template_id):
        g = self.geometry(len(self.T) - 1)
        out = np.zeros((g["H"], g["W"]), np.uint16)
        if lib().orc_coarse_map(self._h, class_id.encode(), template_id, out.ctypes.data) != 0:
            raise KeyError((class_id, template_id))
        return out

    def fetch(self, stage, level, modality):
        n = lib().orc_debug_fetch(self._h, stage, level, modality, None)
        if n < 0:
            raise ValueError("bad tap")
        buf = np.empty(n, np.uint8)
        lib().orc_debug_fetch(self._h, stage, level, modality, buf.ctypes.data)
        g = self.geometry(level, modality)
        if stage in (Stage.QUANTIZED, Stage.SPREAD, Stage.QUANT_RAW):
            return buf.reshape(g["rows"], g["cols"])
        if stage == Stage.RESPONSE:
            return buf.reshape(8, g["rows"], g["cols"])
        if stage == Stage.LINEAR:
            return buf.reshape(8, g["plane_stride"])
        if stage == Stage.MAGNITUDE:
            return buf.view(np.float32).reshape(g["rows"], g["cols"]) if n else np.zeros((0, 0), np.float32)
        raise ValueError(stage)


def sort_unique(recs):
    """The tail of Detector::match (std::sort + std::unique with Match::operator< / ==), in libstdc++."""
    recs = np.ascontiguousarray(recs, dtype=MATCH_DTYPE).copy()
    n = lib().orc_sort_unique(recs.ctypes.data, len(recs))
    return recs[:n]


# -- primitives (pinned against cv2 in tests/test_oracle_primitives.py)
def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def prim_gaussian7(src):
    src = np.ascontiguousarray(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty_like(src)
    lib().orc_prim_gaussian7(_p(src), src.shape[0], src.shape[1], ch, _p(dst))
    return dst


def prim_sobel3(src):
    src = np.ascontiguousarray(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    dx = np.empty(src.shape, np.int16)
    dy = np.empty(src.shape, np.int16)
    lib().orc_prim_sobel3(_p(src), src.shape[0], src.shape[1], ch, _p(dx), _p(dy))
    return dx, dy


def prim_phase_deg(x, y):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    out = np.empty_like(x)
    lib().orc_prim_phase_deg(_p(x), _p(y), x.size, _p(out))
    return out


def prim_pyrdown(src):
    src = np.ascontiguousarray(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    shp = (src.shape[0] // 2, src.shape[1] // 2) + ((ch,) if src.ndim == 3 else ())
    dst = np.empty(shp, np.uint8)
    lib().orc_prim_pyrdown(_p(src), src.shape[0], src.shape[1], ch, _p(dst))
    return dst


def prim_nn_half(src):
    src = np.ascontiguousarray(src)
    dst = np.empty((src.shape[0] // 2, src.shape[1] // 2), np.uint8)
    lib().orc_prim_nn_half(_p(src), src.shape[0], src.shape[1], _p(dst))
    return dst


def prim_median5(src):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_prim_median5(_p(src), src.shape[0], src.shape[1], _p(dst))
    return dst


def prim_median5_fast(src):
    """The CPU baseline's SSE selection network (linemod_fast.inc) on the same input."""
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_prim_median5_fast(_p(src), src.shape[0], src.shape[1], _p(dst))
    return dst


def prim_erode3(src, iterations=1):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_prim_erode3(_p(src), src.shape[0], src.shape[1], iterations, _p(dst))
    return dst


def prim_distance_c3(src):
    src = np.ascontiguousarray(src)
    dst = np.empty(src.shape, np.float32)
    lib().orc_prim_distance_c3(_p(src), src.shape[0], src.shape[1], _p(dst))
    return dst


def prim_cg_quantize(bgr, weak=10.0):
    bgr = np.ascontiguousarray(bgr)
    r, c = bgr.shape[:2]
    mag = np.empty((r, c), np.float32)
    q = np.empty((r, c), np.uint8)
    ang = np.empty((r, c), np.float32)
    lib().orc_prim_cg_quantize(_p(bgr), r, c, float(weak), _p(mag), _p(q), _p(ang))
    return mag, q, ang


def prim_dn_quantize(det, depth, distance=2000, difference=50):
    """quantizedNormals (raw LUT output, and after the 5x5 median) with `det`'s NORMAL_LUT."""
    depth = np.ascontiguousarray(depth, np.uint16)
    raw = np.empty(depth.shape, np.uint8)
    out = np.empty(depth.shape, np.uint8)
    lib().orc_prim_dn_quantize(det._h, _p(depth), depth.shape[0], depth.shape[1], int(distance), int(difference),
                               _p(raw), _p(out))
    return raw, out


def prim_spread(src, T):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_prim_spread(_p(src), src.shape[0], src.shape[1], T, _p(dst))
    return dst


# ---------------------------------------------------------------------------------------------- render_oracle.cpp
def view_sphere(n_points=150, angle_step=10, radius_min=0.5, radius_max=1.0, radius_step=0.1, angle_min=-80, angle_max=80):
    return OrcViewSphere(n_points, angle_min, angle_max, angle_step, radius_min, radius_max, radius_step)


def view_count(vs):
    return lib().orc_view_count(C.byref(vs))


def view_params(vs, index):
    """-> (T[3], up[3], radius, sphere point, angle)"""
    T, up = np.zeros(3), np.zeros(3)
    st, r = (C.c_int32 * 2)(), C.c_float()
    if lib().orc_view_params(C.byref(vs), index, T.ctypes.data, up.ctypes.data, st, C.byref(r)) != 0:
        raise IndexError(index)
    return T, up, r.value, st[0], st[1]


def view_list(vs):
    """Every view in iteration order -> list of (T[3], up[3], radius, sphere point, angle)"""
    n = view_count(vs)
    T, up = np.zeros((n, 3)), np.zeros((n, 3))
    st, r = np.zeros((n, 2), np.int32), np.zeros(n, np.float32)
    assert lib().orc_view_list(C.byref(vs), T.ctypes.data, up.ctypes.data, st.ctypes.data, r.ctypes.data) == n
    return [(T[i], up[i], float(r[i]), int(st[i, 0]), int(st[i, 1])) for i in range(n)]


def look_at(T, up):
    T, up = np.ascontiguousarray(T, np.float64), np.ascontiguousarray(up, np.float64)
    R, t = np.zeros((3, 3), np.float32), np.zeros(3, np.float32)
    if lib().orc_look_at(T.ctypes.data, up.ctypes.data, R.ctypes.data, t.ctypes.data) != 0:
        raise ValueError("degenerate view")
    return R, t


def camera(width=640, height=480, fx=535.566011, fy=537.168115, near=0.1, far=1000.0):
    return OrcCamera(width, height, fx, fy, near, far)


def render(triangles, cam, T, up):
    """-> (bgr, depth, mask, (x, y, w, h)) of one view, see the specification in render_oracle.cpp"""
    tri = np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
    T, up = np.ascontiguousarray(T, np.float64), np.ascontiguousarray(up, np.float64)
    bgr = np.zeros((cam.height, cam.width, 3), np.uint8)
    depth = np.zeros((cam.height, cam.width), np.uint16)
    mask = np.zeros((cam.height, cam.width), np.uint8)
    rect = (C.c_int32 * 4)()
    if lib().orc_render(tri.ctypes.data, len(tri), C.byref(cam), T.ctypes.data, up.ctypes.data, bgr.ctypes.data,
                        depth.ctypes.data, mask.ctypes.data, rect) != 0:
        raise ValueError("degenerate view")
    return bgr, depth, mask, tuple(rect)


def depth_diff(scene, templ, templ_mask, x, y, tx, ty, w, h):
    """rgbdDetector::depth_diff on a scene ROI at (x, y) and a template crop at (tx, ty), both w x h."""
    scene, templ = np.ascontiguousarray(scene, np.uint16), np.ascontiguousarray(templ, np.uint16)
    templ_mask = np.ascontiguousarray(templ_mask, np.uint8)
    return lib().orc_depth_diff(scene.ctypes.data, scene.shape[1], templ.ctypes.data, templ_mask.ctypes.data, templ.shape[1],
                                x, y, tx, ty, w, h)
