"""Host-side mirror of cv::linemod::Detector over the C ABI (include/linemod_b200.h).

Same names, argument meaning and error behaviour as the surface the reference ROS package drives:

    Detector(modalities, T)            /root/reference/src/renderer.cpp:179-185
    addTemplate(sources, class_id, mask)        src/renderer.cpp:308
    match(sources, threshold, class_ids, masks) src/rgbdDetector.cpp:31-34
    read/readClass via Detector.read(path), write/writeClass via Detector.write(path)
                                                 src/rgbdDetector.cpp:1668-1680, src/renderer.cpp:56-70
    getTemplates / classIds / numTemplates       ..._service.cpp:351, linemod_carmine_detect.cpp:319

All pixels are processed by the CUDA library; this module only marshals numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (HDR_DTYPE, LM_COLOR_GRADIENT, LM_DEPTH_NORMAL, MATCH_DTYPE, RAW_DTYPE, LinemodError, LmImage,
                    LmModalityDesc, LmRect, check, image, image_array, lib)


def ColorGradient(weak_threshold=10.0, num_features=63, strong_threshold=55.0):
    """cv::linemod::ColorGradient (defaults of the default constructor used at src/renderer.cpp:180)."""
    return LmModalityDesc(LM_COLOR_GRADIENT, weak_threshold, strong_threshold, 2000, 50, 2, num_features)


def DepthNormal(distance_threshold=2000, difference_threshold=50, num_features=63, extract_threshold=2):
    """cv::linemod::DepthNormal (defaults of the default constructor used at src/renderer.cpp:181)."""
    return LmModalityDesc(LM_DEPTH_NORMAL, 10.0, 55.0, distance_threshold, difference_threshold, extract_threshold,
                          num_features)


class QuantizedPyramid:
    """cv::linemod::QuantizedPyramid as returned by Modality::process ([OCV] linemod.cpp): quantize(), extractTemplate(),
    pyrDown().  The quantisation of all `levels` ran on the GPU when the object was made."""

    def __init__(self, modality, src, mask=None, levels=4, normal_lut=None):
        self._h = C.c_void_p()
        self.level = 0
        simg, sk = image(src)
        mptr = None
        if mask is not None:
            mimg, mk = image(mask)
            mptr = C.pointer(mimg)
        rows, cols = src.shape[:2]
        while levels > 1 and min(rows >> (levels - 1), cols >> (levels - 1)) < 16:
            levels -= 1
        lut = None
        if normal_lut is not None:
            lut = np.ascontiguousarray(normal_lut, np.uint8).reshape(8000)
        check(lib().lm_modality_process(C.byref(modality), C.byref(simg), mptr, levels, None if lut is None else lut.ctypes.data,
                                        C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().lm_qpyramid_destroy(self._h)
            self._h = None

    def pyrDown(self):
        if self.level + 1 >= lib().lm_qpyramid_levels(self._h):
            raise LinemodError(-1, "pyramid has %d levels" % lib().lm_qpyramid_levels(self._h))
        self.level += 1

    def quantize(self):
        r, c = C.c_int(), C.c_int()
        check(lib().lm_qpyramid_size(self._h, self.level, C.byref(r), C.byref(c)))
        dst = np.zeros((r.value, c.value), np.uint8)
        dimg, dk = image(dst)
        check(lib().lm_qpyramid_quantize(self._h, self.level, C.byref(dimg)))
        return dst

    def extractTemplate(self):
        """-> (ok, (width, height, pyramid_level, features[n,3]))"""
        hdr = np.zeros(1, HDR_DTYPE)
        feats = np.zeros((63, 3), np.int32)
        ok = check(lib().lm_qpyramid_extract(self._h, self.level, hdr.ctypes.data, feats.ctypes.data))
        h = hdr[0]
        return bool(ok), (int(h[0]), int(h[1]), int(h[2]), feats[:int(h[3])].copy())


def process(modality, src, mask=None, levels=4, normal_lut=None):
    """cv::linemod::Modality::process(src, mask) for a ColorGradient() / DepthNormal() descriptor."""
    return QuantizedPyramid(modality, src, mask, levels, normal_lut)


class Stage:
    QUANTIZED, SPREAD, RESPONSE, LINEAR, MAGNITUDE, QUANT_RAW, LINEAR_PACKED = range(7)


class Detector:
    """B200-native drop-in for cv::linemod::Detector."""

    def __init__(self, modalities=None, T=(5, 8), _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
            return
        if modalities is None:
            modalities = [ColorGradient(), DepthNormal()]
        Ta = (C.c_int32 * len(T))(*T)
        Ma = (LmModalityDesc * len(modalities))(*modalities)
        check(lib().lm_create(Ta, len(T), Ma, len(modalities), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().lm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ persistence
    @classmethod
    def read(cls, path):
        """readLinemod(): Detector::read(fs.root()) + readClass for every entry of "classes"."""
        h = C.c_void_p()
        check(lib().lm_create_from_yaml(str(path).encode(), C.byref(h)))
        return cls(_handle=h)

    def write(self, path):
        """writeLinemod(): Detector::write + writeClass per class, single file."""
        check(lib().lm_write_yaml(self._h, str(path).encode()))

    @classmethod
    def read_cache(cls, path):
        """Binary template cache written by write_cache(): the same detector without a YAML parse."""
        h = C.c_void_p()
        check(lib().lm_create_from_cache(str(path).encode(), C.byref(h)))
        return cls(_handle=h)

    def write_cache(self, path):
        check(lib().lm_write_cache(self._h, str(path).encode()))

    def readClasses(self, class_ids, fmt="templates_%s.yml.gz"):
        ids = [c.encode() for c in class_ids]
        arr = (C.c_char_p * max(1, len(ids)))(*ids)
        check(lib().lm_read_classes(self._h, arr, len(ids), fmt.encode()))

    def writeClasses(self, fmt="templates_%s.yml.gz"):
        check(lib().lm_write_classes(self._h, fmt.encode()))

    # ------------------------------------------------------------------ introspection
    def pyramidLevels(self):
        return lib().lm_pyramid_levels(self._h)

    def getT(self, level):
        return check(lib().lm_get_T(self._h, level))

    def getModalities(self):
        out = []
        for m in range(lib().lm_num_modalities(self._h)):
            d = LmModalityDesc()
            check(lib().lm_get_modality(self._h, m, C.byref(d)))
            out.append(d)
        return out

    def numClasses(self):
        return lib().lm_num_classes(self._h)

    def numTemplates(self, class_id=None):
        return lib().lm_num_templates(self._h, class_id.encode() if class_id is not None else None)

    def classIds(self):
        return [lib().lm_class_id(self._h, i).decode() for i in range(self.numClasses())]

    def getTemplates(self, class_id, template_id):
        """-> list (index l*M+m) of (width, height, pyramid_level, features[n,3] int32 (x, y, label))."""
        n_t = self.pyramidLevels() * lib().lm_num_modalities(self._h)
        hdr = np.zeros(n_t, HDR_DTYPE)
        total = check(lib().lm_get_templates(self._h, class_id.encode(), template_id, hdr.ctypes.data, None))
        feats = np.zeros((max(total, 1), 3), np.int32)
        check(lib().lm_get_templates(self._h, class_id.encode(), template_id, hdr.ctypes.data, feats.ctypes.data))
        out, k = [], 0
        for i in range(n_t):
            nf = int(hdr[i]["num_features"])
            out.append((int(hdr[i]["width"]), int(hdr[i]["height"]), int(hdr[i]["pyramid_level"]), feats[k:k + nf].copy()))
            k += nf
        return out

    # ------------------------------------------------------------------ training
    def addTemplate(self, sources, class_id, object_mask=None):
        """-> (template_id, (x, y, w, h)); template_id == -1 when a level lacks features (like the reference)."""
        arr, keep = image_array(sources)
        bb = LmRect()
        mptr = None
        if object_mask is not None:
            mimg, mkeep = image(object_mask)
            mptr = C.pointer(mimg)
        r = lib().lm_add_template(self._h, arr, len(sources), class_id.encode(), mptr, C.byref(bb))
        if r < -1:
            raise LinemodError(r + 100, _capi.last_error())
        return r, (bb.x, bb.y, bb.width, bb.height)

    def addTemplates(self, views, class_id):
        """Batched addTemplate on the GPU: views = [(sources, object_mask)] -> (template_ids, bounding_boxes)."""
        from . import training
        return training.add_templates_batch(self, views, class_id)

    def trainViews(self, mesh, cam, T, up, class_id, centre_depth=False):
        """Render + addTemplate for many views (renderer.cpp:239-329) -> (template_ids, bounding_boxes, mask_rects
        [, centre depth])."""
        from . import training
        return training.train_views(self, mesh, cam, T, up, class_id, centre_depth)

    def addTemplateFromQuantized(self, quantized, magnitudes, class_id, object_mask=None):
        """Host half of addTemplate: quantized[l*M+m] u8 maps, magnitudes[l*M+m] f32 maps (None for DepthNormal)."""
        arr, keep = image_array(quantized)
        mags = [None if m is None else np.ascontiguousarray(m, np.float32) for m in magnitudes]
        mp = (C.c_void_p * len(mags))(*[None if m is None else m.ctypes.data for m in mags])
        bb = LmRect()
        mptr = None
        if object_mask is not None:
            mimg, mkeep = image(object_mask)
            mptr = C.pointer(mimg)
        r = lib().lm_add_template_from_quantized(self._h, arr, mp, class_id.encode(), mptr, C.byref(bb))
        if r < -1:
            raise LinemodError(r, _capi.last_error())
        return r, (bb.x, bb.y, bb.width, bb.height)

    def addSyntheticTemplate(self, templates, class_id):
        """templates: list (L*M) of (width, height, pyramid_level, features[n,3])."""
        hdr = np.zeros(len(templates), HDR_DTYPE)
        for i, t in enumerate(templates):
            hdr[i] = (t[0], t[1], t[2], len(t[3]))
        feats = [np.asarray(t[3], np.int32).reshape(-1, 3) for t in templates]
        feats = np.ascontiguousarray(np.concatenate(feats) if feats else np.zeros((0, 3), np.int32), dtype=np.int32)
        if feats.size == 0:
            feats = np.zeros((1, 3), np.int32)
        return check(lib().lm_add_synthetic_template(self._h, class_id.encode(), len(templates), hdr.ctypes.data,
                                                     feats.ctypes.data))

    # ------------------------------------------------------------------ matching
    @staticmethod
    def _ids(class_ids):
        ids = [c.encode() for c in class_ids]
        return (C.c_char_p * max(1, len(ids)))(*ids), len(ids)

    @staticmethod
    def _take(out, n):
        res = np.empty(n, MATCH_DTYPE)
        if n:
            C.memmove(res.ctypes.data, out, n * MATCH_DTYPE.itemsize)
        lib().lm_free_matches(out)
        return res

    def match(self, sources, threshold, class_ids=(), masks=(), quantized_images=False):
        """Detector::match.  Returns a structured array (x, y, template_id, class_index, similarity), or
        (matches, [quantised images index l*M+m]) when quantized_images is True."""
        arr, keep = image_array(sources)
        marr, mkeep = image_array(masks)
        ids, n_ids = self._ids(class_ids)
        out, n = C.c_void_p(), C.c_size_t()
        qarr, qimgs = None, None
        if quantized_images:
            L, M = self.pyramidLevels(), lib().lm_num_modalities(self._h)
            rows, cols = sources[0].shape[:2]
            qimgs = [np.zeros((rows >> l, cols >> l), np.uint8) for l in range(L) for _ in range(M)]
            qarr = (LmImage * (L * M))(*[image(q)[0] for q in qimgs])
        check(lib().lm_match(self._h, arr, len(sources), threshold, ids, n_ids, marr, len(masks), qarr, C.byref(out),
                             C.byref(n)))
        res = self._take(out, n.value)
        return (res, qimgs) if quantized_images else res

    def match_multi(self, sources, queries, masks=()):
        """Several (threshold, class_ids) queries from one front end of the frame (lm_match_multi).
        queries: [(threshold, [class ids])].  -> list of match arrays, one per query."""
        arr, keep = image_array(sources)
        marr, mkeep = image_array(masks)
        qarr, qkeep = _capi.query_array(queries)
        out = C.c_void_p()
        offs = (C.c_size_t * (len(queries) + 1))()
        check(lib().lm_match_multi(self._h, arr, len(sources), qarr, len(queries), marr, len(masks), None,
                                   C.byref(out), offs))
        allm = self._take(out, offs[len(queries)])
        return [allm[offs[i]:offs[i + 1]] for i in range(len(queries))]

    def match_device_multi(self, d_ptrs, rows, cols, queries, stream=0):
        """Device-resident sources, several queries, asynchronous on `stream`.  -> (device pointer of the record block,
        capacity in bytes); a record's query index is order_key >> 28."""
        qarr, qkeep = _capi.query_array(queries)
        ptrs = (C.c_void_p * len(d_ptrs))(*d_ptrs)
        rec, cap = C.c_void_p(), C.c_size_t()
        check(lib().lm_match_device_multi(self._h, ptrs, len(d_ptrs), rows, cols, qarr, len(queries),
                                          C.c_void_p(stream), C.byref(rec), C.byref(cap)))
        return rec.value, cap.value

    def match_batch(self, frames, threshold, class_ids=()):
        """frames: list of per-frame source lists.  -> list of match arrays (chunks of "batch_frames" frames per launch set, chunks pipelined)."""
        flat = [s for f in frames for s in f]
        arr, keep = image_array(flat)
        ids, n_ids = self._ids(class_ids)
        out = C.c_void_p()
        offs = (C.c_size_t * (len(frames) + 1))()
        check(lib().lm_match_batch(self._h, arr, len(frames), len(frames[0]) if frames else 0, threshold, ids, n_ids,
                                   C.byref(out), offs))
        allm = self._take(out, offs[len(frames)])
        return [allm[offs[i]:offs[i + 1]] for i in range(len(frames))]

    def match_batch_multi(self, frames, queries):
        """frames: list of per-frame source lists; queries: [(threshold, [class ids])].
        -> list (per frame) of lists (per query) of match arrays; chunks of frames per launch set, chunks pipelined."""
        flat = [s for f in frames for s in f]
        arr, keep = image_array(flat)
        qarr, qkeep = _capi.query_array(queries)
        out = C.c_void_p()
        n_q = len(queries)
        offs = (C.c_size_t * (len(frames) * n_q + 1))()
        check(lib().lm_match_batch_multi(self._h, arr, len(frames), len(frames[0]) if frames else 0, qarr, n_q,
                                         C.byref(out), offs))
        allm = self._take(out, offs[len(frames) * n_q])
        return [[allm[offs[f * n_q + q]:offs[f * n_q + q + 1]] for q in range(n_q)] for f in range(len(frames))]

    def open_stream(self, queries):
        """A continuous stream of host frames answered for `queries` [(threshold, [class ids])]: see FrameStream."""
        return FrameStream(self, queries)

    def match_device(self, d_ptrs, rows, cols, threshold, stream=0, class_ids=()):
        """Device-resident sources (tightly packed), asynchronous on `stream`.  -> (device pointer of the record
        block {count, capacity, overflow, n_cands} + raw records, capacity in bytes)."""
        ids, n_ids = self._ids(class_ids)
        ptrs = (C.c_void_p * len(d_ptrs))(*d_ptrs)
        rec, cap = C.c_void_p(), C.c_size_t()
        check(lib().lm_match_device(self._h, ptrs, len(d_ptrs), rows, cols, threshold, ids, n_ids,
                                    C.c_void_p(stream), C.byref(rec), C.byref(cap)))
        return rec.value, cap.value

    def finalize_raw(self, raw):
        raw = np.ascontiguousarray(raw, dtype=RAW_DTYPE)
        out, n = C.c_void_p(), C.c_size_t()
        check(lib().lm_finalize_raw(self._h, raw.ctypes.data, len(raw), C.byref(out), C.byref(n)))
        return self._take(out, n.value)

    @staticmethod
    def cluster_matches(matches, obj_origin_dists, rects, vote_step, radius_min, radius_step, cluster_threshold=2,
                        iou_threshold=0.4):
        """rcd_voting -> cluster_filter -> mean-similarity score -> IoU non-maximum suppression (lm_cluster_matches), the
        stage the reference runs right behind Detector::match.  rects: int32 [n_templates, 4] (x, y, width, height).
        -> list of dicts {index, score, rect, matches (indices into `matches`)} in the reference's order."""
        m = np.ascontiguousarray(matches, dtype=MATCH_DTYPE)
        dists = np.ascontiguousarray(obj_origin_dists, dtype=np.float64)
        r = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
        assert len(dists) == len(r)

        class Params(C.Structure):
            _fields_ = [("vote_row_col_step", C.c_int32), ("renderer_radius_min", C.c_double),
                        ("renderer_radius_step", C.c_double), ("cluster_threshold", C.c_int32),
                        ("iou_threshold", C.c_double)]
        CL = np.dtype([("index", "<i4", (3,)), ("score", "<f8"), ("rect", "<i4", (4,)), ("first", "<u4"), ("count", "<u4")],
                      align=True)
        p = Params(vote_step, radius_min, radius_step, cluster_threshold, iou_threshold)
        oc, oi, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(lib().lm_cluster_matches(m.ctypes.data, len(m), dists.ctypes.data, r.ctypes.data, len(r), C.byref(p),
                                       C.byref(oc), C.byref(n), C.byref(oi)))
        out = []
        if n.value:
            cl = np.frombuffer((C.c_char * (n.value * CL.itemsize)).from_address(oc.value), dtype=CL).copy()
            total = int(cl["first"][-1] + cl["count"][-1])
            idx = np.frombuffer((C.c_char * (total * 4)).from_address(oi.value), dtype=np.uint32).copy()
            for c in cl:
                out.append(dict(index=tuple(int(v) for v in c["index"]), score=float(c["score"]),
                                rect=tuple(int(v) for v in c["rect"]),
                                matches=[int(v) for v in idx[c["first"]:c["first"] + c["count"]]]))
        lib().lm_free_clusters(oc, oi)
        return out

    def set_shard(self, rank, world):
        check(lib().lm_set_shard(self._h, rank, world))

    # ------------------------------------------------------------------ tables / options / taps
    def set_similarity_lut(self, lut):
        lut = np.ascontiguousarray(lut, np.uint8)
        assert lut.size == 256
        check(lib().lm_set_similarity_lut(self._h, lut.ctypes.data))

    def similarity_lut(self):
        out = np.empty(256, np.uint8)
        check(lib().lm_get_similarity_lut(self._h, out.ctypes.data))
        return out

    def set_normal_lut(self, lut):
        lut = np.ascontiguousarray(lut, np.uint8)
        assert lut.size == 8000
        check(lib().lm_set_normal_lut(self._h, lut.ctypes.data))

    def load_normal_lut_file(self, path):
        """NORMAL_LUT from OpenCV's own text file (modules/objdetect/src/normal_lut.i)."""
        check(lib().lm_load_normal_lut_file(self._h, str(path).encode()))

    def normal_lut(self):
        out = np.empty(8000, np.uint8)
        check(lib().lm_get_normal_lut(self._h, out.ctypes.data))
        return out

    def set_option(self, key, value):
        check(lib().lm_set_option(self._h, key.encode(), int(value)))

    def build_front(self, sources, masks=()):
        arr, keep = image_array(sources)
        marr, mkeep = image_array(masks)
        check(lib().lm_build_front(self._h, arr, len(sources), marr, len(masks)))

    def geometry(self, level):
        g = (C.c_int32 * 5)()
        ps = C.c_size_t()
        check(lib().lm_level_geometry(self._h, level, g, C.byref(ps)))
        return dict(rows=g[0], cols=g[1], T=g[2], W=g[3], H=g[4], plane_stride=ps.value)

    def fetch(self, stage, level, modality):
        n = check(lib().lm_debug_fetch(self._h, stage, level, modality, None))
        buf = np.empty(n, np.uint8)
        check(lib().lm_debug_fetch(self._h, stage, level, modality, buf.ctypes.data))
        g = self.geometry(level)
        if stage in (Stage.QUANTIZED, Stage.SPREAD, Stage.QUANT_RAW):
            return buf.reshape(g["rows"], g["cols"])
        if stage == Stage.RESPONSE:
            return buf.reshape(8, g["rows"], g["cols"])
        if stage == Stage.LINEAR:
            return buf.reshape(8, g["plane_stride"])
        if stage == Stage.LINEAR_PACKED:
            return buf.reshape(8, g["plane_stride"] // 2)
        return buf.view(np.float32).reshape(g["rows"], g["cols"])

    def coarse_map(self, class_id, template_id):
        g = self.geometry(self.pyramidLevels() - 1)
        out = np.zeros((g["H"], g["W"]), np.uint16)
        check(lib().lm_debug_coarse_map(self._h, class_id.encode(), template_id, out.ctypes.data))
        return out

    def last_presort(self):
        n = lib().lm_debug_presort(self._h, None)
        res = np.empty(n, MATCH_DTYPE)
        lib().lm_debug_presort(self._h, res.ctypes.data)
        return res

    def last_timings(self):
        ms = (C.c_float * 5)()
        k = C.c_int()
        check(lib().lm_last_timings(self._h, ms, C.byref(k)))
        return dict(h2d=ms[0], front=ms[1], coarse=ms[2], refine=ms[3], d2h=ms[4], launches=k.value)

    def last_work(self):
        w = (C.c_uint64 * 8)()
        check(lib().lm_last_work(self._h, w))
        return dict(B_front=w[0], B_coarse=w[1], B_refine=w[2], B_out=w[3], candidates=w[4], evals=w[5],
                    B_coarse_gathered=w[6], frames=w[7])


class DetectorGroup:
    """Several GPUs behind one caller (lm_group): the prototype detector cloned onto `devices`, one worker thread per
    device.  mode "frames": every device holds all templates and takes its share of the frames of a batch over its own PCIe
    link (no exchange between devices); mode "templates": templates sharded by canonical index, every device sees every
    frame, survivors merged before the reference's sort + unique.  Results equal the prototype's in both modes."""
    FRAMES, TEMPLATES = 0, 1

    def __init__(self, prototype, devices, mode="frames", template_shards=None):
        """mode "frames" / "templates", or template_shards = S for the 2-D grid: len(devices) / S sets of S shards."""
        self._h = C.c_void_p()
        self.proto = prototype
        dv = (C.c_int * len(devices))(*devices)
        if template_shards is not None:
            check(lib().lm_group_create_grid(prototype._h, dv, len(devices), int(template_shards), C.byref(self._h)))
            return
        m = {"frames": self.FRAMES, "templates": self.TEMPLATES}[mode]
        check(lib().lm_group_create(prototype._h, dv, len(devices), m, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().lm_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().lm_group_size(self._h)

    def set_option(self, key, value):
        check(lib().lm_group_set_option(self._h, key.encode(), int(value)))

    def match_batch_multi(self, frames, queries):
        """frames: list of per-frame source lists; queries: [(threshold, [class ids])] -> list (per frame) of lists (per query)."""
        flat = [s for f in frames for s in f]
        arr, keep = image_array(flat)
        qarr, qkeep = _capi.query_array(queries)
        out = C.c_void_p()
        n_q = len(queries)
        offs = (C.c_size_t * (len(frames) * n_q + 1))()
        check(lib().lm_group_match_batch_multi(self._h, arr, len(frames), len(frames[0]) if frames else 0, qarr, n_q,
                                               C.byref(out), offs))
        allm = self.proto._take(out, offs[len(frames) * n_q])
        return [[allm[offs[f * n_q + q]:offs[f * n_q + q + 1]] for q in range(n_q)] for f in range(len(frames))]

    def match(self, sources, threshold, class_ids=()):
        return self.match_batch_multi([sources], [(threshold, list(class_ids))])[0][0]


class FrameStream:
    """lm_stream: the chunked pipeline of match_batch_multi kept alive between calls (the device never drains at a call
    boundary).  push() enqueues frames and returns; pop() hands out finished frames in push order as
    [per frame [per query match array]].  The arrays passed to push() must stay unchanged until pop() has returned their
    frames.  While the stream is open the detector refuses other matching calls."""

    def __init__(self, det, queries):
        self._det = det
        self._n_q = len(queries)
        qarr, self._qkeep = _capi.query_array(queries)
        self._h = C.c_void_p()
        check(lib().lm_stream_open(det._h, qarr, self._n_q, C.byref(self._h)))
        self._keep = []   # (frames still in flight, image array + numpy buffers) per push

    def push(self, frames):
        flat = [s for f in frames for s in f]
        arr, keep = image_array(flat)
        self._keep.append([len(frames), (arr, keep, flat)])
        check(lib().lm_stream_push(self._h, arr, len(frames), len(frames[0]) if frames else 0))

    def push_array(self, arr, n_frames, n_sources):
        """A prebuilt lm_image array (image_array of pinned buffers the caller keeps alive): no per-call marshalling."""
        check(lib().lm_stream_push(self._h, arr, n_frames, n_sources))

    def in_flight(self):
        return lib().lm_stream_in_flight(self._h)

    def pop(self, wait_all=False, max_frames=None):
        cap = max(1, self.in_flight() if max_frames is None else max_frames)
        out, n = C.c_void_p(), C.c_int()
        offs = (C.c_size_t * (cap * self._n_q + 1))()
        check(lib().lm_stream_pop(self._h, 1 if wait_all else 0, cap, C.byref(out), offs, C.byref(n)))
        allm = self._det._take(out, offs[n.value * self._n_q])
        left = n.value
        while left > 0 and self._keep:
            take = min(left, self._keep[0][0])
            self._keep[0][0] -= take
            left -= take
            if self._keep[0][0] == 0:
                self._keep.pop(0)
        return [[allm[offs[f * self._n_q + q]:offs[f * self._n_q + q + 1]] for q in range(self._n_q)] for f in range(n.value)]

    def close(self):
        if self._h:
            lib().lm_stream_close(self._h)
            self._h = C.c_void_p()
            self._keep = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
