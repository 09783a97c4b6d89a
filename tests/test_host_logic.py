"""Host-side logic of the product library that needs no GPU: templates.yml persistence (against cv2.FileStorage as the
format oracle), the host half of addTemplate (against the CPU oracle on identical quantised maps), and the final
ordering stage."""
import gzip
import os

import numpy as np
import pytest

import common
from common import O, synth
from linemod_pose_estimation_b200 import Detector, LinemodError, MATCH_DTYPE, RAW_DTYPE, ColorGradient, DepthNormal


def _random_detector(seed=0, classes=("obj",), n=5, T=(5, 8), kinds=("cg", "dn")):
    det = Detector(common.product_modalities(kinds), T)
    rng = np.random.default_rng(seed)
    for cid in classes:
        for _ in range(n):
            det.addSyntheticTemplate(synth.random_pyramid(rng, T=T, M=len(kinds)), cid)
    return det


def _templates_equal(a, b):
    assert a.classIds() == b.classIds()
    for cid in a.classIds():
        assert a.numTemplates(cid) == b.numTemplates(cid)
        for tid in range(a.numTemplates(cid)):
            for (ta, tb) in zip(a.getTemplates(cid, tid), b.getTemplates(cid, tid)):
                assert ta[:3] == tb[:3] and np.array_equal(ta[3], tb[3])


# ---------------------------------------------------------------------------------------------- persistence
def test_yaml_roundtrip_and_layout(tmp_path):
    det = _random_detector(1, classes=("memoryChip2", "cpu_binary"))
    p = tmp_path / "templates.yml"
    det.write(p)
    text = p.read_text()
    lines = text.splitlines()
    assert lines[0] == "%YAML:1.0" and lines[1] == "pyramid_levels: 2" and lines[2] == "T: [ 5, 8 ]"
    assert "      weak_threshold: 10." in lines and "      strong_threshold: 55." in lines
    assert "      modalities: [ ColorGradient, DepthNormal ]" in lines
    assert "---" not in lines                       # OpenCV 2.4 layout
    back = Detector.read(p)
    _templates_equal(det, back)
    assert back.classIds() == ["cpu_binary", "memoryChip2"]  # std::map order
    assert [back.getT(0), back.getT(1)] == [5, 8]
    mods = back.getModalities()
    assert mods[0].type == 0 and mods[0].weak_threshold == 10.0 and mods[1].type == 1 and mods[1].extract_threshold == 2
    p2 = tmp_path / "again.yml"
    back.write(p2)
    assert p2.read_text() == text


def test_yaml_is_readable_by_opencv_filestorage(tmp_path):
    cv2 = pytest.importorskip("cv2")
    det = _random_detector(2)
    p = str(tmp_path / "t.yml")
    det.write(p)
    fs = cv2.FileStorage(p, cv2.FILE_STORAGE_READ)
    assert int(fs.getNode("pyramid_levels").real()) == 2
    T = fs.getNode("T")
    assert [int(T.at(i).real()) for i in range(T.size())] == [5, 8]
    cls = fs.getNode("classes").at(0)
    assert cls.getNode("class_id").string() == "obj"
    tp0 = cls.getNode("template_pyramids").at(0)
    assert int(tp0.getNode("template_id").real()) == 0
    t0 = tp0.getNode("templates").at(0)
    want = det.getTemplates("obj", 0)[0]
    assert int(t0.getNode("width").real()) == want[0]
    f0 = t0.getNode("features").at(0)
    assert [int(f0.at(i).real()) for i in range(3)] == list(want[3][0])
    fs.release()


def test_reads_files_written_by_opencv_filestorage(tmp_path):
    """cv2 4.x emits the '---' document marker and its own float formatting; the reader must accept both forms."""
    cv2 = pytest.importorskip("cv2")
    p = str(tmp_path / "cv.yml")
    fs = cv2.FileStorage(p, cv2.FILE_STORAGE_WRITE)
    fs.write("pyramid_levels", 2)
    fs.startWriteStruct("T", cv2.FILE_NODE_SEQ | cv2.FILE_NODE_FLOW)
    fs.write("", 5); fs.write("", 8)
    fs.endWriteStruct()
    fs.startWriteStruct("modalities", cv2.FILE_NODE_SEQ)
    fs.startWriteStruct("", cv2.FILE_NODE_MAP)
    fs.write("type", "ColorGradient"); fs.write("weak_threshold", 12.5); fs.write("num_features", 63); fs.write("strong_threshold", 55.0)
    fs.endWriteStruct()
    fs.endWriteStruct()
    fs.startWriteStruct("classes", cv2.FILE_NODE_SEQ)
    fs.startWriteStruct("", cv2.FILE_NODE_MAP)
    fs.write("class_id", "obj")
    fs.startWriteStruct("modalities", cv2.FILE_NODE_SEQ | cv2.FILE_NODE_FLOW); fs.write("", "ColorGradient"); fs.endWriteStruct()
    fs.write("pyramid_levels", 2)
    fs.startWriteStruct("template_pyramids", cv2.FILE_NODE_SEQ)
    for tid in range(2):
        fs.startWriteStruct("", cv2.FILE_NODE_MAP)
        fs.write("template_id", tid)
        fs.startWriteStruct("templates", cv2.FILE_NODE_SEQ)
        for lvl in range(2):
            fs.startWriteStruct("", cv2.FILE_NODE_MAP)
            fs.write("width", 100 >> lvl); fs.write("height", 80 >> lvl); fs.write("pyramid_level", lvl)
            fs.startWriteStruct("features", cv2.FILE_NODE_SEQ)
            for k in range(3):
                fs.startWriteStruct("", cv2.FILE_NODE_SEQ | cv2.FILE_NODE_FLOW)
                fs.write("", 10 * k + tid); fs.write("", 7 * k); fs.write("", (k + lvl) % 8)
                fs.endWriteStruct()
            fs.endWriteStruct()
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.endWriteStruct()
    fs.endWriteStruct()
    fs.endWriteStruct()
    fs.endWriteStruct()
    fs.release()
    assert "---" in open(p).read()
    det = Detector.read(p)
    assert det.classIds() == ["obj"] and det.numTemplates("obj") == 2
    assert det.getModalities()[0].weak_threshold == 12.5
    t = det.getTemplates("obj", 1)
    assert t[1][:3] == (50, 40, 1) and list(t[1][3][2]) == [21, 14, 3]


def test_read_errors_follow_reference_asserts(tmp_path):
    det = _random_detector(3)
    p = tmp_path / "t.yml"
    det.write(p)
    text = p.read_text()
    (tmp_path / "dup.yml").write_text(text + text[text.index("   -\n      class_id"):].replace("classes:\n", ""))
    with pytest.raises(LinemodError) as e:   # "Detector should not already have this class"
        Detector.read(tmp_path / "dup.yml")
    assert e.value.code == -3 and "already has class" in str(e.value)
    (tmp_path / "badid.yml").write_text(text.replace("template_id: 1", "template_id: 7", 1))
    with pytest.raises(LinemodError):        # CV_Assert(template_id == expected_id)
        Detector.read(tmp_path / "badid.yml")
    (tmp_path / "badmod.yml").write_text(text.replace("modalities: [ ColorGradient, DepthNormal ]", "modalities: [ DepthNormal, ColorGradient ]"))
    with pytest.raises(LinemodError):        # CV_Assert(modalities[i]->name() == ...)
        Detector.read(tmp_path / "badmod.yml")
    with pytest.raises(LinemodError):
        Detector.read(tmp_path / "does_not_exist.yml")


def test_malformed_template_files_end_in_io_errors(tmp_path):
    """Templates from disk are validated before they reach the packer and the kernels: labels outside 0..7, coordinates
    beyond +-4095, more than 63 features, pyramids of the wrong size -- in templates.yml, in class files and in the binary
    cache (valid checksum, bad contents) -- are LM_E_IO, not device faults."""
    import re
    import struct
    det = _random_detector(3)
    p = tmp_path / "t.yml"
    det.write(p)
    text = p.read_text()
    m = re.search(r"- \[ (-?\d+), (-?\d+), (\d) \]", text)
    assert m, "feature triple not found in the written file"

    def variant(name, new_triple):
        (tmp_path / name).write_text(text[:m.start()] + "- [ %d, %d, %d ]" % new_triple + text[m.end():])
        return tmp_path / name
    for name, triple, what in (("label.yml", (3, 4, 9), "label"), ("coord.yml", (70000, 4, 1), "coordinate"),
                               ("neg.yml", (3, -5000, 1), "coordinate")):
        with pytest.raises(LinemodError) as e:
            Detector.read(variant(name, triple))
        assert e.value.code == -3 and what in str(e.value), str(e.value)
    # too many features: repeat one template's feature list until it exceeds 63
    feats = re.search(r"features:\n((?: +- \[[^\n]*\n)+)", text)
    (tmp_path / "many.yml").write_text(text[:feats.end()] + feats.group(1) * 3 + text[feats.end():])
    with pytest.raises(LinemodError) as e:
        Detector.read(tmp_path / "many.yml")
    assert e.value.code == -3 and "63" in str(e.value)
    # class files take the same path
    fmt = str(tmp_path / "templates_%s.yml")
    det.writeClasses(fmt)
    cls = open(fmt % "obj").read()
    m2 = re.search(r"- \[ (-?\d+), (-?\d+), (\d) \]", cls)
    open(fmt % "bad", "w").write((cls[:m2.start()] + "- [ 1, 2, 8 ]" + cls[m2.end():]).replace("class_id: obj", "class_id: bad"))
    with pytest.raises(LinemodError) as e:
        Detector().readClasses(["bad"], fmt)
    assert e.value.code == -3 and "label" in str(e.value)
    # binary cache with a consistent checksum but a label of 200 / a template count the payload cannot hold
    cache = str(tmp_path / "t.lmb2")
    det.write_cache(cache)
    blob = bytearray(open(cache, "rb").read())

    def fnv1a(b):
        h = 0xcbf29ce484222325
        for x in b:
            h = ((h ^ x) * 0x100000001b3) & 0xffffffffffffffff
        return h
    pay = blob[40:]
    off = 2 * 4 + 2 * 28 + 4 + 3 + 4        # T[2], two modality descriptors, class id "obj", template count
    assert struct.unpack_from("<i", pay, off + 12)[0] == 63   # first template header: width, height, level, num_features
    bad = bytearray(pay)
    bad[off + 16 + 4] = 200                  # first feature: x i16, y i16, label u8
    open(str(tmp_path / "label.lmb2"), "wb").write(bytes(blob[:32]) + struct.pack("<Q", fnv1a(bad)) + bytes(bad))
    with pytest.raises(LinemodError) as e:
        Detector.read_cache(str(tmp_path / "label.lmb2"))
    assert e.value.code == -3 and "label" in str(e.value)
    huge = bytearray(pay)
    struct.pack_into("<I", huge, off - 4, 0x7fffffff)
    open(str(tmp_path / "count.lmb2"), "wb").write(bytes(blob[:32]) + struct.pack("<Q", fnv1a(huge)) + bytes(huge))
    with pytest.raises(LinemodError) as e:
        Detector.read_cache(str(tmp_path / "count.lmb2"))
    assert e.value.code == -3


def test_similarity_lut_is_the_literal_upstream_table():
    """The default SIMILARITY_LUT of the product and of the oracle is the literal 256-entry table of OpenCV 2.4's
    linemod.cpp (tests/golden/similarity_lut_ocv.txt); SURVEY A.5's wrap-around formula is NOT it (6 (i, j) pairs differ)."""
    text = open(os.path.join(common.GOLDEN, "similarity_lut_ocv.txt")).read()
    vals = [int(x) for line in text.splitlines() if not line.startswith("#") for x in line.replace(",", " ").split()]
    want = np.array(vals, np.uint8)
    assert want.size == 256 and want.max() == 4
    assert np.array_equal(Detector().similarity_lut(), want)
    assert np.array_equal(O.OracleDetector().similarity_lut(), want)
    other = common.survey_similarity_lut()
    pairs = {(k // 32, 4 * ((k // 16) % 2) + b) for k in np.flatnonzero(other != want) for b in range(4) if (k % 16) == (1 << b)}
    assert len(pairs) == 6 and all(i - j >= 4 for i, j in pairs)   # orientation i four or more bins above bit j


def test_normal_lut_loader_parses_opencv_text(tmp_path):
    """lm_load_normal_lut_file reads OpenCV's normal_lut.i layout (brace-initialised [20][20][20], comments, dimensions
    in the declaration) and rejects files with a wrong entry count or entries above 255."""
    rng = np.random.default_rng(3)
    lut = (1 << rng.integers(0, 8, 8000)).astype(np.uint8)
    body = ",\n".join("{" + ", ".join("{" + ", ".join(str(int(v)) for v in lut[(a * 20 + b) * 20:(a * 20 + b) * 20 + 20]) + "}"
                                      for b in range(20)) + "}" for a in range(20))
    text = "// generated\nstatic unsigned char NORMAL_LUT[20][20][20] = {\n" + body + "\n}; /* 8000 entries */\n"
    p = tmp_path / "normal_lut.i"
    p.write_text(text)
    det = Detector()
    det.load_normal_lut_file(p)
    assert np.array_equal(det.normal_lut(), lut)
    (tmp_path / "short.i").write_text(text.replace("{" + ", ".join(str(int(v)) for v in lut[:20]) + "}", "{1, 2}", 1))
    (tmp_path / "big.i").write_text(text.replace("= {\n{{", "= {\n{{ 300, ", 1))
    (tmp_path / "nobrace.i").write_text("1, 2, 3")
    for name in ("short.i", "big.i", "nobrace.i", "missing.i"):
        with pytest.raises(LinemodError) as e:
            det.load_normal_lut_file(tmp_path / name)
        assert e.value.code == -3


def test_read_write_classes_gz(tmp_path):
    det = _random_detector(4, classes=("a", "b"))
    fmt = str(tmp_path / "templates_%s.yml.gz")
    det.writeClasses(fmt)
    raw = gzip.open(fmt % "a", "rt").read()
    assert raw.startswith("%YAML:1.0\nclass_id: a\nmodalities: [ ColorGradient, DepthNormal ]\npyramid_levels: 2\ntemplate_pyramids:\n")
    fresh = Detector()
    fresh.readClasses(["b", "a"], fmt)
    _templates_equal(det, fresh)


def test_renderer_params_yaml_parses(tmp_path):
    """The pose table the reference writes next to templates.yml (src/renderer.cpp:72-123): keys with spaces,
    !!opencv-matrix tags, wrapped flow sequences.  Parsed with the same reader through a tiny templates file trick:
    the library only exposes templates persistence, so the layout is checked via cv2 where available."""
    cv2 = pytest.importorskip("cv2")
    ref = "/root/reference/config/data/boxNew_longDistance_linemod_xtion_renderer_params.yml"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present (GPU box)")
    fs = cv2.FileStorage(ref, cv2.FILE_STORAGE_READ)
    n = 0
    while not fs.getNode("Template %d" % n).empty():
        n += 1
    assert n == 2652 and int(fs.getNode("renderer_width").real()) == 640


# ---------------------------------------------------------------------------------------------- addTemplate (host half)
@pytest.mark.parametrize("kinds", [("cg", "dn"), ("cg",)])
def test_extract_template_matches_oracle(kinds):
    """Same quantised inputs (computed by the oracle's primitives) -> product host extraction == oracle addTemplate."""
    T = (5, 8)
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)
    n_ok = 0
    for vi, (bgr, depth, mask) in enumerate(common.rendered_views(10, 17, canvas=(200, 220))):
        want_tid, want_bb = orc.add_template(common.sources_for(kinds, bgr, depth), "obj", mask)
        q, mags = [], []
        levels = [bgr]
        levels.append(O.prim_pyrdown(bgr))
        dn0 = O.prim_dn_quantize(orc, depth)[1]
        dn = [dn0, O.prim_nn_half(dn0)]
        for l in range(2):
            for k in kinds:
                if k == "cg":
                    mag, qq, _ = O.prim_cg_quantize(levels[l], 10.0)
                    q.append(qq); mags.append(mag)
                else:
                    q.append(dn[l]); mags.append(None)
        got_tid, got_bb = det.addTemplateFromQuantized(q, mags, "obj", mask)
        assert got_tid == want_tid, vi
        if want_tid >= 0:
            n_ok += 1
            assert tuple(got_bb) == tuple(want_bb)
            for (a, b) in zip(det.getTemplates("obj", got_tid), orc.get_template("obj", want_tid)):
                assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
    assert n_ok >= 5
    assert det.numTemplates("obj") == orc.num_templates("obj")


def test_extract_without_mask_and_failure_path():
    orc = O.OracleDetector([O.depth_normal()], (4,))
    det = Detector([DepthNormal()], (4,))
    rng = np.random.default_rng(0)
    # four large homogeneous quadrants -> plenty of candidates, no mask
    q = np.zeros((96, 96), np.uint8)
    q[:48, :48], q[:48, 48:], q[48:, :48], q[48:, 48:] = 1, 4, 16, 64
    depth = np.full((96, 96), 700, np.uint16)
    got = det.addTemplateFromQuantized([q], [None], "x", None)
    assert got[0] == 0 and len(det.getTemplates("x", 0)[0][3]) == 63
    # nothing to extract -> -1, and the class entry still exists (the reference creates it up front)
    empty = np.zeros((96, 96), np.uint8)
    assert det.addTemplateFromQuantized([empty], [None], "y", None)[0] == -1
    assert "y" in det.classIds() and det.numTemplates("y") == 0
    del orc, rng, depth


# ---------------------------------------------------------------------------------------------- final ordering
def test_finalize_raw_equals_reference_sort_unique():
    det = _random_detector(6)
    rng = np.random.default_rng(6)
    n = 4000
    raw = np.zeros(n, RAW_DTYPE)
    slots = rng.permutation(40 * 1200)[:n]  # (template, coarse position) pairs are unique in a real frame
    raw["order_key"] = slots // 1200
    raw["coarse_pos"] = slots % 1200
    raw["x"] = rng.integers(0, 12, n) * 5 + 2
    raw["y"] = rng.integers(0, 6, n) * 5 + 2
    raw["score"] = rng.integers(230, 253, n)
    raw["nf"] = 63
    raw["template_id"] = raw["order_key"]
    raw["class_index"] = 0
    got = det.finalize_raw(raw)
    order = np.lexsort((raw["coarse_pos"], raw["order_key"]))
    pre = np.zeros(n, MATCH_DTYPE)
    for k in ("x", "y", "template_id", "class_index"):
        pre[k] = raw[k][order]
    pre["similarity"] = (raw["score"][order].astype(np.float32) * np.float32(100.0)) / np.float32(4 * 63)
    want = O.sort_unique(pre)
    # same input order + same libstdc++ introsort => element-wise identical, including which duplicates merge (D-7)
    common.assert_matches_equal(got, want)
    assert len(got) < n  # duplicates were merged
    assert np.all(np.diff(got["similarity"]) <= 0)


def test_finalize_gathered_equals_per_frame_finalize():
    """lm_finalize_gathered (one call per exchanged chunk of the sharded stream) == lm_finalize_raw on every frame's and
    query's union of the ranks' records; frames whose staged slot is too small are flagged, not silently truncated."""
    import ctypes as C

    from linemod_pose_estimation_b200 import _capi
    from linemod_pose_estimation_b200.sharding import pack_block
    det = _random_detector(8)
    rng = np.random.default_rng(8)
    world, frames, slots, cap, n_q = 3, 5, 6, 64, 2
    block_bytes = 16 + cap * RAW_DTYPE.itemsize
    buf = np.zeros((world, slots, block_bytes), np.uint8)
    raws = {}
    for f in range(frames):
        for r in range(world):
            n = int(rng.integers(0, 50)) if f != 3 else (cap + 9 if r == 1 else 5)   # frame 3: rank 1 outgrows its slot
            raw = np.zeros(n, RAW_DTYPE)
            pos = rng.permutation(20 * 1200)[:n]
            q = rng.integers(0, n_q, n)
            raw["order_key"] = (pos // 1200 * world + r) | (q << 28)
            raw["coarse_pos"] = pos % 1200
            raw["x"], raw["y"] = rng.integers(0, 9, n) * 5 + 2, rng.integers(0, 5, n) * 5 + 2
            raw["score"], raw["nf"] = rng.integers(230, 253, n), 63
            raw["template_id"], raw["class_index"] = raw["order_key"] & 0xfffffff, q
            raws[(f, r)] = raw
            buf[r, f] = pack_block(raw, cap)
    out, offs = C.c_void_p(), (C.c_size_t * (frames * n_q + 1))()
    status = np.zeros(frames, np.uint8)
    _capi.check(_capi.lib().lm_finalize_gathered(det._h, buf.ctypes.data, world, frames, block_bytes, slots * block_bytes, cap, n_q,
                                                 C.byref(out), offs, status.ctypes.data))
    allm = det._take(out, offs[frames * n_q])
    assert list(status) == [0, 0, 0, 1, 0]
    for f in range(frames):
        union = np.concatenate([raws[(f, r)] for r in range(world)])
        for q in range(n_q):
            got = allm[offs[f * n_q + q]:offs[f * n_q + q + 1]]
            if f == 3:
                assert len(got) == 0
            else:
                common.assert_matches_equal(got, det.finalize_raw(union[(union["order_key"] >> 28) == q]), "frame %d query %d" % (f, q))
    assert offs[frames * n_q] > 100


def test_binary_template_cache_roundtrip_and_speed(tmp_path):
    """lm_write_cache / lm_create_from_cache hold exactly the model of the templates.yml, reject corrupted
    files, and load far faster than the YAML parse the reference's service repeats on every request."""
    import time
    rng = np.random.default_rng(17)
    det = Detector()
    for cid, n in (("memoryChip2", 700), ("cpu_binary", 300)):
        for _ in range(n):
            det.addSyntheticTemplate(synth.random_pyramid(rng), cid)
    yml, cache = str(tmp_path / "templates.yml"), str(tmp_path / "templates.lmb2")
    det.write(yml)
    det.write_cache(cache)
    t0 = time.perf_counter(); from_yaml = Detector.read(yml); t_yaml = time.perf_counter() - t0
    t0 = time.perf_counter(); from_cache = Detector.read_cache(cache); t_cache = time.perf_counter() - t0
    assert from_cache.classIds() == from_yaml.classIds() == det.classIds()
    assert from_cache.pyramidLevels() == 2 and [from_cache.getT(l) for l in range(2)] == [5, 8]
    for a, b in zip(from_cache.getModalities(), det.getModalities()):
        assert bytes(a) == bytes(b)
    for cid in det.classIds():
        assert from_cache.numTemplates(cid) == det.numTemplates(cid)
        for tid in range(0, det.numTemplates(cid), 37):
            for x, y, z in zip(from_cache.getTemplates(cid, tid), det.getTemplates(cid, tid), from_yaml.getTemplates(cid, tid)):
                assert x[:3] == y[:3] == z[:3] and np.array_equal(x[3], y[3]) and np.array_equal(x[3], z[3])
    assert t_cache * 5 < t_yaml, (t_cache, t_yaml)
    # the cache of a cache-loaded detector is byte-identical; corruption and truncation are detected
    from_cache.write_cache(str(tmp_path / "again.lmb2"))
    blob = open(cache, "rb").read()
    assert open(str(tmp_path / "again.lmb2"), "rb").read() == blob
    bad = bytearray(blob); bad[len(bad) // 2] ^= 0x40
    open(str(tmp_path / "bad.lmb2"), "wb").write(bytes(bad))
    open(str(tmp_path / "short.lmb2"), "wb").write(blob[:len(blob) - 100])
    open(str(tmp_path / "notcache.lmb2"), "wb").write(b"%YAML:1.0\n" + b" " * 64)
    for name in ("bad.lmb2", "short.lmb2", "notcache.lmb2", "missing.lmb2"):
        with pytest.raises(LinemodError) as e:
            Detector.read_cache(str(tmp_path / name))
        assert e.value.code == -3


def test_match_clustering_equals_reference_restatement():
    """lm_cluster_matches against the pure-Python restatement of rcd_voting / cluster_filter /
    similarity_score_calc / nonMaximaSuppressionUsingIOU / computeIoU (oracle/cluster_oracle.py)."""
    from oracle import cluster_oracle as CO
    rng = np.random.default_rng(5)
    n_templates = 400
    dists = 0.6 + 0.1 * rng.integers(0, 6, n_templates) + rng.uniform(-1e-3, 1e-3, n_templates)   # radii like the trainer's
    rects = np.stack([np.zeros(n_templates), np.zeros(n_templates), rng.integers(55, 194, n_templates),
                      rng.integers(55, 194, n_templates)], 1).astype(np.int32)
    for trial in range(6):
        n = int(rng.integers(0, 400))
        m = np.zeros(n, MATCH_DTYPE)
        centres = rng.integers(40, 440, (8, 2))
        pick = rng.integers(0, 8, n)
        m["x"] = centres[pick, 0] + rng.integers(-14, 15, n)
        m["y"] = centres[pick, 1] + rng.integers(-14, 15, n)
        m["template_id"] = rng.integers(0, n_templates, n)
        m["similarity"] = (rng.integers(8400, 10000, n) / 100.0 + rng.uniform(0, 0.004, n)).astype(np.float32)
        for step, thr in ((8, 2), (16, 1), (25, 0)):
            got = Detector.cluster_matches(m, dists, rects, step, 0.6, 0.1, cluster_threshold=thr, iou_threshold=0.4)
            want = CO.cluster_matches(m, dists, rects, step, 0.6, 0.1, cluster_threshold=thr, iou_threshold=0.4)
            assert len(got) == len(want), (trial, step, len(got), len(want))
            for g, w in zip(got, want):
                assert g["index"] == tuple(w[0]) and g["rect"] == tuple(w[2]) and g["matches"] == list(w[3])
                assert g["score"] == w[1]
        if n > 50:
            assert len(got) > 0
    with pytest.raises(LinemodError):
        bad = np.zeros(1, MATCH_DTYPE); bad["template_id"] = n_templates + 3
        Detector.cluster_matches(bad, dists, rects, 8, 0.6, 0.1)
