"""Parity of the CUDA path (through the C ABI) with the CPU oracle: every stage tap, coarse similarity maps,
candidate-level and final match lists, bit-exact.  Needs a B200; run with `pytest -m gpu`."""
import os

import numpy as np
import pytest

import common
from common import O, synth
from linemod_pose_estimation_b200 import Detector, LinemodError, Stage

pytestmark = pytest.mark.gpu


def _pair(kinds=("cg", "dn"), T=(5, 8), n_views=8, n_random=24, seed=5, classes=("obj",), lut=None, canvas=(240, 240)):
    orc, views = common.build_oracle(kinds, T, n_views, n_random, seed, classes, canvas)
    det = Detector(common.product_modalities(kinds), T)
    common.copy_templates(orc, det)
    if lut is not None:
        orc.set_similarity_lut(lut)
        det.set_similarity_lut(lut)
    return orc, det, views


def _check_stages(orc, det, L, M, kinds):
    for l in range(L):
        for m in range(M):
            for st, name in ((Stage.QUANT_RAW, "quant_raw"), (Stage.QUANTIZED, "quantized"), (Stage.SPREAD, "spread"),
                             (Stage.RESPONSE, "response"), (Stage.LINEAR, "linear")):
                a, b = det.fetch(st, l, m), orc.fetch(st, l, m)
                assert a.shape == b.shape, (name, l, m, a.shape, b.shape)
                bad = np.count_nonzero(a != b)
                assert bad == 0, "%s level %d modality %d: %d mismatching bytes" % (name, l, m, bad)
            if kinds[m] == "cg":
                a, b = det.fetch(Stage.MAGNITUDE, l, m), orc.fetch(O.Stage.MAGNITUDE, l, m)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), ("magnitude", l, m)


# ---------------------------------------------------------------------------------------------- front end
@pytest.mark.parametrize("rows,cols,T,kinds", [
    (480, 640, (5, 8), ("cg", "dn")),    # the reference's trainer configuration (renderer.cpp:179-185)
    (480, 640, (5, 8), ("cg",)),         # Ensenso nodes: RGB only
    (240, 320, (4, 8), ("dn", "cg")),    # modality order swapped, other T
    (96, 160, (2, 4, 8), ("cg", "dn")),  # three pyramid levels
    (100, 180, (5,), ("cg",)),           # single level, W % 4 != 0 -> byte store path
])
def test_front_end_taps_bit_exact(rows, cols, T, kinds):
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)
    det.set_option("debug_taps", 1)
    views = common.rendered_views(3, 41, canvas=(80, 80))
    for seed in (1001, 1002):
        bgr, depth, _ = synth.compose_scene(seed, views, rows=rows, cols=cols)
        src = common.sources_for(kinds, bgr, depth)
        orc.build_front(src)
        det.build_front(src)
        _check_stages(orc, det, len(T), len(kinds), kinds)
        for l in range(len(T)):
            go, gd = orc.geometry(l), det.geometry(l)
            assert go == gd


def test_front_end_with_masks_and_strided_roi():
    """masks argument of Detector::match, and a cropped ROI view with non-contiguous rows like the service passes
    (..._service.cpp:324-326: mat_rgb(crop) at bias_x = 56)."""
    kinds, T = ("cg", "dn"), (5, 8)
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)
    det.set_option("debug_taps", 1)
    views = common.rendered_views(3, 43)
    wide_bgr, wide_depth, _ = synth.compose_scene(7, views, rows=480, cols=752)
    bgr, depth = wide_bgr[:, 56:56 + 640], wide_depth[:, 56:56 + 640]
    assert not bgr.flags["C_CONTIGUOUS"]
    rng = np.random.default_rng(2)
    m0 = (rng.random((480, 640)) < 0.7).astype(np.uint8) * 255
    m1 = np.zeros((480, 640), np.uint8)
    m1[100:400, 150:600] = 1
    orc.build_front([np.ascontiguousarray(bgr), np.ascontiguousarray(depth)], masks=[m0, m1])
    det.build_front([bgr, depth], masks=[m0, m1])
    _check_stages(orc, det, 2, 2, kinds)


@pytest.mark.parametrize("table,dn_count", [("one_hot", 1), ("one_hot", 0), ("any", 1)])
def test_alternative_luts(table, dn_count):
    """Injected tables: the survey's similarity LUT and a random NORMAL_LUT.  A one-hot normal table (every entry 0 or a
    single bit, like upstream's) lets k_dn_fused take medianBlur(5) by counting (`dn_count` 1, default) -- compared here with
    the 99-exchange network (`dn_count` 0) -- and a table with arbitrary bytes must fall back to the network by itself."""
    lut = common.survey_similarity_lut()
    orc, det, views = _pair(lut=lut, n_views=4, n_random=12)
    rng = np.random.default_rng(0)
    if table == "one_hot":
        nlut = np.where(rng.random(8000) < 0.1, 0, 1 << rng.integers(0, 8, 8000)).astype(np.uint8)
    else:
        nlut = rng.integers(0, 256, 8000).astype(np.uint8)
    orc.set_normal_lut(nlut)
    det.set_normal_lut(nlut)
    det.set_option("dn_count", dn_count)
    det.set_option("debug_taps", 1)
    bgr, depth, _ = synth.compose_scene(11, views[:3])
    common.assert_matches_equal(det.match([bgr, depth], 80.0), orc.match([bgr, depth], 80.0))
    _check_stages(orc, det, 2, 2, ("cg", "dn"))


# ---------------------------------------------------------------------------------------------- matching
def test_coarse_similarity_maps_bit_exact():
    orc, det, views = _pair(n_views=8, n_random=120, seed=9)
    bgr, depth, _ = synth.compose_scene(1001, views[:4])
    orc.build_front([bgr, depth])
    det.build_front([bgr, depth])
    n = orc.num_templates("obj")
    assert n >= 120
    for tid in range(n):
        a, b = det.coarse_map("obj", tid), orc.coarse_map("obj", tid)
        assert np.array_equal(a, b), "coarse map of template %d differs in %d cells" % (tid, np.count_nonzero(a != b))


@pytest.mark.parametrize("graphs,narrow", [(1, 1), (0, 1), (1, 0)])
def test_coarse_kernel_maps_lists_and_packed_planes(graphs, narrow):
    """The coarse kernel (nibble-packed linear memories, tile records, exact early termination) must give the oracle's
    coarse maps and match lists, replayed from the lane's CUDA graph or launched plainly, by the u8-only kernel of requests
    whose tiles have at most 63 features (31 + 31 here; `coarse_narrow` 1, default) and by the general one; the packed
    planes must be the reference's byte planes two positions per byte."""
    variant = 2 * graphs + narrow
    orc, det, views = _pair(n_views=8, n_random=90, seed=21, classes=("a", "b"))
    det.set_option("graphs", graphs)
    det.set_option("coarse_narrow", narrow)
    bgr, depth, _ = synth.compose_scene(1003, views[:4])
    orc.build_front([bgr, depth])
    det.build_front([bgr, depth])
    for m in range(2):
        packed = det.fetch(Stage.LINEAR_PACKED, 1, m)
        want = orc.fetch(Stage.LINEAR, 1, m)
        assert np.array_equal(packed & 15, want[:, 0::2]) and np.array_equal(packed >> 4, want[:, 1::2])
    for cid in ("a", "b"):
        for tid in range(orc.num_templates(cid)):
            a, b = det.coarse_map(cid, tid), orc.coarse_map(cid, tid)
            assert np.array_equal(a, b), "variant %d: coarse map of %s/%d differs in %d cells" % (variant, cid, tid, np.count_nonzero(a != b))
    for thr in (90.0, 65.0):
        want = orc.match([bgr, depth], thr, keep_candidates=True)
        got = det.match([bgr, depth], thr)
        common.assert_matches_equal(got, want, "variant %d thr %g" % (variant, thr))
        assert det.last_work()["candidates"] == len(orc.last_candidates())
    det.set_option("coarse_narrow", 1)   # process-wide switch: back to the default
    assert len(want) > 0


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("kinds", [("cg", "dn"), ("dn", "cg"), ("cg", "dn", "cg")])
def test_modality_order_of_the_coarse_sum_does_not_change_results(order, kinds):
    """The coarse kernel may sum the modalities in template order (0), reversed (1) or in the order the front end's
    spread-bit counters suggest for the frame (2, default): the early-termination bound is exact for any order, so the
    candidate counts and the match lists are the oracle's in every case, while the bytes gathered differ."""
    orc, det, views = _pair(kinds=kinds, n_views=6, n_random=60, seed=33)
    det.set_option("mod_order", order)
    bgr, depth, _ = synth.compose_scene(1007, views[:4])
    src = common.sources_for(kinds, bgr, depth)
    for thr in (92.0, 70.0):
        want = orc.match(src, thr, keep_candidates=True)
        got = det.match(src, thr)
        common.assert_matches_equal(got, want, "order %d thr %g" % (order, thr))
        assert det.last_work()["candidates"] == len(orc.last_candidates())
        w = det.last_work()
        assert 0 < w["B_coarse_gathered"] <= w["B_coarse"]
    assert len(want) > 0


@pytest.mark.parametrize("variant,tiled", [(3, 1), (0, 1), (3, 0)])
def test_refine_kernel_many_candidates(variant, tiled):
    """Refinement on nibble-packed planes with loose thresholds so that thousands of candidates are refined (warp per
    candidate) and a tight one (block per candidate), with the exact early termination of hopeless candidates on
    (`prune` 3, default) and off (0), on column-blocked planes (`refine_tiled` 1, default) and on flat ones (0); the lists
    must be the oracle's, and the packed planes of the refinement level -- handed out in the reference's flat order, with the
    column-blocked layout's halo rows and zero run verified on the way -- must be its byte planes two positions per byte."""
    orc, det, views = _pair(n_views=8, n_random=60, seed=23, classes=("a", "b"))
    det.set_option("prune", variant)
    det.set_option("refine_tiled", tiled)
    for seed, thr in ((1005, 86.0), (1006, 58.0)):
        bgr, depth, _ = synth.compose_scene(seed, views[:5])
        want = orc.match([bgr, depth], thr, keep_candidates=True)
        got = det.match([bgr, depth], thr)
        common.assert_matches_equal(det.last_presort(), orc.last_presort(), "variant %d thr %g pre-sort" % (variant, thr))
        common.assert_matches_equal(got, want, "variant %d thr %g" % (variant, thr))
    assert len(orc.last_candidates()) > 500
    for m in range(2):
        want = orc.fetch(Stage.LINEAR, 0, m)
        assert np.array_equal(det.fetch(Stage.LINEAR, 0, m), want)   # unpacked from the nibble planes when variant == 0
        if True:
            packed = det.fetch(Stage.LINEAR_PACKED, 0, m)
            assert np.array_equal(packed & 15, want[:, 0::2]) and np.array_equal(packed >> 4, want[:, 1::2])


@pytest.mark.parametrize("threshold", [92.0, 80.0, 60.0])
def test_match_lists_identical(threshold):
    orc, det, views = _pair(n_views=10, n_random=60, seed=13, classes=("cpu_binary", "memoryChip2"))
    for seed in (1001, 2000, 2001):
        bgr, depth, _ = synth.compose_scene(seed, views[:5])
        want = orc.match([bgr, depth], threshold, keep_candidates=True)
        got = det.match([bgr, depth], threshold)
        common.assert_matches_equal(det.last_presort(), orc.last_presort(), "pre-sort list")
        common.assert_matches_equal(got, want)
        assert det.last_work()["candidates"] == len(orc.last_candidates())
    assert len(want) > 0


def test_golden_scene_fixture():
    """CUDA path against the committed fixture (no oracle involved at run time)."""
    G = np.load(os.path.join(common.GOLDEN, "oracle_scene.npz"))
    det = Detector()
    det.set_option("debug_taps", 1)
    flat, k = G["templates_flat"], 0
    for _ in range(int(G["n_templates"][0])):
        pyr = []
        for _ in range(4):
            w, h, lvl, nf = flat[k:k + 4]
            pyr.append((int(w), int(h), int(lvl), flat[k + 4:k + 4 + 3 * nf].reshape(-1, 3)))
            k += 4 + 3 * nf
        det.addSyntheticTemplate(pyr, "obj")
    got = det.match([G["bgr"], G["depth"]], 80.0)
    common.assert_matches_equal(got, G["matches"])
    common.assert_matches_equal(det.last_presort(), G["presort"], "pre-sort list")
    hashes = []
    for l in range(2):
        for m in range(2):
            for st in (Stage.QUANTIZED, Stage.SPREAD, Stage.RESPONSE, Stage.LINEAR):
                hashes.append("%d/%d/%d:%s" % (l, m, st, common.sha(det.fetch(st, l, m))))
    assert hashes == list(G["stage_hashes"])


def test_class_filter_order_and_unknown_ids():
    orc, det, views = _pair(n_views=6, n_random=20, seed=17, classes=("a", "b", "c"))
    bgr, depth, _ = synth.compose_scene(5, views[:4])
    for ids in (["b"], ["c", "a"], ["a", "zzz", "a"], ["zzz"]):
        common.assert_matches_equal(det.match([bgr, depth], 75.0, class_ids=ids), orc.match([bgr, depth], 75.0, class_ids=ids), str(ids))
        common.assert_matches_equal(det.last_presort(), orc.last_presort(), "pre-sort " + str(ids))


def test_single_modality_and_single_level():
    for kinds, T in ((("cg",), (5, 8)), (("cg",), (8,)), (("dn",), (4, 8)), (("cg", "dn"), (8,))):
        orc, det, views = _pair(kinds=kinds, T=T, n_views=6, n_random=30, seed=23)
        bgr, depth, _ = synth.compose_scene(31, views[:3], rows=240, cols=320)
        src = common.sources_for(kinds, bgr, depth)
        thr = 70.0 if len(T) > 1 else 85.0
        common.assert_matches_equal(det.match(src, thr), orc.match(src, thr), "%s %s" % (kinds, T))
        common.assert_matches_equal(det.last_presort(), orc.last_presort(), "pre-sort %s %s" % (kinds, T))


def test_three_levels():
    """Two refinement levels: level 0 (W = 240) is column-blocked, level 1 (W = 60, not a multiple of 16) stays flat; switching
    the layout off between two calls rebuilds the workspace and returns the same lists."""
    orc, det, views = _pair(T=(2, 4, 8), n_views=6, n_random=20, seed=29, canvas=(160, 160))
    bgr, depth, _ = synth.compose_scene(3, views[:3], rows=320, cols=480)
    for tiled in (1, 0, 1):
        det.set_option("refine_tiled", tiled)
        common.assert_matches_equal(det.match([bgr, depth], 70.0), orc.match([bgr, depth], 70.0))
        common.assert_matches_equal(det.last_presort(), orc.last_presort(), "pre-sort")
        for l in range(3):
            for m in range(2):
                assert np.array_equal(det.fetch(Stage.LINEAR, l, m), orc.fetch(Stage.LINEAR, l, m)), (tiled, l, m)


def test_refinement_level_too_short_for_column_blocks_stays_flat():
    """150x320 with T = (10, 5): the refinement level has W/T = 32 columns (a multiple of 16) but only H/T = 15 rows -- fewer
    than a window -- so its planes stay in the reference's flat order (and T = 10 takes the spread kernel's generic path).
    Linear memories, candidate counts and the (empty: no 16-cell window fits the image) match lists equal the oracle's."""
    orc, det, views = _pair(T=(10, 5), n_views=8, n_random=12, seed=31, canvas=(96, 96))
    bgr, depth, _ = synth.compose_scene(77, views[:3], rows=150, cols=320)
    for thr in (60.0, 40.0):
        want = orc.match([bgr, depth], thr, keep_candidates=True)
        got = det.match([bgr, depth], thr)
        common.assert_matches_equal(got, want, "thr %g" % thr)
        assert det.last_work()["candidates"] == len(orc.last_candidates()) > 0
    for l in range(2):
        for m in range(2):
            assert np.array_equal(det.fetch(Stage.LINEAR, l, m), orc.fetch(Stage.LINEAR, l, m)), (l, m)


def test_edge_templates():
    """Template larger than the image (P <= 0), features outside the image, the (width, height) spill corner,
    negative coordinates, a template with one feature, an empty class."""
    kinds, T = ("cg", "dn"), (5, 8)
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)

    def add(pyr):
        orc.add_synthetic_template("obj", pyr)
        det.addSyntheticTemplate(pyr, "obj")

    rng = np.random.default_rng(1)
    f = lambda pts: np.array(pts, np.int32)
    add([(700, 500, 0, f([[0, 0, 1], [700, 500, 2]])), (700, 500, 0, f([[5, 5, 0]])),
         (350, 250, 1, f([[0, 0, 1], [349, 249, 3]])), (350, 250, 1, f([[3, 3, 0]]))])          # bigger than 640x480
    add([(120, 80, 0, f([[0, 0, 1], [120, 80, 2], [639, 479, 3], [640, 10, 1], [10, 480, 1], [-3, 4, 2]])),
         (120, 80, 0, f([[60, 40, 4]])),
         (60, 40, 1, f([[0, 0, 1], [60, 40, 2], [319, 239, 5], [320, 3, 1], [-1, -1, 0]])),
         (60, 40, 1, f([[30, 20, 6]]))])                                                        # out-of-bounds features
    add([(80, 80, 0, f([[80, 80, 7]])), (80, 80, 0, f([[0, 0, 0]])), (40, 40, 1, f([[40, 40, 7]])), (40, 40, 1, f([[0, 0, 0]]))])
    add([(8, 8, 0, np.zeros((0, 3), np.int32)), (8, 8, 0, f([[1, 1, 1]])), (4, 4, 1, np.zeros((0, 3), np.int32)), (4, 4, 1, f([[1, 1, 1]]))])
    for _ in range(10):
        add(synth.random_pyramid(rng, T=T, M=2, wh_range=(8, 600)))
    views = common.rendered_views(3, 47)
    bgr, depth, _ = synth.compose_scene(13, views)
    orc.build_front([bgr, depth])
    det.build_front([bgr, depth])
    for tid in range(orc.num_templates("obj")):
        assert np.array_equal(det.coarse_map("obj", tid), orc.coarse_map("obj", tid)), tid
    for thr in (50.0, 10.0):
        common.assert_matches_equal(det.match([bgr, depth], thr), orc.match([bgr, depth], thr), "thr %g" % thr)
        common.assert_matches_equal(det.last_presort(), orc.last_presort(), "pre-sort thr %g" % thr)


def test_config3_1280x960_refinement_heavy():
    """BASELINE config 3: Ensenso-resolution frames (1280x960 with T = {5, 8}; 1280x1024 violates rows % T, see
    test_geometry_asserts_like_the_reference), both modalities, a threshold loose enough for > 1000 coarse candidates per
    frame so that similarityLocal dominates.  Every stage tap and the match list against the oracle."""
    orc, det, views = _pair(n_views=8, n_random=40, seed=83)
    det.set_option("debug_taps", 1)
    bgr, depth, _ = synth.compose_scene(4001, views[:6], rows=960, cols=1280)
    want = orc.match([bgr, depth], 55.0, keep_candidates=True)
    got = det.match([bgr, depth], 55.0)
    assert len(orc.last_candidates()) > 1000
    assert det.last_work()["candidates"] == len(orc.last_candidates())
    common.assert_matches_equal(det.last_presort(), orc.last_presort(), "pre-sort list")
    common.assert_matches_equal(got, want)
    _check_stages(orc, det, 2, 2, ("cg", "dn"))
    assert len(want) > 0


def test_no_templates_and_empty_results():
    det = Detector()
    bgr = np.zeros((480, 640, 3), np.uint8)
    depth = np.zeros((480, 640), np.uint16)
    assert len(det.match([bgr, depth], 90.0)) == 0
    rng = np.random.default_rng(0)
    det.addSyntheticTemplate(synth.random_pyramid(rng), "obj")
    assert len(det.match([bgr, depth], 90.0)) == 0  # flat black frame: no responses above threshold


def test_geometry_asserts_like_the_reference():
    det = Detector()
    with pytest.raises(LinemodError) as e:  # linearize: rows % T == 0 (1024 % 5 = 4)
        det.match([np.zeros((1024, 1280, 3), np.uint8), np.zeros((1024, 1280), np.uint16)], 90.0)
    assert e.value.code == -1 and "% T" in str(e.value)


def test_many_candidates_overflow_and_growth():
    """A loose threshold on a big template set overflows the initial candidate / result buffers; the library must
    grow and still return the exact list."""
    orc, det, views = _pair(n_views=4, n_random=300, seed=37)
    bgr, depth, _ = synth.compose_scene(21, views[:3])
    want = orc.match([bgr, depth], 20.0, keep_candidates=True)
    got = det.match([bgr, depth], 20.0)
    assert len(orc.last_candidates()) > (1 << 16)
    common.assert_matches_equal(got, want)


def test_add_template_on_gpu_matches_oracle():
    kinds, T = ("cg", "dn"), (5, 8)
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)
    ok = 0
    for (bgr, depth, mask) in common.rendered_views(12, 53, canvas=(240, 256)):
        wt, wbb = orc.add_template([bgr, depth], "obj", mask)
        gt, gbb = det.addTemplate([bgr, depth], "obj", mask)
        assert gt == wt
        if wt >= 0:
            ok += 1
            assert tuple(gbb) == tuple(wbb)
            for (a, b) in zip(det.getTemplates("obj", gt), orc.get_template("obj", wt)):
                assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
    assert ok >= 6
    # without a mask (whole image) as well
    bgr, depth, _ = synth.compose_scene(3, [], rows=120, cols=160)
    wt, _ = orc.add_template([bgr, depth], "scene", None)
    gt, _ = det.addTemplate([bgr, depth], "scene", None)
    assert gt == wt
    if wt >= 0:
        for (a, b) in zip(det.getTemplates("scene", gt), orc.get_template("scene", wt)):
            assert np.array_equal(a[3], b[3])


def test_modality_process_and_quantized_pyramid_equal_the_oracle():
    """cv::linemod::Modality::process -> QuantizedPyramid::{quantize, extractTemplate, pyrDown} ([OCV] linemod.cpp), the
    surface under addTemplate and match: per level the masked quantised image and the uncropped template."""
    from linemod_pose_estimation_b200 import ColorGradient, DepthNormal, process
    kinds, T = ("cg", "dn"), (5, 8, 8)
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    mods = (ColorGradient(), DepthNormal())
    cases = [(b, d, m) for (b, d, m) in common.rendered_views(4, 77, canvas=(240, 256))]
    scene_b, scene_d, _ = synth.compose_scene(11, [], rows=120, cols=160)
    cases.append((scene_b, scene_d, None))                            # no mask: the whole image
    flat = (np.full((96, 128, 3), 90, np.uint8), np.full((96, 128), 800, np.uint16), None)
    cases.append(flat)                                                # nothing to extract: extractTemplate returns false
    extracted = 0
    for (bgr, depth, mask) in cases:
        for m, src in enumerate((bgr, depth)):
            qp = process(mods[m], src, mask, levels=3)
            for level in range(3):
                if level > 0:
                    qp.pyrDown()
                want_q, want_ok, want_t = orc.modality_process(m, src, mask, level)
                got_q = qp.quantize()
                assert got_q.shape == want_q.shape and np.array_equal(got_q, want_q), (m, level)
                ok, t = qp.extractTemplate()
                assert ok == want_ok, (m, level)
                if ok:
                    extracted += 1
                    assert t[:3] == want_t[:3] == (-1, -1, level) and np.array_equal(t[3], want_t[3]), (m, level)
    assert extracted >= 16
    qp = process(mods[0], flat[0], None, levels=2)
    qp.pyrDown()
    with pytest.raises(LinemodError):
        qp.pyrDown()
    with pytest.raises(LinemodError):
        process(mods[1], flat[0])                                    # DepthNormal needs CV_16UC1


def test_batch_equals_single_and_quantized_images(tmp_path):
    orc, det, views = _pair(n_views=6, n_random=40, seed=59)
    frames = [list(synth.compose_scene(2000 + i, views[:4])[:2]) for i in range(5)]
    singles = [det.match(f, 85.0) for f in frames]
    batch = det.match_batch(frames, 85.0)
    for a, b, f in zip(singles, batch, frames):
        common.assert_matches_equal(b, a)
        common.assert_matches_equal(a, orc.match(f, 85.0))
    m, q = det.match(frames[0], 85.0, quantized_images=True)
    for l in range(2):
        for mod in range(2):
            assert np.array_equal(q[l * 2 + mod], orc.fetch(O.Stage.QUANTIZED, l, mod)) or True
    orc.match(frames[0], 85.0)
    for l in range(2):
        for mod in range(2):
            assert np.array_equal(q[l * 2 + mod], orc.fetch(O.Stage.QUANTIZED, l, mod))
    # persistence round trip keeps results identical
    p = tmp_path / "t.yml"
    det.write(p)
    again = Detector.read(p)
    common.assert_matches_equal(again.match(frames[1], 85.0), singles[1])


# ---------------------------------------------------------------------------------------------- full size
def test_full_size_properties_20k_templates():
    """BASELINE config 4 size (20 000 templates, 640x480): too slow for the scalar oracle in a test, so parity is
    checked through size-independent properties: determinism, shard-union == whole, class-split == whole, and the
    oracle on a random sample of templates."""
    rng = np.random.default_rng(99)
    det = Detector()
    pyrs = [synth.random_pyramid(rng) for _ in range(20000)]
    for i, p in enumerate(pyrs):
        det.addSyntheticTemplate(p, "c%02d" % (i % 15))
    views = common.rendered_views(4, 61)
    bgr, depth, _ = synth.compose_scene(1001, views)
    thr = 62.0
    whole = det.match([bgr, depth], thr)
    assert len(whole) > 0
    common.assert_matches_equal(det.match([bgr, depth], thr), whole, "determinism")
    # union of class-filtered runs (std::map order) == whole, before the global sort
    pre_whole = det.last_presort()
    pre_parts = []
    for cid in det.classIds():
        det.match([bgr, depth], thr, class_ids=[cid])
        pre_parts.append(det.last_presort())
    common.assert_matches_equal(np.concatenate(pre_parts), pre_whole, "class split")
    # oracle on a sample of templates
    orc = O.OracleDetector()
    sample = sorted(rng.choice(20000, 200, replace=False))
    keep = {}
    for k, i in enumerate(sample):
        orc.add_synthetic_template("s", pyrs[i])
        keep[k] = ("c%02d" % (i % 15), i // 15)
    orc.match([bgr, depth], thr)
    pre_o = orc.last_presort()
    ids = det.classIds()
    for k, (cid, tid) in keep.items():
        a = pre_whole[(pre_whole["class_index"] == ids.index(cid)) & (pre_whole["template_id"] == tid)]
        b = pre_o[pre_o["template_id"] == k]
        assert len(a) == len(b) and np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and \
            np.array_equal(a["similarity"], b["similarity"]), (cid, tid)


def test_shard_union_equals_whole():
    """lm_set_shard + lm_match_device + lm_finalize_raw: four template shards evaluated one after another on one GPU
    (the N-rank layout without the collective) reproduce the unsharded match list exactly."""
    import torch
    from linemod_pose_estimation_b200 import RAW_DTYPE
    from linemod_pose_estimation_b200.sharding import device_view
    orc, det, views = _pair(n_views=8, n_random=80, seed=67, classes=("a", "b"))
    bgr, depth, _ = synth.compose_scene(77, views[:4])
    want = orc.match([bgr, depth], 70.0)
    dev = torch.device("cuda", 0)
    d_bgr = torch.from_numpy(bgr).to(dev)
    d_depth = torch.from_numpy(depth.view(np.int16)).to(dev)
    raws = []
    for rank in range(4):
        shard = Detector()
        common.copy_templates(orc, shard)
        shard.set_shard(rank, 4)
        rec, cap = shard.match_device([d_bgr.data_ptr(), d_depth.data_ptr()], 480, 640, 70.0,
                                      stream=torch.cuda.current_stream().cuda_stream)
        block = device_view(rec, cap, dev).cpu().numpy()
        hdr = block[:16].view(np.uint32)
        assert hdr[2] == 0, "overflow"
        raws.append(block[16:16 + int(hdr[0]) * RAW_DTYPE.itemsize].view(RAW_DTYPE).copy())
    assert all(len(r) > 0 for r in raws)
    common.assert_matches_equal(det.finalize_raw(np.concatenate(raws)), want)
    common.assert_matches_equal(det.match([bgr, depth], 70.0), want)


def test_match_multi_equals_separate_matches():
    """lm_match_multi: several (class list, threshold) queries from one front end == separate match calls."""
    orc, det, views = _pair(n_views=8, n_random=60, seed=73, classes=("cpu_binary", "memoryChip2"))
    queries = [(92.0, ["memoryChip2"]), (74.0, ["cpu_binary"]), (60.0, []), (80.0, ["memoryChip2", "cpu_binary"])]
    for seed in (2000, 2001):
        bgr, depth, _ = synth.compose_scene(seed, views[:5])
        got = det.match_multi([bgr, depth], queries)
        assert len(got) == len(queries)
        for g, (thr, ids) in zip(got, queries):
            common.assert_matches_equal(g, orc.match([bgr, depth], thr, class_ids=ids), "query %s" % ((thr, ids),))
            common.assert_matches_equal(g, det.match([bgr, depth], thr, class_ids=ids), "separate call %s" % ((thr, ids),))
    assert sum(len(g) for g in got) > 0


def test_batch_multi_equals_per_frame_multi():
    """lm_match_batch_multi: a stream of frames, every frame answering every query == lm_match_multi per frame."""
    orc, det, views = _pair(n_views=8, n_random=40, seed=79, classes=("cpu_binary", "memoryChip2"))
    queries = [(90.0, ["memoryChip2"]), (70.0, ["cpu_binary"]), (62.0, [])]
    frames = []
    for seed in (3000, 3001, 3002, 3003, 3004):
        bgr, depth, _ = synth.compose_scene(seed, views[:5])
        frames.append([bgr, depth])
    got = det.match_batch_multi(frames, queries)
    assert len(got) == len(frames)
    total = 0
    for f, per_query in zip(frames, got):
        single = det.match_multi(f, queries)
        for q, (g, s) in enumerate(zip(per_query, single)):
            common.assert_matches_equal(g, s, "frame/query %d" % q)
            common.assert_matches_equal(g, orc.match(f, queries[q][0], class_ids=queries[q][1]), "oracle, query %d" % q)
            total += len(g)
    assert total > 0
    assert det.match_batch_multi([], queries) == []


@pytest.mark.parametrize("batch_frames,lanes", [(16, 4), (8, 4), (3, 2), (1, 1)])
def test_frame_stream_equals_the_oracle(batch_frames, lanes):
    """lm_stream: frames pushed in uneven pieces with pops in between (blocking and not) come back in push order with the
    oracle's lists for every query; a frame whose survivors outgrow the head of its result block is redone inside the stream;
    the detector refuses other matching calls while the stream is open and answers again after close()."""
    orc, det, views = _pair(n_views=8, n_random=40, seed=79, classes=("cpu_binary", "memoryChip2"))
    det.set_option("stream_frames", batch_frames)    # frames per chunk of the stream (default 16)
    det.set_option("batch_lanes", lanes)
    queries = [(90.0, ["memoryChip2"]), (62.0, [])]
    frames = []
    for seed in range(3100, 3123):
        bgr, depth, _ = synth.compose_scene(seed, views[:5])
        frames.append([bgr, depth])
    want = [[orc.match(f, thr, class_ids=ids) for thr, ids in queries] for f in frames]
    got = []
    with det.open_stream(queries) as st:
        with pytest.raises(LinemodError):
            det.match(frames[0], 90.0)
        at = 0
        for piece in (1, 5, 2, 9, 6):
            st.push(frames[at:at + piece])
            at += piece
            assert st.in_flight() == at - len(got)
            got += st.pop()                      # what is ready, without blocking
        assert at == len(frames)
        got += st.pop(wait_all=True, max_frames=2)   # at most two of the rest ...
        got += st.pop(wait_all=True)                 # ... then all of it
        assert st.in_flight() == 0 and st.pop() == []
    assert len(got) == len(frames)
    total = 0
    for f, (g, w) in enumerate(zip(got, want)):
        for q in range(len(queries)):
            common.assert_matches_equal(g[q], w[q], "frame %d query %d" % (f, q))
            total += len(g[q])
    assert total > 0
    common.assert_matches_equal(det.match(frames[0], 62.0), want[0][1], "after close")


def test_sharded_stream_single_rank():
    """ShardedDetector.match_stream (chunked upload, lanes, staged survivor exchange) at world size 1 == lm_match_multi per
    frame; the collectives themselves are exercised at world size 2 on CPU (tests/test_sharding_gloo.py) and by bench.py."""
    from linemod_pose_estimation_b200.sharding import ShardedDetector
    orc, det, views = _pair(n_views=8, n_random=40, seed=89, classes=("cpu_binary", "memoryChip2"))
    queries = [(88.0, ["memoryChip2"]), (66.0, [])]
    frames = [list(synth.compose_scene(5000 + i, views[:5])[:2]) for i in range(11)]
    sd = ShardedDetector(det, capacity=256)   # 66 % over all classes outgrows 256 records: those frames take the fallback
    got = sd.match_stream(frames, queries, chunk=4, lanes=3)
    assert len(got) == len(frames)
    total = 0
    for f, per_query in zip(frames, got):
        for g, s in zip(per_query, det.match_multi(f, queries)):
            common.assert_matches_equal(g, s, "streamed vs single")
            total += len(g)
    assert total > 0


def test_sharded_stream_large_chunks_forced_fallback_and_defaults():
    """(i) the per-frame fallback of frames whose survivors outgrow the exchange slot runs only after the stream
    has drained (it uses lane 0, which later chunks are still using while their predecessors finish) -- large chunks, several
    chunks in flight, most frames forced into the fallback; (ii) the default capacity (4 096) works on a fresh detector:
    the library's device record blocks follow the exchange capacity."""
    from linemod_pose_estimation_b200.sharding import ShardedDetector
    orc, det, views = _pair(n_views=8, n_random=40, seed=89, classes=("cpu_binary", "memoryChip2"))
    queries = [(88.0, ["memoryChip2"]), (60.0, [])]
    frames = [list(synth.compose_scene(5200 + i, views[i % 4:i % 4 + 4])[:2]) for i in range(40)]
    want = [det.match_multi(f, queries) for f in frames]
    assert max(len(w[1]) for w in want) > 64            # the loose query outgrows a 64-record slot
    sd = ShardedDetector(det, capacity=64)
    got = sd.match_stream(frames, queries, chunk=16, lanes=4)
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            common.assert_matches_equal(a, b, "forced fallback")
    fresh = Detector()
    common.copy_templates(orc, fresh)
    sd2 = ShardedDetector(fresh)                        # defaults
    got = sd2.match_stream(frames[:9], queries)
    for g, w in zip(got, want[:9]):
        for a, b in zip(g, w):
            common.assert_matches_equal(a, b, "defaults")


def test_chunk_result_blocks_are_one_region():
    """lm_match_device_stream runs a chunk of device-resident frames as ONE launch set on a lane (no copies: the kernels
    read the caller's buffers through the frame table); lm_device_result_region: the chunk's record blocks lie frame stride
    apart; lm_copy_result_block copies the first block's head; the staged heads equal the blocks."""
    import ctypes as C

    import torch
    from linemod_pose_estimation_b200 import RAW_DTYPE, _capi
    orc, det, views = _pair(n_views=6, n_random=20, seed=97)
    dev = torch.device("cuda", 0)
    frames = [synth.compose_scene(6001 + i, views[i % 3:i % 3 + 3])[:2] for i in range(5)]
    d_frames = [(torch.from_numpy(b).to(dev), torch.from_numpy(d.view(np.int16)).to(dev)) for (b, d) in frames]
    lib = _capi.lib()
    qarr, _keep = _capi.query_array([(88.0, [])])
    flat = [p for (b, d) in d_frames for p in (b.data_ptr(), d.data_ptr())]
    ptrs = (C.c_void_p * len(flat))(*flat)
    s = torch.cuda.current_stream().cuda_stream
    streams = (C.c_void_p * 1)(s)
    slot = 16 + 1024 * 32
    stage = torch.zeros(5 * slot, dtype=torch.uint8, device=dev)
    _capi.check(lib.lm_match_device_stream(det._h, ptrs, 5, 2, 480, 640, qarr, 1, streams, 1, stage.data_ptr(), slot))
    base, stride, n_frames = C.c_void_p(), C.c_size_t(), C.c_int()
    _capi.check(lib.lm_device_result_region(det._h, 0, C.byref(base), C.byref(stride), C.byref(n_frames)))
    assert n_frames.value >= 5 and stride.value % 256 == 0
    head = torch.zeros(slot, dtype=torch.uint8, device=dev)
    _capi.check(lib.lm_copy_result_block(det._h, 0, head.data_ptr(), slot, s))
    torch.cuda.synchronize()
    from linemod_pose_estimation_b200.sharding import device_view
    region = device_view(base.value, stride.value * 5, dev).cpu().numpy()
    staged = stage.cpu().numpy()
    assert np.array_equal(head.cpu().numpy()[:16], staged[:16])
    for f, (bgr, depth) in enumerate(frames):
        block = staged[f * slot:(f + 1) * slot]
        hdr = block[:16].view(np.uint32)
        assert hdr[2] == 0 and 0 < hdr[0] <= 1024
        n = int(hdr[0])
        assert np.array_equal(region[f * stride.value:f * stride.value + 16 + n * 32], block[:16 + n * 32])
        raw = block[16:16 + n * 32].view(RAW_DTYPE).copy()
        common.assert_matches_equal(det.finalize_raw(raw), orc.match([bgr, depth], 88.0), "frame %d" % f)


@pytest.mark.parametrize("batch_frames,lanes", [(1, 2), (3, 1), (8, 4), (16, 3), (32, 2)])
def test_chunked_batches_equal_the_oracle(batch_frames, lanes):
    """lm_match_batch_multi cuts the frames into chunks of `batch_frames` (one launch set per chunk) pipelined over
    `batch_lanes` lanes: every chunk size, ragged tails included, must give each frame the oracle's lists; a frame whose
    survivors outgrow its record block (66 % over all classes) is redone alone."""
    orc, det, views = _pair(n_views=8, n_random=40, seed=89, classes=("cpu_binary", "memoryChip2"))
    det.set_option("batch_frames", batch_frames)
    det.set_option("batch_lanes", lanes)
    queries = [(88.0, ["memoryChip2"]), (66.0, [])]
    frames = [list(synth.compose_scene(5100 + i, views[i % 4:i % 4 + 4])[:2]) for i in range(21)]
    got = det.match_batch_multi(frames, queries)
    assert len(got) == len(frames)
    total = 0
    for f, per_query in zip(frames, got):
        for (thr, ids), g in zip(queries, per_query):
            common.assert_matches_equal(g, orc.match(f, thr, class_ids=ids), "chunk %d" % batch_frames)
            total += len(g)
    assert total > 0


@pytest.mark.parametrize("share", [1, 0])
def test_shared_tail_tiles_equal_the_oracle(share):
    """Templates small enough that their coarse span exceeds one 1 024-position pass leave a tail pass of a few dozen positions;
    the coarse kernel scores such tails for eight frames per warp (`coarse_share` 1, default) or like any other tile (0).
    Chunks of 8, 3 and 1 frames (all, some and one of a warp's frame slots in use), loose thresholds so that tail positions
    become candidates: every frame must get the oracle's lists and candidate counts."""
    kinds, T = ("cg", "dn"), (5, 8)
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)
    rng = np.random.default_rng(97)
    for i in range(70):
        pyr = synth.random_pyramid(rng, T=T, M=2, wh_range=(24, 90) if i % 3 else (100, 190))
        orc.add_synthetic_template("small", pyr)
        det.addSyntheticTemplate(pyr, "small")
    det.set_option("coarse_share", share)
    views = common.rendered_views(4, 53)
    frames = [list(synth.compose_scene(6100 + i, views)[:2]) for i in range(12)]
    n_tail, want = 0, []
    for f in frames:
        want.append([orc.match(f, thr, keep_candidates=True) for thr in (75.0, 68.0)])
        n_tail += int(np.count_nonzero(orc.last_candidates()["pos"] >= 1024))   # candidates of the loose query in tail passes
    assert n_tail > 0
    for bf in (8, 3, 1):
        det.set_option("batch_frames", bf)
        got = det.match_batch_multi(frames, [(75.0, []), (68.0, [])])
        for f, (g, w) in enumerate(zip(got, want)):
            for q in range(2):
                common.assert_matches_equal(g[q], w[q], "share %d, chunks of %d, frame %d query %d" % (share, bf, f, q))
    single = det.match(frames[0], 68.0)
    common.assert_matches_equal(single, orc.match(frames[0], 68.0, keep_candidates=True), "single frame")
    assert det.last_work()["candidates"] == len(orc.last_candidates())
    det.set_option("coarse_share", 1)


def test_config5_64_frame_batch_15_classes():
    """BASELINE configs[4] at test scale: a 64-frame synthetic 640x480 video against 15 object classes, end to end
    (quantise -> spread -> response -> match) through lm_match_batch; every frame's match list equals the oracle's."""
    classes = tuple("obj%02d" % i for i in range(15))
    orc, det, views = _pair(n_views=2, n_random=6, seed=131, classes=classes, canvas=(200, 200))
    assert det.numClasses() == 15
    frames = [list(synth.compose_scene(7000 + i, views[(i % 5):(i % 5) + 4])[:2]) for i in range(64)]
    got = det.match_batch(frames, 80.0)
    assert len(got) == 64
    total = 0
    for f, g in zip(frames, got):
        common.assert_matches_equal(g, orc.match(f, 80.0))
        total += len(g)
    assert total > 64
    # a class subset in caller order, same stream
    sub = [classes[11], classes[3]]
    for f, g in zip(frames[:8], det.match_batch(frames[:8], 80.0, class_ids=sub)):
        common.assert_matches_equal(g, orc.match(f, 80.0, class_ids=sub))


def test_cache_loaded_detector_matches_identically(tmp_path):
    """SURVEY 8f N1 on the GPU: a detector loaded from the binary template cache (lm_create_from_cache) and one loaded
    from the templates.yml it was written from return the original detector's -- and the oracle's -- match lists."""
    orc, det, views = _pair(n_views=8, n_random=50, seed=171, classes=("cpu_binary", "memoryChip2"))
    yml, cache = str(tmp_path / "templates.yml"), str(tmp_path / "templates.lmb2")
    det.write(yml)
    det.write_cache(cache)
    from_cache, from_yaml = Detector.read_cache(cache), Detector.read(yml)
    total = 0
    for seed in (8101, 8102, 8103):
        bgr, depth, _ = synth.compose_scene(seed, views[:5])
        for thr, ids in ((90.0, []), (72.0, ["memoryChip2"])):
            want = orc.match([bgr, depth], thr, class_ids=ids)
            for d in (det, from_cache, from_yaml):
                common.assert_matches_equal(d.match([bgr, depth], thr, class_ids=ids), want, "thr %g" % thr)
            total += len(want)
    assert total > 0
    # ... and through the chunked batch path of the cache-loaded handle
    frames = [list(synth.compose_scene(8200 + i, views[:4])[:2]) for i in range(9)]
    for f, g in zip(frames, from_cache.match_batch(frames, 85.0)):
        common.assert_matches_equal(g, orc.match(f, 85.0))


def test_clusters_of_a_gpu_match_list_equal_the_restatement():
    """SURVEY 8f N2 behind the real matcher: lm_cluster_matches (rcd_voting, cluster_filter, mean-similarity score, IoU
    NMS -- /root/reference/src/rgbdDetector.cpp:36-145,462-574) on the match list the GPU returns, against
    oracle/cluster_oracle.py on the oracle's list; the pose tables are the trainer's (radius and mask rectangle per view)."""
    from oracle import cluster_oracle as CO
    from linemod_pose_estimation_b200 import Mesh, training
    det, orc = Detector(), O.OracleDetector()
    mesh = Mesh(synth.bracket_mesh())
    cam = training.camera()
    sphere = training.ViewSphere(n_points=12, angle_step=40, radius_min=0.5, radius_max=0.7, radius_step=0.1)
    T, up = sphere.views()
    radii = np.array([sphere.view(i)[2] for i in range(len(sphere))])
    tids, bbs, rects = det.trainViews(mesh, cam, T, up, "obj")[:3]
    ok = tids >= 0
    assert ok.sum() >= 20
    common.copy_templates(_DetAsSource(det), _OrcAsSink(orc))
    dists = radii[ok].astype(np.float64)
    trects = np.stack([rects["x"][ok], rects["y"][ok], rects["width"][ok], rects["height"][ok]], 1).astype(np.int32)
    r = training.render_views(det, mesh, cam, T[ok][:3], up[ok][:3])
    planted = [(r["bgr"][k], r["depth"][k], r["mask"][k]) for k in range(3)]
    n_clusters = 0
    for seed in (8301, 8302):
        bgr, depth, _ = synth.compose_scene(seed, planted, noise=False)
        got, want = det.match([bgr, depth], 80.0), orc.match([bgr, depth], 80.0)
        common.assert_matches_equal(got, want)
        for step, thr in ((8, 2), (16, 0)):
            a = Detector.cluster_matches(got, dists, trects, step, 0.5, 0.1, cluster_threshold=thr, iou_threshold=0.4)
            b = CO.cluster_matches(want, dists, trects, step, 0.5, 0.1, cluster_threshold=thr, iou_threshold=0.4)
            assert len(a) == len(b)
            for g, w in zip(a, b):
                assert g["index"] == tuple(w[0]) and g["rect"] == tuple(w[2]) and g["matches"] == list(w[3]) and g["score"] == w[1]
            n_clusters += len(a)
    assert n_clusters > 0


class _DetAsSource:
    def __init__(self, det):
        self.det = det

    def class_ids(self):
        return self.det.classIds()

    def num_templates(self, cid):
        return self.det.numTemplates(cid)

    def get_template(self, cid, tid):
        return self.det.getTemplates(cid, tid)


class _OrcAsSink:
    def __init__(self, orc):
        self.orc = orc

    def addSyntheticTemplate(self, templates, cid):
        return self.orc.add_synthetic_template(cid, templates)


@pytest.mark.parametrize("mode,members", [("frames", 2), ("frames", 3), ("templates", 2), ("templates", 3), ("grid2", 4), ("grid3", 6)])
def test_device_group_equals_the_oracle(mode, members):
    """lm_group (several GPUs behind one C++ caller): frames dealt out in launch sets / templates sharded with the shards'
    survivors merged on the host.  On a one-GPU box the members share device 0 -- the dealing, the per-member worker
    threads, the shard merge and the ordering are the same code; bench.py and tools/groupbench.py use distinct devices."""
    import torch
    from linemod_pose_estimation_b200 import DetectorGroup
    orc, det, views = _pair(n_views=8, n_random=40, seed=211, classes=("cpu_binary", "memoryChip2"))
    det.set_option("batch_frames", 4)
    n_dev = torch.cuda.device_count()
    if mode.startswith("grid"):   # 2-D: members / S sets of S template shards
        group = DetectorGroup(det, [i % n_dev for i in range(members)], template_shards=int(mode[4:]))
    else:
        group = DetectorGroup(det, [i % n_dev for i in range(members)], mode)
    assert len(group) == members
    queries = [(88.0, ["memoryChip2"]), (70.0, [])]
    frames = [list(synth.compose_scene(8400 + i, views[i % 3:i % 3 + 4])[:2]) for i in range(19)]
    got = group.match_batch_multi(frames, queries)
    total = 0
    for f, per_query in zip(frames, got):
        for (thr, ids), g in zip(queries, per_query):
            common.assert_matches_equal(g, orc.match(f, thr, class_ids=ids), "%s x%d" % (mode, members))
            total += len(g)
    assert total > 0
    common.assert_matches_equal(group.match(frames[0], 85.0), orc.match(frames[0], 85.0))
    group.close()
