"""The oracle's matching path: regression against its committed golden scene, an independent numpy restatement of
spread / response / linearize / similarity, and the quirks of App. D that parity must keep."""
import os

import numpy as np

import common
from common import O, synth

G = np.load(os.path.join(common.GOLDEN, "oracle_scene.npz"))


def _golden_oracle():
    orc = O.OracleDetector()
    flat = G["templates_flat"]
    k = 0
    for _ in range(int(G["n_templates"][0])):
        pyr = []
        for _ in range(4):
            w, h, lvl, nf = flat[k:k + 4]
            f = flat[k + 4:k + 4 + 3 * nf].reshape(-1, 3)
            k += 4 + 3 * nf
            pyr.append((int(w), int(h), int(lvl), f))
        orc.add_synthetic_template("obj", pyr)
    return orc


def test_golden_scene_regression():
    orc = _golden_oracle()
    got = orc.match([G["bgr"], G["depth"]], 80.0, keep_candidates=True)
    common.assert_matches_equal(got, G["matches"].view(O.MATCH_DTYPE) if G["matches"].dtype != O.MATCH_DTYPE else G["matches"])
    common.assert_matches_equal(orc.last_presort(), G["presort"], "presort")
    assert np.array_equal(orc.last_candidates(), G["candidates"])
    hashes = []
    for l in range(2):
        for m in range(2):
            for st in (O.Stage.QUANTIZED, O.Stage.SPREAD, O.Stage.RESPONSE, O.Stage.LINEAR):
                hashes.append("%d/%d/%d:%s" % (l, m, st, common.sha(orc.fetch(st, l, m))))
    assert hashes == list(G["stage_hashes"])


def _numpy_coarse(quantized, T, lut, template, size):
    """Plain numpy restatement of spread -> response -> linearize -> similarity for one modality."""
    rows, cols = quantized.shape
    spread = np.zeros_like(quantized)
    for r in range(T):
        for c in range(T):
            spread[:rows - r, :cols - c] |= quantized[r:, c:]
    W, H = cols // T, rows // T
    resp = [np.maximum(lut[32 * o + (spread & 15)], lut[32 * o + 16 + (spread >> 4)]) for o in range(8)]
    pad = W * H + 16 * W + 16
    planes = []
    for o in range(8):
        mems = [resp[o][rs::T, cs::T].reshape(-1) for rs in range(T) for cs in range(T)]
        planes.append(np.concatenate(mems + [np.zeros(pad, np.uint8)]))
    w, h, _, feats = template
    wf, hf = (w - 1) // T + 1, (h - 1) // T + 1
    P = (H - hf) * W + (W - wf) + 1
    dst = np.zeros(W * H, np.uint16)
    for (x, y, label) in feats:
        if x < 0 or y < 0 or x >= size[0] or y >= size[1] or P <= 0:
            continue
        a = ((y % T) * T + (x % T)) * W * H + (y // T) * W + x // T
        dst[:P] += planes[label][a:a + P]
    return dst.reshape(H, W)


def test_numpy_restatement_agrees():
    orc, views = common.build_oracle(kinds=("cg", "dn"), T=(5, 8), n_views=4, n_random=6, seed=21)
    bgr, depth, _ = synth.compose_scene(77, views[:2], rows=240, cols=320)
    orc.build_front([bgr, depth])
    lut = orc.similarity_lut()
    g = orc.geometry(1)
    for tid in range(orc.num_templates("obj")):
        tp = orc.get_template("obj", tid)
        total = np.zeros((g["H"], g["W"]), np.uint16)
        for m in range(2):
            q = orc.fetch(O.Stage.QUANTIZED, 1, m)
            total += _numpy_coarse(q, 8, lut, tp[2 + m], (g["cols"], g["rows"]))
        assert np.array_equal(total, orc.coarse_map("obj", tid)), tid


def test_planted_view_is_found_at_its_offset():
    orc = O.OracleDetector()
    views = []
    boxes = []
    for v in common.rendered_views(4, 3):
        tid, bb = orc.add_template([v[0], v[1]], "obj", v[2])
        if tid >= 0:
            views.append(v)
            boxes.append(bb)
    bgr, depth, placements = synth.compose_scene(5, views[:2], noise=False)
    m = orc.match([bgr, depth], 90.0)
    for (vi, ox, oy) in placements:
        # template origin = canvas offset + crop box origin; reported position adds T/2 + (T%2-1) = 2 at T=5
        ex, ey = ox + boxes[vi][0] + 2, oy + boxes[vi][1] + 2
        hit = m[(m["template_id"] == vi) & (np.abs(m["x"] - ex) <= 5) & (np.abs(m["y"] - ey) <= 5)]
        assert len(hit) >= 1, (vi, ex, ey, m[:5])
        assert hit["similarity"].max() > 90.0  # the silhouette gradients change against clutter instead of black


def test_sort_unique_semantics():
    """Match::operator< = similarity desc, template_id asc; operator== ignores template_id (App. A.1, D-7)."""
    recs = np.zeros(5, O.MATCH_DTYPE)
    recs["x"] = [10, 10, 10, 20, 10]
    recs["y"] = [5, 5, 5, 5, 5]
    recs["similarity"] = [90.0, 95.0, 95.0, 95.0, 90.0]
    recs["template_id"] = [7, 3, 1, 2, 7]
    out = O.sort_unique(recs)
    assert list(out["similarity"]) == [95.0, 95.0, 95.0, 90.0]
    assert list(out["template_id"][:3]) == [1, 2, 3]          # x=20 (tid 2) separates the two x=10 records
    assert out["template_id"][3] == 7 and len(out) == 4      # exact duplicates collapse


def test_single_level_keeps_coarse_score_with_half_percent():
    """L == 1: no refinement, similarity = raw*100/(4 nf) + 0.5f and no re-threshold (App. D-3)."""
    orc = O.OracleDetector([O.color_gradient()], T=(8,))
    rng = np.random.default_rng(8)
    for _ in range(12):
        orc.add_synthetic_template("obj", synth.random_pyramid(rng, T=(8,), M=1, nf0=31, wh_range=(20, 60)))
    views = common.rendered_views(2, 12)
    bgr, _, _ = synth.compose_scene(9, views, rows=240, cols=320)
    m = orc.match([bgr], 50.0, keep_candidates=True)
    cands = orc.last_candidates()
    assert len(cands) == len(orc.last_presort()) and len(m) > 0
    pre = orc.last_presort()
    want = (cands["raw"].astype(np.float32) * np.float32(100.0)) / np.float32(4 * 31) + np.float32(0.5)
    assert np.array_equal(pre["similarity"], want)
    assert np.array_equal(pre["x"], (cands["pos"] % 40) * 8 + 3)


def test_wraparound_positions_are_scored():
    """Positions with c > span_x in rows < span_y belong to the flat run and must be evaluated (App. D-1)."""
    orc = O.OracleDetector([O.color_gradient()], T=(8,))
    f = np.array([[0, 0, 0], [24, 16, 0], [8, 8, 1]], np.int32)
    orc.add_synthetic_template("obj", [(24, 16, 0, f)])
    bgr = np.zeros((64, 96, 3), np.uint8)
    bgr[:, 40:] = 255  # a vertical edge: orientation label 0 responses everywhere along it
    orc.build_front([bgr])
    cm = orc.coarse_map("obj", 0)
    W, H = 12, 8
    wf, hf = (24 - 1) // 8 + 1, (16 - 1) // 8 + 1
    P = (H - hf) * W + (W - wf) + 1
    flat = cm.reshape(-1)
    assert flat[P:].max() == 0
    lm = orc.fetch(O.Stage.LINEAR, 0, 0)
    # recompute one wrapped position by hand: row 0, column W-1 (> span_x)
    j = W - 1
    want = 0
    for (x, y, label) in f:
        a = ((y % 8) * 8 + x % 8) * W * H + (y // 8) * W + x // 8
        want += int(lm[label, a + j])
    assert flat[j] == want


def test_modality_process_is_add_template_before_cropping():
    """orc_modality_process (Modality::process + pyrDown + quantize / extractTemplate) against orc_add_template: the stored
    templates are the extracted ones shifted by the bounding-box origin, and quantize() is the masked QUANTIZED tap."""
    kinds, T = ("cg", "dn"), (5, 8)
    det = O.OracleDetector(common.oracle_modalities(kinds), T)
    checked = 0
    for (bgr, depth, mask) in common.rendered_views(4, 91, canvas=(240, 240)):
        tid, bb = det.add_template([bgr, depth], "obj", mask)
        det.build_front([bgr, depth], [mask, mask])
        for m, src in enumerate((bgr, depth)):
            for level in range(2):
                q, ok, t = det.modality_process(m, src, mask, level)
                assert np.array_equal(q, det.fetch(O.Stage.QUANTIZED, level, m))
                assert ok == (tid >= 0) or tid < 0
                if tid >= 0:
                    stored = det.get_template("obj", tid)[level * 2 + m]
                    assert t[:3] == (-1, -1, level)
                    assert np.array_equal(t[3] - np.array([bb[0] >> level, bb[1] >> level, 0]), stored[3])
                    checked += 1
    assert checked >= 8
