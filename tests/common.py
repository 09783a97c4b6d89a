"""Shared helpers for the test-suite: seeded scenes / templates, oracle <-> product plumbing."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from linemod_pose_estimation_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def oracle_modalities(kinds):
    return [O.color_gradient() if k == "cg" else O.depth_normal() for k in kinds]


def product_modalities(kinds):
    from linemod_pose_estimation_b200 import ColorGradient, DepthNormal
    return [ColorGradient() if k == "cg" else DepthNormal() for k in kinds]


def sources_for(kinds, bgr, depth):
    return [bgr if k == "cg" else depth for k in kinds]


def rendered_views(n, seed, canvas=(240, 240)):
    out = []
    for (s, scale, rot, tilt) in synth.view_params(n, seed=seed):
        out.append(synth.render_view(s, scale, rot, canvas=canvas, tilt=tilt))
    return out


def build_oracle(kinds=("cg", "dn"), T=(5, 8), n_views=8, n_random=24, seed=5, classes=("obj",), canvas=(240, 240)):
    """Oracle detector with `n_views` extracted templates + `n_random` random stress templates per class.
    Returns (oracle, views) where views are the renders whose extraction succeeded (for planting)."""
    det = O.OracleDetector(oracle_modalities(kinds), T)
    views = []
    rng = np.random.default_rng(seed)
    for ci, cid in enumerate(classes):
        for (bgr, depth, mask) in rendered_views(n_views, seed + 31 * ci, canvas):
            tid, _ = det.add_template(sources_for(kinds, bgr, depth), cid, mask)
            if tid >= 0:
                views.append((bgr, depth, mask))
        for _ in range(n_random):
            det.add_synthetic_template(cid, synth.random_pyramid(rng, T=T, M=len(kinds)))
    return det, views


def copy_templates(orc, dst):
    """Copies every template of an oracle detector into a product Detector (same class ids, same order)."""
    for cid in orc.class_ids():
        for tid in range(orc.num_templates(cid)):
            got = dst.addSyntheticTemplate(orc.get_template(cid, tid), cid)
            assert got == tid


def assert_matches_equal(got, want, what="matches"):
    assert len(got) == len(want), "%s: %d vs %d" % (what, len(got), len(want))
    for name in ("x", "y", "template_id", "class_index"):
        assert np.array_equal(got[name], want[name]), "%s differ in %s" % (what, name)
    # similarity must be the same f32 bit pattern
    assert np.array_equal(got["similarity"].view(np.uint32), want["similarity"].view(np.uint32)), "%s differ in similarity bits" % what


def survey_similarity_lut():
    """The alternative SIMILARITY_LUT of SURVEY.md A.5 (asymmetric wrap-around), used to show parity holds for any table."""
    lut = np.zeros(256, np.uint8)
    for i in range(8):
        for h in range(2):
            for v in range(16):
                best = 0
                for b in range(4):
                    if v & (1 << b):
                        j = 4 * h + b
                        best = max(best, max(0, 4 - min(abs(i - j), 8 - (i - j))))
                lut[32 * i + 16 * h + v] = best
    return lut
