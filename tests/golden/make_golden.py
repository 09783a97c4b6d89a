"""Generates the committed golden fixtures.  Run from the repo root in the build container:

    python tests/golden/make_golden.py

1. primitives_cv2.npz -- outputs of the in-container OpenCV (cv2 4.13) for the imgproc primitives the LINEMOD path
   is built on (GaussianBlur 7x7, Sobel 3x3, phase, pyrDown, medianBlur 5, NN resize, erode, distanceTransform) on
   seeded inputs.  These are REAL OpenCV outputs: they pin the oracle's primitives (tests/test_oracle_primitives.py).
2. oracle_scene.npz -- the oracle's own outputs on a seeded scene (stage hashes + match list).  The reference's hot
   path (OpenCV 2.4.x linemod.cpp) is not available anywhere in this environment, so this fixture is a regression pin
   of the restatement, not a reference output; the CUDA path is compared against it on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HERE = os.path.dirname(os.path.abspath(__file__))


def primitives():
    import cv2
    rng = np.random.default_rng(20261018)
    noise = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (48, 64, 3), dtype=np.uint8), (9, 9), 3)
    out = {"in_noise": noise, "in_smooth": smooth}
    for name, im in (("noise", noise), ("smooth", smooth)):
        g = cv2.GaussianBlur(im, (7, 7), 0, borderType=cv2.BORDER_REPLICATE)
        out["gauss_" + name] = g
        out["sobelx_" + name] = cv2.Sobel(g, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE)
        out["sobely_" + name] = cv2.Sobel(g, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE)
        out["pyrdown_" + name] = cv2.pyrDown(im)
        g1 = np.ascontiguousarray(im[..., 0])
        out["median_" + name] = cv2.medianBlur(g1, 5)
        out["nn_" + name] = cv2.resize(g1, (32, 24), interpolation=cv2.INTER_NEAREST)
        m = (g1 > 110).astype(np.uint8) * 255
        out["mask_" + name] = m
        out["erode1_" + name] = cv2.erode(m, None, iterations=1, borderType=cv2.BORDER_REPLICATE)
        out["erode2_" + name] = cv2.erode(m, None, iterations=2, borderType=cv2.BORDER_REPLICATE)
        m2 = (cv2.GaussianBlur(g1, (15, 15), 5) > 120).astype(np.uint8)
        out["dtin_" + name] = m2
        out["dist_" + name] = cv2.distanceTransform(m2, cv2.DIST_C, 3)
    gx = rng.integers(-1020, 1021, 4096).astype(np.float32)
    gy = rng.integers(-1020, 1021, 4096).astype(np.float32)
    gx[:8] = [0, 0, 1, -1, 0, 5, -5, 7]
    gy[:8] = [0, 1, 0, 0, -1, 5, 5, -7]
    cv2.setUseOptimized(False)  # the non-FMA evaluation == OpenCV 2.4.8's SSE2 fastAtan2
    out["phase_x"], out["phase_y"] = gx, gy
    out["phase_deg"] = cv2.phase(gx, gy, angleInDegrees=True).ravel()
    cv2.setUseOptimized(True)
    np.savez_compressed(os.path.join(HERE, "primitives_cv2.npz"), **out)
    print("primitives_cv2.npz:", len(out), "arrays")


def oracle_scene():
    import common
    from oracle import oracle as O
    orc, views = common.build_oracle(kinds=("cg", "dn"), T=(5, 8), n_views=8, n_random=16, seed=5)
    bgr, depth, _ = common.synth.compose_scene(1001, views[:4], rows=240, cols=320)
    matches = orc.match([bgr, depth], 80.0, keep_candidates=True)
    out = {"bgr": bgr, "depth": depth, "matches": matches, "presort": orc.last_presort(),
           "candidates": orc.last_candidates()}
    hashes = []
    for l in range(2):
        for m in range(2):
            for st in (O.Stage.QUANTIZED, O.Stage.SPREAD, O.Stage.RESPONSE, O.Stage.LINEAR):
                hashes.append("%d/%d/%d:%s" % (l, m, st, common.sha(orc.fetch(st, l, m))))
    out["stage_hashes"] = np.array(hashes)
    out["quant_l0_cg"] = orc.fetch(O.Stage.QUANTIZED, 0, 0)
    out["quant_l0_dn"] = orc.fetch(O.Stage.QUANTIZED, 0, 1)
    tpl = []
    for tid in range(orc.num_templates("obj")):
        for (w, h, lvl, f) in orc.get_template("obj", tid):
            tpl.append(np.concatenate([[w, h, lvl, len(f)], f.ravel()]))
    out["templates_flat"] = np.concatenate(tpl).astype(np.int32)
    out["n_templates"] = np.array([orc.num_templates("obj")])
    np.savez_compressed(os.path.join(HERE, "oracle_scene.npz"), **out)
    print("oracle_scene.npz: %d matches, %d candidates, %d templates" % (len(matches), len(out["candidates"]), out["n_templates"][0]))


if __name__ == "__main__":
    primitives()
    oracle_scene()
