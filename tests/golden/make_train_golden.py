#!/usr/bin/env python
"""Freezes the template-generation path on a small case: 12 views of synth.gear_mesh() rendered by the scalar oracle
(oracle/render_oracle.cpp) and passed through the oracle's addTemplate -> tests/golden/train_gear.npz (silhouette boxes,
centre depths, per-view template ids and every template's features).  Regression fixture for the oracle (CPU test) and for
lm_train_views (GPU test).

    python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from linemod_pose_estimation_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

SPHERE = (6, 80, 0.35, 0.45, 0.1)     # n_points, angle_step, radius_min, radius_max, radius_step -> 6 x 3 x 2 = 36 views
PICK = list(range(0, 36, 3))           # 12 of them
CAMERA = (320, 240, 420.0, 415.0)


def build():
    tri = synth.gear_mesh() * 1.6
    vs = O.view_sphere(*SPHERE)
    cam = O.camera(*CAMERA)
    orc = O.OracleDetector()
    out = {"triangles": tri, "T": [], "up": [], "rect": [], "centre_mm": [], "tid": [], "hdr": [], "feats": []}
    for k in PICK:
        T, up = O.view_params(vs, k)[:2]
        bgr, depth, mask, rect = O.render(tri, cam, T, up)
        tid, bb = orc.add_template([bgr, depth], "gear", mask)
        out["T"].append(T); out["up"].append(up); out["rect"].append(rect)
        out["centre_mm"].append(int(depth[CAMERA[1] // 2, CAMERA[0] // 2])); out["tid"].append(tid)
        if tid >= 0:
            for (w, h, lvl, f) in orc.get_template("gear", tid):
                out["hdr"].append((tid, w, h, lvl, len(f)))
                out["feats"].append(f)
    return {"triangles": tri, "T": np.array(out["T"]), "up": np.array(out["up"]), "rect": np.array(out["rect"], np.int32),
            "centre_mm": np.array(out["centre_mm"], np.int32), "tid": np.array(out["tid"], np.int32),
            "hdr": np.array(out["hdr"], np.int32), "feats": np.concatenate(out["feats"]).astype(np.int32)}


if __name__ == "__main__":
    d = build()
    path = os.path.join(HERE, "train_gear.npz")
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path), "bytes; template ids", d["tid"].tolist())
