#!/usr/bin/env python
"""The matcher on TRAINED template sets (what DESIGN.md section 8 names as the next bench workload): the reference's own
2 652 views of config/stl/boxNew.stl (tests/golden/renderer_params_boxnew.npz) plus the same views of a second mesh, trained
with lm_train_views, matched at the reference's thresholds (92 / 94) against 640x480 frames that hold rendered instances
of both objects in clutter.

    python tools/realbench.py [--frames 32] [--reps 6] [--check 2]

Prints one JSON line: templates, frames/s through lm_match_batch_multi from pinned host frames, matches and coarse
candidates per frame, and whether the match lists of the first --check frames equal the CPU oracle's (same templates)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from linemod_pose_estimation_b200 import Detector, Mesh, _capi, synth, training  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)
import common  # noqa: E402
import test_oracle_render as golden  # noqa: E402

QUERIES = [(92.0, ["boxNew"]), (94.0, ["bracket"])]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--check", type=int, default=2)
    a = ap.parse_args()
    G = golden.G
    views, idx = golden._oracle_views()
    cam = golden._golden_camera(training)
    T = np.array([views[i][0] for i in idx])
    up = np.array([views[i][1] for i in idx])
    meshes = {"boxNew": Mesh(G["triangles"]), "bracket": Mesh(synth.bracket_mesh())}
    det = Detector()
    t0 = time.perf_counter()
    ok = {cid: det.trainViews(m, cam, T, up, cid)[0] >= 0 for cid, m in meshes.items()}
    t_train = time.perf_counter() - t0
    rng = np.random.default_rng(7)
    frames, planted = [], 0
    for f in range(a.frames):
        bgr, depth = synth.make_background(4000 + f)
        bgr = np.clip(bgr, 0, 255).astype(np.uint8)
        depth = np.clip(depth, 1, 65535).astype(np.uint16)
        for cid, m in meshes.items():
            for _ in range(2):
                k = int(rng.choice(np.flatnonzero(ok[cid])))
                r = training.render_views(det, m, cam, T[k:k + 1], up[k:k + 1])
                x, y, w, h = (int(v) for v in r["rects"][0])
                dx = int(rng.integers(-x, cam.width - (x + w) + 1))
                dy = int(rng.integers(-y, cam.height - (y + h) + 1))
                ys, xs = np.nonzero(r["mask"][0])
                bgr[ys + dy, xs + dx] = r["bgr"][0][ys, xs]
                depth[ys + dy, xs + dx] = r["depth"][0][ys, xs]
                planted += 1
        pb, pd = _capi.pinned_empty(bgr.shape, np.uint8), _capi.pinned_empty(depth.shape, np.uint16)
        pb[...] = bgr
        pd[...] = depth
        frames.append([pb, pd])
    res = det.match_batch_multi(frames, QUERIES)          # warm-up: pack upload, graph capture
    times = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        res = det.match_batch_multi(frames, QUERIES)
        times.append(time.perf_counter() - t0)
    det.match_multi(frames[0], QUERIES)
    work, tm = det.last_work(), det.last_timings()
    same = None
    if a.check:
        orc = O.OracleDetector()
        orc.set_threads(O.OracleDetector.max_threads())
        common.copy_templates(det_to_oracle_source(det), orc_adapter(orc))
        same = True
        for f in range(min(a.check, a.frames)):
            for q, (thr, ids) in enumerate(QUERIES):
                want = orc.match(frames[f], thr, class_ids=ids)
                got = res[f][q]
                same &= len(got) == len(want) and all(np.array_equal(got[n], want[n]) for n in ("x", "y", "template_id", "class_index")) \
                    and np.array_equal(got["similarity"].view(np.uint32), want["similarity"].view(np.uint32))
    best = min(times)
    print(json.dumps({"workload": "trained sets: boxNew.stl (the reference's 2652 views) + bracket mesh (same views), thr 92/94, "
                                  "640x480 frames with 4 rendered instances in clutter",
                      "templates": det.numTemplates(), "per_class": {c: int(v.sum()) for c, v in ok.items()}, "train_s": t_train,
                      "frames": a.frames, "fps_e2e_batch": a.frames / best, "ms_per_frame": 1e3 * best / a.frames,
                      "matches_per_frame": float(np.mean([sum(len(q) for q in fr) for fr in res])),
                      "coarse_candidates_frame0": int(work["candidates"]), "gathered_frac_frame0": work["B_coarse_gathered"] / max(1, work["B_coarse"]),
                      "stage_ms_frame0": {k: tm[k] for k in ("front", "coarse", "refine")},
                      "identical_to_oracle_on_checked_frames": same}))


class det_to_oracle_source:
    """common.copy_templates reads (class_ids, num_templates, get_template) from its first argument."""

    def __init__(self, det):
        self.det = det

    def class_ids(self):
        return self.det.classIds()

    def num_templates(self, cid):
        return self.det.numTemplates(cid)

    def get_template(self, cid, tid):
        return self.det.getTemplates(cid, tid)


class orc_adapter:
    """... and writes through addSyntheticTemplate(templates, class_id) on its second."""

    def __init__(self, orc):
        self.orc = orc

    def addSyntheticTemplate(self, templates, cid):
        return self.orc.add_synthetic_template(cid, templates)


if __name__ == "__main__":
    main()
