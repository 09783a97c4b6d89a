#!/usr/bin/env python
"""The reference's trainer (renderer_node, /root/reference/src/renderer.cpp:170-354) on this library: every view of the
view sphere around an STL mesh is rendered and handed to addTemplate on the GPU, then templates.yml (writeLinemod) and
renderer_params.yml (writeLinemodTemplateParams) are written in the reference's formats.

    python tools/train.py --stl mesh.stl --templates out_templates.yml --params out_renderer_params.yml
                          [--fx 535.566011 --fy 537.168115 --width 640 --height 480]
                          [--n-points 150 --angle-step 10 --radius-min 0.5 --radius-max 1.0 --radius-step 0.1] [--class-id obj]

Defaults are the reference's (renderer.cpp:199-213).  The reference's fork of the renderer additionally drops views it
considers invalid for planar objects (`is_restricted` / `is_image_valid`, renderer.cpp:249-255); that code is not part of
the reference tree, so every view whose extraction succeeds becomes a template here.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from linemod_pose_estimation_b200 import Detector, Mesh, ViewSphere, camera, training  # noqa: E402
from linemod_pose_estimation_b200._capi import LmRendererParams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stl", required=True)
    ap.add_argument("--templates", required=True)
    ap.add_argument("--params", required=True)
    ap.add_argument("--fx", type=float, default=535.566011)
    ap.add_argument("--fy", type=float, default=537.168115)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--near", type=float, default=0.1)
    ap.add_argument("--far", type=float, default=1000.0)
    ap.add_argument("--n-points", type=int, default=150)
    ap.add_argument("--angle-step", type=int, default=10)
    ap.add_argument("--radius-min", type=float, default=0.5)
    ap.add_argument("--radius-max", type=float, default=1.0)
    ap.add_argument("--radius-step", type=float, default=0.1)
    ap.add_argument("--class-id", default="obj")
    a = ap.parse_args()
    mesh = Mesh.load_stl(a.stl)
    cam = camera(a.width, a.height, a.fx, a.fy, a.near, a.far)
    vs = ViewSphere(a.n_points, a.angle_step, a.radius_min, a.radius_max, a.radius_step)
    det = Detector()   # ColorGradient + DepthNormal, T = {5, 8} (renderer.cpp:179-185)
    n = len(vs)
    views = [vs.view(i) for i in range(n)]
    T, up = np.array([v[0] for v in views]), np.array([v[1] for v in views])
    t0 = time.perf_counter()
    tids, bbs, rects, centre = det.trainViews(mesh, cam, T, up, a.class_id, centre_depth=True)
    dt = time.perf_counter() - t0
    ok = np.flatnonzero(tids >= 0)
    # like the reference, the recorded rectangle keeps GL's bottom-left origin and one pixel of margin
    rr = np.array([(r["x"] - 1, a.height - (r["y"] + r["height"]) - 1, r["width"] + 2, r["height"] + 2) for r in rects[ok]])
    poses = training.poses_for_views(T[ok], up[ok], cam, [views[i][2] for i in ok], rr.reshape(-1, 4), centre[ok])
    det.write(a.templates)
    training.write_renderer_params(a.params, poses, LmRendererParams(a.n_points, a.angle_step, a.radius_min, a.radius_max,
                                                                      a.radius_step, a.width, a.height, a.fx, a.fy, a.near, a.far))
    print(json.dumps({"triangles": len(mesh), "views": n, "templates": int(len(ok)), "train_s": dt, "views_per_s": n / dt,
                      "templates_file": a.templates, "params_file": a.params}))


if __name__ == "__main__":
    main()
