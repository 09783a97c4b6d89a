#!/usr/bin/env python
"""Throughput of the template-generation path on the reference's own training run: the 2 652 valid views of
config/stl/boxNew.stl recorded in tests/golden/renderer_params_boxnew.npz (640x480, ColorGradient + DepthNormal, T = {5, 8}).

    python tools/trainbench.py [--cpu-views 48]

Prints one JSON line: views/s of lm_train_views (render + quantise + extract on the GPU), of the sequential product path
(lm_render_views + one lm_add_template per view: GPU quantisation, host feature selection) and of the CPU oracle (scalar
render + addTemplate) on a sample, and whether the three produced identical templates on that sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from linemod_pose_estimation_b200 import Detector, Mesh, training  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker / CPU baseline only)
import test_oracle_render as golden  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-views", type=int, default=48)
    args = ap.parse_args()
    G = golden.G
    views, idx = golden._oracle_views()
    cam, ocam = golden._golden_camera(training), golden._golden_camera(O)
    T = np.array([views[i][0] for i in idx])
    up = np.array([views[i][1] for i in idx])
    mesh = Mesh(G["triangles"])
    det = Detector()
    det.trainViews(mesh, cam, T[:64], up[:64], "warm")   # allocations, first-launch costs
    det = Detector()
    t0 = time.perf_counter()
    tids, bbs, rects = det.trainViews(mesh, cam, T, up, "obj")
    t_gpu = time.perf_counter() - t0
    n = args.cpu_views
    sample = np.linspace(0, len(T) - 1, n).astype(int)
    seq = Detector()
    t0 = time.perf_counter()
    r = training.render_views(seq, mesh, cam, T[sample], up[sample])
    for v in range(n):
        seq.addTemplate([r["bgr"][v], r["depth"][v]], "obj", r["mask"][v])
    t_seq = time.perf_counter() - t0
    orc = O.OracleDetector()
    t0 = time.perf_counter()
    for v in sample:
        bgr, depth, mask, rect = O.render(G["triangles"], ocam, T[v], up[v])
        orc.add_template([bgr, depth], "obj", mask)
    t_cpu = time.perf_counter() - t0
    same = True
    k = 0
    for j, v in enumerate(sample):
        if tids[v] < 0:
            continue
        a, b, c = det.getTemplates("obj", int(tids[v])), orc.get_template("obj", k), seq.getTemplates("obj", k)
        k += 1
        for x, y, z in zip(a, b, c):
            same &= x[:3] == y[:3] == z[:3] and np.array_equal(x[3], y[3]) and np.array_equal(x[3], z[3])
    print(json.dumps({"workload": "boxNew.stl, 2652 views of the reference's training run, 640x480, CG+DN, T={5,8}",
                      "views": len(T), "templates": int((tids >= 0).sum()),
                      "gpu_batched_views_per_s": len(T) / t_gpu, "gpu_batched_s": t_gpu,
                      "gpu_sequential_views_per_s": n / t_seq, "cpu_oracle_views_per_s": n / t_cpu, "sample": n,
                      "identical_templates_on_sample": bool(same)}))


if __name__ == "__main__":
    main()
