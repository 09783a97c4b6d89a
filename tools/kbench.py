#!/usr/bin/env python
"""Per-stage device timings of the bench workload for every kernel variant (A/B tool for GPU sessions).

    python tools/kbench.py [--templates-per-class 2652] [--frames 64] [--variants 0,1,2]

Uses lm_last_timings (CUDA events on the library's stream) around lm_match_multi; prints one line per variant and checks
that every variant returns the same match lists as variant 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from linemod_pose_estimation_b200 import Detector, _capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--templates-per-class", type=int, default=bench.TEMPLATES_PER_CLASS)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--configs", default="coarse_variant=0,prune=1;coarse_variant=0,prune=0;coarse_variant=2;coarse_variant=1",
                    help="';'-separated configurations, each a ','-separated list of lm_set_option key=value pairs")
    ap.add_argument("--threshold-scale", type=float, default=1.0, help="scales the queries' thresholds (pruning sensitivity)")
    args = ap.parse_args()
    views = bench.rendered_views()
    det = Detector()
    bench.fill_templates(lambda cid, b, d, m: det.addTemplate([b, d], cid, m)[0],
                         lambda cid, pyr: det.addSyntheticTemplate(pyr, cid), views, args.templates_per_class)
    frames = bench.make_frames(views, min(args.frames, 32))
    host = []
    for (b, d) in frames:
        pb, pd = _capi.pinned_empty(b.shape, np.uint8), _capi.pinned_empty(d.shape, np.uint16)
        pb[...] = b
        pd[...] = d
        host.append((pb, pd))
    ref = None
    queries = [(thr * args.threshold_scale, ids) for thr, ids in bench.QUERIES]
    for cfg in args.configs.split(";"):
        opts = dict(kv.split("=") for kv in cfg.split(","))
        for k, v in opts.items():
            det.set_option(k, int(v))
        v = cfg
        res = [det.match_multi(list(host[i % len(host)]), queries) for i in range(len(host))]
        if ref is None:
            ref = res
        else:
            for a, b in zip(ref, res):
                for qa, qb in zip(a, b):
                    assert np.array_equal(qa, qb), "configuration %s disagrees with the first one" % v
        stages = {k: [] for k in ("h2d", "front", "coarse", "refine", "d2h")}
        for i in range(args.frames):
            det.match_multi(list(host[i % len(host)]), queries)
            t = det.last_timings()
            for k in stages:
                stages[k].append(t[k])
        w = det.last_work()
        out = {"config": cfg, "templates": det.numTemplates(), "launches": t["launches"], "B_coarse": w["B_coarse"],
               "gathered_frac": round(w["B_coarse_gathered"] / max(1, w["B_coarse"]), 4), "candidates": w["candidates"],
               "matches": int(sum(len(q) for q in res[-1]))}
        out.update({k + "_us": round(1e3 * float(np.median(x)), 2) for k, x in stages.items()})
        out["coarse_GBps"] = round(w["B_coarse"] / (np.median(stages["coarse"]) * 1e-3) / 1e9, 1)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
