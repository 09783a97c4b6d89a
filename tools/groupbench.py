#!/usr/bin/env python
"""One process, N GPUs (lm_group): end-to-end frames/s of the bench workload from pinned host frames through
lm_group_match_batch_multi, for both group modes and every device count the box offers.

    gpurun --gpus 8 -- python tools/groupbench.py [--frames 512] [--reps 6] [--check 2]

One JSON line per (mode, devices).  "frames": every device holds all templates and takes its share of the frames over its own
PCIe link.  "templates": templates sharded, every device sees every frame (the north-star layout).  The first --check
frames of every configuration are compared with the single-device lists."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    import torch
    from linemod_pose_estimation_b200 import Detector, DetectorGroup, Mesh, _capi, training
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--pool", type=int, default=64)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--check", type=int, default=2)
    ap.add_argument("--batch-frames", type=int, default=8)
    args = ap.parse_args()
    n_dev = torch.cuda.device_count()
    det = Detector()
    det.set_option("batch_frames", args.batch_frames)
    cam = training.camera()
    tri = bench.meshes()
    mesh = {cid: Mesh(tri[cid]) for cid, _, _ in bench.CLASSES}
    views = bench.class_views(lambda r0, r1, rs: training.ViewSphere(radius_min=r0, radius_max=r1, radius_step=rs).views())
    for cid, _, _ in bench.CLASSES:
        det.trainViews(mesh[cid], cam, views[cid][0], views[cid][1], cid)

    def render(cid, T, up):
        r = training.render_views(det, mesh[cid], cam, T[None], up[None])
        return r["bgr"][0], r["depth"][0], r["mask"][0], tuple(int(v) for v in r["rects"][0])
    frames = bench.make_frames(render, views, args.pool)
    host = []
    for (b, d) in frames:
        pb, pd = _capi.pinned_empty(b.shape, np.uint8), _capi.pinned_empty(d.shape, np.uint16)
        pb[...] = b
        pd[...] = d
        host.append([pb, pd])
    stream = [host[i % args.pool] for i in range(args.frames)]
    import ctypes as C
    lib = _capi.lib()
    arr, _keep = _capi.image_array([a for fr in stream for a in fr])
    qarr, _qk = _capi.query_array(bench.QUERIES)
    n_q = len(bench.QUERIES)
    offs = (C.c_size_t * (args.frames * n_q + 1))()
    out_p = C.c_void_p()
    ref = det.match_batch_multi(host[:max(1, args.check)], bench.QUERIES)
    counts = [n for n in (1, 2, 4, 8) if n <= n_dev]
    for mode in ("frames", "templates"):
        for n in counts:
            group = DetectorGroup(det, list(range(n)), mode)
            got = group.match_batch_multi(stream, bench.QUERIES)       # warm-up: packs, graphs, buffers on every device
            same = all(np.array_equal(a, b) for fa, fb in zip(got[:args.check], ref) for a, b in zip(fa, fb))
            times = []
            for _ in range(args.reps):   # the C ABI call itself: pinned host frames in, host match lists out
                t0 = time.perf_counter()
                _capi.check(lib.lm_group_match_batch_multi(group._h, arr, args.frames, 2, qarr, n_q, C.byref(out_p), offs))
                times.append(time.perf_counter() - t0)
                lib.lm_free_matches(out_p)
            best, med = min(times), float(np.median(times))
            print(json.dumps({"mode": mode, "devices": n, "templates": det.numTemplates(), "frames_per_call": args.frames,
                              "fps_best": args.frames / best, "fps_median": args.frames / med, "us_per_frame_median": 1e6 * med / args.frames,
                              "identical_to_single_device": bool(same)}), flush=True)
            group.close()


if __name__ == "__main__":
    main()
