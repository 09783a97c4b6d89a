#!/usr/bin/env python
"""A/B tool for GPU sessions: the bench workload (bench.py, configs[1], trained templates) under different chunk sizes /
lane counts / kernel options.

    python tools/kbench.py [--configs "batch_frames=8,streams=4;batch_frames=1,streams=8;..."] [--frames 1024]

Per configuration one JSON line: device-resident us per frame (CUDA events around lm_match_device_stream), end-to-end us
per frame (lm_match_batch_multi from pinned host frames), per-stage us per frame of a chunk (timing option) and the coarse
kernel's gathered fraction.  Keys other than `streams` are lm_set_option options.  Checks that every configuration returns
the first one's match lists."""
import argparse
import ctypes as C
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
    from linemod_pose_estimation_b200 import Detector, Mesh, _capi, training
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="batch_frames=8,streams=4;batch_frames=1,streams=8;batch_frames=4,streams=4;"
                                         "batch_frames=16,streams=3;batch_frames=8,streams=4,mod_order=0;batch_frames=8,streams=4,prune=0")
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--workload", default="trained", choices=["trained", "stress"],
                    help="trained: bench.py's workload; stress: round 1's (24 extracted + 2 628 random stress templates per class, "
                         "SURVEY 8d's random-template set) for like-for-like comparisons with profiles/r01_*")
    ap.add_argument("--pool", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    det = Detector()
    if args.workload == "trained":
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
    else:
        from linemod_pose_estimation_b200 import synth
        sv = {}
        for ci, (cid, _, _) in enumerate(bench.CLASSES):
            sv[cid] = [synth.render_view(s, scale, rot, canvas=(200, 200), tilt=tilt)
                       for (s, scale, rot, tilt) in synth.view_params(24, seed=900 + ci)]
            n_ok = sum(det.addTemplate([b, d], cid, m)[0] >= 0 for (b, d, m) in sv[cid])
            rng = np.random.default_rng(4242 + ci)
            for _ in range(2652 - n_ok):
                det.addSyntheticTemplate(synth.random_pyramid(rng), cid)
        planted = [sv[cid][k] for cid, _, _ in bench.CLASSES for k in (0, 1)]
        frames = [synth.compose_scene(2000 + i, planted, rows=bench.ROWS, cols=bench.COLS)[:2] for i in range(args.pool)]
    lib = _capi.lib()
    host = []
    for (b, d) in frames:
        pb, pd = _capi.pinned_empty(b.shape, np.uint8), _capi.pinned_empty(d.shape, np.uint16)
        pb[...] = b
        pd[...] = d
        host.append((pb, pd))
    dev_frames = [(torch.from_numpy(b).to(dev), torch.from_numpy(d.view(np.int16)).to(dev)) for (b, d) in frames]
    qarr, _qk = _capi.query_array(bench.QUERIES)
    n_q = len(bench.QUERIES)
    flat_dev = [p for (fb, fd) in dev_frames for p in (fb.data_ptr(), fd.data_ptr())]
    dev_ptrs = (C.c_void_p * len(flat_dev))(*flat_dev)
    harr, _hk = _capi.image_array([a for fr in host for a in fr])
    out_p = C.c_void_p()
    offs = (C.c_size_t * (args.pool * n_q + 1))()
    ref = None
    defaults = {"batch_frames": 8, "prune": 3, "mod_order": 2, "graphs": 1, "coarse_grid_limit": 0, "refine_tiled": 1, "coarse_narrow": 1, "coarse_share": 1, "dn_count": 1}
    for cfg in args.configs.split(";"):
        opts = dict(defaults)
        opts.update({k: int(v) for k, v in (kv.split("=") for kv in cfg.split(","))})
        n_streams = opts.pop("streams", 4)
        for k, v in opts.items():
            det.set_option(k, v)
        det.set_option("batch_lanes", n_streams)
        det.set_option("stream_frames", opts["batch_frames"])
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        sp = (C.c_void_p * n_streams)(*[s.cuda_stream for s in streams])

        def device_pass():
            _capi.check(lib.lm_match_device_stream(det._h, dev_ptrs, args.pool, 2, bench.ROWS, bench.COLS, qarr, n_q, sp, n_streams, None, 0))
        for _ in range(3):
            device_pass()
        torch.cuda.synchronize()
        reps = max(1, args.frames // args.pool)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for s in streams:
            s.wait_stream(cur)
        for _ in range(reps):
            device_pass()
        for s in streams:
            cur.wait_stream(s)
        e1.record(cur)
        torch.cuda.synchronize()
        dev_us = 1e3 * e0.elapsed_time(e1) / (reps * args.pool)

        def host_pass(keep=False):
            _capi.check(lib.lm_match_batch_multi(det._h, harr, args.pool, 2, qarr, n_q, C.byref(out_p), offs))
            if keep:
                allm = det._take(out_p, offs[args.pool * n_q])
                return [allm[offs[i]:offs[i + 1]] for i in range(args.pool * n_q)]
            lib.lm_free_matches(out_p)
        res = host_pass(True)
        if ref is None:
            ref = res
        else:
            assert all(np.array_equal(a, b) for a, b in zip(ref, res)), "configuration %s disagrees with the first one" % cfg
        host_pass()
        t0 = time.perf_counter()
        for _ in range(reps):
            host_pass()
        e2e_us = 1e6 * (time.perf_counter() - t0) / (reps * args.pool)
        # the same through lm_stream: the pipeline stays alive between pushes of `pool` frames
        sh = C.c_void_p()
        _capi.check(lib.lm_stream_open(det._h, qarr, n_q, C.byref(sh)))
        soffs = (C.c_size_t * (4 * args.pool * n_q + 1))()
        n_pop = C.c_int()

        def stream_pass(n_push):
            for _ in range(n_push):
                _capi.check(lib.lm_stream_push(sh, harr, args.pool, 2))
                _capi.check(lib.lm_stream_pop(sh, 0, 4 * args.pool, C.byref(out_p), soffs, C.byref(n_pop)))
                lib.lm_free_matches(out_p)
            while lib.lm_stream_in_flight(sh) > 0:
                _capi.check(lib.lm_stream_pop(sh, 1, 4 * args.pool, C.byref(out_p), soffs, C.byref(n_pop)))
                lib.lm_free_matches(out_p)
        stream_pass(1)
        t0 = time.perf_counter()
        stream_pass(reps)
        stream_us = 1e6 * (time.perf_counter() - t0) / (reps * args.pool)
        lib.lm_stream_close(sh)
        one, _k1 = _capi.image_array(list(host[0]))
        offs1 = (C.c_size_t * (n_q + 1))()
        lat = []
        for i in range(60):     # one blocking call per frame: the ROS service's shape
            t0 = time.perf_counter()
            _capi.check(lib.lm_match_batch_multi(det._h, one, 1, 2, qarr, n_q, C.byref(out_p), offs1))
            lat.append(time.perf_counter() - t0)
            lib.lm_free_matches(out_p)
        single_us = 1e6 * float(np.median(lat[10:]))
        det.set_option("timing", 1)
        det.set_option("batch_lanes", 1)    # one chunk at a time: stage times of a chunk's kernels running alone
        host_pass()
        t, w = det.last_timings(), det.last_work()
        det.set_option("timing", 0)
        fr = max(1, w["frames"])
        out = {"workload": args.workload, "config": cfg, "templates": det.numTemplates(), "device_us_per_frame": round(dev_us, 2), "e2e_us_per_frame": round(e2e_us, 2), "stream_e2e_us_per_frame": round(stream_us, 2),
               "single_call_us": round(single_us, 1), "chunk_frames": fr, "launches_per_chunk": t["launches"],
               "gathered_frac": round(w["B_coarse_gathered"] / max(1, w["B_coarse"]), 4), "candidates_per_frame": w["candidates"] / fr,
               "matches_per_frame": sum(len(x) for x in res) / args.pool}
        out.update({k + "_us_per_frame": round(1e3 * t[k] / fr, 2) for k in ("h2d", "front", "coarse", "refine", "d2h")})
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
