"""The CUDA path at BASELINE.json's full size -- bench.py's workload: 5 236 templates trained on the GPU from the reference's
own meshes (memoryChip2, cpu_binary), thresholds 92 / 94, 640x480 frames with rendered instances in clutter -- held to the
oracle on whole frames and to the size-independent properties of the path: the exact early terminations change nothing,
chunked batches equal single calls, the union of template shards equals the whole set, dealing frames out changes nothing."""
import os
import sys

import numpy as np
import pytest

import common
from common import O

sys.path.insert(0, common.ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workload():
    from linemod_pose_estimation_b200 import Detector, Mesh, training
    det = Detector()
    cam = training.camera()
    tri = bench.meshes()
    mesh = {cid: Mesh(tri[cid]) for cid, _, _ in bench.CLASSES}
    views = bench.class_views(lambda r0, r1, rs: training.ViewSphere(radius_min=r0, radius_max=r1, radius_step=rs).views())
    for cid, _, _ in bench.CLASSES:
        det.trainViews(mesh[cid], cam, views[cid][0], views[cid][1], cid)

    def render(cid, T, up):
        r = training.render_views(det, mesh[cid], cam, T[None], up[None])
        return r["bgr"][0], r["depth"][0], r["mask"][0], tuple(int(v) for v in r["rects"][0])
    frames = [list(f) for f in bench.make_frames(render, views, 10)]
    assert det.numTemplates() > 5000
    return det, frames


def _same(a, b, what):
    assert len(a) == len(b), what
    for qa, qb in zip(a, b):
        common.assert_matches_equal(qa, qb, what)


def test_whole_frames_equal_the_oracle(workload):
    det, frames = workload
    orc = O.OracleDetector()
    orc.set_fast(True)
    orc.set_threads(O.OracleDetector.max_threads())
    bench.copy_templates_to_oracle(det, orc)
    total = 0
    for f in frames[:3]:
        got = det.match_multi(f, bench.QUERIES)
        for (thr, ids), g in zip(bench.QUERIES, got):
            want = orc.match(f, thr, class_ids=ids, keep_candidates=True)
            common.assert_matches_equal(g, want, "thr %g" % thr)
            total += len(want)
    assert total > 20
    # candidate counts of a single-query request are the oracle's too
    det.match(frames[0], 92.0, class_ids=["memoryChip2"])
    orc.match(frames[0], 92.0, class_ids=["memoryChip2"], keep_candidates=True)
    assert det.last_work()["candidates"] == len(orc.last_candidates()) > 100


def test_early_terminations_change_nothing(workload):
    det, frames = workload
    det.set_option("prune", 3)
    ref = [det.match_multi(f, bench.QUERIES) for f in frames]
    gathered = det.last_work()["B_coarse_gathered"]
    for prune in (0, 1, 2):
        det.set_option("prune", prune)
        for f, r in zip(frames, ref):
            _same(det.match_multi(f, bench.QUERIES), r, "prune %d" % prune)
        if prune == 0:
            w = det.last_work()
            assert w["B_coarse_gathered"] == w["B_coarse"] > gathered   # exhaustive: every byte the reference loads
    det.set_option("prune", 3)


@pytest.mark.parametrize("batch_frames", [1, 8, 32])
def test_chunked_batches_equal_single_calls(workload, batch_frames):
    det, frames = workload
    ref = [det.match_multi(f, bench.QUERIES) for f in frames]
    det.set_option("batch_frames", batch_frames)
    for r, g in zip(ref, det.match_batch_multi(frames, bench.QUERIES)):
        _same(g, r, "chunks of %d" % batch_frames)
    det.set_option("batch_frames", 8)


@pytest.mark.parametrize("mode,members", [("templates", 3), ("frames", 2)])
def test_shards_and_dealt_frames_equal_the_whole(workload, mode, members):
    import torch
    from linemod_pose_estimation_b200 import DetectorGroup
    det, frames = workload
    ref = [det.match_multi(f, bench.QUERIES) for f in frames]
    n_dev = torch.cuda.device_count()
    group = DetectorGroup(det, [i % n_dev for i in range(members)], mode)
    for r, g in zip(ref, group.match_batch_multi(frames, bench.QUERIES)):
        _same(g, r, "%s x%d" % (mode, members))
    group.close()
