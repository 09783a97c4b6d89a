"""The CPU baseline's front end (oracle/linemod_fast.inc: separable integer filters, SSE response maps / ORs / median
network, threaded row bands) against the plain restatement (oracle/linemod_oracle.cpp), which stays the parity oracle:
every stage both produce must be bit-identical, for every geometry class the GPU tests use, with and without masks,
at 1 and several threads."""
import numpy as np
import pytest

import common
from linemod_pose_estimation_b200 import synth
from oracle import oracle as O
from oracle.oracle import Stage


def _same_front(plain, fast, levels, n_mod, kinds):
    for l in range(levels):
        assert plain.geometry(l) == fast.geometry(l)
        for m in range(n_mod):
            for st in (Stage.QUANT_RAW, Stage.QUANTIZED, Stage.SPREAD, Stage.RESPONSE, Stage.LINEAR):
                assert np.array_equal(plain.fetch(st, l, m), fast.fetch(st, l, m)), (st, l, m)
            if kinds[m] == "cg":
                a, b = plain.fetch(Stage.MAGNITUDE, l, m), fast.fetch(Stage.MAGNITUDE, l, m)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), ("magnitude", l, m)


@pytest.mark.parametrize("rows,cols,T,kinds,threads", [
    (480, 640, (5, 8), ("cg", "dn"), 4),
    (240, 320, (4, 8), ("dn", "cg"), 1),
    (96, 160, (2, 4, 8), ("cg", "dn"), 3),
    (100, 180, (5,), ("cg",), 2),
    (64, 48, (4,), ("dn",), 5),
])
def test_fast_front_end_equals_plain(rows, cols, T, kinds, threads):
    plain = O.OracleDetector(common.oracle_modalities(kinds), T)
    fast = O.OracleDetector(common.oracle_modalities(kinds), T)
    fast.set_fast(True)
    fast.set_threads(threads)
    views = common.rendered_views(3, 41, canvas=(80, 80))
    for seed in (1001, 1002):
        bgr, depth, _ = synth.compose_scene(seed, views, rows=rows, cols=cols)
        src = common.sources_for(kinds, bgr, depth)
        plain.build_front(src)
        fast.build_front(src)
        _same_front(plain, fast, len(T), len(kinds), kinds)


def test_fast_front_end_with_masks_and_match_lists():
    kinds, T = ("cg", "dn"), (5, 8)
    plain, fast = O.OracleDetector(), O.OracleDetector()
    fast.set_fast(True)
    fast.set_threads(4)
    plain.set_threads(4)
    views = common.rendered_views(6, 43)
    for (b, d, m) in views:
        assert plain.add_template([b, d], "obj", m)[0] == fast.add_template([b, d], "obj", m)[0]
    bgr, depth, _ = synth.compose_scene(7, views[:3])
    rng = np.random.default_rng(2)
    m0 = (rng.random((480, 640)) < 0.7).astype(np.uint8) * 255
    m1 = np.zeros((480, 640), np.uint8)
    m1[100:400, 150:600] = 1
    plain.build_front([bgr, depth], masks=[m0, m1])
    fast.build_front([bgr, depth], masks=[m0, m1])
    _same_front(plain, fast, 2, 2, kinds)
    for thr in (90.0, 70.0):
        a, b = plain.match([bgr, depth], thr), fast.match([bgr, depth], thr)
        common.assert_matches_equal(a, b)
    assert len(a) > 0


def test_fast_median_network_on_random_bytes():
    """The baseline's pminub / pmaxub median-of-25 against numpy on random bytes (arbitrary LUT values, borders)."""
    rng = np.random.default_rng(5)
    det = O.OracleDetector([O.depth_normal()], (4,))
    img = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    want = O.prim_median5(img)
    pad = np.pad(img, 2, mode="edge")
    win = np.stack([pad[j:j + 37, i:i + 53] for j in range(5) for i in range(5)])
    assert np.array_equal(want, np.sort(win, axis=0)[12])
    assert np.array_equal(O.prim_median5_fast(img), want)
