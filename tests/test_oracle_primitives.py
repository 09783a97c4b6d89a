"""The oracle's imgproc primitives against real OpenCV outputs: committed fixtures (tests/golden/primitives_cv2.npz,
made by tests/golden/make_golden.py with cv2 4.13) and, when cv2 is importable, live on fresh seeds.
These are the routines [OCV] linemod.cpp delegates to cv:: (App. C probes)."""
import os

import numpy as np
import pytest

import common
from common import O

G = np.load(os.path.join(common.GOLDEN, "primitives_cv2.npz"))


@pytest.mark.parametrize("name", ["noise", "smooth"])
def test_golden_primitives(name):
    im = G["in_" + name]
    g = O.prim_gaussian7(im)
    assert np.array_equal(g, G["gauss_" + name])
    dx, dy = O.prim_sobel3(g)
    assert np.array_equal(dx, G["sobelx_" + name]) and np.array_equal(dy, G["sobely_" + name])
    assert np.array_equal(O.prim_pyrdown(im), G["pyrdown_" + name])
    g1 = np.ascontiguousarray(im[..., 0])
    assert np.array_equal(O.prim_median5(g1), G["median_" + name])
    assert np.array_equal(O.prim_nn_half(g1), G["nn_" + name])
    m = G["mask_" + name]
    assert np.array_equal(O.prim_erode3(m, 1), G["erode1_" + name])
    assert np.array_equal(O.prim_erode3(m, 2), G["erode2_" + name])
    assert np.array_equal(O.prim_distance_c3(G["dtin_" + name]), G["dist_" + name])


def test_golden_phase_bits():
    got = O.prim_phase_deg(G["phase_x"], G["phase_y"])
    assert np.array_equal(got.view(np.uint32), G["phase_deg"].view(np.uint32))


def test_live_cv2_primitives():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(77)
    for shape in ((37, 53, 3), (64, 80, 3)):
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        ref = cv2.GaussianBlur(im, (7, 7), 0, borderType=cv2.BORDER_REPLICATE)
        assert np.array_equal(O.prim_gaussian7(im), ref)
        dx, dy = O.prim_sobel3(ref)
        assert np.array_equal(dx, cv2.Sobel(ref, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE))
        assert np.array_equal(dy, cv2.Sobel(ref, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE))
        if shape[0] % 2 == 0:
            assert np.array_equal(O.prim_pyrdown(im), cv2.pyrDown(im))
        g1 = np.ascontiguousarray(im[..., 1])
        assert np.array_equal(O.prim_median5(g1), cv2.medianBlur(g1, 5))
    # phase over the whole Sobel range, non-optimised build == the SSE2 (non-FMA) formula
    x = rng.integers(-1020, 1021, 20000).astype(np.float32)
    y = rng.integers(-1020, 1021, 20000).astype(np.float32)
    cv2.setUseOptimized(False)
    ref = cv2.phase(x, y, angleInDegrees=True).ravel()
    cv2.setUseOptimized(True)
    assert np.array_equal(O.prim_phase_deg(x, y).view(np.uint32), ref.view(np.uint32))


def test_cg_quantize_structure():
    """hysteresisGradient: border ring is zero, outputs are one-hot, weak pixels are dropped."""
    rng = np.random.default_rng(3)
    views = common.rendered_views(1, 9, canvas=(96, 128))
    bgr = views[0][0]
    mag, q, ang = O.prim_cg_quantize(bgr, 10.0)
    assert q[0].max() == 0 and q[-1].max() == 0 and q[:, 0].max() == 0 and q[:, -1].max() == 0
    nz = q[q > 0]
    assert nz.size > 0 and np.all((nz & (nz - 1)) == 0)
    assert np.all(q[mag <= 100.0] == 0)
    assert ang.min() >= 0 and ang.max() <= 360.0
    flat = np.full((32, 32, 3), 128, np.uint8)
    _, qf, _ = O.prim_cg_quantize(flat, 10.0)
    assert qf.max() == 0
    del rng


def test_spread_definition():
    rng = np.random.default_rng(4)
    src = (rng.random((40, 48)) < 0.05).astype(np.uint8) * (1 << rng.integers(0, 8, (40, 48))).astype(np.uint8)
    for T in (1, 4, 5, 8):
        want = np.zeros_like(src)
        for r in range(T):
            for c in range(T):
                want[:40 - r, :48 - c] |= src[r:, c:]
        assert np.array_equal(O.prim_spread(src, T), want)
