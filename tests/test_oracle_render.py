"""Pins the template-generation oracle (oracle/render_oracle.cpp) and the product's host-side view logic against the
one training run the reference ships (tests/golden/renderer_params_boxnew.npz, made by tests/golden/make_renderer_golden.py
from config/data/boxNew_longDistance_linemod_xtion_renderer_params.yml + config/stl/boxNew.stl).  CPU only."""
import os

import numpy as np
import pytest

import common
from common import O, synth
from linemod_pose_estimation_b200 import LinemodError, Mesh, ViewSphere, training

with np.load(os.path.join(common.GOLDEN, "renderer_params_boxnew.npz")) as _z:
    G = {k: _z[k] for k in _z.files}   # materialised once: NpzFile re-reads the archive on every access
PARAMS = dict(zip([str(k) for k in G["param_names"]], G["param_values"]))


def _golden_sphere(kind):
    args = (int(PARAMS["renderer_n_points"]), int(PARAMS["renderer_angle_step"]), PARAMS["renderer_radius_min"],
            PARAMS["renderer_radius_max"], PARAMS["renderer_radius_step"])
    return O.view_sphere(*args) if kind == "oracle" else ViewSphere(*args)


def _golden_camera(mod):
    return mod.camera(int(PARAMS["renderer_width"]), int(PARAMS["renderer_height"]), PARAMS["renderer_focal_length_x"],
                      PARAMS["renderer_focal_length_y"], PARAMS["renderer_near"], PARAMS["renderer_far"])


_VIEW_CACHE = {}


def _oracle_views():
    """Every view of the golden run's iterator (oracle restatement), cached: (list of view tuples, golden template -> view index)."""
    if "v" not in _VIEW_CACHE:
        vs = _golden_sphere("oracle")
        views = O.view_list(vs)
        _VIEW_CACHE["v"] = (views, _golden_view_indices(views))
    return _VIEW_CACHE["v"]


def _golden_view_indices(vs_views):
    """Index of every recorded template in the iterator's enumeration, matched on T (= -camera position) and R."""
    allT = np.array([v[0] for v in vs_views])
    order = np.lexsort((allT[:, 2], allT[:, 1], allT[:, 0]))
    key = allT[order]
    idx = []
    pose = {}
    for k in range(len(G["T"])):
        want = -G["T"][k]
        lo = np.searchsorted(key[:, 0], want[0] - 1e-12)
        cand = []
        while lo < len(key) and key[lo, 0] <= want[0] + 1e-12:
            if np.abs(key[lo] - want).max() < 1e-12:
                cand.append(int(order[lo]))
            lo += 1
        assert cand, "template %d: no view with this camera position" % k
        for c in cand:
            if c not in pose:
                pose[c] = training.view_pose(vs_views[c][0], vs_views[c][1])[0]
        errs = [np.abs(pose[c] - G["R"][k]).max() for c in cand]
        j = int(np.argmin(errs))
        assert errs[j] < 1e-12, "template %d: orientation differs by %g" % (k, errs[j])
        idx.append(cand[j])
    return np.array(idx)


def check_rect_against_golden(k, rect, H, stats):
    """Silhouette box (x, y, w, h; top-left origin) of golden template k vs the recorded Rect: ORK grows its box by one pixel
    per side and keeps GL's bottom-left origin.  Tolerance (OpenGL's rasteriser is not this one): every component within
    one pixel; exactness is counted in `stats` and asserted over the whole set by the caller."""
    x, y, w, h = rect
    mine = np.array([x - 1, H - (y + h) - 1, w + 2, h + 2])
    d = mine - G["rect"][k]
    assert np.abs(d).max() <= 1, (k, rect, G["rect"][k])
    stats["n"] = stats.get("n", 0) + 1
    stats["exact"] = stats.get("exact", 0) + int(not d.any())
    stats["x_exact"] = stats.get("x_exact", 0) + int(d[0] == 0 and d[2] == 0)


def check_rect_stats(stats):
    # measured over all 2 652 views: x / width exact in 99.8 %, all four components exact in 87 % (the rest: one extra
    # silhouette row at the image-bottom edge)
    assert stats["x_exact"] >= 0.99 * stats["n"] and stats["exact"] >= 0.85 * stats["n"], stats


def test_view_iterator_reproduces_the_shipped_training_run():
    """All 2 652 recorded (R, T) are views of RendererIterator(150 points, angle step 10, radius 0.5..1.0 step 0.1), in
    iteration order: T_file = -camera position, R_file = [left; -up; -view direction] = this library's camera rotation."""
    views, idx = _oracle_views()
    n = len(views)
    assert n == 150 * 17 * 6   # the shipped run holds six radii: the sweep tolerates 1.0000001 > radius_max
    assert np.all(np.diff(idx) > 0), "recorded templates are not in iteration order"
    radii = sorted({views[i][2] for i in idx})
    assert np.allclose(radii, sorted(set(G["ori_dist"])), rtol=0, atol=0)
    assert sorted({views[i][4] for i in idx}) == list(range(-80, 81, 10))
    # the product's host-side iterator is the same function (no GPU needed for it)
    pv = _golden_sphere("product")
    assert len(pv) == n
    for i in list(idx[::53]) + [0, n - 1]:
        a, b = pv.view(int(i)), views[int(i)]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]
    with pytest.raises(LinemodError):
        pv.view(n)


def test_oracle_rasteriser_geometry_matches_recorded_rectangles():
    """Silhouette bounding boxes of boxNew.stl vs the recorded `Rect`s (renderer.cpp flips the images but not the
    rectangle), centre depth vs D = D_obj - depth(centre) (renderer.cpp:271) within 2 mm (two u16 roundings)."""
    views, idx = _oracle_views()
    cam = _golden_camera(O)
    H = cam.height
    stats = {}
    for k in range(0, len(idx), 7):   # 379 views spread over every radius / angle (the GPU test covers all 2 652)
        T, up = views[idx[k]][:2]
        bgr, depth, mask, (x, y, w, h) = O.render(G["triangles"], cam, T, up)
        check_rect_against_golden(k, (x, y, w, h), H, stats)
        centre_m = depth[H // 2, cam.width // 2] / 1000.0
        assert abs((G["ori_dist"][k] - centre_m) - G["D"][k]) <= 2.001e-3, (k, centre_m, G["D"][k])
        assert mask.max() == 255 and np.array_equal(mask > 0, depth > 0)
        ys, xs = np.nonzero(mask)
        assert (xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1) == (x, y, w, h)
    check_rect_stats(stats)


def test_rasteriser_properties():
    cam = O.camera(320, 240, 300.0, 300.0)
    tri = synth.box_mesh(0.1, 0.1, 0.1)
    vs = O.view_sphere(20, 40, 0.4, 0.6, 0.2)
    T, up = O.view_params(vs, 2)[:2]   # point 0, radius 0.4, angle 0
    bgr, depth, mask, rect = O.render(tri, cam, T, up)
    # triangle order does not matter (ties go to the lower index only for equal depth: same surface, same pixels)
    perm = np.random.default_rng(3).permutation(len(tri))
    b2, d2, m2, r2 = O.render(tri[perm], cam, T, up)
    assert np.array_equal(depth, d2) and np.array_equal(mask, m2) and rect == r2
    # nearest surface: the cube's centre pixel is at radius - half diagonal .. radius - half edge
    r = np.linalg.norm(T)
    c = depth[120, 160] / 1000.0
    assert r - 0.0867 <= c <= r - 0.0499
    # camera behind / object out of view -> empty render
    far = O.render(tri + 100.0, cam, T, up)
    assert far[3] == (0, 0, 0, 0) and far[2].max() == 0
    # scaling the radius scales the silhouette
    T2, up2 = O.view_params(vs, 7)[:2]   # next radius, same point and angle
    assert np.linalg.norm(T2) > np.linalg.norm(T)
    small = O.render(tri, cam, T2, up2)[3]
    assert small[2] < rect[2] and small[3] < rect[3]


def test_stl_loader_binary_and_ascii(tmp_path):
    tri = synth.gear_mesh()
    for binary in (True, False):
        p = tmp_path / ("g_%d.stl" % binary)
        synth.write_stl(p, tri, binary=binary)
        m = Mesh.load_stl(p)
        assert len(m) == len(tri)
        assert np.array_equal(m.triangles, tri) if binary else np.allclose(m.triangles, tri, rtol=1e-7, atol=0)
    bad = tmp_path / "bad.stl"
    bad.write_text("solid x\nendsolid x\n")
    with pytest.raises(LinemodError):
        Mesh.load_stl(bad)
    with pytest.raises(LinemodError):
        Mesh.load_stl(tmp_path / "missing.stl")
    assert np.array_equal(Mesh(tri).triangles, tri)


def test_depth_diff_restatement():
    """rgbdDetector::depth_diff (rgbdDetector.cpp:236-283) against a direct numpy transcription of its OpenCV calls."""
    rng = np.random.default_rng(11)
    scene = rng.integers(0, 1200, (60, 80)).astype(np.uint16)
    scene[rng.random(scene.shape) < 0.2] = 0
    templ = rng.integers(300, 1300, (50, 70)).astype(np.uint16)
    tmask = (rng.random(templ.shape) < 0.6).astype(np.uint8) * 255
    x, y, tx, ty, w, h = 5, 7, 3, 2, 40, 30
    roi = scene[y:y + h, x:x + w]
    t = templ[ty:ty + h, tx:tx + w]
    m = tmask[ty:ty + h, tx:tx + w] & np.minimum(roi, 255).astype(np.uint8)       # convertTo(CV_8UC1) saturates
    sub = np.where(t > roi, t - roi, 0).astype(np.uint16).view(np.int16)          # u16 saturating subtract read as short
    want = np.abs(sub.astype(np.int64))[m > 0].sum() / ((m > 0).sum() * 1000.0)
    assert O.depth_diff(scene, templ, tmask, x, y, tx, ty, w, h) == want


def test_pose_table_writer_reproduces_the_shipped_file_byte_for_byte(tmp_path):
    """writeLinemodTemplateParams (renderer.cpp:72-123): the 2 652 recorded poses written by lm_write_renderer_params give
    the reference's own 2 MB renderer_params.yml back exactly (sha256 recorded by make_renderer_golden.py), the reader
    returns what was written, and OpenCV's FileStorage loads the file."""
    import hashlib

    from linemod_pose_estimation_b200._capi import LmRendererParams, POSE_DTYPE
    n = len(G["R"])
    poses = np.zeros(n, POSE_DTYPE)
    poses["R"], poses["T"], poses["D"], poses["ori_dist"] = G["R"], G["T"], G["D"], G["ori_dist"]
    K = np.array([[PARAMS["renderer_focal_length_x"], 0, PARAMS["renderer_width"] / 2.0],
                  [0, PARAMS["renderer_focal_length_y"], PARAMS["renderer_height"] / 2.0], [0, 0, 1]], np.float32)
    poses["K"] = K
    for i, name in enumerate(("x", "y", "width", "height")):
        poses["rect"][name] = G["rect"][:, i]
    params = LmRendererParams(int(PARAMS["renderer_n_points"]), int(PARAMS["renderer_angle_step"]), PARAMS["renderer_radius_min"],
                              PARAMS["renderer_radius_max"], PARAMS["renderer_radius_step"], int(PARAMS["renderer_width"]),
                              int(PARAMS["renderer_height"]), PARAMS["renderer_focal_length_x"],
                              PARAMS["renderer_focal_length_y"], PARAMS["renderer_near"], PARAMS["renderer_far"])
    p = tmp_path / "renderer_params.yml"
    training.write_renderer_params(p, poses, params)
    raw = open(p, "rb").read()
    assert len(raw) == int(G["yml_bytes"]) and hashlib.sha256(raw).hexdigest() == str(G["yml_sha256"])
    back, bp = training.read_renderer_params(p)
    assert back.tobytes() == poses.tobytes()
    assert (bp.n_points, bp.angle_step, bp.width, bp.height, bp.fx, bp.far_) == (150, 10, 640, 480, 535.566011, 1000.0)
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(str(p), cv2.FILE_STORAGE_READ)
    assert np.array_equal(fs.getNode("Template 2651").getNode("R").mat(), G["R"][2651])
    assert fs.getNode("Template 2652").empty() and fs.getNode("renderer_radius_step").real() == 0.1
    with pytest.raises(LinemodError):
        training.read_renderer_params(tmp_path / "missing.yml")


def test_training_fixture_is_reproduced_by_the_oracle():
    """tests/golden/train_gear.npz (made by tests/golden/make_train_golden.py): oracle render + addTemplate of 12 views of a
    gear mesh, frozen.  Guards the oracle against drift; tests/test_gpu_train.py holds lm_train_views to the same file."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_train_golden", os.path.join(common.GOLDEN, "make_train_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    with np.load(os.path.join(common.GOLDEN, "train_gear.npz")) as z:
        assert sorted(z.files) == sorted(now)
        for k in z.files:
            assert np.array_equal(z[k], now[k]), k
        assert (z["tid"] >= 0).sum() >= 10 and len(z["hdr"]) == 4 * (z["tid"] >= 0).sum()
