"""GPU parity of the template-generation path through the C ABI: CUDA rasteriser vs the scalar oracle
(bit-exact), batched addTemplate vs the oracle's addTemplate (identical templates), the trainer loop, depth_diff.
Needs a B200; run with `pytest -m gpu`."""
import numpy as np
import pytest

import common
from common import O, synth
from linemod_pose_estimation_b200 import Detector, LinemodError, Mesh, ViewSphere, camera, training
import test_oracle_render as golden

pytestmark = pytest.mark.gpu

MESHES = {"box": synth.box_mesh(), "bracket": synth.bracket_mesh(), "gear": synth.gear_mesh()}


def _views(n_points=12, angle_step=80, radius=(0.35, 0.55, 0.2)):
    vs = ViewSphere(n_points, angle_step, radius[0], radius[1], radius[2])
    return vs.views()


def _same_templates(det, orc, cid, what):
    assert det.numTemplates(cid) == orc.num_templates(cid), what
    for tid in range(orc.num_templates(cid)):
        a, b = det.getTemplates(cid, tid), orc.get_template(cid, tid)
        for (w0, h0, l0, f0), (w1, h1, l1, f1) in zip(a, b):
            assert (w0, h0, l0) == (w1, h1, l1), (what, tid)
            assert np.array_equal(f0, f1), (what, tid, l0)


# ---------------------------------------------------------------------------------------------- rasteriser
@pytest.mark.parametrize("name", sorted(MESHES))
def test_rasteriser_bit_exact_vs_oracle(name):
    tri = MESHES[name]
    det, mesh = Detector(), Mesh(tri)
    cam, ocam = camera(320, 240, 420.0, 415.0), O.camera(320, 240, 420.0, 415.0)
    T, up = _views()
    got = training.render_views(det, mesh, cam, T, up)
    assert len(T) > 32   # more than one internal batch
    for v in range(len(T)):
        bgr, depth, mask, rect = O.render(tri, ocam, T[v], up[v])
        assert np.array_equal(got["depth"][v], depth), (name, v)
        assert np.array_equal(got["mask"][v], mask), (name, v)
        assert np.array_equal(got["bgr"][v], bgr), (name, v)
        assert tuple(got["rects"][v]) == rect, (name, v)
    assert got["mask"].max() == 255


def test_rasteriser_geometry_matches_the_reference_run():
    """All 2 652 views of the reference's shipped training run (boxNew.stl), rendered by the CUDA rasteriser: silhouette
    boxes vs the recorded Rects and centre depth vs D (tolerances stated in tests/test_oracle_render.py)."""
    G = golden.G
    views, idx = golden._oracle_views()
    cam = golden._golden_camera(training)
    det, mesh = Detector(), Mesh(G["triangles"])
    T = np.array([views[i][0] for i in idx])
    up = np.array([views[i][1] for i in idx])
    stats = {}
    for k0 in range(0, len(idx), 256):
        got = training.render_views(det, mesh, cam, T[k0:k0 + 256], up[k0:k0 + 256], want=("depth",))
        for j in range(len(got["rects"])):
            k = k0 + j
            golden.check_rect_against_golden(k, tuple(got["rects"][j]), cam.height, stats)
            centre_m = got["depth"][j][cam.height // 2, cam.width // 2] / 1000.0
            assert abs((G["ori_dist"][k] - centre_m) - G["D"][k]) <= 2.001e-3, k
    assert stats["n"] == 2652
    golden.check_rect_stats(stats)


def test_render_edge_cases():
    det, mesh = Detector(), Mesh(MESHES["box"])
    cam = camera(160, 120, 200.0, 200.0)
    T, up = _views(4, 160)
    # object behind the far plane / off screen -> empty masks and zero rects
    off = training.render_views(det, Mesh(MESHES["box"] + 50.0), cam, T, up)
    assert off["mask"].max() == 0 and not off["rects"]["width"].any()
    # empty mesh, zero views
    assert training.render_views(det, Mesh(np.zeros((0, 3, 3), np.float32)), cam, T, up)["mask"].max() == 0
    assert len(training.render_views(det, mesh, cam, np.zeros((0, 3)), np.zeros((0, 3)))["rects"]) == 0
    # camera inside the object: every triangle crosses the near plane or faces away -> still well defined, equals the oracle
    Tn = np.array([[0.0, 0.0, 0.02]])
    upn = np.array([[0.0, 1.0, 0.0]])
    a = training.render_views(det, mesh, cam, Tn, upn)
    b = O.render(MESHES["box"], O.camera(160, 120, 200.0, 200.0), Tn[0], upn[0])
    assert np.array_equal(a["depth"][0], b[1])
    with pytest.raises(LinemodError):   # up parallel to the viewing direction
        training.render_views(det, mesh, cam, np.array([[0.0, 0.0, 0.5]]), np.array([[0.0, 0.0, 1.0]]))


# ---------------------------------------------------------------------------------------------- batched addTemplate
@pytest.mark.parametrize("kinds,T", [(("cg", "dn"), (5, 8)), (("dn", "cg"), (4, 8)), (("cg",), (5,)), (("cg", "dn"), (2, 4, 8))])
def test_batch_add_template_equals_oracle(kinds, T):
    """lm_add_templates_batch on synthetic object views == the oracle's addTemplate view by view (features, sizes, boxes,
    failures), and == the product's own sequential lm_add_template."""
    canvas = (240, 240) if len(T) < 3 else (256, 256)
    views = common.rendered_views(40, 171, canvas)
    # a view too small to yield 63 features and an all-zero mask: both must fail with -1, not disturb the others
    tiny = synth.render_view(3, 0.12, 10.0, canvas=canvas)
    views.insert(5, tiny)
    views.insert(17, (views[0][0], views[0][1], np.zeros_like(views[0][2])))
    orc = O.OracleDetector(common.oracle_modalities(kinds), T)
    det = Detector(common.product_modalities(kinds), T)
    seq = Detector(common.product_modalities(kinds), T)
    want, want_bb = [], []
    for bgr, depth, mask in views:
        tid, bb = orc.add_template(common.sources_for(kinds, bgr, depth), "obj", mask)
        want.append(tid)
        want_bb.append(bb if tid >= 0 else (0, 0, 0, 0))
    tids, bbs = det.addTemplates([(common.sources_for(kinds, b, d), m) for b, d, m in views], "obj")
    assert list(tids) == want
    assert -1 in want and max(want) > 20
    assert [tuple(b) for b in bbs] == [tuple(b) for b in want_bb]
    _same_templates(det, orc, "obj", "batch vs oracle")
    for bgr, depth, mask in views[:12]:
        seq.addTemplate(common.sources_for(kinds, bgr, depth), "obj", mask)
    for tid in range(seq.numTemplates("obj")):
        for a, b in zip(seq.getTemplates("obj", tid), det.getTemplates("obj", tid)):
            assert a[:3] == b[:3] and np.array_equal(a[3], b[3])


def test_batch_add_template_large_object_and_full_mask():
    """A 640x480 view whose object fills most of the frame (tens of thousands of DepthNormal candidates: the global-memory
    sort path, long distance searches) and, for a DepthNormal-only detector, a full-frame mask over a flat depth image:
    one normal bin everywhere, no zero pixel in the bin's plane (the distance transform's "far border" values), 307 200
    candidates."""
    big = synth.render_view(7, 3.2, 25.0, canvas=(480, 640), tilt=0.3)
    flat_depth = np.full((480, 640), 700, np.uint16)
    flat_bgr = np.clip(synth.make_background(5)[0], 0, 255).astype(np.uint8)
    full = np.full((480, 640), 255, np.uint8)
    views = [big, (flat_bgr, flat_depth, full)]
    orc, det = O.OracleDetector(), Detector()
    want = [orc.add_template([b, d], "obj", m)[0] for b, d, m in views]
    tids, _ = det.addTemplates([([b, d], m) for b, d, m in views], "obj")
    assert list(tids) == want and want == [0, -1]   # a full mask has no silhouette ring: ColorGradient finds no candidates
    _same_templates(det, orc, "obj", "large view")
    orc, det = O.OracleDetector(common.oracle_modalities(("dn",)), (5, 8)), Detector(common.product_modalities(("dn",)), (5, 8))
    want = [orc.add_template([d], "obj", m)[0] for b, d, m in views]
    tids, _ = det.addTemplates([([d], m) for b, d, m in views], "obj")
    assert list(tids) == want and want == [0, 1]
    _same_templates(det, orc, "obj", "DepthNormal only, full mask")


def test_batch_add_template_with_multi_bit_normal_lut():
    """An injected NORMAL_LUT whose entries are not one-hot (the stock table is): the DepthNormal distance falls back from
    the run-table separation to the ring search; templates still equal the oracle's."""
    views = common.rendered_views(16, 181, (240, 240))
    orc, det = O.OracleDetector(), Detector()
    lut = orc.normal_lut().copy()
    lut[lut == 4] = 12     # bins 2 and 3 merged into a two-bit value: such pixels are in two planes and never candidates
    orc.set_normal_lut(lut)
    det.set_normal_lut(lut)
    want = [orc.add_template([b, d], "obj", m)[0] for b, d, m in views]
    tids, _ = det.addTemplates([([b, d], m) for b, d, m in views], "obj")
    assert list(tids) == want and max(want) >= 3 and -1 in want
    _same_templates(det, orc, "obj", "multi-bit LUT")


def test_add_templates_batch_argument_errors():
    det = Detector()
    bgr, depth, mask = common.rendered_views(1, 5)[0]
    with pytest.raises(LinemodError):   # mask of another size
        training.add_templates_batch(Detector(common.product_modalities(("cg",)), (5, 8)), [([bgr], mask[:100])], "obj")
    with pytest.raises(LinemodError):   # sources of different sizes
        training.add_templates_batch(det, [([bgr, depth], mask), ([bgr[:200], depth[:200]], mask[:200])], "obj")
    tids, bbs = training.add_templates_batch(det, [], "obj")
    assert len(tids) == 0


# ---------------------------------------------------------------------------------------------- trainer loop
def test_train_views_equals_oracle_pipeline_and_matches():
    """lm_train_views (render + addTemplate on the device) == oracle render + oracle addTemplate per view; the trained
    detector then finds a planted view at the planted position."""
    tri = MESHES["bracket"]
    cam, ocam = camera(), O.camera()
    vs = ViewSphere(24, 40, 0.45, 0.65, 0.1)
    T, up = vs.views()
    assert len(T) == 24 * 5 * 3
    det, orc, mesh = Detector(), O.OracleDetector(), Mesh(tri)
    tids, bbs, rects = det.trainViews(mesh, cam, T, up, "bracket")
    want = []
    for v in range(len(T)):
        bgr, depth, mask, rect = O.render(tri, ocam, T[v], up[v])
        tid, bb = orc.add_template([bgr, depth], "bracket", mask)
        want.append(tid)
        assert tuple(rects[v]) == rect
        if tid >= 0:
            assert tuple(bbs[v]) == bb, v
    assert list(tids) == want and max(want) > 200
    _same_templates(det, orc, "bracket", "trainer")
    # plant one training view in clutter and find it
    v = int(np.flatnonzero(tids >= 0)[57])
    bgr, depth, mask, rect = O.render(tri, ocam, T[v], up[v])
    scene_bgr = np.clip(synth.make_background(9)[0], 0, 255).astype(np.uint8)
    scene_depth = np.full((480, 640), 1500, np.uint16)
    dx, dy = 40, -30
    ys, xs = np.nonzero(mask)
    scene_bgr[ys + dy, xs + dx] = bgr[ys, xs]
    scene_depth[ys + dy, xs + dx] = depth[ys, xs]
    got = det.match([scene_bgr, scene_depth], 85.0)
    ref = orc.match([scene_bgr, scene_depth], 85.0)
    common.assert_matches_equal(got, ref)
    assert len(got) > 0
    best = got[0]
    assert best["similarity"] > 90 and abs(best["x"] - (bbs[v]["x"] + dx)) <= 8 and abs(best["y"] - (bbs[v]["y"] + dy)) <= 8


def test_depth_diff_equals_reference_restatement():
    tri = MESHES["gear"]
    cam, ocam = camera(), O.camera()
    T, up = ViewSphere(10, 80, 0.5, 0.5, 0.1).views()
    det, mesh = Detector(), Mesh(tri)
    rng = np.random.default_rng(5)
    scene = rng.integers(400, 900, (480, 640)).astype(np.uint16)
    scene[rng.random(scene.shape) < 0.1] = 0
    scene[100:140, 200:260] = 80   # depths below 256 mm exercise the byte saturation of the validity mask
    xs = rng.integers(0, 300, len(T)).astype(np.int32)
    ys = rng.integers(0, 200, len(T)).astype(np.int32)
    got = training.depth_diff(det, scene, mesh, cam, T, up, xs, ys)
    for v in range(len(T)):
        _, depth, mask, (rx, ry, rw, rh) = O.render(tri, ocam, T[v], up[v])
        want = O.depth_diff(scene, depth, mask, int(xs[v]), int(ys[v]), rx, ry, rw, rh)
        assert got[v] == want, (v, got[v], want)
    with pytest.raises(LinemodError):   # crop leaves the scene
        training.depth_diff(det, scene, mesh, cam, T[:1], up[:1], [630], [470])


def test_trainer_tool_writes_the_reference_file_pair(tmp_path):
    """tools/train.py == renderer_node (renderer.cpp:170-354): templates.yml + renderer_params.yml for an STL file; and on
    the reference's own run (boxNew.stl, the 2 652 recorded views) the pose table it derives agrees with the recorded one:
    R, T to 1e-15, D within 2 mm, Rect within the rasteriser tolerance."""
    import json
    import os
    import subprocess
    import sys
    stl = tmp_path / "gear.stl"
    synth.write_stl(stl, MESHES["gear"] * 2.0, binary=True)
    out_t, out_p = tmp_path / "gear_templates.yml", tmp_path / "gear_renderer_params.yml"
    r = subprocess.run([sys.executable, os.path.join(common.ROOT, "tools", "train.py"), "--stl", str(stl), "--templates", str(out_t),
                        "--params", str(out_p), "--n-points", "16", "--angle-step", "40", "--radius-min", "0.4", "--radius-max", "0.6",
                        "--radius-step", "0.1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["views"] == 16 * 5 * 3 and info["templates"] > 100
    det = Detector.read(out_t)
    poses, params = training.read_renderer_params(out_p)
    assert det.numTemplates("obj") == len(poses) == info["templates"] and params.n_points == 16
    assert np.allclose(np.linalg.norm(poses["T"], axis=1), poses["ori_dist"], atol=1e-6)
    # the reference's run
    G = golden.G
    views, idx = golden._oracle_views()
    cam = golden._golden_camera(training)
    T = np.array([views[i][0] for i in idx])
    up = np.array([views[i][1] for i in idx])
    d2 = Detector()
    tids, bbs, rects, centre = d2.trainViews(Mesh(G["triangles"]), cam, T, up, "obj", centre_depth=True)
    assert (tids >= 0).all() and len(tids) == 2652   # the reference kept exactly these 2 652 views
    rr = np.array([(q["x"] - 1, cam.height - (q["y"] + q["height"]) - 1, q["width"] + 2, q["height"] + 2) for q in rects])
    mine = training.poses_for_views(T, up, cam, [views[i][2] for i in idx], rr, centre)
    assert np.abs(mine["R"] - G["R"]).max() < 1e-12 and np.abs(mine["T"] - G["T"]).max() < 1e-12
    assert np.array_equal(mine["ori_dist"], G["ori_dist"]) and np.abs(mine["D"] - G["D"]).max() <= 2.001e-3
    d = np.abs(rr - G["rect"])
    assert d.max() <= 1 and (d.max(axis=1) == 0).mean() >= 0.85


def test_train_views_reproduces_the_frozen_fixture():
    """lm_train_views against tests/golden/train_gear.npz (oracle render + addTemplate, frozen): rectangles, centre depths,
    template ids and every feature."""
    import os
    with np.load(os.path.join(common.GOLDEN, "train_gear.npz")) as z:
        F = {k: z[k] for k in z.files}
    det, mesh = Detector(), Mesh(F["triangles"])
    cam = camera(320, 240, 420.0, 415.0)
    tids, bbs, rects, centre = det.trainViews(mesh, cam, F["T"], F["up"], "gear", centre_depth=True)
    assert np.array_equal(tids, F["tid"])
    assert np.array_equal(np.array([tuple(r) for r in rects], np.int32), F["rect"])
    assert np.array_equal(centre.astype(np.int32), F["centre_mm"])
    k = f0 = 0
    for tid in range(int(F["tid"].max()) + 1):
        for (w, h, lvl, feats) in det.getTemplates("gear", tid):
            t, fw, fh, fl, n = F["hdr"][k]
            assert (t, fw, fh, fl, n) == (tid, w, h, lvl, len(feats)), (tid, k)
            assert np.array_equal(feats, F["feats"][f0:f0 + n]), (tid, k)
            k += 1
            f0 += n
    assert k == len(F["hdr"]) and f0 == len(F["feats"])
