"""Seeded synthetic RGB-D data for the LINEMOD path (numpy only).

The reference ships no images, no trained templates.yml and no recorded bags (SURVEY.md section 0.2 / 4;
/root/reference/.MISSING_LARGE_BLOBS), so tests and bench.py drive the matcher with procedurally rendered
"object views" (what renderer_node feeds to Detector::addTemplate, /root/reference/src/renderer.cpp:262-329)
and with scenes that have those views planted in clutter (what detect_cb feeds to Detector::match,
/root/reference/src/linemod_carmine_detect.cpp:329-348).  Everything is a pure function of its seed.
"""
import numpy as np


def _smooth_noise(rng, rows, cols, sigma, lo, hi):
    """Low-frequency field in [lo, hi]: box-filtered white noise (separable cumulative sums, no scipy needed)."""
    k = max(1, int(sigma))
    a = rng.random((rows + 4 * k, cols + 4 * k))
    for _ in range(2):
        c = np.cumsum(a, axis=0)
        a = c[2 * k:] - c[:-2 * k]
        c = np.cumsum(a, axis=1)
        a = c[:, 2 * k:] - c[:, :-2 * k]
    a = a[:rows, :cols]
    a = (a - a.min()) / max(1e-9, a.max() - a.min())
    return lo + a * (hi - lo)


def _point_in_polygon(px, py, vx, vy):
    inside = np.zeros(px.shape, bool)
    n = len(vx)
    j = n - 1
    for i in range(n):
        xi, yi, xj, yj = vx[i], vy[i], vx[j], vy[j]
        cond = ((yi > py) != (yj > py)) & (px < (xj - xi) * (py - yi) / (yj - yi + 1e-12) + xi)
        inside ^= cond
        j = i
    return inside


def render_view(shape_seed, scale=1.0, rot_deg=0.0, canvas=(256, 256), base_depth=600, tilt=0.0):
    """One rendered object view: (bgr u8 [H,W,3], depth u16 [H,W] in mm, mask u8 [H,W] 0/255).

    A faceted polygonal "part": 5-8 silhouette vertices, 3 planar faces meeting at the centroid, every face with its
    own colour, stripe texture and depth gradient (so both ColorGradient and DepthNormal find features).  Background
    is black / depth 0, like the off-screen GL renders the reference trains from.
    """
    rng = np.random.default_rng(shape_seed)
    rows, cols = canvas
    cy, cx = rows / 2.0, cols / 2.0
    n = int(rng.integers(5, 9))
    ang = np.sort(rng.uniform(0, 2 * np.pi, n) * 0.35 + np.arange(n) * 2 * np.pi / n * 0.65 + rng.uniform(0, 2 * np.pi))
    rad = rng.uniform(0.62, 1.0, n) * 60.0 * scale
    rot = np.deg2rad(rot_deg)
    vx = cx + rad * np.cos(ang + rot)
    vy = cy + rad * np.sin(ang + rot) * (1.0 - 0.35 * tilt)
    yy, xx = np.mgrid[0:rows, 0:cols]
    px, py = xx + 0.5, yy + 0.5
    mask = _point_in_polygon(px, py, vx, vy)

    # three faces = angular sectors around the centroid
    cuts = np.sort(rng.uniform(0, 2 * np.pi, 3))
    theta = (np.arctan2(py - cy, px - cx) - rot) % (2 * np.pi)
    face = np.searchsorted(cuts, theta) % 3

    base = rng.permutation(np.array([[40, 90, 230], [220, 200, 60], [70, 220, 110], [200, 70, 200], [240, 240, 240]]))[:3]
    bgr = np.zeros((rows, cols, 3), np.float64)
    depth = np.zeros((rows, cols), np.float64)
    face_dir = rng.uniform(0, 2 * np.pi) + np.arange(3) * (2 * np.pi / 3) + rng.uniform(-0.3, 0.3, 3)
    face_slope = rng.uniform(1.6, 3.2, 3) / max(scale, 0.3)
    for f in range(3):
        sel = face == f
        sdir = rng.uniform(0, np.pi) + rot
        period = rng.uniform(9, 16) * scale
        stripes = 0.5 + 0.5 * np.sign(np.sin(((px - cx) * np.cos(sdir) + (py - cy) * np.sin(sdir)) * 2 * np.pi / period))
        col = base[f][None, None, :] * (0.55 + 0.45 * stripes[..., None])
        bgr[sel] = col[sel]
        gx, gy = face_slope[f] * np.cos(face_dir[f] + rot), face_slope[f] * np.sin(face_dir[f] + rot)
        d = base_depth + gx * (px - cx) + gy * (py - cy)
        depth[sel] = d[sel]
    bgr[~mask] = 0
    depth[~mask] = 0
    return (np.clip(bgr, 0, 255).astype(np.uint8), np.clip(depth, 0, 65535).astype(np.uint16),
            (mask.astype(np.uint8) * 255))


def make_background(seed, rows=480, cols=640):
    """Cluttered background: low-frequency colour field + a few textured quads, tilted ground-plane depth."""
    rng = np.random.default_rng(seed)
    bgr = np.stack([_smooth_noise(rng, rows, cols, 12, 40, 200) for _ in range(3)], axis=-1)
    yy, xx = np.mgrid[0:rows, 0:cols]
    depth = 800.0 + rng.uniform(-0.25, 0.25) * (xx - cols / 2) + rng.uniform(-0.25, 0.25) * (yy - rows / 2)
    for _ in range(int(rng.integers(3, 9))):
        w, h = rng.integers(30, 140, 2)
        w, h = int(min(w, cols // 2)), int(min(h, rows // 2))
        x0, y0 = int(rng.integers(0, cols - w)), int(rng.integers(0, rows - h))
        colour = rng.uniform(20, 235, 3)
        period = rng.uniform(6, 20)
        ph = rng.uniform(0, np.pi)
        sub = (np.sin((xx[y0:y0 + h, x0:x0 + w] * np.cos(ph) + yy[y0:y0 + h, x0:x0 + w] * np.sin(ph)) * 2 * np.pi / period) > 0)
        bgr[y0:y0 + h, x0:x0 + w] = colour[None, None, :] * (0.6 + 0.4 * sub[..., None])
        depth[y0:y0 + h, x0:x0 + w] = rng.uniform(400, 1000) + rng.uniform(-1.5, 1.5) * (xx[y0:y0 + h, x0:x0 + w] - x0)
    return bgr, depth


def compose_scene(seed, views, rows=480, cols=640, n_instances=None, noise=True):
    """Scene with object views pasted in.  views: list of (bgr, depth, mask).  Returns (bgr u8, depth u16, placements)
    where placements = [(view_index, x_off, y_off)] gives the paste offset of each view canvas."""
    rng = np.random.default_rng(seed)
    bgr, depth = make_background(seed + 7919, rows, cols)
    placements = []
    k = len(views) if n_instances is None else n_instances
    for i in range(k):
        vi = i % len(views) if views else 0
        vb, vd, vm = views[vi]
        ys, xs = np.nonzero(vm)
        y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
        h, w = y1 - y0, x1 - x0
        if h + 4 >= rows or w + 4 >= cols:
            continue
        oy = int(rng.integers(2, rows - h - 2))
        ox = int(rng.integers(2, cols - w - 2))
        m = vm[y0:y1, x0:x1] > 0
        bgr[oy:oy + h, ox:ox + w][m] = vb[y0:y1, x0:x1][m]
        depth[oy:oy + h, ox:ox + w][m] = vd[y0:y1, x0:x1][m]
        placements.append((vi, ox - x0, oy - y0))
    if noise:
        bgr = bgr + rng.normal(0, 2.0, bgr.shape)
        depth = depth + rng.integers(-2, 3, depth.shape)
        holes = rng.random(depth.shape) < 0.02
        depth[holes] = 0
    return (np.clip(np.rint(bgr), 0, 255).astype(np.uint8), np.clip(np.rint(depth), 0, 65535).astype(np.uint16),
            placements)


def random_pyramid(rng, T=(5, 8), M=2, nf0=63, wh_range=(55, 194)):
    """A random (stress-set) template pyramid: list of L*M tuples (width, height, level, features[n,3]).
    Feature coordinates obey the cropTemplates invariant 0 <= x <= width, 0 <= y <= height with the extreme
    corners present at level 0."""
    L = len(T)
    w0 = int(rng.integers(wh_range[0], wh_range[1] + 1)) & ~1
    h0 = int(rng.integers(wh_range[0], wh_range[1] + 1)) & ~1
    out = []
    for l in range(L):
        w, h = w0 >> l, h0 >> l
        nf = max(1, nf0 >> l)
        for m in range(M):
            f = np.stack([rng.integers(0, w + 1, nf), rng.integers(0, h + 1, nf), rng.integers(0, 8, nf)], axis=1)
            if l == 0 and m == 0:
                f[0, :2] = (0, 0)
                f[1, :2] = (w, h)
            out.append((w, h, l, f.astype(np.int32)))
    return out


def view_params(n_views, seed=0, scales=(0.6, 0.7, 0.8, 0.9, 1.0, 1.1)):
    """(shape_seed, scale, rot_deg, tilt) tuples for a renderer-style sweep: views x in-plane rotations x scales,
    ordered like the reference trainer (viewpoint-major, /root/reference/src/renderer.cpp:262)."""
    rng = np.random.default_rng(seed)
    out = []
    v = 0
    while len(out) < n_views:
        rot = float(rng.uniform(0, 360))
        tilt = float(rng.uniform(0, 1))
        for s in scales:
            if len(out) < n_views:
                out.append((seed * 100003 + v, float(s), rot, tilt))
        v += 1
    return out


# ------------------------------------------------------------------------------------------------ meshes
def _quad(a, b, c, d):
    return [[a, b, c], [a, c, d]]


def box_mesh(sx=0.093, sy=0.133, sz=0.08, centre=(0.0, 0.0, 0.0)):
    """Axis-aligned box (12 triangles), metres; the default is the extent of the reference's config/stl/boxNew.stl."""
    hx, hy, hz = sx / 2.0, sy / 2.0, sz / 2.0
    cx, cy, cz = centre
    v = np.array([[cx + dx * hx, cy + dy * hy, cz + dz * hz] for dz in (-1, 1) for dy in (-1, 1) for dx in (-1, 1)], np.float32)
    faces = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
    tris = []
    for f in faces:
        tris += _quad(*[v[i] for i in f])
    return np.array(tris, np.float32)


def bracket_mesh(length=0.12, width=0.06, height=0.08, thickness=0.02):
    """L-shaped bracket: two boxes sharing an edge (24 triangles, interior faces included -- the z-buffer hides them)."""
    base = box_mesh(length, width, thickness, (0.0, 0.0, -height / 2.0 + thickness / 2.0))
    wall = box_mesh(thickness, width, height, (-length / 2.0 + thickness / 2.0, 0.0, 0.0))
    return np.concatenate([base, wall])


def gear_mesh(teeth=9, r_in=0.035, r_out=0.05, height=0.02):
    """Extruded star polygon: a non-convex silhouette with many short edges (4 * 2 * teeth triangles)."""
    n = 2 * teeth
    ang = np.arange(n) * (2.0 * np.pi / n)
    rad = np.where(np.arange(n) % 2 == 0, r_out, r_in)
    ring = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    top = np.concatenate([ring, np.full((n, 1), height / 2.0)], 1)
    bot = np.concatenate([ring, np.full((n, 1), -height / 2.0)], 1)
    ct, cb = np.array([0.0, 0.0, height / 2.0]), np.array([0.0, 0.0, -height / 2.0])
    tris = []
    for i in range(n):
        j = (i + 1) % n
        tris.append([ct, top[i], top[j]])
        tris.append([cb, bot[j], bot[i]])
        tris += _quad(bot[i], bot[j], top[j], top[i])
    return np.array(tris, np.float32)


def write_stl(path, triangles, binary=True, name="synth"):
    """Writes a mesh as binary or ASCII STL (facet normals recomputed, zero for degenerate facets)."""
    tri = np.asarray(triangles, np.float32).reshape(-1, 3, 3)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(ln > 0, nrm / np.maximum(ln, 1e-30), 0).astype(np.float32)
    if binary:
        with open(path, "wb") as f:
            f.write(name.encode().ljust(80, b" "))
            f.write(np.uint32(len(tri)).tobytes())
            for t, n in zip(tri, nrm):
                f.write(n.tobytes() + t.tobytes() + b"\0\0")
    else:
        with open(path, "w") as f:
            f.write("solid %s\n" % name)
            for t, n in zip(tri, nrm):
                f.write("   facet normal %.6e %.6e %.6e\n      outer loop\n" % tuple(n))
                for v in t:
                    f.write("         vertex %.9e %.9e %.9e\n" % tuple(v))
                f.write("      endloop\n   endfacet\n")
            f.write("endsolid %s\n" % name)
