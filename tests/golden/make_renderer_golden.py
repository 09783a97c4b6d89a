#!/usr/bin/env python
"""Builds tests/golden/renderer_params_boxnew.npz from the one training run the reference ships:

    /root/reference/config/data/boxNew_longDistance_linemod_xtion_renderer_params.yml   (written by writeLinemodTemplateParams,
                                                                                         src/renderer.cpp:332-349)
    /root/reference/config/stl/boxNew.stl                                               (the mesh that run rendered)

Per template the file records R (Rs_), T (Ts_), D (distance), Ori_dist (D_obj) and Rect (the render rectangle), plus the
renderer parameters.  Stored here verbatim (f64 / i32) together with the mesh triangles, so the tests can pin the view
iterator and the rasteriser's geometry on the GPU box, where /root/reference does not exist.  Run in the build container:

    python tests/golden/make_renderer_golden.py
"""
import hashlib
import os
import re
import struct

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "renderer_params_boxnew.npz")


def main():
    yml = os.path.join(REF, "config/data/boxNew_longDistance_linemod_xtion_renderer_params.yml")
    s = open(yml).read()
    sha = hashlib.sha256(open(yml, "rb").read()).hexdigest()

    def mats(key):
        out = re.findall(key + r": !!opencv-matrix\s+rows: \d\s+cols: \d\s+dt: \w\s+data: \[([^\]]+)\]", s)
        return np.array([[float(x) for x in t.replace("\n", " ").split(",")] for t in out])

    R = mats("R").reshape(-1, 3, 3)
    T = mats("T")
    D = np.array([float(x) for x in re.findall(r"\n   D: ([0-9.e+-]+)", s)])
    ori = np.array([float(x) for x in re.findall(r"Ori_dist: ([0-9.e+-]+)", s)])
    rect = np.array([[int(v) for v in r.split(",")] for r in re.findall(r"Rect: \[([^\]]+)\]", s)], np.int32)
    params = {k: float(v) for k, v in re.findall(r"\n(renderer_\w+): ([0-9.e+-]+)", s)}
    assert len(R) == len(T) == len(D) == len(ori) == len(rect) == 2652
    raw = open(os.path.join(REF, "config/stl/boxNew.stl"), "rb").read()
    n = struct.unpack("<I", raw[80:84])[0]
    assert 84 + 50 * n == len(raw)
    tri = np.zeros((n, 9), np.float32)
    for i in range(n):
        tri[i] = np.frombuffer(raw[84 + 50 * i + 12:84 + 50 * i + 48], np.float32)
    np.savez_compressed(OUT, R=R, T=T, D=D, ori_dist=ori, rect=rect, triangles=tri.reshape(n, 3, 3),
                        yml_sha256=np.array(sha), yml_bytes=np.array(os.path.getsize(yml)),
                        param_names=np.array(sorted(params)), param_values=np.array([params[k] for k in sorted(params)]))
    print(OUT, os.path.getsize(OUT), "bytes;", len(R), "templates,", n, "triangles;", params)


if __name__ == "__main__":
    main()
