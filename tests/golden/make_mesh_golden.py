#!/usr/bin/env python
"""Writes tests/golden/meshes_config2.npz: the triangle lists (float32 [n, 3, 3], metres, object frame) of the two objects
BASELINE.json configs[1] names -- /root/reference/config/stl/memoryChip2.stl (ASCII STL) and cpu_binary.stl (binary STL) --
so that bench.py and the tests can train the reference's own objects on the GPU box, where /root/reference does not exist.
Mesh data only (what `Renderer3d(stl_file)` loads at /root/reference/src/renderer.cpp:239); run in the build container:

    python tests/golden/make_mesh_golden.py
"""
import hashlib
import os
import re
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STL_DIR = "/root/reference/config/stl"


def load_stl(path):
    b = open(path, "rb").read()
    if b[:5] == b"solid" and b"facet" in b[:2000]:
        v = np.array(re.findall(rb"vertex\s+(\S+)\s+(\S+)\s+(\S+)", b), dtype=np.float64).astype(np.float32)
        return v.reshape(-1, 3, 3)
    n = struct.unpack("<I", b[80:84])[0]
    rec = np.frombuffer(b[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    return rec["v"].copy()


def main():
    out = {}
    for name in ("memoryChip2", "cpu_binary"):
        path = os.path.join(STL_DIR, name + ".stl")
        out[name] = load_stl(path)
        out[name + "_sha256"] = np.array(hashlib.sha256(open(path, "rb").read()).hexdigest())
        print(name, out[name].shape, out[name].reshape(-1, 3).min(0), out[name].reshape(-1, 3).max(0))
    np.savez_compressed(os.path.join(HERE, "meshes_config2.npz"), **out)


if __name__ == "__main__":
    main()
