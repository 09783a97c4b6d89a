"""Exhaustive proof (zero-one principle) that the compare-exchange network in csrc/lm_median_net.h selects the median
of 25: all 2^25 binary inputs are pushed through the list parsed from the header the CUDA kernel compiles."""
import os
import re

import numpy as np

import common

HEADER = os.path.join(common.ROOT, "linemod_pose_estimation_b200", "csrc", "lm_median_net.h")


def network():
    text = open(HEADER).read()
    text = text[text.index("#define LM_MEDIAN25_NET"):]
    return [(int(a), int(b)) for a, b in re.findall(r"X\((\d+),\s*(\d+)\)", text)]


def test_network_selects_median_of_every_binary_input():
    net = network()
    assert len(net) == 99 and all(0 <= a < b < 25 for a, b in net)
    chunk = 1 << 21
    for base in range(0, 1 << 25, chunk):
        idx = np.arange(base, base + chunk, dtype=np.uint32)
        p = [((idx >> k) & 1).astype(np.uint8) for k in range(25)]
        ones = np.zeros(chunk, np.uint8)
        for k in range(25):
            ones += p[k]
        for a, b in net:
            lo, hi = np.minimum(p[a], p[b]), np.maximum(p[a], p[b])
            p[a], p[b] = lo, hi
        want = (ones >= 13).astype(np.uint8)  # 13th smallest of 25 is 1 iff at least 13 ones
        assert np.array_equal(p[12], want), "network fails near input %d" % base


def test_network_on_random_bytes():
    net = network()
    rng = np.random.default_rng(0)
    v = rng.integers(0, 256, (25, 20000)).astype(np.uint8)
    p = [v[k].copy() for k in range(25)]
    for a, b in net:
        lo, hi = np.minimum(p[a], p[b]), np.maximum(p[a], p[b])
        p[a], p[b] = lo, hi
    assert np.array_equal(p[12], np.sort(v, axis=0)[12])
