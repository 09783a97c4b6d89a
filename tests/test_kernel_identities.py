"""Arithmetic identities the CUDA kernels rely on, restated in numpy / Python integers (no GPU).  Each one replaces a slower
but obviously right formulation inside a kernel; the kernels themselves are held to the oracle by tests/test_gpu_parity.py.

* k_refine_nib's address phase takes x / T as the high word of x * ceil(2^32 / T) (lm_match.cu refine_feature_address).
* k_dn_fused takes medianBlur(5) of one-hot normal codes by counting: 5-bit counters per rank, prefix sums by one
  multiplication, "prefix >= 13" as bit 4 of (prefix + 3) (lm_frontend_fused.cu dn_tile_count).
* k_refine_nib finds the smallest score that passes `score * 100 / (4 nf) < threshold` (f32, two roundings) with one ballot
  over 32 candidate scores instead of a sequential search (lm_match.cu min_passing_score).
"""
import numpy as np


def test_division_by_multiply_high_is_exact_for_the_coordinates_that_occur():
    for T in range(2, 17):
        m = 0xFFFFFFFF // T + 1                      # RefineLevel::inv_T as lm_detector.cu computes it
        x = np.arange(0, 16384, dtype=np.uint64)      # feature coordinates are < 8192
        assert np.array_equal((x * np.uint64(m)) >> np.uint64(32), x // np.uint64(T)), T


def _ffs(v):
    return 0 if v == 0 else (int(v) & -int(v)).bit_length()


def _median_by_counting(codes):
    h0 = h1 = 0
    for c in codes:
        k = _ffs(c)                                   # rank: 0 for code 0, j + 1 for 1 << j
        if k < 6:
            h0 += 1 << (5 * k)
        else:
            h1 += 1 << (5 * (k - 6))
    p0 = (h0 * 0x02108421) & 0xFFFFFFFF               # field j: count of ranks <= j
    t0 = (p0 + 0x06318C63) & 0x21084210               # bit 5j + 4: that count >= 13
    if t0:
        rank = ((_ffs(t0) - 1) * 13) >> 6
    else:
        p1 = ((h1 + ((p0 >> 25) & 31)) * 0x421) & 0xFFFFFFFF
        t1 = (p1 + 0xC63) & 0x4210
        assert t1
        rank = 6 + (((_ffs(t1) - 1) * 13) >> 6)
    return 0 if rank == 0 else 1 << (rank - 1)


def test_median_of_25_one_hot_codes_by_counting():
    rng = np.random.default_rng(3)
    values = np.array([0, 1, 2, 4, 8, 16, 32, 64, 128])
    for _ in range(20000):
        p = rng.dirichlet(np.ones(9) * rng.choice([0.2, 1.0, 5.0]))
        codes = values[rng.choice(9, size=25, p=p)]
        assert _median_by_counting(codes) == np.sort(codes)[12]
    for v in values:                                  # a constant window, and 12 / 13 splits around every rank boundary
        assert _median_by_counting([v] * 25) == v
    for lo, hi in zip(values[:-1], values[1:]):
        assert _median_by_counting([lo] * 13 + [hi] * 12) == lo
        assert _median_by_counting([lo] * 12 + [hi] * 13) == hi


f32 = np.float32


def _fails(s, den, thr):
    return f32(f32(f32(s) * f32(100.0)) / den) < thr


def _sequential(thr, nf):
    den, cap = f32(4 * nf), 4 * nf + 1
    s = max(0, int(np.floor(f32(f32(thr * den) * f32(0.01)))) - 2)
    while s < cap and _fails(s, den, thr):
        s += 1
    while s > 0 and not _fails(s - 1, den, thr):
        s -= 1
    return s


def _ballot(thr, nf):
    den, cap = f32(4 * nf), 4 * nf + 1
    s0 = max(0, int(np.floor(f32(f32(thr * den) * f32(0.01)))) - 2)
    m = [(s0 + lane >= cap) or not _fails(s0 + lane, den, thr) for lane in range(32)]
    if any(m) and (not m[0] or s0 == 0):
        return min(cap, s0 + m.index(True))
    return _sequential(thr, nf)


def test_smallest_passing_score_by_one_ballot():
    rng = np.random.default_rng(0)
    cases = [(f32(t), nf) for t in (0.0, 0.5, 50.0, 88.0, 92.0, 94.0, 99.9, 100.0) for nf in (1, 2, 31, 62, 63, 126, 189, 252)]
    cases += [(f32(rng.uniform(0, 100)), int(rng.integers(1, 253))) for _ in range(5000)]
    for thr, nf in cases:
        assert _ballot(thr, nf) == _sequential(thr, nf), (thr, nf)
