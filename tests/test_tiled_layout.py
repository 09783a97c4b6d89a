"""The column-blocked layout of refinement-level planes (linemod_pose_estimation_b200/csrc/lm_kernels.cuh::tiled_nibble_index)
restated in numpy: a 16 x 16 window read through the two-chunk addressing of k_refine_nib's address phase must equal the
reference's flat read of the same window -- including its over-read past the end of a row (next row), past the bottom of a
phase matrix (next phase) and past the last phase (zero tail).  This pins the layout contract on the CPU;
the kernels themselves are held to the oracle by tests/test_gpu_parity.py."""
import numpy as np
import pytest


def tiled_index(W, Hh, phase, row, col):
    return phase * W * Hh + (col >> 4) * Hh * 16 + row * 16 + (col & 15)


def build(T, W, H, rng):
    WH, Hh = W * H, H + 16
    plane_stride = (T * T * WH + WH + 16 * W + 16 + 15) // 16 * 16          # the reference's flat plane + zero tail
    flat = np.zeros(plane_stride, np.uint8)
    flat[:T * T * WH] = rng.integers(0, 5, T * T * WH)
    nib_plane = (T * T * W * Hh + 256 + 31) // 32 * 32
    blocked = np.zeros(nib_plane, np.uint8)
    for g in range(T * T):                                                    # what k_spread_all stores, word by word
        for a in range(H):
            for c8 in range(0, W, 8):
                word = flat[g * WH + a * W + c8:g * WH + a * W + c8 + 8]
                n0 = tiled_index(W, Hh, g, a, c8)
                blocked[n0:n0 + 8] = word
                if a < 16 and g > 0:                                          # halo of the previous phase
                    h = n0 - (W * Hh - H * 16)
                    blocked[h:h + 8] = word
    return flat, blocked, Hh, nib_plane


@pytest.mark.parametrize("T,W,H", [(5, 128, 96), (5, 64, 48), (2, 80, 48), (4, 32, 16), (8, 16, 17)])
def test_windows_read_through_the_blocked_layout_equal_flat_reads(T, W, H):
    rng = np.random.default_rng(T * 1000 + W)
    flat, blocked, Hh, nib_plane = build(T, W, H, rng)
    WH = W * H
    block_bytes, phase_bytes = Hh * 8, W * Hh // 2
    cases = [(ph, row, col) for ph in (0, T * T - 1) for row in (0, H - 16, H - 1) for col in (0, 7, 8, 15, W - 16, W - 9, W - 1)]
    cases += [(int(rng.integers(0, T * T)), int(rng.integers(0, H)), int(rng.integers(0, W))) for _ in range(1500)]
    for ph, row, col in cases:
        base = ph * WH + row * W + col
        want = np.stack([flat[base + r * W:base + r * W + 16] for r in range(16)])
        cb, sh = col >> 4, col & 15
        phase0 = ph * phase_bytes
        b0 = phase0 + cb * block_bytes + row * 8
        b1 = b0 + block_bytes if cb + 1 < (W >> 4) else phase0 + (row + 1) * 8   # past the last block: block 0, one row down
        got = []
        for r in range(16):
            c0 = blocked[(b0 + r * 8) * 2:(b0 + r * 8) * 2 + 16]
            c1 = blocked[(b1 + r * 8) * 2:(b1 + r * 8) * 2 + 16]
            got.append(np.concatenate([c0, c1])[sh:sh + 16])
        assert np.array_equal(np.stack(got), want), (ph, row, col)
    zero_run = nib_plane // 2 - 128                                           # where features outside the image point
    assert not blocked[zero_run * 2:zero_run * 2 + 256].any()
