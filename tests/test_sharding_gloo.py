"""Multi-rank plumbing of the sharded matcher (frame broadcast, fixed-capacity all-gather of survivor blocks,
finalisation on rank 0) on CPU with torch.distributed / gloo, world_size 2.  The rank-local CUDA matcher is replaced
by the oracle restricted to the rank's template shard (tests may use the oracle; the product never does); the
exchange and lm_finalize_raw are the product's own code."""
import os
import socket
import sys

import numpy as np
import pytest

import common

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, capacity, out_dir):
    sys.path.insert(0, common.ROOT)
    sys.path.insert(0, os.path.join(common.ROOT, "tests"))
    import torch.distributed as dist
    from common import O, synth
    from linemod_pose_estimation_b200 import Detector
    from linemod_pose_estimation_b200.sharding import ShardedMatcher, pack_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc, views = common.build_oracle(n_views=6, n_random=40, seed=71, classes=("a", "b"))
        det = Detector()                      # host-only use: template bookkeeping + lm_finalize_raw
        common.copy_templates(orc, det)
        if rank == 0:
            bgr, depth, _ = synth.compose_scene(1001, views[:4], rows=240, cols=320)
        else:
            bgr, depth = np.zeros((240, 320, 3), np.uint8), np.zeros((240, 320), np.uint16)
        frame = [torch.from_numpy(bgr), torch.from_numpy(depth.view(np.int16))]

        def local_match(tensors, queries):
            b = tensors[0].numpy()
            d = tensors[1].numpy().view(np.uint16)
            parts = []
            for q, (thr, ids) in enumerate(queries):
                orc.match([b, d], thr, class_ids=ids, keep_candidates=True)
                raw = orc.last_raw()
                # this rank's template shard: canonical index % world == rank ("b" alone starts at its own offset 0,
                # so recover the canonical index from class + template id)
                canon = raw["template_id"] + np.where(raw["class_index"] == 1, orc.num_templates("a"), 0)
                mine = raw[canon % world == rank].copy()
                mine["order_key"] |= np.uint32(q << 28)   # query tag, as the CUDA matcher emits it
                parts.append(mine)
            return torch.from_numpy(pack_block(np.concatenate(parts), 1 << 14))

        sm = ShardedMatcher(local_match, det.finalize_raw, rank, world, capacity=capacity)
        queries = [(70.0, []), (60.0, ["b"])]
        got = sm.match(frame, queries)
        if rank == 0:
            for g, (thr, ids) in zip(got, queries):
                want = orc.match([bgr, depth], thr, class_ids=ids)
                assert len(want) > 1
                common.assert_matches_equal(g, want)
            np.save(os.path.join(out_dir, "ok_%d.npy" % capacity), np.array([len(got[0]), sm.capacity]))
        else:
            assert got is None
            assert np.array_equal(frame[0].numpy(), synth.compose_scene(1001, views[:4], rows=240, cols=320)[0])  # broadcast arrived
        # streamed exchange: the survivor blocks of several frames travel in one all-gather
        if capacity >= 64:
            from linemod_pose_estimation_b200.sharding import unpack_blocks
            thresholds = (75.0, 68.0, 62.0)
            for slot, thr in enumerate(thresholds):
                sm.stage_block(local_match(frame, [(thr, [])]), slot, len(thresholds))
            recv = sm.gather_staged().numpy()
            assert recv.shape[:2] == (world, len(thresholds))
            for slot, thr in enumerate(thresholds):
                raws, need = unpack_blocks(np.ascontiguousarray(recv[:, slot]).reshape(-1), world, sm.capacity)
                assert need == 0
                common.assert_matches_equal(det.finalize_raw(np.concatenate(raws)),
                                            orc.match([frame[0].numpy(), frame[1].numpy().view(np.uint16)], thr), "slot %d" % slot)
            # ... and rank 0 finalises the whole gathered chunk in one library call (what match_stream does)
            import ctypes as C

            from linemod_pose_estimation_b200 import _capi
            two = [(70.0, []), (60.0, ["b"])]
            for slot in range(len(thresholds)):
                sm.stage_block(local_match(frame, two), slot, len(thresholds))
            recv = np.ascontiguousarray(sm.gather_staged().numpy())
            n_slots, block_bytes = recv.shape[1], recv.shape[2]
            outp, offs = C.c_void_p(), (C.c_size_t * (n_slots * 2 + 1))()
            status = np.zeros(n_slots, np.uint8)
            _capi.check(_capi.lib().lm_finalize_gathered(det._h, recv.ctypes.data, world, n_slots, block_bytes, n_slots * block_bytes,
                                                         sm.capacity, 2, C.byref(outp), offs, status.ctypes.data))
            allm = det._take(outp, offs[n_slots * 2])
            assert not status.any()
            for slot in range(n_slots):
                for q, (thr, ids) in enumerate(two):
                    common.assert_matches_equal(allm[offs[slot * 2 + q]:offs[slot * 2 + q + 1]],
                                                orc.match([frame[0].numpy(), frame[1].numpy().view(np.uint16)], thr, class_ids=ids),
                                                "gathered chunk, slot %d query %d" % (slot, q))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("capacity", [4096, 2])
def test_sharded_match_world2(tmp_path, capacity):
    """capacity=2 forces the second, larger gather round."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, capacity, str(tmp_path)), nprocs=2, join=True)
    res = np.load(os.path.join(str(tmp_path), "ok_%d.npy" % capacity))
    assert res[0] > 4
    if capacity == 2:
        assert res[1] > 2


def test_block_roundtrip():
    from linemod_pose_estimation_b200 import RAW_DTYPE
    from linemod_pose_estimation_b200.sharding import pack_block, unpack_blocks
    rng = np.random.default_rng(0)
    raws = []
    for n in (0, 3, 10):
        r = np.zeros(n, RAW_DTYPE)
        r["order_key"] = rng.integers(0, 100, n)
        r["score"] = rng.integers(0, 252, n)
        raws.append(r)
    blocks = np.concatenate([pack_block(r, 16) for r in raws])
    out, need = unpack_blocks(blocks, 3, 16)
    assert need == 0 and all(np.array_equal(a, b) for a, b in zip(out, raws))
    blocks = np.concatenate([pack_block(r, 4) for r in raws])
    out, need = unpack_blocks(np.concatenate([b for b in [pack_block(r, 16)[:16 + 4 * 32] for r in raws]]), 3, 4)
    assert need == 10
