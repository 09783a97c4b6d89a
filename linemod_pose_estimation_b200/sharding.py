"""Multi-GPU layout of the matcher: template set sharded across ranks, frame broadcast, matches gathered to rank 0.

One process per GPU (torch.distributed, NCCL over NVLink / NVSwitch; gloo for the CPU tests of this plumbing).
Templates are independent given a frame's linear memories ([OCV] Detector::matchClass has no cross-template state,
SURVEY.md section 8e), so the only exchanges are
    C1  broadcast of the raw frame (1.5 MB at 640x480; every rank rebuilds the identical front end itself), and
    C2  a fixed-capacity all-gather of each rank's survivor block {count, capacity, overflow, n_cands | raw records},
after which rank 0 restores the reference's emission order and runs the same std::sort + std::unique
(lm_finalize_raw) -- global, because std::unique acts across templates (App. D-7) -- for the unchanged NMS / ICP
stage of the reference (/root/reference/src/rgbdDetector.cpp:36-144, 462-574).
torch is plumbing here (device tensors, streams, collectives); all matching runs in liblinemod_b200.so.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._capi import RAW_DTYPE, RESULT_HEADER_BYTES

RECORD_BYTES = RAW_DTYPE.itemsize
CANDIDATE_OVERFLOW = 1 << 40   # "records needed" value that stands for: a rank's coarse candidate list was truncated


class _DevicePtr:
    """Zero-copy torch view of library-owned device memory (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_view(ptr, nbytes, device):
    return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)


def pack_block(raw, capacity):
    """Host-side constructor of a survivor block (used by the CPU tests that stand in for the CUDA matcher)."""
    raw = np.ascontiguousarray(raw, dtype=RAW_DTYPE)
    block = np.zeros(RESULT_HEADER_BYTES + capacity * RECORD_BYTES, np.uint8)
    hdr = block[:RESULT_HEADER_BYTES].view(np.uint32)
    hdr[0], hdr[1], hdr[2], hdr[3] = len(raw), capacity, int(len(raw) > capacity), len(raw)
    n = min(len(raw), capacity)
    block[RESULT_HEADER_BYTES:RESULT_HEADER_BYTES + n * RECORD_BYTES] = raw[:n].view(np.uint8).reshape(-1)
    return block


def unpack_blocks(gathered, world, capacity):
    """[world x block] bytes -> (list of raw record arrays, needed capacity if some rank had more than `capacity`)."""
    stride = RESULT_HEADER_BYTES + capacity * RECORD_BYTES
    raws, need = [], 0
    for r in range(world):
        blk = gathered[r * stride:(r + 1) * stride]
        hdr = blk[:RESULT_HEADER_BYTES].view(np.uint32)
        count = int(hdr[0])
        if hdr[2] != 0 and count <= capacity:   # not the record block: the rank's candidate list was truncated
            need = max(need, CANDIDATE_OVERFLOW)
            continue
        if count > capacity:                    # (also when the device block itself overflowed: count keeps counting)
            need = max(need, count)
            continue
        raws.append(blk[RESULT_HEADER_BYTES:RESULT_HEADER_BYTES + count * RECORD_BYTES].view(RAW_DTYPE).copy())
    return raws, need


class ShardedMatcher:
    """Exchange + finalisation around a rank-local matcher.

    local_match(frame_tensors, queries) must return a uint8 tensor holding the rank's survivor block for the whole
    request (records tagged with their query index in order_key >> 28) with room for at least `capacity` records (the
    CUDA library's block is used in place; the CPU tests build one with pack_block)."""

    def __init__(self, local_match, finalize, rank, world, group=None, capacity=4096):
        self.local_match, self.finalize = local_match, finalize
        self.rank, self.world, self.group = rank, world, group
        self.capacity = capacity
        self._gather_buf = None
        self._send_buf = self._recv_buf = None

    def broadcast_frame(self, tensors):
        if self.world > 1:
            for t in tensors:  # bytes on the wire: neither NCCL nor gloo carries 16-bit integers
                dist.broadcast(t.view(torch.uint8), src=0, group=self.group)  # C1
        return tensors

    def gather_async(self, block, slot=0):
        """C2, device side only: all-gather of the header + first `capacity` records of every rank's survivor block.
        Nothing is synchronised, so this is what the device-timed path runs.  `slot` selects the receive buffer (one per
        frame in flight)."""
        nbytes = RESULT_HEADER_BYTES + self.capacity * RECORD_BYTES
        mine = block[:nbytes]
        if self.world == 1:
            return mine
        total = nbytes * self.world
        if self._gather_buf is None:
            self._gather_buf = {}
        buf = self._gather_buf.get(slot)
        if buf is None or buf.numel() != total or buf.device != mine.device:
            buf = self._gather_buf[slot] = torch.empty(total, dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(buf, mine.contiguous(), group=self.group)
        return buf

    def _ensure_send(self, n_slots, device):
        nbytes = RESULT_HEADER_BYTES + self.capacity * RECORD_BYTES
        if self._send_buf is None or self._send_buf.shape != (n_slots, nbytes) or self._send_buf.device != device:
            self._send_buf = torch.empty((n_slots, nbytes), dtype=torch.uint8, device=device)
            self._recv_buf = torch.empty((self.world, n_slots, nbytes), dtype=torch.uint8, device=device)
            self._send_slots = [self._send_buf[i] for i in range(n_slots)]
            self._send_ptrs = [self._send_buf[i].data_ptr() for i in range(n_slots)]
            self._heads = {}
        return nbytes

    def stage_block(self, block, slot, n_slots):
        """Streamed exchange, step 1 (device side, asynchronous): park a frame's survivor block in slot `slot` of this
        rank's send buffer.  One collective then moves `n_slots` frames at once (gather_staged): the per-frame payload
        is a few hundred bytes, so the exchange is launch-latency bound and is batched over frames, not over links."""
        nbytes = self._ensure_send(n_slots, block.device)
        head = self._heads.get(block.data_ptr())   # the library's blocks are few and stable: keep their sliced views
        if head is None:
            head = self._heads[block.data_ptr()] = block[:nbytes]
        self._send_slots[slot].copy_(head, non_blocking=True)

    def gather_staged(self):
        """Streamed exchange, step 2: all-gather of every rank's staged blocks -> tensor [world][n_slots][block bytes]
        (device, asynchronous).  unpack_blocks() on recv[:, slot].reshape(-1) yields frame `slot`'s raw records."""
        if self.world == 1:
            return self._send_buf.unsqueeze(0)
        dist.all_gather_into_tensor(self._recv_buf.view(-1), self._send_buf.view(-1), group=self.group)
        return self._recv_buf

    def gather(self, block):
        """C2 + download: every rank's raw records.  Every rank sees every header, so all ranks agree on whether a
        (rare) second round with a larger capacity is needed without an extra collective."""
        while True:
            host = self.gather_async(block).cpu().numpy()
            raws, need = unpack_blocks(host, self.world, self.capacity)
            if need == 0:
                return raws, 0
            if need > (block.numel() - RESULT_HEADER_BYTES) // RECORD_BYTES:
                return None, need          # the rank-local block (or candidate list) is too small: redo the local match
            self.capacity = int(need * 1.25) + 64

    def grow_local(self, need):
        """Called on every rank when some rank's local survivor block overflowed (`need` records)."""
        raise RuntimeError("survivor block too small for %d records" % need)

    def match(self, frame_tensors, queries):
        """frame_tensors: per-modality tensors, valid on rank 0 (other ranks pass same-shaped buffers).
        queries: [(threshold, [class ids])].  Returns the finalised match lists (one per query) on rank 0, None
        elsewhere."""
        self.broadcast_frame(frame_tensors)
        while True:
            block = self.local_match(frame_tensors, queries)
            raws, need = self.gather(block)
            if need == 0:
                break
            self.grow_local(need)
        if self.rank != 0:
            return None
        raw = np.concatenate(raws) if raws else np.zeros(0, RAW_DTYPE)
        tag = raw["order_key"] >> 28
        return [self.finalize(raw[tag == q]) for q in range(len(queries))]


class ShardedDetector(ShardedMatcher):
    """The CUDA matcher sharded over torch.distributed ranks: rank r keeps templates with canonical index % world == r."""

    def __init__(self, detector, group=None, capacity=4096):
        self.det = detector
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        detector.set_shard(rank, world)
        # the library's device record blocks must hold at least what one exchange slot carries
        self._device_cap = max(int(capacity), 2048)
        self._cand_per_frame = 1 << 16
        detector.set_option("device_out_cap", self._device_cap)
        self.device = torch.device("cuda", torch.cuda.current_device())
        super().__init__(self._local, detector.finalize_raw, rank, world, group, capacity)

    def _local(self, frame_tensors, queries):
        rows, cols = frame_tensors[0].shape[:2]
        ptrs = [t.data_ptr() for t in frame_tensors]
        rec, cap = self.det.match_device_multi(ptrs, rows, cols, queries, stream=torch.cuda.current_stream().cuda_stream)
        return device_view(rec, cap, self.device)

    def grow_local(self, need):
        if need >= CANDIDATE_OVERFLOW:   # a rank's coarse candidate list was truncated: every rank grows it alike
            self._cand_per_frame *= 4
            self.det.set_option("cand_per_frame", self._cand_per_frame)
            return
        self._device_cap = max(self._device_cap, int(need * 1.25) + 64)
        self.det.set_option("device_out_cap", self._device_cap)   # the library re-allocates its record blocks

    def frame_buffers(self, rows, cols, kinds):
        """Device buffers for one frame: uint8 [rows, cols, 3] for ColorGradient, int16-typed uint16 storage for depth."""
        return [torch.empty((rows, cols, 3), dtype=torch.uint8, device=self.device) if k == "cg"
                else torch.empty((rows, cols), dtype=torch.int16, device=self.device) for k in kinds]

    def stage_lane(self, lane, slot, n_slots, stream):
        """stage_block for the CUDA matcher without torch ops: the library copies lane's survivor block into the send
        slot on `stream` (lm_copy_result_block)."""
        from . import _capi
        nbytes = self._ensure_send(n_slots, self.device)
        _capi.check(_capi.lib().lm_copy_result_block(self.det._h, lane, self._send_ptrs[slot], nbytes, stream))

    # ------------------------------------------------------------------ streamed frames
    def match_stream(self, host_frames, queries, kinds=("cg", "dn"), chunk=16, lanes=4):
        """A stream of frames through the sharded matcher, pipelined: per chunk of `chunk` frames ONE broadcast per
        modality (rank 0 uploads its pinned host frames into a device chunk buffer first) and ONE all-gather of the
        survivor blocks; within a chunk the library runs launch sets of "batch_frames" frames on `lanes` streams; the
        upload + broadcast of chunk c+1 overlaps the matching of chunk c (two chunk buffers).

        host_frames: list of per-frame lists of numpy arrays (pinned for full copy speed); their contents matter on
        rank 0 only, every rank must pass the same number of frames of the same shape.
        Returns on rank 0 a list (per frame) of lists (per query) of finalised match arrays; None elsewhere."""
        import ctypes as C

        from . import _capi
        n = len(host_frames)
        if n == 0:
            return [] if self.rank == 0 else None
        rows, cols = host_frames[0][0].shape[:2]
        if getattr(self, "_stream_state", None) is None or self._stream_state["key"] != (rows, cols, kinds, chunk, lanes):
            bufs = [[torch.empty((chunk, rows, cols, 3), dtype=torch.uint8, device=self.device) if k == "cg"
                     else torch.empty((chunk, rows, cols), dtype=torch.int16, device=self.device) for k in kinds]
                    for _ in range(2)]
            self._stream_state = {
                "key": (rows, cols, kinds, chunk, lanes), "bufs": bufs,
                "copy": torch.cuda.Stream(device=self.device),
                "lanes": [torch.cuda.Stream(device=self.device) for _ in range(lanes)],
                "ready": [torch.cuda.Event() for _ in range(2)], "free": [None, None], "keep": [None, None],
            }
            st0 = self._stream_state
            st0["stream_ptrs"] = (C.c_void_p * lanes)(*[s.cuda_stream for s in st0["lanes"]])
            st0["ptrs"] = [(C.c_void_p * (chunk * len(kinds)))(*[bufs[b][m][j].data_ptr() for j in range(chunk)
                                                                 for m in range(len(kinds))]) for b in range(2)]
        st = self._stream_state
        qarr, _qkeep = _capi.query_array(queries)
        lib = _capi.lib()
        n_q = len(queries)
        n_chunks = (n + chunk - 1) // chunk

        def stage(c):   # upload (rank 0) + broadcast of chunk c into buffer c & 1, on the copy stream
            b, lo = c & 1, c * chunk
            g = min(chunk, n - lo)
            with torch.cuda.stream(st["copy"]):
                if st["free"][b] is not None:
                    st["copy"].wait_event(st["free"][b])      # the matching of chunk c-2 has finished reading this buffer
                if self.rank == 0:   # the whole chunk's host -> device copies in one library call (lm_upload_images)
                    imgs, keep = _capi.image_array([host_frames[lo + j][m] for j in range(g) for m in range(len(kinds))])
                    _capi.check(lib.lm_upload_images(self.det._h, imgs, g * len(kinds), st["ptrs"][b],
                                                     C.c_void_p(st["copy"].cuda_stream)))
                    st["keep"][b] = keep   # borrowed host images stay referenced until the buffer is reused
                if self.world > 1:
                    for m in range(len(kinds)):
                        dist.broadcast(st["bufs"][b][m][:g].view(torch.uint8), src=0, group=self.group)
                st["ready"][b].record(st["copy"])

        def run(c):     # matching of chunk c on the lane streams, survivors staged, one all-gather -> device tensor
            b, lo = c & 1, c * chunk
            g = min(chunk, n - lo)
            for s in st["lanes"]:
                s.wait_event(st["ready"][b])
            nbytes = self._ensure_send(chunk, self.device)
            ptrs = st["ptrs"][b]
            _capi.check(lib.lm_match_device_stream(self.det._h, ptrs, g, len(kinds), rows, cols, qarr, n_q, st["stream_ptrs"],
                                                   lanes, self._send_ptrs[0], nbytes))
            main = st["lanes"][0]
            for s in st["lanes"][1:]:
                main.wait_stream(s)
            with torch.cuda.stream(main):
                recv = self.gather_staged()
                host = torch.empty(recv.shape, dtype=torch.uint8, pin_memory=True)
                host.copy_(recv, non_blocking=True)
                done = torch.cuda.Event()
                done.record(main)
            for s in st["lanes"][1:]:
                s.wait_stream(main)                           # the send buffer is reused by the next chunk
            st["free"][b] = done
            return host, done, g, lo

        def redo(frame):   # rare: this frame again through the per-frame path, which grows its own exchange
            keep = self.capacity
            if not hasattr(self, "_redo_bufs") or self._redo_bufs[0].shape[:2] != (rows, cols):
                self._redo_bufs = self.frame_buffers(rows, cols, kinds)
            if self.rank == 0:
                for m, k in enumerate(kinds):
                    a = host_frames[frame][m]
                    self._redo_bufs[m].copy_(torch.from_numpy(a if k == "cg" else a.view(np.int16)))
            res = self.match(self._redo_bufs, queries)
            self.capacity = keep
            return res

        deferred = []   # (index into results, frame) of frames that take the per-frame fallback once the stream has drained:
                        # the fallback uses lane 0 of the same handle, which frames of later chunks are still using now

        def finish(host, done, g, lo):
            done.synchronize()
            arr = host.numpy()                                   # [world][chunk][block bytes]
            world, slots, block_bytes = arr.shape
            hdr = np.ascontiguousarray(arr[:, :g, :RESULT_HEADER_BYTES]).view(np.uint32).reshape(world, g, 4)
            # every rank sees every header, so all ranks agree on which frames outgrew the staged capacity (or, flagged by
            # the device, a rank's candidate list) and defer the same frames in the same order
            over = (hdr[:, :, 0] > self.capacity).any(axis=0) | (hdr[:, :, 2] != 0).any(axis=0)
            base = len(results)
            for j in range(g):
                if over[j]:
                    deferred.append((base + j, lo + j))
            if self.rank != 0:
                return [None] * g
            # rank 0: the whole chunk is ordered / de-duplicated in one library call (lm_finalize_gathered)
            outp = C.c_void_p()
            offs = (C.c_size_t * (g * n_q + 1))()
            status = np.zeros(g, np.uint8)
            _capi.check(lib.lm_finalize_gathered(self.det._h, arr.ctypes.data, world, g, block_bytes, slots * block_bytes,
                                                 self.capacity, n_q, C.byref(outp), offs, status.ctypes.data))
            allm = self.det._take(outp, offs[g * n_q])
            assert all((status[j] != 0) == bool(over[j]) for j in range(g))
            return [None if over[j] else [allm[offs[j * n_q + q]:offs[j * n_q + q + 1]] for q in range(n_q)] for j in range(g)]

        results = []
        stage(0)
        pending = None
        for c in range(n_chunks):
            if c + 1 < n_chunks:
                stage(c + 1)
            cur = run(c)
            if pending is not None:
                results.extend(finish(*pending))
            pending = cur
        results.extend(finish(*pending))
        if deferred:
            torch.cuda.synchronize(self.device)    # every lane stream idle: lane 0 is free for the per-frame path
            for at, frame in deferred:
                results[at] = redo(frame)
        return results if self.rank == 0 else None
