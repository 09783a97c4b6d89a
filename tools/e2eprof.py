# where the host time of lm_match_batch_multi goes (LM_HOST_PROFILE=1 prints the per-frame split to stderr)
import os, sys, time
os.environ["LM_HOST_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from linemod_pose_estimation_b200 import Detector, _capi
views = bench.rendered_views(); det = Detector()
bench.fill_templates(lambda cid, b, d, m: det.addTemplate([b, d], cid, m)[0], lambda cid, p: det.addSyntheticTemplate(p, cid), views, bench.TEMPLATES_PER_CLASS)
frames = bench.make_frames(views, 128)
pinned = []
for b, d in frames:
    pb = _capi.pinned_empty(b.shape, b.dtype); pb[...] = b
    pd = _capi.pinned_empty(d.shape, d.dtype); pd[...] = d
    pinned.append([pb, pd])
for rep in range(4):
    t0 = time.perf_counter()
    for c in range(4):
        det.match_batch_multi(pinned, bench.QUERIES)
    dt = time.perf_counter() - t0
    print("rep", rep, "us/frame", dt / (4 * 128) * 1e6, flush=True)

# pinned host -> device copy rate of one frame's two images (what bounds e2e at N = 1) and of a large block
import torch
for nbytes in (921600, 614400, 64 << 20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); dv = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3): dv.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dv.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D", nbytes, "bytes:", nbytes * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9, "GB/s", flush=True)
