# host-side cost per device-path frame (no GPU wait): time 2000 enqueue calls, sync at the end only
import sys,time,ctypes as C
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch, bench
from linemod_pose_estimation_b200 import Detector,_capi
dev=torch.device('cuda',0); torch.cuda.set_device(0)
views=bench.rendered_views(); det=Detector()
bench.fill_templates(lambda cid,b,d,m: det.addTemplate([b,d],cid,m)[0], lambda cid,p: det.addSyntheticTemplate(p,cid), views, 300)
frames=bench.make_frames(views,8)
dev_frames=[(torch.from_numpy(b).to(dev), torch.from_numpy(d.view(np.int16)).to(dev)) for b,d in frames]
ptrs=[(C.c_void_p*2)(fb.data_ptr(),fd.data_ptr()) for fb,fd in dev_frames]
qarr,qk=_capi.query_array(bench.QUERIES); lib=_capi.lib()
st=[torch.cuda.current_stream(), torch.cuda.Stream(device=dev)]
rec,cap=C.c_void_p(),C.c_size_t()
def step(i,use_ctx):
    k=i&1
    if use_ctx:
        with torch.cuda.stream(st[k]):
            lib.lm_match_device_multi_lane(det._h,k,ptrs[i%8],2,480,640,qarr,2,C.c_void_p(st[k].cuda_stream),C.byref(rec),C.byref(cap))
    else:
        lib.lm_match_device_multi_lane(det._h,k,ptrs[i%8],2,480,640,qarr,2,C.c_void_p(st[k].cuda_stream),C.byref(rec),C.byref(cap))
for use_ctx in (True,False):
    for i in range(20): step(i,use_ctx)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for i in range(2000): step(i,use_ctx)
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print('ctx' if use_ctx else 'noctx','enqueue us/frame',(t1-t0)/2000*1e6,'total us/frame',(t2-t0)/2000*1e6)
