import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
import basic_video_codec_b200 as bvc
from tests import synth
W,H,n,bs,qp,ip,nref,lanes=352,288,300,16,3,8,4,38
frames=synth.moving_clip(77,H,W,n,step=2,clamp=24)
out=np.empty(n*W*H//2+(1<<20),np.uint8)
with bvc.Context(W,H,bs,16,qp,nref,True,False,ip,device=0,max_lanes=lanes) as ctx:
    ctx.set_lane_groups(1)
    for _ in range(2): ctx.encode_clip_into(frames,out)
    kt,clip=ctx.last_kernel_times()
    print(os.environ.get("BVC_DBG_SKIPMAP"),os.environ.get("BVC_DBG_SKIPWALK"),"clip",round(clip,2),{k:round(v[0],2) for k,v in kt.items()})
    ctx.set_fastme_direct(True)
    for _ in range(2): ctx.encode_clip_into(frames,out)
    kt,clip=ctx.last_kernel_times()
    print("direct: clip",round(clip,2),{k:round(v[0],2) for k,v in kt.items()})
