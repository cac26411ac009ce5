import sys, os, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import basic_video_codec_b200 as bvc
from tests import synth
W,H,BS,R,QP,IP,N=1920,1088,16,32,4,30,600
lanes=int(sys.argv[1]) if len(sys.argv)>1 else 20
frames=synth.moving_clip(1080,H,W,N,step=6,clamp=96,noise=2)
out=np.empty(N*W*H//2,np.uint8)
ctx=bvc.Context(W,H,BS,R,QP,1,False,False,IP,device=0,max_lanes=lanes)
ctx.clip_upload(frames)
ref=None
for g in (1, 2, 3, 4):
    ctx.set_lane_groups(g)
    for _ in range(2): ctx.encode_clip_resident(N,out)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(3): _,ln=ctx.encode_clip_resident(N,out)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
    h=hashlib.sha256(out[:ln].tobytes()).hexdigest()[:12]
    if ref is None: ref=h
    kt,clip=ctx.last_kernel_times()
    print(f"lanes={lanes} groups={g}: {dt*1e3:.2f} ms/clip {N/dt:.0f} f/s same={h==ref} dev={clip:.2f} me={kt['me'][0]:.1f} tq={kt['tq_p'][0]:.1f}",flush=True)
