"""Motion-search time of the 4K r=64 4-reference configuration for a given tile shape (BVC_ME_SHAPE=nb,nby)."""
import os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import basic_video_codec_b200 as bvc
from tests import synth
W, H, BS, R, QP, IP, N, LANES = 3840, 2160, 16, 64, 4, 8, 32, 4
frames = synth.moving_clip(2160, H, W, N, step=3, clamp=48, noise=2)
out = np.empty(N * W * H // 2, np.uint8)
with bvc.Context(W, H, BS, R, QP, 4, False, False, IP, device=0, max_lanes=LANES) as ctx:
    ctx.set_lane_groups(1)
    ctx.clip_upload(frames)
    ctx.encode_clip_resident(N, out)
    _, ln = ctx.encode_clip_resident(N, out)
    kt, clip = ctx.last_kernel_times()
    work = sum(ctx.me_work_per_frame(min(k, 4)) for k in range(1, IP)) * (N // IP)
    print(f"shape={os.environ.get('BVC_ME_SHAPE','default')}: clip {clip:.1f} ms, me {kt['me'][0]:.1f} ms = {work/kt['me'][0]/1e9:.1f} Tpx/s "
          f"({work/kt['me'][0]/1e9/72.1:.3f} of peak), sha={hashlib.sha256(out[:ln].tobytes()).hexdigest()[:12]}", flush=True)
