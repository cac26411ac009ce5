"""FastME search class against the number of references and half-pel, window walk (mode 3) and transfer tables (mode 4),
CIF, 38 lanes.  Usage: python profiles/exp_fastme_scale.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import basic_video_codec_b200 as bvc
from tests import synth
W, H, n, bs, qp, ip, lanes = 352, 288, 300, 16, 3, 8, 38
frames = synth.moving_clip(77, H, W, n, step=2, clamp=24)
out = np.empty(n * W * H // 2 + (1 << 20), np.uint8)
for frac in (False, True):
    for nref in (1, 2, 4):
        with bvc.Context(W, H, bs, 16, qp, nref, True, frac, ip, device=0, max_lanes=lanes) as ctx:
            ctx.set_lane_groups(1)
            for mode in (3, 4):
                ctx.set_fastme_direct(mode)
                for _ in range(2):
                    ctx.encode_clip_into(frames, out)
                kt, clip = ctx.last_kernel_times()
                print(f"frac={frac} nref={nref} mode={mode}: me {kt['me'][0]:.2f} ms  ({kt['me'][0] / 7 / 396 * 1e3:.2f} us per block)", flush=True)
