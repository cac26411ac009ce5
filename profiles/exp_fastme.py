"""FastME (BASELINE config 2: CIF, i=16, 4 references): clip time of every evaluation mode (bvc_set_fastme_direct: 0 auto,
3 window walk, 4 transfer tables, 2 serial SAD-map walk, 1 direct candidate evaluation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import basic_video_codec_b200 as bvc
from tests import synth
W, H, n, bs, qp, ip, nref, lanes = 352, 288, 300, 16, 3, 8, 4, 38
frames = synth.moving_clip(77, H, W, n, step=2, clamp=24)
out = np.empty(n * W * H // 2 + (1 << 20), np.uint8)
with bvc.Context(W, H, bs, 16, qp, nref, True, False, ip, device=0, max_lanes=lanes) as ctx:
    ctx.set_lane_groups(1)
    for direct in (0, 3, 4, 2, 1):   # auto, window walk, transfer tables, serial walk on the SAD map, direct evaluation
        ctx.set_fastme_direct(direct)
        for _ in range(2):
            ctx.encode_clip_into(frames, out)
        kt, clip = ctx.last_kernel_times()
        print(f"direct={direct}: clip {clip:.2f} ms", {k: round(v[0], 2) for k, v in kt.items()}, flush=True)
