#!/usr/bin/env python3
"""I-frame wavefront: one warp per block pair (BVC_IQUAD=0) against four (default).  I-step time (kernel class tq_i, one lane
group so that the per-class events are valid) and the stream hash, 1080p, i = 16 and 8.
Usage: python profiles/exp_iframe.py [lanes ...]"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import basic_video_codec_b200 as bvc  # noqa: E402
from tests import synth  # noqa: E402

W, H, IP = 1920, 1088, 30
lanes_list = [int(x) for x in sys.argv[1:]] or [20, 3]
base = synth.moving_clip(1080, H, W, IP * max(lanes_list), step=6, clamp=96, noise=2)
for bs, r, qp in ((16, 4, 4), (8, 2, 3)):
    for lanes in lanes_list:
        n = lanes * IP
        frames = base[:n]
        out = np.empty(n * W * H, np.uint8)
        ref = None
        for quad in (0, 1, 2):   # 2 = four warps whatever the number of CTAs in flight
            os.environ["BVC_IQUAD"] = str(quad)
            with bvc.Context(W, H, bs, r, qp, 1, False, False, IP, device=0, max_lanes=lanes) as ctx:
                ctx.clip_upload(frames)
                ctx.set_lane_groups(1)
                for _ in range(4):
                    try:
                        _, ln = ctx.encode_clip_resident(n, out)
                    except MemoryError:   # first call on busy content: the library enlarges its staging buffer and asks again
                        pass
                kt, clip_ms = ctx.last_kernel_times()
                h = hashlib.sha256(out[:ln].tobytes()).hexdigest()[:16]
                ref = ref or h
                print(json.dumps({"bs": bs, "lanes": lanes, "quad": quad, "i_step_ms": round(kt["tq_i"][0], 4), "clip_ms": round(clip_ms, 3),
                                  "same_stream": h == ref}), flush=True)
